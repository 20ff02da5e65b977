// Host-side mirror of the reference's `zkproofs` call surface (reference src/zkproofs/mod.rs:29-43), in
// C++ because the reference's own toolchain (Rust) is absent from the build image.  Same names, argument
// meaning, error behaviour and serde wire format as the Rust structs; all big-integer loops go through the
// C ABI of include/zkp_b200.h (CUDA, no CPU fallback).  Every proof type also has a `*_batch` form -- the
// one thing the engine adds (the reference takes exactly one statement per call and fans out on rayon).
//
//   reference                                      here
//   RangeProofNi::prove / verify / verify_self      RangeProofNi::prove / verify / verify_self (+ _batch)
//   NiCorrectKeyProof::proof / verify               NiCorrectKeyProof::proof / verify (+ verify_batch)
//   {Zero,Ciphertext,Mul,Verlin}Proof::prove/verify same (+ _batch)
//   Err(IncorrectProof)                             throw IncorrectProof   (batch forms return 0/1 per proof)
//   panic (assert_eq!, unwrap, index out of range)  throw ReferencePanic
#pragma once
#include <cstdio>
#include <cstring>
#include <exception>
#include <memory>
#include <stdexcept>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/zkp_b200.h"
#include "bigint.hpp"
#include "json.hpp"
#include "sha256.hpp"

namespace zkproofs {

using zkhost::BigInt;
using zkhost::Json;
using ByteSource = BigInt::ByteSource;

struct IncorrectProof : std::exception {  // errors.rs:5-13
  const char* what() const noexcept override { return "given proof doesn't match a statement"; }
};
struct ReferencePanic : std::logic_error {  // where the reference panics instead of returning Err
  using std::logic_error::logic_error;
};

constexpr size_t SECURITY_PARAMETER = 128;  // range_proof_ni.rs:23
constexpr size_t M2 = 11;                   // correct_key_ni.rs:29
static const uint8_t SALT_STRING[4] = {75, 90, 101, 110};  // correct_key_ni.rs:28

inline ByteSource os_rng() {  // OsRng
  return [](uint8_t* p, size_t n) {
    FILE* f = fopen("/dev/urandom", "rb");
    const size_t got = f ? fread(p, 1, n, f) : 0;
    if (f) fclose(f);
    if (got != n) throw std::runtime_error("cannot read /dev/urandom");
  };
}

struct EncryptionKey {  // kzen-paillier EncryptionKey {n, nn}
  BigInt n, nn;
  EncryptionKey() {}
  explicit EncryptionKey(const BigInt& n_) : n(n_), nn(n_ * n_) {}
  bool operator==(const EncryptionKey& o) const { return n == o.n; }
  Json to_json() const { return Json::object().set("n", Json::string(n.to_dec())); }  // minimal form (RECALLED)
  static EncryptionKey from_json(const Json& j) { return EncryptionKey(BigInt::from_dec(j.at("n").as_str())); }
};
struct DecryptionKey {
  BigInt p, q;
};

// fn(i) for i in [0, count) on up to `threads` host threads (0 = hardware concurrency): per-item host BigInt work around a device call
template <class F>
inline void parallel_for(size_t count, unsigned threads, F fn) {
  if (threads == 0) threads = std::max(1u, std::thread::hardware_concurrency());
  threads = (unsigned)std::min<size_t>(threads, std::max<size_t>(count, 1));
  if (threads <= 1) {
    for (size_t i = 0; i < count; ++i) fn(i);
    return;
  }
  std::vector<std::thread> pool;
  std::atomic<size_t> next(0);
  for (unsigned t = 0; t < threads; ++t)
    pool.emplace_back([&] {
      for (size_t i = next.fetch_add(1); i < count; i = next.fetch_add(1)) fn(i);
    });
  for (auto& th : pool) th.join();
}

inline size_t round4(size_t limbs) { return (limbs + 3) / 4 * 4; }
inline size_t limbs_for_bits(size_t bits) { return round4((bits + 31) / 32); }

// serde helpers: src/serialize.rs (decimal strings) and curv's native BigInt serde (hex of to_bytes, RECALLED)
inline Json ser_dec(const BigInt& x) { return Json::string(x.to_dec()); }
inline BigInt de_dec(const Json& j, bool panic_on_error) {
  BigInt v;
  if (!BigInt::parse_dec(j.as_str(), v)) {
    if (panic_on_error) throw ReferencePanic("from_str_radix(..).unwrap() on a malformed decimal string (serialize.rs:69)");
    throw std::runtime_error("invalid decimal BigInt");
  }
  return v;
}
inline Json ser_native(const BigInt& x) { return Json::string(x.to_hex_bytes()); }
inline BigInt de_native(const Json& j) {
  BigInt v;
  if (!BigInt::parse_hex_bytes(j.as_str(), v)) throw std::runtime_error("invalid hex BigInt");
  return v;
}
inline Json ser_vec(const std::vector<BigInt>& v) {
  Json a = Json::array();
  for (auto& x : v) a.push(ser_dec(x));
  return a;
}
inline std::vector<BigInt> de_vec(const Json& j) {
  std::vector<BigInt> v;
  for (auto& x : j.arr) v.push_back(de_dec(x, true));
  return v;
}

// One engine context = one GPU.  Single-threaded, like zkp_ctx.
class Engine {
 public:
  explicit Engine(int device = 0) {
    int rc = zkp_ctx_create(device, nullptr, &h_);
    if (rc != ZKP_OK) throw std::runtime_error("zkp_ctx_create failed: no usable CUDA device (the engine has no CPU fallback)");
  }
  ~Engine() { zkp_ctx_destroy(h_); }
  Engine(const Engine&) = delete;
  Engine& operator=(const Engine&) = delete;
  zkp_ctx* handle() { return h_; }
  void check(int rc) const {
    if (rc != ZKP_OK) throw std::runtime_error(std::string("zkp_b200: ") + zkp_last_error(h_));
  }
  void use_key(const EncryptionKey& ek) {
    if (have_key_ && key_ == ek.n) return;
    nl_ = limbs_for_bits(ek.n.bit_length());
    std::vector<uint32_t> n = ek.n.to_limbs(nl_);
    check(zkp_set_key(h_, n.data(), (int)nl_));
    key_ = ek.n;
    have_key_ = true;
  }
  size_t nl() const { return nl_; }
  size_t nnl() const { return 2 * nl_; }
  size_t zl() const { return nl_ + 12; }

 private:
  zkp_ctx* h_ = nullptr;
  BigInt key_;
  bool have_key_ = false;
  size_t nl_ = 0;
};

// [rows] BigInts -> dense limb matrix
inline std::vector<uint32_t> pack(const std::vector<BigInt>& v, size_t limbs) {
  std::vector<uint32_t> out(v.size() * limbs);
  for (size_t i = 0; i < v.size(); ++i) v[i].to_limbs(out.data() + i * limbs, limbs);
  return out;
}
inline std::vector<BigInt> unpack(const std::vector<uint32_t>& m, size_t limbs) {
  std::vector<BigInt> v;
  for (size_t i = 0; i + limbs <= m.size(); i += limbs) v.push_back(BigInt::from_limbs(m.data() + i, limbs));
  return v;
}

// ------------------------------------------------------------------------------------ RangeProofNi
struct EncryptedPairs {  // range_proof.rs:32-39
  std::vector<BigInt> c1, c2;
};
struct Response {  // range_proof.rs:53-78
  bool open = true;
  BigInt w1, r1, w2, r2;        // Open
  uint8_t j = 0;                // Mask
  BigInt masked_x, masked_r;
};
struct Proof {  // range_proof.rs:80-81
  std::vector<Response> responses;
};
struct RangeStatement {
  BigInt range, ciphertext, secret_x, secret_r;
};

class RangeProofNi {  // range_proof_ni.rs:36-44
 public:
  EncryptionKey ek;
  BigInt range, ciphertext;
  EncryptedPairs encrypted_pairs;
  Proof proof;
  size_t error_factor = 0;

  // range_proof_ni.rs:47-82.  Randomness: per proof, w1[0..ef) <- sample_range(third, two_thirds), then ef coin
  // bytes (low bit), then r1[0..ef), r2[0..ef) <- sample_below(n)  (range_proof.rs:136-159; the reference's draw
  // order under rayon is nondeterministic, so any fixed order is as faithful as another).
  static std::vector<RangeProofNi> prove_batch(Engine& eng, const EncryptionKey& ek, const std::vector<RangeStatement>& st,
                                               const ByteSource& rng = os_rng(), size_t ef = SECURITY_PARAMETER) {
    eng.use_key(ek);
    const size_t B = st.size(), nl = eng.nl(), nnl = eng.nnl();
    if (B == 0) return {};
    size_t wbits = 0;
    for (auto& s : st) wbits = std::max({wbits, s.range.bit_length() + 2, s.secret_x.bit_length() + 2});
    const size_t wl = limbs_for_bits(wbits);
    if (wl > 64) throw std::length_error("range / secret_x wider than 2048 bits");
    std::vector<BigInt> range, x, r, w1, r1, r2;
    std::vector<uint8_t> swap(B * ef);
    for (auto& s : st) {
      range.push_back(s.range);
      x.push_back(s.secret_x);
      r.push_back(s.secret_r);
      BigInt third = s.range.div_floor(BigInt(3)), two = third + third;
      if (third.is_zero()) throw ReferencePanic("sample_range on an empty interval (range < 3)");
      for (size_t i = 0; i < ef; ++i) w1.push_back(BigInt::sample_range(rng, third, two));
      rng(swap.data() + (&s - &st[0]) * ef, ef);
      for (size_t i = 0; i < ef; ++i) r1.push_back(BigInt::sample_below(rng, ek.n));
      for (size_t i = 0; i < ef; ++i) r2.push_back(BigInt::sample_below(rng, ek.n));
    }
    for (auto& b : swap) b &= 1;
    std::vector<uint32_t> c1(B * ef * nnl), c2(B * ef * nnl), resp_w(B * ef * 2 * wl), resp_r(B * ef * 2 * nl);
    std::vector<uint8_t> kind(B * ef), digest(B * 32);
    eng.check(zkp_rangeproof_ni_prove(eng.handle(), (int)B, (int)ef, (int)wl, pack(range, wl).data(), pack(x, wl).data(),
                                      pack(r, nl).data(), pack(w1, wl).data(), swap.data(), pack(r1, nl).data(), pack(r2, nl).data(),
                                      c1.data(), c2.data(), digest.data(), kind.data(), resp_w.data(), resp_r.data()));
    std::vector<RangeProofNi> out(B);
    for (size_t b = 0; b < B; ++b) {
      RangeProofNi& p = out[b];
      p.ek = ek;
      p.range = st[b].range;
      p.ciphertext = st[b].ciphertext;
      p.error_factor = ef;
      for (size_t i = 0; i < ef; ++i) {
        const size_t t = b * ef + i;
        p.encrypted_pairs.c1.push_back(BigInt::from_limbs(&c1[t * nnl], nnl));
        p.encrypted_pairs.c2.push_back(BigInt::from_limbs(&c2[t * nnl], nnl));
        Response rs;
        const uint32_t* w = &resp_w[t * 2 * wl];
        const uint32_t* rr = &resp_r[t * 2 * nl];
        if (kind[t] == ZKP_RP_OPEN) {
          rs.open = true;
          rs.w1 = BigInt::from_limbs(w, wl);
          rs.w2 = BigInt::from_limbs(w + wl, wl);
          rs.r1 = BigInt::from_limbs(rr, nl);
          rs.r2 = BigInt::from_limbs(rr + nl, nl);
        } else {
          rs.open = false;
          rs.j = kind[t];
          rs.masked_x = BigInt::from_limbs(w, wl);
          rs.masked_r = BigInt::from_limbs(rr, nl);
        }
        p.proof.responses.push_back(rs);
      }
    }
    return out;
  }
  static RangeProofNi prove(Engine& eng, const EncryptionKey& ek, const BigInt& range, const BigInt& ciphertext,
                            const BigInt& secret_x, const BigInt& secret_r, const ByteSource& rng = os_rng()) {
    return prove_batch(eng, ek, {RangeStatement{range, ciphertext, secret_x, secret_r}}, rng)[0];
  }

  // range_proof_ni.rs:109-128 for many proofs under ONE key and error factor.  1 = Ok(()), 0 = Err(IncorrectProof).
  // `interactive_e`: the verifier's raw ChallengeBits (interactive RangeProof), same bytes for every proof of the batch
  static std::vector<int> verify_batch(Engine& eng, const std::vector<const RangeProofNi*>& ps, const std::vector<uint8_t>* interactive_e = nullptr) {
    const size_t B = ps.size();
    if (B == 0) return {};
    const EncryptionKey& ek = ps[0]->ek;
    const size_t ef = ps[0]->error_factor;
    for (auto* p : ps)
      if (!(p->ek == ek) || p->error_factor != ef) throw std::invalid_argument("verify_batch: proofs must share key and error factor");
    if (ef == 0) return std::vector<int>(B, 1);  // (0..0).all(..) is true: the reference returns Ok(()) (range_proof.rs:350-354)
    eng.use_key(ek);
    const size_t nl = eng.nl(), nnl = eng.nnl();
    // Row width of range / w1 / w2 / masked_x: from the RANGES only (the statement side).  A response value wider than its
    // proof's range can never pass the interval predicates (w < 2 * third, masked_x <= 2 * third; range_proof.rs:301-307,341),
    // so it rejects that one proof and does not widen - or poison - the batch.  Ranges beyond 2047 bits are outside what the
    // device rows hold (documented limit): such a proof is rejected on its own as well.
    constexpr size_t kMaxRangeBits = 64 * 32 - 1;
    size_t wbits = 1;
    for (auto* p : ps) {
      // bits_of_e[i] / responses[i] index out of range in the reference (range_proof.rs:273-274)
      if (p->proof.responses.size() < ef || p->encrypted_pairs.c1.size() < ef || p->encrypted_pairs.c2.size() < ef || ef > 256)
        throw ReferencePanic("index out of bounds: proof shorter than error_factor");
      if (p->range.bit_length() <= kMaxRangeBits) wbits = std::max(wbits, p->range.bit_length() + 1);
    }
    const size_t wl = limbs_for_bits(wbits);
    std::vector<uint32_t> range(B * wl), cx(B * nnl), c1(B * ef * nnl), c2(B * ef * nnl), resp_w(B * ef * 2 * wl, 0), resp_r(B * ef * 2 * nl, 0);
    std::vector<uint8_t> kind(B * ef), accept(B), fault(B);
    std::vector<int> out(B, -1);
    auto fits = [](const BigInt& v, size_t limbs) { return v.d.size() <= limbs; };
    for (size_t b = 0; b < B; ++b) {
      const RangeProofNi& p = *ps[b];
      bool representable = p.range.bit_length() <= kMaxRangeBits && fits(p.ciphertext, nnl);
      for (size_t i = 0; i < ef && representable; ++i) {
        const Response& r = p.proof.responses[i];
        representable = fits(p.encrypted_pairs.c1[i], nnl) && fits(p.encrypted_pairs.c2[i], nnl) &&
                        (r.open ? fits(r.w1, wl) && fits(r.w2, wl) : fits(r.masked_x, wl));
      }
      if (!representable) {  // wider than its row: never equal to a canonical ciphertext / never inside the interval: Err(IncorrectProof)
        out[b] = 0;
        continue;
      }
      p.range.to_limbs(&range[b * wl], wl);
      p.ciphertext.to_limbs(&cx[b * nnl], nnl);
      for (size_t i = 0; i < ef; ++i) {
        const size_t t = b * ef + i;
        const Response& r = p.proof.responses[i];
        p.encrypted_pairs.c1[i].to_limbs(&c1[t * nnl], nnl);
        p.encrypted_pairs.c2[i].to_limbs(&c2[t * nnl], nnl);
        if (r.open) {
          kind[t] = ZKP_RP_OPEN;
          r.w1.to_limbs(&resp_w[t * 2 * wl], wl);
          r.w2.to_limbs(&resp_w[t * 2 * wl + wl], wl);
          // randomness enters only as r^n mod nn, which depends on r mod n: reduced as the reference's mod_pow does
          (fits(r.r1, nl) ? r.r1 : r.r1 % ek.n).to_limbs(&resp_r[t * 2 * nl], nl);
          (fits(r.r2, nl) ? r.r2 : r.r2 % ek.n).to_limbs(&resp_r[t * 2 * nl + nl], nl);
        } else {
          kind[t] = r.j == 1 ? ZKP_RP_MASK1 : ZKP_RP_MASK2;  // `if *j == 1 { c1 } else { c2 }` (range_proof.rs:321-325)
          r.masked_x.to_limbs(&resp_w[t * 2 * wl], wl);
          (fits(r.masked_r, nl) ? r.masked_r : r.masked_r % ek.n).to_limbs(&resp_r[t * 2 * nl], nl);
        }
      }
    }
    if (!interactive_e) {
      eng.check(zkp_rangeproof_ni_verify(eng.handle(), (int)B, (int)ef, (int)wl, range.data(), cx.data(), c1.data(), c2.data(), kind.data(),
                                         resp_w.data(), resp_r.data(), accept.data(), fault.data(), nullptr));
    } else {
      std::vector<uint8_t> ch;
      for (size_t b = 0; b < B; ++b) ch.insert(ch.end(), interactive_e->begin(), interactive_e->end());
      static const uint8_t none = 0;
      eng.check(zkp_rp_verify_stage(eng.handle(), (int)B, (int)ef, (int)wl, range.data(), cx.data(), c1.data(), c2.data(), kind.data(),
                                    resp_w.data(), resp_r.data()));
      eng.check(zkp_rp_verify_run_with_challenge(eng.handle(), ch.empty() ? &none : ch.data(), (int)std::max<size_t>(1, interactive_e->size())));
      eng.check(zkp_rp_verify_fetch(eng.handle(), accept.data(), fault.data(), nullptr));
    }
    for (size_t b = 0; b < B; ++b) {
      if (out[b] == 0) continue;
      if (fault[b]) throw ReferencePanic("index out of bounds in verifier_output (range_proof.rs:273)");
      out[b] = accept[b];
    }
    return out;
  }
  void verify_self(Engine& eng) const {
    if (!verify_batch(eng, {this})[0]) throw IncorrectProof();
  }
  // range_proof_ni.rs:84-107
  void verify(Engine& eng, const EncryptionKey& ek_, const BigInt& ciphertext_) const {
    if (!(ek_ == ek)) throw ReferencePanic("assertion failed: `(left == right)` ek (range_proof_ni.rs:86)");
    if (ciphertext_ != ciphertext) throw ReferencePanic("assertion failed: `(left == right)` ciphertext (range_proof_ni.rs:88)");
    verify_self(eng);
  }

  Json to_json_value() const {
    Json j = Json::object();
    j.set("ek", ek.to_json());
    j.set("range", ser_native(range));
    j.set("ciphertext", ser_native(ciphertext));
    j.set("encrypted_pairs", Json::object().set("c1", ser_vec(encrypted_pairs.c1)).set("c2", ser_vec(encrypted_pairs.c2)));
    Json arr = Json::array();
    for (auto& r : proof.responses) {
      if (r.open)
        arr.push(Json::object().set("Open", Json::object().set("w1", ser_dec(r.w1)).set("r1", ser_dec(r.r1)).set("w2", ser_dec(r.w2)).set("r2", ser_dec(r.r2))));
      else
        arr.push(Json::object().set("Mask", Json::object().set("j", Json::number(r.j)).set("masked_x", ser_dec(r.masked_x)).set("masked_r", ser_dec(r.masked_r))));
    }
    j.set("proof", arr);
    j.set("error_factor", Json::number((int64_t)error_factor));
    return j;
  }
  std::string to_json() const { return to_json_value().dump(); }
  static RangeProofNi from_json_value(const Json& j) {
    RangeProofNi p;
    p.ek = EncryptionKey::from_json(j.at("ek"));
    p.range = de_native(j.at("range"));
    p.ciphertext = de_native(j.at("ciphertext"));
    p.encrypted_pairs.c1 = de_vec(j.at("encrypted_pairs").at("c1"));
    p.encrypted_pairs.c2 = de_vec(j.at("encrypted_pairs").at("c2"));
    for (auto& o : j.at("proof").arr) {
      Response r;
      if (const Json* v = o.find("Open")) {
        r.open = true;
        r.w1 = de_dec(v->at("w1"), false); r.r1 = de_dec(v->at("r1"), false);
        r.w2 = de_dec(v->at("w2"), false); r.r2 = de_dec(v->at("r2"), false);
      } else {
        const Json& m = o.at("Mask");
        r.open = false;
        int64_t jj = m.at("j").as_num();
        if (jj < 0 || jj > 255) throw std::runtime_error("j out of range for u8");
        r.j = (uint8_t)jj;
        r.masked_x = de_dec(m.at("masked_x"), false);
        r.masked_r = de_dec(m.at("masked_r"), false);
      }
      p.proof.responses.push_back(r);
    }
    p.error_factor = (size_t)j.at("error_factor").as_num();
    return p;
  }
  static RangeProofNi from_json(const std::string& s) { return from_json_value(Json::parse(s)); }
};

// ------------------------------------------------------------------------------------- RangeProof
// The interactive proof (range_proof.rs:100-363), STATISTICAL_ERROR_FACTOR = 40 (:30).  The verifier commits
// to its challenge, the prover sends the encrypted pairs, the verifier opens, the prover answers.
constexpr size_t STATISTICAL_ERROR_FACTOR = 40;
struct ChallengeBits { std::vector<uint8_t> bytes; };          // range_proof.rs:49-50
struct Commitment { BigInt com; };                             // :83
struct ChallengeRandomness { BigInt r; };                      // :100
struct DataRandomnessPairs { std::vector<BigInt> w1, w2, r1, r2; };  // :41-47

class RangeProof {
 public:
  // local compute_digest(bytes) (range_proof.rs:365-369): SHA-256 of the raw bytes as a BigInt
  static BigInt digest_of_bytes(const std::vector<uint8_t>& b) { return BigInt::from_bytes(zkhost::sha256(b.data(), b.size())); }
  // get_paillier_commitment (range_proof.rs:359-363): Enc(ek, x, r)
  static BigInt paillier_commitment(Engine& eng, const EncryptionKey& ek, const BigInt& x, const BigInt& r) {
    eng.use_key(ek);
    const size_t nl = eng.nl(), nnl = eng.nnl();
    std::vector<uint32_t> out(nnl);
    eng.check(zkp_paillier_enc(eng.handle(), x.to_limbs(nl).data(), (int)nl, r.to_limbs(nl).data(), (int)nl, 1, out.data()));
    return BigInt::from_limbs(out.data(), nnl);
  }
  struct VerifierCommit { Commitment com; ChallengeRandomness r; ChallengeBits e; };
  // range_proof.rs:118-126
  static VerifierCommit verifier_commit(Engine& eng, const EncryptionKey& ek, const ByteSource& rng = os_rng()) {
    VerifierCommit v;
    v.e.bytes.resize(STATISTICAL_ERROR_FACTOR / 8);            // ChallengeBits::sample (:86-91)
    rng(v.e.bytes.data(), v.e.bytes.size());
    BigInt m = digest_of_bytes(v.e.bytes);
    v.r.r = BigInt::sample_below(rng, ek.n);
    v.com.com = paillier_commitment(eng, ek, m, v.r.r);
    return v;
  }
  // range_proof.rs:195-208
  static void verify_commit(Engine& eng, const EncryptionKey& ek, const Commitment& com, const ChallengeRandomness& r, const ChallengeBits& e) {
    if (paillier_commitment(eng, ek, digest_of_bytes(e.bytes), r.r) != com.com) throw IncorrectProof();
  }
  // range_proof.rs:128-193: 2 * error_factor Paillier encryptions on the device
  static std::pair<EncryptedPairs, DataRandomnessPairs> generate_encrypted_pairs(Engine& eng, const EncryptionKey& ek, const BigInt& range,
                                                                                 size_t error_factor, const ByteSource& rng = os_rng()) {
    eng.use_key(ek);
    const size_t nl = eng.nl(), nnl = eng.nnl(), ef = error_factor;
    const BigInt third = range.div_floor(BigInt(3)), two = third + third;
    DataRandomnessPairs d;
    for (size_t i = 0; i < ef; ++i) d.w1.push_back(BigInt::sample_range(rng, third, two));     // :136-139
    for (size_t i = 0; i < ef; ++i) d.w2.push_back(d.w1[i] - third);                           // :141
    std::vector<uint8_t> coin(ef);
    rng(coin.data(), ef);
    for (size_t i = 0; i < ef; ++i)
      if (coin[i] & 1) std::swap(d.w1[i], d.w2[i]);                                            // :144-149
    for (size_t i = 0; i < ef; ++i) d.r1.push_back(BigInt::sample_below(rng, ek.n));           // :151-154
    for (size_t i = 0; i < ef; ++i) d.r2.push_back(BigInt::sample_below(rng, ek.n));           // :156-159
    std::vector<BigInt> m(d.w1), r(d.r1);
    m.insert(m.end(), d.w2.begin(), d.w2.end());
    r.insert(r.end(), d.r2.begin(), d.r2.end());
    std::vector<uint32_t> out(2 * ef * nnl);
    eng.check(zkp_paillier_enc(eng.handle(), pack(m, nl).data(), (int)nl, pack(r, nl).data(), (int)nl, (int)(2 * ef), out.data()));
    EncryptedPairs p;
    for (size_t i = 0; i < ef; ++i) p.c1.push_back(BigInt::from_limbs(&out[i * nnl], nnl));
    for (size_t i = 0; i < ef; ++i) p.c2.push_back(BigInt::from_limbs(&out[(ef + i) * nnl], nnl));
    return {p, d};
  }
  // range_proof.rs:210-252.  No modexp here: error_factor comparisons and (on average half as many) products
  // secret_r * r_j % n, done with the host BigInt.
  static Proof generate_proof(const EncryptionKey& ek, const BigInt& secret_x, const BigInt& secret_r, const ChallengeBits& e, const BigInt& range,
                              const DataRandomnessPairs& data, size_t error_factor) {
    const BigInt third = range.div_floor(BigInt(3)), two = third + third;
    Proof pr;
    for (size_t i = 0; i < error_factor; ++i) {
      if (i / 8 >= e.bytes.size()) throw ReferencePanic("index out of bounds: bits_of_e[i] (range_proof.rs:225)");
      const bool ei = (e.bytes[i / 8] >> (7 - i % 8)) & 1;
      Response r;
      if (!ei) {
        r.open = true;
        r.w1 = data.w1[i]; r.r1 = data.r1[i]; r.w2 = data.w2[i]; r.r2 = data.r2[i];
      } else {
        r.open = false;
        const BigInt s1 = secret_x + data.w1[i];
        if (s1 > third && s1 < two) {
          r.j = 1; r.masked_x = s1; r.masked_r = (secret_r * data.r1[i]) % ek.n;
        } else {
          r.j = 2; r.masked_x = secret_x + data.w2[i]; r.masked_r = (secret_r * data.r2[i]) % ek.n;
        }
      }
      pr.responses.push_back(r);
    }
    return pr;
  }
  // range_proof.rs:254-355 against the verifier's own ChallengeBits, on the device
  static void verifier_output(Engine& eng, const EncryptionKey& ek, const ChallengeBits& e, const EncryptedPairs& pairs, const Proof& proof,
                              const BigInt& range, const BigInt& cipher_x, size_t error_factor) {
    RangeProofNi p;
    p.ek = ek; p.range = range; p.ciphertext = cipher_x; p.encrypted_pairs = pairs; p.proof = proof; p.error_factor = error_factor;
    if (!RangeProofNi::verify_batch(eng, {&p}, &e.bytes)[0]) throw IncorrectProof();
  }
};

// ------------------------------------------------------------------------------- NiCorrectKeyProof
class NiCorrectKeyProof {  // correct_key_ni.rs:34-39
 public:
  std::vector<BigInt> sigma_vec;

  // correct_key_ni.rs:42-71.  rho on the device (zkp_correct_key_ni_rho); extract_nroot(dk, rho_i) =
  // rho_i^(n^-1 mod phi) mod n through CRT: two half-width K2 modexps per i, recombined on the host.
  static NiCorrectKeyProof proof(Engine& eng, const DecryptionKey& dk, const uint8_t* salt = nullptr, size_t salt_len = 0) {
    if (!salt) { salt = SALT_STRING; salt_len = sizeof(SALT_STRING); }
    const BigInt n = dk.p * dk.q;
    const size_t nl = limbs_for_bits(n.bit_length());
    std::vector<uint32_t> rho(M2 * nl), nrow = n.to_limbs(nl);
    eng.check(zkp_correct_key_ni_rho(eng.handle(), 1, (int)nl, nrow.data(), salt, (int)salt_len, rho.data()));
    const BigInt pm1 = dk.p - BigInt(1), qm1 = dk.q - BigInt(1);
    BigInt dp, dq, pinv;
    if (!BigInt::mod_inv(n % pm1, pm1, dp) || !BigInt::mod_inv(n % qm1, qm1, dq) || !BigInt::mod_inv(dk.p % dk.q, dk.q, pinv))
      throw ReferencePanic("extract_nroot: n is not invertible mod phi(n)");
    const size_t hl = limbs_for_bits(std::max(dk.p.bit_length(), dk.q.bit_length()));
    std::vector<BigInt> bases, exps{dp, dq}, mods{dk.p, dk.q};
    std::vector<BigInt> rhos = unpack(rho, nl);
    for (auto& r : rhos) bases.push_back(r % dk.p);
    for (auto& r : rhos) bases.push_back(r % dk.q);
    std::vector<uint32_t> out(2 * M2 * hl);
    eng.check(zkp_modexp_var(eng.handle(), pack(bases, hl).data(), pack(exps, hl).data(), (int)hl, (int)(32 * hl), (int)M2,
                             pack(mods, hl).data(), (int)hl, (int)M2, (int)(2 * M2), out.data()));
    std::vector<BigInt> s = unpack(out, hl);
    NiCorrectKeyProof pr;
    for (size_t i = 0; i < M2; ++i) {
      const BigInt& sp = s[i];
      const BigInt& sq = s[M2 + i];
      BigInt diff = (sq + dk.q - (sp % dk.q)) % dk.q;      // (sq - sp) mod q
      BigInt h = (diff * pinv) % dk.q;
      pr.sigma_vec.push_back(sp + dk.p * h);
    }
    return pr;
  }

  // The same for many keys in three device calls (rho of every key; the half-width roots mod every p; mod every q).
  // The per-key host work (three modular inversions and the CRT recombinations) is spread over host threads.
  static std::vector<NiCorrectKeyProof> proof_batch(Engine& eng, const std::vector<DecryptionKey>& dks, const uint8_t* salt = nullptr,
                                                    size_t salt_len = 0, unsigned host_threads = 0) {
    if (!salt) { salt = SALT_STRING; salt_len = sizeof(SALT_STRING); }
    const size_t B = dks.size();
    if (B == 0) return {};
    std::vector<BigInt> n(B);
    size_t nbits = 0, hbits = 0;
    for (size_t b = 0; b < B; ++b) {
      n[b] = dks[b].p * dks[b].q;
      nbits = std::max(nbits, n[b].bit_length());
      hbits = std::max({hbits, dks[b].p.bit_length(), dks[b].q.bit_length()});
    }
    const size_t nl = limbs_for_bits(nbits), hl = limbs_for_bits(hbits);
    std::vector<uint32_t> rho(B * M2 * nl);
    eng.check(zkp_correct_key_ni_rho(eng.handle(), (int)B, (int)nl, pack(n, nl).data(), salt, (int)salt_len, rho.data()));
    std::vector<BigInt> dp(B), dq(B), pinv(B);
    std::vector<uint32_t> base_p(B * M2 * hl), base_q(B * M2 * hl);
    std::vector<int> bad(B, 0);
    auto per_key = [&](size_t b) {
      const BigInt one(1), pm1 = dks[b].p - one, qm1 = dks[b].q - one;
      if (!BigInt::mod_inv(n[b] % pm1, pm1, dp[b]) || !BigInt::mod_inv(n[b] % qm1, qm1, dq[b]) || !BigInt::mod_inv(dks[b].p % dks[b].q, dks[b].q, pinv[b])) {
        bad[b] = 1;
        return;
      }
      for (size_t i = 0; i < M2; ++i) {
        const BigInt r = BigInt::from_limbs(&rho[(b * M2 + i) * nl], nl);
        (r % dks[b].p).to_limbs(&base_p[(b * M2 + i) * hl], hl);
        (r % dks[b].q).to_limbs(&base_q[(b * M2 + i) * hl], hl);
      }
    };
    parallel_for(B, host_threads, per_key);
    for (size_t b = 0; b < B; ++b)
      if (bad[b]) throw ReferencePanic("extract_nroot: n is not invertible mod phi(n)");
    std::vector<BigInt> ps, qs;
    for (auto& dk : dks) { ps.push_back(dk.p); qs.push_back(dk.q); }
    std::vector<uint32_t> sp(B * M2 * hl), sq(B * M2 * hl);
    eng.check(zkp_modexp_var(eng.handle(), base_p.data(), pack(dp, hl).data(), (int)hl, (int)(32 * hl), (int)M2, pack(ps, hl).data(), (int)hl, (int)M2,
                             (int)(B * M2), sp.data()));
    eng.check(zkp_modexp_var(eng.handle(), base_q.data(), pack(dq, hl).data(), (int)hl, (int)(32 * hl), (int)M2, pack(qs, hl).data(), (int)hl, (int)M2,
                             (int)(B * M2), sq.data()));
    std::vector<NiCorrectKeyProof> out(B);
    parallel_for(B, host_threads, [&](size_t b) {
      const BigInt &p = dks[b].p, &q = dks[b].q;
      for (size_t i = 0; i < M2; ++i) {
        const BigInt a = BigInt::from_limbs(&sp[(b * M2 + i) * hl], hl), c = BigInt::from_limbs(&sq[(b * M2 + i) * hl], hl);
        const BigInt diff = (c + q - (a % q)) % q;
        out[b].sigma_vec.push_back(a + p * ((diff * pinv[b]) % q));
      }
    });
    return out;
  }

  // correct_key_ni.rs:73-100 for many (proof, key) pairs with one salt; every modulus padded to a common width.
  static std::vector<int> verify_batch(Engine& eng, const std::vector<const NiCorrectKeyProof*>& ps, const std::vector<EncryptionKey>& eks,
                                       const uint8_t* salt, size_t salt_len) {
    const size_t B = ps.size();
    if (B == 0) return {};
    size_t bits = 0;
    for (auto& ek : eks) bits = std::max(bits, ek.n.bit_length());
    const size_t nl = limbs_for_bits(bits);
    std::vector<uint32_t> n(B * nl), sigma(B * M2 * nl);
    std::vector<uint8_t> accept(B);
    std::vector<int> out(B, -1);
    for (size_t b = 0; b < B; ++b) {
      if (ps[b]->sigma_vec.size() < M2) throw ReferencePanic("index out of bounds: sigma_vec shorter than M2 (correct_key_ni.rs:92)");
      eks[b].n.to_limbs(&n[b * nl], nl);
      for (size_t i = 0; i < M2; ++i) {
        BigInt s = ps[b]->sigma_vec[i];
        if (s.d.size() > nl) s = s % eks[b].n;  // mod_pow reduces its base; same residue
        s.to_limbs(&sigma[(b * M2 + i) * nl], nl);
      }
    }
    eng.check(zkp_correct_key_ni_verify(eng.handle(), (int)B, (int)nl, n.data(), sigma.data(), salt, (int)salt_len, accept.data(), nullptr));
    for (size_t b = 0; b < B; ++b) out[b] = accept[b];
    return out;
  }
  void verify(Engine& eng, const EncryptionKey& ek, const uint8_t* salt, size_t salt_len) const {
    if (!verify_batch(eng, {this}, {ek}, salt, salt_len)[0]) throw IncorrectProof();
  }
  std::string to_json() const { return Json::object().set("sigma_vec", ser_vec(sigma_vec)).dump(); }
  static NiCorrectKeyProof from_json(const std::string& s) {
    NiCorrectKeyProof p;
    p.sigma_vec = de_vec(Json::parse(s).at("sigma_vec"));
    return p;
  }
};

// ------------------------------------------------------------------------------ shared device helpers
// BigInt::mod_pow(base_i, exp_(i or shared), mod) for a batch under ONE modulus (K2 through zkp_modexp_var).
inline std::vector<BigInt> powm_batch(Engine& eng, const std::vector<BigInt>& bases, const std::vector<BigInt>& exps, const BigInt& mod) {
  const size_t B = bases.size();
  if (B == 0) return {};
  if (exps.size() != 1 && exps.size() != B) throw std::invalid_argument("powm_batch: one exponent, or one per base");
  const size_t nl = limbs_for_bits(mod.bit_length());
  size_t ebits = 1;
  for (auto& e : exps) ebits = std::max(ebits, e.bit_length());
  const size_t el = limbs_for_bits(ebits);
  std::vector<BigInt> b2;
  for (auto& b : bases) b2.push_back(b.d.size() > nl ? b % mod : b);
  std::vector<uint32_t> out(B * nl);
  eng.check(zkp_modexp_var(eng.handle(), pack(b2, nl).data(), pack(exps, el).data(), (int)el, (int)(32 * el), exps.size() == 1 ? (int)B : 1,
                           mod.to_limbs(nl).data(), (int)nl, (int)B, (int)B, out.data()));
  return unpack(out, nl);
}
// compute_digest (utils.rs:9-22) of one item list, on the device (K4)
inline BigInt compute_digest(Engine& eng, const std::vector<BigInt>& items) {
  size_t bits = 32;
  for (auto& v : items) bits = std::max(bits, v.bit_length());
  const size_t l = limbs_for_bits(bits);
  uint8_t dig[32];
  eng.check(zkp_sha256_transcript(eng.handle(), pack(items, l).data(), (int)l, (int)items.size(), 1, dig));
  return BigInt::from_bytes(dig, 32);
}
// kzen-paillier extract_nroot(dk, z) = z^(n^-1 mod phi) mod n for a batch: two half-width K2 batches + CRT
inline std::vector<BigInt> extract_nroots(Engine& eng, const DecryptionKey& dk, const std::vector<BigInt>& zs) {
  const BigInt n = dk.p * dk.q, pm1 = dk.p - BigInt(1), qm1 = dk.q - BigInt(1);
  BigInt dp, dq, pinv;
  if (!BigInt::mod_inv(n % pm1, pm1, dp) || !BigInt::mod_inv(n % qm1, qm1, dq) || !BigInt::mod_inv(dk.p % dk.q, dk.q, pinv))
    throw ReferencePanic("extract_nroot: n is not invertible mod phi(n)");
  std::vector<BigInt> bp, bq;
  for (auto& z : zs) { bp.push_back(z % dk.p); bq.push_back(z % dk.q); }
  std::vector<BigInt> sp = powm_batch(eng, bp, {dp}, dk.p), sq = powm_batch(eng, bq, {dq}, dk.q), out;
  for (size_t i = 0; i < zs.size(); ++i) {
    BigInt diff = (sq[i] + dk.q - (sp[i] % dk.q)) % dk.q;
    out.push_back(sp[i] + dk.p * ((diff * pinv) % dk.q));
  }
  return out;
}

// --------------------------------------------------------------------------------------- CorrectKey
// The interactive proof of co-primality of n and phi(n) (correct_key.rs:64-172).
struct Challenge {  // correct_key.rs:28-38
  std::vector<BigInt> sn;
  BigInt e;
  std::vector<BigInt> z;
  std::string to_json() const { return Json::object().set("sn", ser_vec(sn)).set("e", ser_dec(e)).set("z", ser_vec(z)).dump(); }
  static Challenge from_json(const std::string& s) {
    Json j = Json::parse(s);
    Challenge c;
    c.sn = de_vec(j.at("sn")); c.e = de_dec(j.at("e"), false); c.z = de_vec(j.at("z"));
    return c;
  }
};
struct VerificationAid { BigInt s_digest; };   // :40-44
struct CorrectKeyProof { BigInt s_digest; };   // :46-50
struct CorrectKeyProveError : std::runtime_error {  // :174-184
  enum Kind { SniNotCoprimeWithN, ZiNotCoprimeWithN, RniNotCoprimeWithN, EWasntComputedCorrectly } kind;
  CorrectKeyProveError(Kind k, const char* what) : std::runtime_error(what), kind(k) {}
};

class CorrectKey {
 public:
  // correct_key.rs:65-107: s_i, then r_i <- sample_below(n); 80 x^n mod n and 40 s^e mod n on the device
  static std::pair<Challenge, VerificationAid> challenge(Engine& eng, const EncryptionKey& ek, const ByteSource& rng = os_rng()) {
    const size_t k = STATISTICAL_ERROR_FACTOR;
    std::vector<BigInt> s, r;
    for (size_t i = 0; i < k; ++i) s.push_back(BigInt::sample_below(rng, ek.n));
    for (size_t i = 0; i < k; ++i) r.push_back(BigInt::sample_below(rng, ek.n));
    std::vector<BigInt> both(s);
    both.insert(both.end(), r.begin(), r.end());
    std::vector<BigInt> pw = powm_batch(eng, both, {ek.n}, ek.n);
    Challenge ch;
    ch.sn.assign(pw.begin(), pw.begin() + k);
    std::vector<BigInt> items{ek.n};
    items.insert(items.end(), pw.begin(), pw.end());          // n, sn.., rn..
    ch.e = compute_digest(eng, items);
    std::vector<BigInt> se = powm_batch(eng, s, {ch.e}, ek.n);
    for (size_t i = 0; i < k; ++i) ch.z.push_back((r[i] * se[i]) % ek.n);
    return {ch, VerificationAid{compute_digest(eng, s)}};
  }
  // correct_key.rs:109-162
  static CorrectKeyProof prove(Engine& eng, const DecryptionKey& dk, const Challenge& ch) {
    const BigInt n = dk.q * dk.p, one(1);
    for (auto& v : ch.sn)
      if (BigInt::gcd(n, v) != one) throw CorrectKeyProveError(CorrectKeyProveError::SniNotCoprimeWithN, "`challenge.sn[i]` isn't co-prime with `n`");
    for (auto& v : ch.z)
      if (BigInt::gcd(n, v) != one) throw CorrectKeyProveError(CorrectKeyProveError::ZiNotCoprimeWithN, "`challenge.z[i]` isn't co-prime with `n`");
    const BigInt phi = (dk.q - one) * (dk.p - one);
    const BigInt phimine = phi - (ch.e % phi);
    std::vector<BigInt> zn = powm_batch(eng, ch.z, {n}, n), snphi = powm_batch(eng, ch.sn, {phimine}, n), rn;
    const size_t k = std::min(ch.z.size(), ch.sn.size());     // zip
    for (size_t i = 0; i < k; ++i) rn.push_back((zn[i] * snphi[i]) % n);
    for (auto& v : rn)
      if (BigInt::gcd(n, v) != one) throw CorrectKeyProveError(CorrectKeyProveError::RniNotCoprimeWithN, "`rn[i]` isn't co-prime with `n`");
    std::vector<BigInt> items{n};
    items.insert(items.end(), ch.sn.begin(), ch.sn.end());
    items.insert(items.end(), rn.begin(), rn.end());
    if (ch.e != compute_digest(eng, items))
      throw CorrectKeyProveError(CorrectKeyProveError::EWasntComputedCorrectly, "`challenge.e` wasn't computed correctly");
    return CorrectKeyProof{compute_digest(eng, extract_nroots(eng, dk, ch.sn))};
  }
  // correct_key.rs:164-171
  static void verify(const CorrectKeyProof& proof, const VerificationAid& va) {
    if (proof.s_digest != va.s_digest) throw IncorrectProof();
  }
};

// ----------------------------------------------------------------------------------- sigma protocols
// Field names and order follow the reference structs; BigInt fields use curv's native serde.
//
// Batches.  The device verifies one key per launch (zkp_set_key), but every Statement carries its own ek, and a batch
// handed to verify_batch is untrusted input.  So verify_batch groups the statements by key and runs one device call
// per key; a proof whose fields cannot be laid out in the device rows is decided on its own (rows the reference only
// uses under a reduction are reduced here; rows that enter the transcript hash or an exponent wider than their device
// row make THAT proof 0) and a proof on which the reference would panic is reported as -1 - nothing one proof carries
// can change the verdict of another.  verify_batch returns 1 = Ok(()), 0 = Err(IncorrectProof), -1 = the reference
// panics; the single-proof verify() turns those into IncorrectProof / ReferencePanic.  prove_batch is the prover's own
// batch: it requires one key and throws std::invalid_argument otherwise.
template <class St>
inline std::vector<std::vector<size_t>> group_by_key(const std::vector<St>& st) {
  std::vector<std::vector<size_t>> groups;
  for (size_t b = 0; b < st.size(); ++b) {
    size_t k = 0;
    while (k < groups.size() && !(st[groups[k][0]].ek == st[b].ek)) ++k;
    if (k == groups.size()) groups.emplace_back();
    groups[k].push_back(b);
  }
  return groups;
}
template <class St>
inline void require_one_key(const std::vector<St>& st, const char* who) {
  for (auto& s : st)
    if (!(s.ek == st[0].ek)) throw std::invalid_argument(std::string(who) + ": the statements of one proving batch must share the key");
}
inline bool fits_limbs(const BigInt& v, size_t limbs) { return v.d.size() <= limbs; }
// verify_batch over mixed keys: `one_key(ps, st)` verifies a sub-batch whose statements share st[0].ek
template <class Proof, class St, class F>
inline std::vector<int> verify_by_key(const std::vector<const Proof*>& ps, const std::vector<St>& st, F one_key) {
  if (ps.size() != st.size()) throw std::invalid_argument("verify_batch: one statement per proof");
  std::vector<int> out(ps.size(), 0);
  for (auto& idx : group_by_key(st)) {
    std::vector<const Proof*> p2;
    std::vector<St> s2;
    for (size_t b : idx) {
      p2.push_back(ps[b]);
      s2.push_back(st[b]);
    }
    const std::vector<int> r = one_key(p2, s2);
    for (size_t k = 0; k < idx.size(); ++k) out[idx[k]] = r[k];
  }
  return out;
}
inline void throw_for_verdict(int v, const char* panic_text) {
  if (v < 0) throw ReferencePanic(panic_text);
  if (!v) throw IncorrectProof();
}
struct ZeroStatement { EncryptionKey ek; BigInt c; };        // zero_enc_proof.rs:37-41
struct ZeroWitness { BigInt r; };                            // :32-35
class ZeroProof {                                            // :26-30
 public:
  BigInt z, a;
  static std::vector<ZeroProof> prove_batch(Engine& eng, const std::vector<ZeroWitness>& w, const std::vector<ZeroStatement>& st,
                                            const ByteSource& rng = os_rng()) {
    const size_t B = st.size();
    if (B == 0) return {};
    require_one_key(st, "ZeroProof::prove_batch");
    eng.use_key(st[0].ek);
    const size_t nl = eng.nl(), nnl = eng.nnl();
    std::vector<BigInt> r, c, rp;
    for (size_t b = 0; b < B; ++b) {
      r.push_back(w[b].r);
      c.push_back(st[b].c);
      rp.push_back(BigInt::sample_below(rng, st[b].ek.n));  // :45
    }
    std::vector<uint32_t> z(B * nnl), a(B * nnl);
    eng.check(zkp_zero_prove(eng.handle(), (int)B, pack(r, nl).data(), pack(c, nnl).data(), pack(rp, nl).data(), z.data(), a.data()));
    std::vector<ZeroProof> out(B);
    for (size_t b = 0; b < B; ++b) {
      out[b].z = BigInt::from_limbs(&z[b * nnl], nnl);
      out[b].a = BigInt::from_limbs(&a[b * nnl], nnl);
    }
    return out;
  }
  static ZeroProof prove(Engine& eng, const ZeroWitness& w, const ZeroStatement& st, const ByteSource& rng = os_rng()) {
    return prove_batch(eng, {w}, {st}, rng)[0];
  }
  static std::vector<int> verify_batch(Engine& eng, const std::vector<const ZeroProof*>& ps, const std::vector<ZeroStatement>& st) {
    return verify_by_key(ps, st, [&](const std::vector<const ZeroProof*>& p, const std::vector<ZeroStatement>& s) { return verify_one_key(eng, p, s); });
  }
  void verify(Engine& eng, const ZeroStatement& st) const { throw_for_verdict(verify_batch(eng, {this}, {st})[0], "unreachable"); }

 private:
  static std::vector<int> verify_one_key(Engine& eng, const std::vector<const ZeroProof*>& ps, const std::vector<ZeroStatement>& st) {
    const size_t B = ps.size();
    eng.use_key(st[0].ek);
    const size_t nnl = eng.nnl();
    const BigInt& nn = st[0].ek.nn;
    std::vector<BigInt> c, z, a;
    std::vector<int> out(B, 1);
    for (size_t b = 0; b < B; ++b) {
      // c and a enter the transcript hash as given (and mod_pow / mod_mul reduce them): they must fit their rows
      const bool ok = fits_limbs(st[b].c, nnl) && fits_limbs(ps[b]->a, nnl);
      if (!ok) out[b] = 0;
      c.push_back(ok ? st[b].c : BigInt(0));
      a.push_back(ok ? ps[b]->a : BigInt(0));
      z.push_back(ps[b]->z % nn);  // Enc(0, z) = z^n mod nn (zero_enc_proof.rs:73-79): reduced by mod_pow
    }
    std::vector<uint8_t> acc(B);
    eng.check(zkp_zero_verify(eng.handle(), (int)B, pack(c, nnl).data(), pack(z, nnl).data(), pack(a, nnl).data(), acc.data()));
    for (size_t b = 0; b < B; ++b)
      if (out[b]) out[b] = acc[b];
    return out;
  }

 public:
  std::string to_json() const { return Json::object().set("z", ser_native(z)).set("a", ser_native(a)).dump(); }
  static ZeroProof from_json(const std::string& s) {
    Json j = Json::parse(s);
    ZeroProof p;
    p.z = de_native(j.at("z"));
    p.a = de_native(j.at("a"));
    return p;
  }
};

struct CiphertextStatement { EncryptionKey ek; BigInt c; };  // correct_ciphertext.rs:35-39
struct CiphertextWitness { BigInt x, r; };                   // :29-33
class CiphertextProof {                                      // :22-27
 public:
  BigInt z1, z2, c_prime;
  static std::vector<CiphertextProof> prove_batch(Engine& eng, const std::vector<CiphertextWitness>& w,
                                                  const std::vector<CiphertextStatement>& st, const ByteSource& rng = os_rng()) {
    const size_t B = st.size();
    if (B == 0) return {};
    require_one_key(st, "CiphertextProof::prove_batch");
    eng.use_key(st[0].ek);
    const size_t nl = eng.nl(), nnl = eng.nnl(), zl = eng.zl();
    std::vector<BigInt> x, r, c, xp, rp;
    for (size_t b = 0; b < B; ++b) {
      x.push_back(w[b].x); r.push_back(w[b].r); c.push_back(st[b].c);
      xp.push_back(BigInt::sample_below(rng, st[b].ek.n));  // :43
      rp.push_back(BigInt::sample_below(rng, st[b].ek.n));  // :44
    }
    std::vector<uint32_t> z1(B * zl), z2(B * nnl), cp(B * nnl);
    eng.check(zkp_ciphertext_prove(eng.handle(), (int)B, (int)zl, pack(x, nl).data(), pack(r, nl).data(), pack(c, nnl).data(),
                                   pack(xp, nl).data(), pack(rp, nl).data(), z1.data(), z2.data(), cp.data()));
    std::vector<CiphertextProof> out(B);
    for (size_t b = 0; b < B; ++b) {
      out[b].z1 = BigInt::from_limbs(&z1[b * zl], zl);
      out[b].z2 = BigInt::from_limbs(&z2[b * nnl], nnl);
      out[b].c_prime = BigInt::from_limbs(&cp[b * nnl], nnl);
    }
    return out;
  }
  static CiphertextProof prove(Engine& eng, const CiphertextWitness& w, const CiphertextStatement& st, const ByteSource& rng = os_rng()) {
    return prove_batch(eng, {w}, {st}, rng)[0];
  }
  static std::vector<int> verify_batch(Engine& eng, const std::vector<const CiphertextProof*>& ps, const std::vector<CiphertextStatement>& st) {
    return verify_by_key(ps, st, [&](const std::vector<const CiphertextProof*>& p, const std::vector<CiphertextStatement>& s) { return verify_one_key(eng, p, s); });
  }
  void verify(Engine& eng, const CiphertextStatement& st) const { throw_for_verdict(verify_batch(eng, {this}, {st})[0], "unreachable"); }

 private:
  static std::vector<int> verify_one_key(Engine& eng, const std::vector<const CiphertextProof*>& ps, const std::vector<CiphertextStatement>& st) {
    const size_t B = ps.size();
    eng.use_key(st[0].ek);
    const size_t nnl = eng.nnl(), zl = eng.zl();
    const EncryptionKey& ek = st[0].ek;
    std::vector<BigInt> c, z1, z2, cp;
    std::vector<int> out(B, 1);
    for (size_t b = 0; b < B; ++b) {
      const bool ok = fits_limbs(st[b].c, nnl) && fits_limbs(ps[b]->c_prime, nnl);  // hashed as given
      if (!ok) out[b] = 0;
      c.push_back(ok ? st[b].c : BigInt(0));
      cp.push_back(ok ? ps[b]->c_prime : BigInt(0));
      z1.push_back(fits_limbs(ps[b]->z1, zl) ? ps[b]->z1 : ps[b]->z1 % ek.n);  // (m*n + 1) % nn depends on m mod n only
      z2.push_back(ps[b]->z2 % ek.nn);                                            // randomness: reduced by mod_pow
    }
    std::vector<uint8_t> acc(B);
    eng.check(zkp_ciphertext_verify(eng.handle(), (int)B, (int)zl, pack(c, nnl).data(), pack(z1, zl).data(), pack(z2, nnl).data(),
                                    pack(cp, nnl).data(), acc.data()));
    for (size_t b = 0; b < B; ++b)
      if (out[b]) out[b] = acc[b];
    return out;
  }

 public:
  std::string to_json() const {
    return Json::object().set("z1", ser_native(z1)).set("z2", ser_native(z2)).set("c_prime", ser_native(c_prime)).dump();
  }
  static CiphertextProof from_json(const std::string& s) {
    Json j = Json::parse(s);
    CiphertextProof p;
    p.z1 = de_native(j.at("z1")); p.z2 = de_native(j.at("z2")); p.c_prime = de_native(j.at("c_prime"));
    return p;
  }
};

struct MulStatement { EncryptionKey ek; BigInt e_a, e_b, e_c; };        // multiplication_proof.rs:51-57
struct MulWitness { BigInt a, b, c, r_a, r_b, r_c; };                   // :41-49
class MulProof {                                                        // :32-39
 public:
  BigInt f, z1, z2, e_d, e_db;
  static BigInt sample_paillier_random(const ByteSource& rng, const BigInt& modulo) {  // :148-154
    for (;;) {
      BigInt r = BigInt::sample_below(rng, modulo);
      if (BigInt::gcd(r, modulo) == BigInt(1)) return r;
    }
  }
  static std::vector<MulProof> prove_batch(Engine& eng, const std::vector<MulWitness>& w, const std::vector<MulStatement>& st,
                                           const ByteSource& rng = os_rng()) {
    const size_t B = st.size();
    if (B == 0) return {};
    require_one_key(st, "MulProof::prove_batch");
    eng.use_key(st[0].ek);
    const size_t nl = eng.nl(), nnl = eng.nnl();
    std::vector<BigInt> a, b, ra, rb, rc, ea, eb, ec, d, rd;
    for (size_t i = 0; i < B; ++i) {
      a.push_back(w[i].a); b.push_back(w[i].b); ra.push_back(w[i].r_a); rb.push_back(w[i].r_b); rc.push_back(w[i].r_c);
      ea.push_back(st[i].e_a); eb.push_back(st[i].e_b); ec.push_back(st[i].e_c);
      d.push_back(BigInt::sample_below(rng, st[i].ek.n));          // :61
      rd.push_back(sample_paillier_random(rng, st[i].ek.n));       // :62
    }
    std::vector<uint32_t> f(B * nl), z1(B * nnl), z2(B * nnl), ed(B * nnl), edb(B * nnl);
    std::vector<uint8_t> fault(B);
    eng.check(zkp_mul_prove(eng.handle(), (int)B, pack(a, nl).data(), pack(b, nl).data(), pack(ra, nl).data(), pack(rb, nl).data(),
                            pack(rc, nl).data(), pack(ea, nnl).data(), pack(eb, nnl).data(), pack(ec, nnl).data(), pack(d, nl).data(),
                            pack(rd, nl).data(), f.data(), z1.data(), z2.data(), ed.data(), edb.data(), fault.data()));
    std::vector<MulProof> out(B);
    for (size_t i = 0; i < B; ++i) {
      if (fault[i]) throw ReferencePanic("called `Option::unwrap()` on a `None` value (mod_inv, multiplication_proof.rs:96)");
      out[i].f = BigInt::from_limbs(&f[i * nl], nl);
      out[i].z1 = BigInt::from_limbs(&z1[i * nnl], nnl);
      out[i].z2 = BigInt::from_limbs(&z2[i * nnl], nnl);
      out[i].e_d = BigInt::from_limbs(&ed[i * nnl], nnl);
      out[i].e_db = BigInt::from_limbs(&edb[i * nnl], nnl);
    }
    return out;
  }
  static MulProof prove(Engine& eng, const MulWitness& w, const MulStatement& st, const ByteSource& rng = os_rng()) {
    return prove_batch(eng, {w}, {st}, rng)[0];
  }
  // -1: BigInt::mod_inv(..).unwrap() panics for that proof (multiplication_proof.rs:137)
  static std::vector<int> verify_batch(Engine& eng, const std::vector<const MulProof*>& ps, const std::vector<MulStatement>& st) {
    return verify_by_key(ps, st, [&](const std::vector<const MulProof*>& p, const std::vector<MulStatement>& s) { return verify_one_key(eng, p, s); });
  }
  void verify(Engine& eng, const MulStatement& st) const {
    throw_for_verdict(verify_batch(eng, {this}, {st})[0], "called `Option::unwrap()` on a `None` value (mod_inv, multiplication_proof.rs:137)");
  }

 private:
  static std::vector<int> verify_one_key(Engine& eng, const std::vector<const MulProof*>& ps, const std::vector<MulStatement>& st) {
    const size_t B = ps.size();
    eng.use_key(st[0].ek);
    const size_t nl = eng.nl(), nnl = eng.nnl();
    const BigInt& nn = st[0].ek.nn;
    std::vector<BigInt> ea, eb, ec, f, z1, z2, ed, edb;
    std::vector<int> out(B, 1);
    const BigInt zero(0);
    for (size_t i = 0; i < B; ++i) {
      // e_a, e_b, e_c, e_d, e_db are hashed as given; f is the exponent of e_b^f (:138): they must fit their rows
      const bool ok = fits_limbs(st[i].e_a, nnl) && fits_limbs(st[i].e_b, nnl) && fits_limbs(st[i].e_c, nnl) && fits_limbs(ps[i]->e_d, nnl) &&
                      fits_limbs(ps[i]->e_db, nnl) && fits_limbs(ps[i]->f, nl);
      if (!ok) out[i] = 0;
      ea.push_back(ok ? st[i].e_a : zero); eb.push_back(ok ? st[i].e_b : zero); ec.push_back(ok ? st[i].e_c : zero);
      f.push_back(ok ? ps[i]->f : zero); ed.push_back(ok ? ps[i]->e_d : zero); edb.push_back(ok ? ps[i]->e_db : zero);
      z1.push_back(ps[i]->z1 % nn);  // randomness of Enc(f, z1) / Enc(0, z2) (:118-131): reduced by mod_pow
      z2.push_back(ps[i]->z2 % nn);
    }
    std::vector<uint8_t> acc(B), fault(B);
    eng.check(zkp_mul_verify(eng.handle(), (int)B, pack(ea, nnl).data(), pack(eb, nnl).data(), pack(ec, nnl).data(), pack(f, nl).data(),
                             pack(z1, nnl).data(), pack(z2, nnl).data(), pack(ed, nnl).data(), pack(edb, nnl).data(), acc.data(), fault.data()));
    for (size_t i = 0; i < B; ++i)
      if (out[i]) out[i] = fault[i] ? -1 : acc[i];
    return out;
  }

 public:
  std::string to_json() const {
    return Json::object().set("f", ser_native(f)).set("z1", ser_native(z1)).set("z2", ser_native(z2)).set("e_d", ser_native(e_d)).set("e_db", ser_native(e_db)).dump();
  }
  static MulProof from_json(const std::string& s) {
    Json j = Json::parse(s);
    MulProof p;
    p.f = de_native(j.at("f")); p.z1 = de_native(j.at("z1")); p.z2 = de_native(j.at("z2")); p.e_d = de_native(j.at("e_d")); p.e_db = de_native(j.at("e_db"));
    return p;
  }
};

struct VerlinStatement { EncryptionKey ek; BigInt c, c_prime, phi_x; };         // verlin_proof.rs:51-57
struct VerlinWitness { BigInt x, x_prime, x_double_prime, r_x; };               // :43-49
class VerlinProof {                                                            // :34-41
 public:
  BigInt phi_a, z, z_prime, z_double_prime, r_z;
  static std::vector<VerlinProof> prove_batch(Engine& eng, const std::vector<VerlinWitness>& w, const std::vector<VerlinStatement>& st,
                                              const ByteSource& rng = os_rng()) {
    const size_t B = st.size();
    if (B == 0) return {};
    require_one_key(st, "VerlinProof::prove_batch");
    eng.use_key(st[0].ek);
    const size_t nl = eng.nl(), nnl = eng.nnl(), zl = eng.zl();
    std::vector<BigInt> x, xp, xdp, rx, c, cp, phix, a, ap, adp, ra;
    for (size_t i = 0; i < B; ++i) {
      x.push_back(w[i].x); xp.push_back(w[i].x_prime); xdp.push_back(w[i].x_double_prime); rx.push_back(w[i].r_x);
      c.push_back(st[i].c); cp.push_back(st[i].c_prime); phix.push_back(st[i].phi_x);
      const BigInt& n = st[i].ek.n;
      a.push_back(BigInt::sample_below(rng, n));             // :61
      ap.push_back(BigInt::sample_below(rng, n));            // :62
      adp.push_back(BigInt::sample_below(rng, n));           // :63
      ra.push_back(MulProof::sample_paillier_random(rng, n));  // :64-67
    }
    std::vector<uint32_t> phia(B * nnl), z(B * zl), zp(B * zl), zdp(B * zl), rz(B * nnl);
    eng.check(zkp_verlin_prove(eng.handle(), (int)B, (int)zl, pack(x, nl).data(), pack(xp, nl).data(), pack(xdp, nl).data(), pack(rx, nl).data(),
                               pack(c, nnl).data(), pack(cp, nnl).data(), pack(phix, nnl).data(), pack(a, nl).data(), pack(ap, nl).data(),
                               pack(adp, nl).data(), pack(ra, nl).data(), phia.data(), z.data(), zp.data(), zdp.data(), rz.data()));
    std::vector<VerlinProof> out(B);
    for (size_t i = 0; i < B; ++i) {
      out[i].phi_a = BigInt::from_limbs(&phia[i * nnl], nnl);
      out[i].z = BigInt::from_limbs(&z[i * zl], zl);
      out[i].z_prime = BigInt::from_limbs(&zp[i * zl], zl);
      out[i].z_double_prime = BigInt::from_limbs(&zdp[i * zl], zl);
      out[i].r_z = BigInt::from_limbs(&rz[i * nnl], nnl);
    }
    return out;
  }
  static VerlinProof prove(Engine& eng, const VerlinWitness& w, const VerlinStatement& st, const ByteSource& rng = os_rng()) {
    return prove_batch(eng, {w}, {st}, rng)[0];
  }
  static std::vector<int> verify_batch(Engine& eng, const std::vector<const VerlinProof*>& ps, const std::vector<VerlinStatement>& st) {
    return verify_by_key(ps, st, [&](const std::vector<const VerlinProof*>& p, const std::vector<VerlinStatement>& s) { return verify_one_key(eng, p, s); });
  }
  void verify(Engine& eng, const VerlinStatement& st) const { throw_for_verdict(verify_batch(eng, {this}, {st})[0], "unreachable"); }

 private:
  static std::vector<int> verify_one_key(Engine& eng, const std::vector<const VerlinProof*>& ps, const std::vector<VerlinStatement>& st) {
    const size_t B = ps.size();
    eng.use_key(st[0].ek);
    const size_t nnl = eng.nnl(), zl = eng.zl();
    const EncryptionKey& ek = st[0].ek;
    std::vector<BigInt> c, cp, phix, phia, z, zp, zdp, rz;
    std::vector<int> out(B, 1);
    const BigInt zero(0);
    for (size_t i = 0; i < B; ++i) {
      // c, c', phi_x, phi_a are hashed as given; z, z' are the exponents of gen_phi (:147-155): they must fit their rows
      const bool ok = fits_limbs(st[i].c, nnl) && fits_limbs(st[i].c_prime, nnl) && fits_limbs(st[i].phi_x, nnl) && fits_limbs(ps[i]->phi_a, nnl) &&
                      fits_limbs(ps[i]->z, zl) && fits_limbs(ps[i]->z_prime, zl);
      if (!ok) out[i] = 0;
      c.push_back(ok ? st[i].c : zero); cp.push_back(ok ? st[i].c_prime : zero); phix.push_back(ok ? st[i].phi_x : zero);
      phia.push_back(ok ? ps[i]->phi_a : zero); z.push_back(ok ? ps[i]->z : zero); zp.push_back(ok ? ps[i]->z_prime : zero);
      const BigInt& zdd = ps[i]->z_double_prime;
      zdp.push_back(fits_limbs(zdd, zl) ? zdd : zdd % ek.n);  // plaintext of Enc(z'', r_z) (:157-163): only z'' mod n matters
      rz.push_back(ps[i]->r_z % ek.nn);                        // its randomness: reduced by mod_pow
    }
    std::vector<uint8_t> acc(B);
    eng.check(zkp_verlin_verify(eng.handle(), (int)B, (int)zl, pack(c, nnl).data(), pack(cp, nnl).data(), pack(phix, nnl).data(),
                                pack(phia, nnl).data(), pack(z, zl).data(), pack(zp, zl).data(), pack(zdp, zl).data(), pack(rz, nnl).data(),
                                acc.data()));
    for (size_t i = 0; i < B; ++i)
      if (out[i]) out[i] = acc[i];
    return out;
  }

 public:
  std::string to_json() const {
    return Json::object().set("phi_a", ser_native(phi_a)).set("z", ser_native(z)).set("z_prime", ser_native(z_prime))
        .set("z_double_prime", ser_native(z_double_prime)).set("r_z", ser_native(r_z)).dump();
  }
  static VerlinProof from_json(const std::string& s) {
    Json j = Json::parse(s);
    VerlinProof p;
    p.phi_a = de_native(j.at("phi_a")); p.z = de_native(j.at("z")); p.z_prime = de_native(j.at("z_prime"));
    p.z_double_prime = de_native(j.at("z_double_prime")); p.r_z = de_native(j.at("r_z"));
    return p;
  }
};

// ------------------------------------------------------------------------------------ CorrectOpening
// `impl CorrectOpening for Paillier` (correct_opening.rs:17-30): c == encrypt_with_chosen_randomness(ek, m, r)
struct Paillier {
  // Paillier::encrypt_with_chosen_randomness (kzen-paillier; K1m on the device)
  static std::vector<BigInt> encrypt_with_chosen_randomness_batch(Engine& eng, const EncryptionKey& ek, const std::vector<BigInt>& m,
                                                                  const std::vector<BigInt>& r) {
    const size_t B = m.size();
    if (B == 0) return {};
    eng.use_key(ek);
    const size_t nl = eng.nl(), nnl = eng.nnl();
    std::vector<BigInt> mr, rr;
    for (size_t b = 0; b < B; ++b) { mr.push_back(m[b] % ek.n); rr.push_back(r[b] % ek.n); }
    std::vector<uint32_t> out(B * nnl);
    eng.check(zkp_paillier_enc(eng.handle(), pack(mr, nl).data(), (int)nl, pack(rr, nl).data(), (int)nl, (int)B, out.data()));
    return unpack(out, nnl);
  }
  // Paillier::encrypt: r = sample_below(n) per plaintext (kzen-paillier RECALLED), then the above.  Returns (c, r).
  static std::pair<std::vector<BigInt>, std::vector<BigInt>> encrypt_batch(Engine& eng, const EncryptionKey& ek, const std::vector<BigInt>& m,
                                                                           const ByteSource& rng = os_rng()) {
    std::vector<BigInt> r;
    for (size_t b = 0; b < m.size(); ++b) r.push_back(BigInt::sample_below(rng, ek.n));
    return {encrypt_with_chosen_randomness_batch(eng, ek, m, r), r};
  }
  // Paillier::decrypt by CRT (kzen-paillier RECALLED): m_p = L_p(c^(p-1) mod p^2) h_p mod p, likewise q, recombined.
  // The two half-width modexps per ciphertext run on the device (K2); the plaintext is the unique m in [0, n) either way.
  static std::vector<BigInt> decrypt_batch(Engine& eng, const DecryptionKey& dk, const std::vector<BigInt>& c) {
    const size_t B = c.size();
    if (B == 0) return {};
    const BigInt &p = dk.p, &q = dk.q, one(1);
    const BigInt n = p * q, pp = p * p, qq = q * q, pm1 = p - one, qm1 = q - one;
    auto L = [&](const BigInt& x, const BigInt& pr) { return (x - one) / pr; };
    auto h = [&](const BigInt& pr, const BigInt& prpr) {  // h_p = L_p((1 + n)^(p-1) mod p^2)^-1 mod p, and (1 + n)^(p-1) = 1 + (p-1) n mod p^2
      BigInt gp = (one + ((pr - one) * (n % prpr)) % prpr) % prpr, inv;
      if (!BigInt::mod_inv(L(gp, pr) % pr, pr, inv)) throw ReferencePanic("decrypt: L_p(g^(p-1)) is not invertible mod p");
      return inv;
    };
    const BigInt hp = h(p, pp), hq = h(q, qq);
    BigInt pinv;
    if (!BigInt::mod_inv(p % q, q, pinv)) throw ReferencePanic("decrypt: p is not invertible mod q");
    std::vector<BigInt> cp, cq;
    for (auto& x : c) { cp.push_back(x % pp); cq.push_back(x % qq); }
    std::vector<BigInt> dp = powm_batch(eng, cp, {pm1}, pp), dq = powm_batch(eng, cq, {qm1}, qq), out;
    for (size_t i = 0; i < B; ++i) {
      const BigInt mp = (L(dp[i], p) * hp) % p, mq = (L(dq[i], q) * hq) % q;
      const BigInt diff = (mq + q - (mp % q)) % q;
      out.push_back(mp + p * ((diff * pinv) % q));
    }
    return out;
  }
  // Paillier::keypair_with_modulus_size (kzen-paillier RECALLED: two primes of bits/2 bits drawn by rejection sampling with
  // a probabilistic primality test).  The primality tests are what costs: candidates are drawn a wave at a time, sieved by
  // small primes on the host, and the Miller-Rabin exponentiations a^d mod candidate of the whole wave (a distinct modulus per
  // job) run as one K2 launch.  Top two bits of each prime are set so that n has exactly `bits` bits.  `rounds` bases per
  // candidate, base 2 first.
  static DecryptionKey keypair_with_modulus_size(Engine& eng, size_t bits, const ByteSource& rng = os_rng(), int rounds = 24, size_t wave = 96) {
    if (bits < 256 || bits % 64) throw std::invalid_argument("keypair: modulus size must be a multiple of 64, at least 256");
    const size_t hb = bits / 2;
    static const std::vector<uint32_t> small = [] {
      std::vector<uint32_t> v;
      for (uint32_t x = 3; x < 4000; x += 2) {
        bool pr = true;
        for (uint32_t d = 3; d * d <= x; d += 2) if (x % d == 0) { pr = false; break; }
        if (pr) v.push_back(x);
      }
      return v;
    }();
    const BigInt one(1), two(2);
    std::vector<BigInt> primes;
    while (primes.size() < 2) {
      std::vector<BigInt> cand;
      while (cand.size() < wave) {
        BigInt c = BigInt::sample(rng, hb);
        c = c | one | one.shl(hb - 1) | one.shl(hb - 2);
        bool ok = true;
        for (uint32_t sp : small) if (c.mod_small(sp) == 0) { ok = false; break; }
        if (ok) cand.push_back(c);
      }
      // Miller-Rabin: cand - 1 = d 2^s; a^d on the device, the s - 1 squarings (s is 1 here: the low bits are random) on the host
      std::vector<BigInt> d, alive = cand;
      for (int round = 0; round < rounds && !alive.empty(); ++round) {
        std::vector<BigInt> bases, exps, mods;
        std::vector<size_t> ss;
        for (auto& c : alive) {
          BigInt dd = c - one;
          size_t sft = 0;
          while (!dd.is_odd()) { dd = dd.shr(1); ++sft; }
          ss.push_back(sft);
          exps.push_back(dd);
          mods.push_back(c);
          bases.push_back(round == 0 ? two : two + BigInt::sample_below(rng, c - BigInt(3)));
        }
        const size_t nl = limbs_for_bits(hb);
        std::vector<uint32_t> out(alive.size() * nl);
        eng.check(zkp_modexp_var(eng.handle(), pack(bases, nl).data(), pack(exps, nl).data(), (int)nl, (int)(32 * nl), 1, pack(mods, nl).data(),
                                 (int)nl, 1, (int)alive.size(), out.data()));
        std::vector<BigInt> x = unpack(out, nl), next;
        for (size_t i = 0; i < alive.size(); ++i) {
          const BigInt cm1 = alive[i] - one;
          bool pass = x[i] == one || x[i] == cm1;
          for (size_t k = 1; !pass && k < ss[i]; ++k) {
            x[i] = (x[i] * x[i]) % alive[i];
            pass = x[i] == cm1;
          }
          if (pass) next.push_back(alive[i]);
        }
        alive.swap(next);
      }
      for (auto& c : alive)
        if (primes.size() < 2 && (primes.empty() || primes[0] != c)) primes.push_back(c);
    }
    return DecryptionKey{primes[0], primes[1]};
  }
  // `count` key pairs of `bits`-bit moduli at once (synthetic workloads with a distinct modulus per proof: BASELINE configs[2]).
  // One search window per prime: a random start (top two bits set, = 3 mod 4, so that c - 1 = 2 d with d odd and one
  // Miller-Rabin exponentiation a^d decides a round), the window c = start + 4 k sieved by the primes below 4000 on the host,
  // and the surviving candidates of ALL unresolved windows tested together, base 2 first, as one zkp_modexp_var launch per wave
  // (a distinct modulus per job; ~70 candidates per prime at 1536 bits).  The first candidate of a window that passes is then
  // confirmed with `rounds - 1` random bases, again one launch per round over every window.
  static std::vector<DecryptionKey> keypairs_batch(Engine& eng, size_t bits, size_t count, const ByteSource& rng = os_rng(), int rounds = 8,
                                                   size_t wave_jobs = 131072) {
    if (bits < 256 || bits % 64) throw std::invalid_argument("keypairs_batch: modulus size must be a multiple of 64, at least 256");
    const size_t hb = bits / 2, nl = limbs_for_bits(hb), P = 2 * count;
    constexpr size_t W = 4096;  // candidates per window
    static const std::vector<uint32_t> small = [] {
      std::vector<uint32_t> v;
      for (uint32_t x = 3; x < 4000; x += 2) {
        bool pr = true;
        for (uint32_t d = 3; d * d <= x; d += 2) if (x % d == 0) { pr = false; break; }
        if (pr) v.push_back(x);
      }
      return v;
    }();
    struct Window {
      BigInt start;
      std::vector<uint16_t> alive;  // surviving k, ascending
      size_t next = 0;              // first untested survivor
      int confirmed = 0;            // Miller-Rabin rounds the current candidate has passed
      BigInt cand;
    };
    const BigInt one(1), two(2), four(4);
    auto fresh = [&](Window& w) {
      BigInt c = BigInt::sample(rng, hb);
      c = c | one | two | one.shl(hb - 1) | one.shl(hb - 2);  // = 3 mod 4
      w.start = c;
      std::vector<uint8_t> dead(W, 0);
      for (uint32_t sp : small) {
        const uint32_t r = c.mod_small(sp);
        // smallest k with r + 4 k = 0 mod sp
        uint32_t inv4 = 1;
        while ((4ull * inv4) % sp != 1) ++inv4;
        for (size_t k = (size_t)(((uint64_t)(sp - r) % sp) * inv4 % sp); k < W; k += sp) dead[k] = 1;
      }
      w.alive.clear();
      for (size_t k = 0; k < W; ++k) if (!dead[k]) w.alive.push_back((uint16_t)k);
      w.next = 0;
      w.confirmed = 0;
    };
    std::vector<Window> win(P);
    for (auto& w : win) fresh(w);
    auto candidate = [&](const Window& w, size_t j) { return w.start + BigInt((uint64_t)4 * w.alive[j]); };
    for (;;) {
      // ---- wave: base-2 tests for the windows without a candidate
      std::vector<size_t> open;
      for (size_t i = 0; i < P; ++i) if (win[i].confirmed == 0) open.push_back(i);
      if (!open.empty()) {
        const size_t per = std::max<size_t>(1, std::min<size_t>(64, wave_jobs / open.size()));
        std::vector<uint32_t> bases, exps, mods;
        std::vector<std::pair<size_t, size_t>> job;  // (window, survivor index)
        for (size_t i : open) {
          Window& w = win[i];
          if (w.next >= w.alive.size()) fresh(w);
          for (size_t j = w.next; j < std::min(w.next + per, w.alive.size()); ++j) job.emplace_back(i, j);
        }
        bases.assign(job.size() * nl, 0u);
        exps.resize(job.size() * nl);
        mods.resize(job.size() * nl);
        parallel_for(job.size(), 0, [&](size_t t) {
          const BigInt c = candidate(win[job[t].first], job[t].second);
          c.to_limbs(&mods[t * nl], nl);
          (c - one).shr(1).to_limbs(&exps[t * nl], nl);
          bases[t * nl] = 2u;
        });
        std::vector<uint32_t> out(job.size() * nl);
        eng.check(zkp_modexp_var(eng.handle(), bases.data(), exps.data(), (int)nl, (int)hb, 1, mods.data(), (int)nl, 1, (int)job.size(), out.data()));
        for (size_t t = 0; t < job.size(); ++t) {
          Window& w = win[job[t].first];
          if (w.confirmed) continue;  // an earlier candidate of this window already passed
          const BigInt x = BigInt::from_limbs(&out[t * nl], nl), c = candidate(w, job[t].second);
          w.next = job[t].second + 1;
          if (x == one || x == c - one) {
            w.confirmed = 1;
            w.cand = c;
          }
        }
      }
      // ---- confirmation rounds with random bases for every window that holds a candidate short of `rounds`
      std::vector<size_t> todo;
      for (size_t i = 0; i < P; ++i) if (win[i].confirmed > 0 && win[i].confirmed < rounds) todo.push_back(i);
      if (todo.empty() && open.empty()) break;
      if (!todo.empty()) {
        std::vector<BigInt> bases, exps, mods;
        for (size_t i : todo) {
          const BigInt& c = win[i].cand;
          bases.push_back(two + BigInt::sample_below(rng, c - BigInt(3)));
          exps.push_back((c - one).shr(1));
          mods.push_back(c);
        }
        std::vector<uint32_t> out(todo.size() * nl);
        eng.check(zkp_modexp_var(eng.handle(), pack(bases, nl).data(), pack(exps, nl).data(), (int)nl, (int)hb, 1, pack(mods, nl).data(), (int)nl, 1,
                                 (int)todo.size(), out.data()));
        for (size_t t = 0; t < todo.size(); ++t) {
          Window& w = win[todo[t]];
          const BigInt x = BigInt::from_limbs(&out[t * nl], nl);
          if (x == one || x == w.cand - one) ++w.confirmed;
          else w.confirmed = 0;  // a base-2 pseudoprime: the window goes on from the next survivor
        }
      }
    }
    std::vector<DecryptionKey> keys(count);
    for (size_t k = 0; k < count; ++k) {
      keys[k] = DecryptionKey{win[2 * k].cand, win[2 * k + 1].cand};
      if (keys[k].p == keys[k].q) throw std::runtime_error("keypairs_batch: equal primes (broken randomness source)");
    }
    return keys;
  }
  // Paillier::open (kzen-paillier RECALLED): (m, r) with c = Enc(m, r); r = extract_nroot(dk, c (1 + m n)^-1 mod n)
  // and (1 + m n)^-1 = 1 (mod n), so r is the n-th root of c mod n.  Used at correct_opening.rs:52-53.
  static std::pair<std::vector<BigInt>, std::vector<BigInt>> open_batch(Engine& eng, const DecryptionKey& dk, const std::vector<BigInt>& c);
  static std::vector<int> verify_opening_batch(Engine& eng, const EncryptionKey& ek, const std::vector<BigInt>& m, const std::vector<BigInt>& r,
                                               const std::vector<BigInt>& c) {
    const size_t B = m.size();
    if (B == 0) return {};
    eng.use_key(ek);
    const size_t nl = eng.nl(), nnl = eng.nnl();
    std::vector<BigInt> mr, rr, cc;
    std::vector<int> out(B, 1);
    for (size_t b = 0; b < B; ++b) {
      mr.push_back(m[b] % ek.n);  // (m n + 1) % nn only depends on m mod n
      rr.push_back(r[b] % ek.n);  // r^n mod nn only depends on r mod n
      if (!fits_limbs(c[b], nnl)) out[b] = 0;  // wider than n^2: never equal to a canonical ciphertext
      cc.push_back(out[b] ? c[b] : BigInt(0));
    }
    std::vector<uint8_t> ok(B);
    eng.check(zkp_verify_opening(eng.handle(), (int)B, (int)nl, pack(mr, nl).data(), pack(rr, nl).data(), pack(cc, nnl).data(), ok.data()));
    for (size_t b = 0; b < B; ++b)
      if (out[b]) out[b] = ok[b];
    return out;
  }
  static bool verify_opening(Engine& eng, const EncryptionKey& ek, const BigInt& m, const BigInt& r, const BigInt& c) {
    return verify_opening_batch(eng, ek, {m}, {r}, {c})[0] != 0;
  }
};

inline std::pair<std::vector<BigInt>, std::vector<BigInt>> Paillier::open_batch(Engine& eng, const DecryptionKey& dk, const std::vector<BigInt>& c) {
  const BigInt n = dk.p * dk.q;
  std::vector<BigInt> cn;
  for (auto& x : c) cn.push_back(x % n);
  return {decrypt_batch(eng, dk, c), extract_nroots(eng, dk, cn)};
}

// ------------------------------------------------------------------------------------ CompositeDLogProof
struct DLogStatement {  // wi_dlog_proof.rs:34-39
  BigInt N, g, ni;
  std::string to_json() const { return Json::object().set("N", ser_native(N)).set("g", ser_native(g)).set("ni", ser_native(ni)).dump(); }
  static DLogStatement from_json(const std::string& s) {
    Json j = Json::parse(s);
    return {de_native(j.at("N")), de_native(j.at("g")), de_native(j.at("ni"))};
  }
};
class CompositeDLogProof {  // wi_dlog_proof.rs:28-32
 public:
  static constexpr size_t K = 128, K_PRIME = 128, SAMPLE_S = 256;  // :19-21
  BigInt x, y;
  static size_t width(const std::vector<DLogStatement>& st) {
    size_t bits = 0;
    for (auto& s : st) bits = std::max(bits, std::max(s.N.bit_length(), std::max(s.g.bit_length(), s.ni.bit_length())));
    return limbs_for_bits(bits);
  }
  static std::vector<CompositeDLogProof> prove_batch(Engine& eng, const std::vector<DLogStatement>& st, const std::vector<BigInt>& secret,
                                                     const ByteSource& rng = os_rng()) {
    const size_t B = st.size();
    if (B == 0) return {};
    const size_t nl = width(st);
    BigInt R = BigInt(1).shl(K + K_PRIME + SAMPLE_S);                                    // :51
    std::vector<BigInt> N, g, ni, r;
    size_t sbits = 1;
    for (size_t b = 0; b < B; ++b) {
      N.push_back(st[b].N); g.push_back(st[b].g); ni.push_back(st[b].ni);
      r.push_back(BigInt::sample_below(rng, R));                                          // :52
      sbits = std::max(sbits, secret[b].bit_length());
    }
    const size_t sl = limbs_for_bits(sbits), rl = limbs_for_bits(K + K_PRIME + SAMPLE_S);
    const size_t yl = round4(std::max(rl, sl + 8) + 1);                                   // y = r + e * secret, e < 2^256
    std::vector<uint32_t> x(B * nl), y(B * yl);
    std::vector<uint8_t> fault(B);
    eng.check(zkp_dlog_prove(eng.handle(), (int)B, (int)nl, pack(N, nl).data(), pack(g, nl).data(), pack(ni, nl).data(), pack(secret, sl).data(),
                             (int)sl, pack(r, rl).data(), (int)rl, (int)yl, x.data(), y.data(), fault.data()));
    std::vector<CompositeDLogProof> out(B);
    for (size_t b = 0; b < B; ++b) {
      if (fault[b]) throw std::length_error("y wider than its rows");
      out[b].x = BigInt::from_limbs(&x[b * nl], nl);
      out[b].y = BigInt::from_limbs(&y[b * yl], yl);
    }
    return out;
  }
  static CompositeDLogProof prove(Engine& eng, const DLogStatement& st, const BigInt& secret, const ByteSource& rng = os_rng()) {
    return prove_batch(eng, {st}, {secret}, rng)[0];
  }
  // 1 accept, 0 Err(IncorrectProof), -1 where the reference's assert! / assert_eq! panics (:68-72)
  static std::vector<int> verify_batch(Engine& eng, const std::vector<const CompositeDLogProof*>& ps, const std::vector<DLogStatement>& st) {
    const size_t B = ps.size();
    if (B == 0) return {};
    const size_t nl = width(st);
    size_t ybits = 1;
    std::vector<BigInt> N, g, ni, x, y;
    std::vector<int> out(B, 0);
    std::vector<size_t> idx;
    for (size_t b = 0; b < B; ++b) {
      if (!st[b].N.is_odd() || ps[b]->x.d.size() > nl) {  // an even N has no Montgomery form: settle these few on the host rules
        if (!st[b].N.is_odd()) throw std::domain_error("CompositeDLogProof: even modulus is not supported by the engine");
        out[b] = 0;  // x >= 2^(32 nl) > N can never equal a residue modulo N; the asserts are still checked below
        if (!(st[b].N > BigInt(1).shl(K)) || BigInt::gcd(st[b].g, st[b].N) != BigInt(1) || BigInt::gcd(st[b].ni, st[b].N) != BigInt(1)) out[b] = -1;
        continue;
      }
      idx.push_back(b);
      N.push_back(st[b].N); g.push_back(st[b].g); ni.push_back(st[b].ni); x.push_back(ps[b]->x); y.push_back(ps[b]->y);
      ybits = std::max(ybits, ps[b]->y.bit_length());
    }
    if (idx.empty()) return out;
    const size_t yl = limbs_for_bits(ybits);
    std::vector<uint8_t> acc(idx.size()), fault(idx.size());
    eng.check(zkp_dlog_verify(eng.handle(), (int)idx.size(), (int)nl, pack(N, nl).data(), pack(g, nl).data(), pack(ni, nl).data(),
                              pack(x, nl).data(), pack(y, yl).data(), (int)yl, acc.data(), fault.data()));
    for (size_t k = 0; k < idx.size(); ++k) out[idx[k]] = fault[k] ? -1 : acc[k];
    return out;
  }
  void verify(Engine& eng, const DLogStatement& st) const {
    const int v = verify_batch(eng, {this}, {st})[0];
    if (v < 0) throw ReferencePanic("assertion failed: N > 2^K, gcd(g, N) == 1, gcd(ni, N) == 1");
    if (!v) throw IncorrectProof();
  }
  std::string to_json() const { return Json::object().set("x", ser_native(x)).set("y", ser_native(y)).dump(); }
  static CompositeDLogProof from_json(const std::string& s) {
    Json j = Json::parse(s);
    CompositeDLogProof p;
    p.x = de_native(j.at("x"));
    p.y = de_native(j.at("y"));
    return p;
  }
};

// ------------------------------------------------------------------------------------ CorrectMessageProof
class CorrectMessageProof {  // correct_message.rs:25-32 (no serde derive in the reference)
 public:
  static constexpr size_t B_BITS = 256;  // :19
  std::vector<BigInt> e_vec, z_vec, a_vec;
  BigInt ciphertext;
  std::vector<BigInt> valid_messages;
  EncryptionKey ek;
  CorrectMessageProof() : ek(BigInt(1)) {}
  static std::vector<CorrectMessageProof> prove_batch(Engine& eng, const EncryptionKey& ek, const std::vector<BigInt>& valid_messages,
                                                      const std::vector<BigInt>& messages, const ByteSource& rng = os_rng()) {
    const size_t B = messages.size(), M = valid_messages.size();
    if (B == 0) return {};
    if (M == 0) throw ReferencePanic("attempt to subtract with overflow (no valid messages)");
    eng.use_key(ek);
    const size_t nl = eng.nl(), nnl = eng.nnl();
    std::vector<BigInt> valid, msg, r, er, zr, w;
    for (size_t b = 0; b < B; ++b) {
      r.push_back(BigInt::sample_below(rng, ek.n));                                       // :41
      for (auto& v : valid_messages) valid.push_back(v % ek.n);                            // (m n + 1) % nn depends on m mod n only
      msg.push_back(messages[b]);
      for (size_t j = 0; j + 1 < M; ++j) er.push_back(BigInt::sample(rng, B_BITS));        // :58-60
      for (size_t j = 0; j + 1 < M; ++j) zr.push_back(BigInt::sample_below(rng, ek.n));    // :61-63
      w.push_back(BigInt::sample_below(rng, ek.n));                                       // :65
    }
    // the engine matches slots by comparing rows: compare what the reference compares (the unreduced values)
    std::vector<uint32_t> valid_rows(B * M * nl), msg_rows(B * nl);
    for (size_t b = 0; b < B; ++b) {
      bool any = false;
      for (size_t i = 0; i < M; ++i) {
        const bool eq = valid_messages[i] == messages[b];
        any = any || eq;
        // a slot that does not match must not compare equal after reduction either: give it the reduced value only when it
        // differs from the reduced message, otherwise the proof below would take a branch the reference does not take
        BigInt v = valid_messages[i] % ek.n;
        if (!eq && v == messages[b] % ek.n) throw std::domain_error("CorrectMessageProof: two distinct messages congruent modulo n");
        v.to_limbs(&valid_rows[(b * M + i) * nl], nl);
      }
      (messages[b] % ek.n).to_limbs(&msg_rows[b * nl], nl);
      if (!any) throw ReferencePanic("index out of bounds: the message is not one of the valid messages");
    }
    std::vector<uint32_t> c(B * nnl), e(B * M * 8), z(B * M * nl), a(B * M * nnl);
    std::vector<uint8_t> fault(B);
    eng.check(zkp_correct_message_prove(eng.handle(), (int)B, (int)M, (int)nl, valid_rows.data(), msg_rows.data(), pack(r, nl).data(),
                                        M > 1 ? pack(er, 8).data() : nullptr, M > 1 ? pack(zr, nl).data() : nullptr, pack(w, nl).data(), c.data(),
                                        e.data(), z.data(), a.data(), fault.data()));
    std::vector<CorrectMessageProof> out(B);
    for (size_t b = 0; b < B; ++b) {
      if (fault[b]) throw ReferencePanic("called `Option::unwrap()` on a `None` value (mod_inv)");
      out[b].ciphertext = BigInt::from_limbs(&c[b * nnl], nnl);
      for (size_t i = 0; i < M; ++i) {
        out[b].e_vec.push_back(BigInt::from_limbs(&e[(b * M + i) * 8], 8));
        out[b].z_vec.push_back(BigInt::from_limbs(&z[(b * M + i) * nl], nl));
        out[b].a_vec.push_back(BigInt::from_limbs(&a[(b * M + i) * nnl], nnl));
      }
      out[b].valid_messages = valid_messages;
      out[b].ek = ek;
    }
    return out;
  }
  static CorrectMessageProof prove(Engine& eng, const EncryptionKey& ek, const std::vector<BigInt>& valid_messages, const BigInt& message,
                                   const ByteSource& rng = os_rng()) {
    return prove_batch(eng, ek, valid_messages, {message}, rng)[0];
  }
  // 1 accept, 0 Err(IncorrectProof), -1 where assert_eq!(chal, ei_sum) panics (:133); proofs of one batch share ek and M
  static std::vector<int> verify_batch(Engine& eng, const std::vector<const CorrectMessageProof*>& ps) {
    const size_t B = ps.size();
    if (B == 0) return {};
    const EncryptionKey& ek = ps[0]->ek;
    const size_t M = ps[0]->valid_messages.size();
    eng.use_key(ek);
    const size_t nl = eng.nl(), nnl = eng.nnl();
    size_t ebits = 256;
    for (auto* p : ps) {
      if (!(p->ek == ek) || p->valid_messages.size() != M) throw std::invalid_argument("verify_batch: proofs must share the key and the number of messages");
      if (p->e_vec.size() < M || p->z_vec.size() < M || p->a_vec.size() < M) throw ReferencePanic("index out of bounds: short proof vector");
      // the reference hashes ALL of a_vec and folds ALL of e_vec (correct_message.rs:128-131); the device rows hold M of each
      if (p->e_vec.size() != M || p->a_vec.size() != M)
        throw std::invalid_argument("CorrectMessageProof::verify_batch: e_vec / a_vec longer than the message list are not supported");
      for (auto& e : p->e_vec) ebits = std::max(ebits, e.bit_length());
    }
    const size_t el = limbs_for_bits(ebits);
    std::vector<BigInt> c, valid, e, z, a;
    for (auto* p : ps) {
      c.push_back(p->ciphertext % ek.nn);
      for (size_t i = 0; i < M; ++i) {
        valid.push_back(p->valid_messages[i] % ek.n);
        e.push_back(p->e_vec[i]);
        z.push_back(p->z_vec[i] % ek.n);   // z^n mod nn depends on z mod n only
        a.push_back(p->a_vec[i]);
      }
    }
    // the transcript hashes a_vec as given: a proof with a row wider than n^2 cannot be laid out and is rejected on its own
    std::vector<int> wide(B, 0);
    for (size_t b = 0; b < B; ++b)
      for (size_t i = 0; i < M; ++i)
        if (a[b * M + i].d.size() > nnl) {
          wide[b] = 1;
          a[b * M + i] = BigInt(0);
        }
    std::vector<uint8_t> acc(B), fault(B);
    eng.check(zkp_correct_message_verify(eng.handle(), (int)B, (int)M, (int)nl, (int)el, pack(c, nnl).data(), pack(valid, nl).data(),
                                         pack(e, el).data(), pack(z, nl).data(), pack(a, nnl).data(), acc.data(), fault.data()));
    std::vector<int> out(B);
    for (size_t b = 0; b < B; ++b) out[b] = wide[b] ? 0 : (fault[b] ? -1 : acc[b]);
    return out;
  }
  void verify(Engine& eng) const {
    const int v = verify_batch(eng, {this})[0];
    if (v < 0) throw ReferencePanic("assertion failed: `(left == right)` (chal, ei_sum)");
    if (!v) throw IncorrectProof();
  }
};

}  // namespace zkproofs
