// JSON-in / JSON-out C entry point over the C++ host mirror (zkproofs.hpp), so the pytest suite can drive
// the reference-shaped interface (prove / verify, serde wire format, panics vs Err) end to end:
//     char* zkh_call(const char* op, const char* request_json);   // caller frees with zkh_free
// Every response is {"ok": true, ...} or {"ok": false, "error": "...", "kind": "panic" | "error"}.
#include <cstdlib>
#include <cstring>
#include <map>

#include "zkproofs.hpp"

using namespace zkproofs;

namespace {

Engine& engine(int device) {
  static std::map<int, std::unique_ptr<Engine>> engines;
  auto& e = engines[device];
  if (!e) e.reset(new Engine(device));
  return *e;
}

// deterministic byte source from a hex string (tests); falls back to OsRng when absent
ByteSource rng_from(const Json& req) {
  const Json* h = req.find("rng_hex");
  if (!h) return os_rng();
  auto buf = std::make_shared<std::vector<uint8_t>>();
  const std::string& s = h->as_str();
  for (size_t i = 0; i + 1 < s.size(); i += 2) buf->push_back((uint8_t)std::stoi(s.substr(i, 2), nullptr, 16));
  auto pos = std::make_shared<size_t>(0);
  return [buf, pos](uint8_t* p, size_t n) {
    if (*pos + n > buf->size()) throw std::runtime_error("rng_hex exhausted");
    memcpy(p, buf->data() + *pos, n);
    *pos += n;
  };
}

BigInt dec(const Json& j, const char* k) { return BigInt::from_dec(j.at(k).as_str()); }
std::vector<uint8_t> unhex(const std::string& s) {
  std::vector<uint8_t> v;
  for (size_t i = 0; i + 1 < s.size(); i += 2) v.push_back((uint8_t)std::stoi(s.substr(i, 2), nullptr, 16));
  return v;
}

// run `fn` per item, mapping the three outcomes of a reference verify
template <class F>
Json outcomes(size_t n, F fn) {
  Json arr = Json::array();
  for (size_t i = 0; i < n; ++i) {
    try {
      fn(i);
      arr.push(Json::string("ok"));
    } catch (const IncorrectProof&) {
      arr.push(Json::string("incorrect"));
    } catch (const ReferencePanic& e) {
      arr.push(Json::string(std::string("panic: ") + e.what()));
    }
  }
  return arr;
}

Json dispatch(const std::string& op, const Json& req) {
  const int device = req.find("device") ? (int)req.at("device").as_num() : 0;
  Json out = Json::object();
  out.set("ok", Json::boolean(true));
  if (op == "bigint.selftest") {
    BigInt a = dec(req, "a"), b = dec(req, "b");
    out.set("sum", ser_dec(a + b));
    out.set("prod", ser_dec(a * b));
    if (a >= b) out.set("diff", ser_dec(a - b));
    if (!b.is_zero()) {
      out.set("quot", ser_dec(a / b));
      out.set("rem", ser_dec(a % b));
      BigInt inv;
      if (BigInt::mod_inv(a, b, inv)) out.set("inv", ser_dec(inv));
    }
    out.set("gcd", ser_dec(BigInt::gcd(a, b)));
    out.set("or", ser_dec(a | b));
    out.set("shl", ser_dec(a.shl(37)));
    out.set("shr", ser_dec(a.shr(37)));
    out.set("mod_small", Json::number((int64_t)a.mod_small(4093u)));
    out.set("hex", ser_native(a));
    out.set("bits", Json::number((int64_t)a.bit_length()));
    BigInt back;
    BigInt::parse_hex_bytes(a.to_hex_bytes(), back);
    out.set("roundtrip", Json::boolean(back == a && BigInt::from_dec(a.to_dec()) == a && BigInt::from_limbs(a.to_limbs(a.d.size() + 3).data(), a.d.size() + 3) == a));
    return out;
  }
  if (op == "sample") {  // the sampling rules of SURVEY 8a a15 on a given byte stream
    ByteSource rng = rng_from(req);
    Json arr = Json::array();
    BigInt lo = dec(req, "lo"), hi = dec(req, "hi");
    for (int64_t i = 0; i < req.at("count").as_num(); ++i) arr.push(ser_dec(BigInt::sample_range(rng, lo, hi)));
    out.set("values", arr);
    return out;
  }
  Engine& eng = engine(device);
  if (op == "rangeproof_ni.prove") {
    EncryptionKey ek(dec(req, "n"));
    std::vector<RangeStatement> st;
    for (auto& s : req.at("statements").arr) st.push_back({dec(s, "range"), dec(s, "ciphertext"), dec(s, "x"), dec(s, "r")});
    size_t ef = req.find("error_factor") ? (size_t)req.at("error_factor").as_num() : SECURITY_PARAMETER;
    auto proofs = RangeProofNi::prove_batch(eng, ek, st, rng_from(req), ef);
    Json arr = Json::array();
    for (auto& p : proofs) arr.push(Json::string(p.to_json()));
    out.set("proofs", arr);
    return out;
  }
  if (op == "rangeproof_ni.verify") {  // one proof at a time through verify(ek, ciphertext), like a caller of the crate
    EncryptionKey ek(dec(req, "n"));
    std::vector<RangeProofNi> ps;
    for (auto& s : req.at("proofs").arr) ps.push_back(RangeProofNi::from_json(s.as_str()));
    const Json& cts = req.at("ciphertexts");
    out.set("results", outcomes(ps.size(), [&](size_t i) { ps[i].verify(eng, ek, BigInt::from_dec(cts.arr[i].as_str())); }));
    return out;
  }
  if (op == "rangeproof_ni.verify_batch") {
    std::vector<RangeProofNi> ps;
    for (auto& s : req.at("proofs").arr) ps.push_back(RangeProofNi::from_json(s.as_str()));
    std::vector<const RangeProofNi*> ptrs;
    for (auto& p : ps) ptrs.push_back(&p);
    Json arr = Json::array();
    for (int v : RangeProofNi::verify_batch(eng, ptrs)) arr.push(Json::number(v));
    out.set("accept", arr);
    return out;
  }
  if (op == "rangeproof.flow") {  // the whole interactive protocol (range_proof.rs:431-525), both parties on this engine
    EncryptionKey ek(dec(req, "n"));
    const BigInt range = dec(req, "range"), x = dec(req, "x"), r = dec(req, "r"), cx = dec(req, "ciphertext");
    ByteSource rng = rng_from(req);
    const size_t ef = STATISTICAL_ERROR_FACTOR;
    auto vc = RangeProof::verifier_commit(eng, ek, rng);                        // verifier
    auto pd = RangeProof::generate_encrypted_pairs(eng, ek, range, ef, rng);    // prover
    RangeProof::verify_commit(eng, ek, vc.com, vc.r, vc.e);                     // prover checks the opening
    Proof proof = RangeProof::generate_proof(ek, x, r, vc.e, range, pd.second, ef);
    std::string result = "ok";
    try {
      RangeProof::verifier_output(eng, ek, vc.e, pd.first, proof, range, cx, ef);
    } catch (const IncorrectProof&) {
      result = "incorrect";
    }
    bool bad_open = false;
    try {
      ChallengeBits e2 = vc.e;
      e2.bytes[0] ^= 1;
      RangeProof::verify_commit(eng, ek, vc.com, vc.r, e2);
    } catch (const IncorrectProof&) {
      bad_open = true;
    }
    RangeProofNi carrier;  // reuse the serde of the NI struct to ship pairs + responses to the test
    carrier.ek = ek; carrier.range = range; carrier.ciphertext = cx; carrier.encrypted_pairs = pd.first; carrier.proof = proof; carrier.error_factor = ef;
    std::string ehex;
    for (uint8_t b : vc.e.bytes) { char t[3]; snprintf(t, 3, "%02x", b); ehex += t; }
    out.set("result", Json::string(result));
    out.set("e_hex", Json::string(ehex));
    out.set("com", ser_dec(vc.com.com));
    out.set("com_r", ser_dec(vc.r.r));
    out.set("tampered_opening_rejected", Json::boolean(bad_open));
    out.set("transcript", Json::string(carrier.to_json()));
    return out;
  }
  if (op == "correct_key.challenge") {
    EncryptionKey ek(dec(req, "n"));
    auto cv = CorrectKey::challenge(eng, ek, rng_from(req));
    out.set("challenge", Json::string(cv.first.to_json()));
    out.set("verification_aid", Json::object().set("s_digest", ser_dec(cv.second.s_digest)));
    return out;
  }
  if (op == "correct_key.prove") {
    DecryptionKey dk{dec(req, "p"), dec(req, "q")};
    try {
      CorrectKeyProof pr = CorrectKey::prove(eng, dk, Challenge::from_json(req.at("challenge").as_str()));
      out.set("proof", Json::object().set("s_digest", ser_dec(pr.s_digest)));
      try {
        CorrectKey::verify(pr, VerificationAid{dec(req, "s_digest")});
        out.set("verify", Json::string("ok"));
      } catch (const IncorrectProof&) {
        out.set("verify", Json::string("incorrect"));
      }
    } catch (const CorrectKeyProveError& e) {
      out.set("prove_error", Json::string(e.what()));
    }
    return out;
  }
  if (op == "correct_key_ni.proof") {
    DecryptionKey dk{dec(req, "p"), dec(req, "q")};
    NiCorrectKeyProof pr;
    if (req.find("salt_hex")) {
      auto salt = unhex(req.at("salt_hex").as_str());
      static const uint8_t none = 0;
      pr = NiCorrectKeyProof::proof(eng, dk, salt.empty() ? &none : salt.data(), salt.size());
    } else {
      pr = NiCorrectKeyProof::proof(eng, dk);
    }
    out.set("proof", Json::string(pr.to_json()));
    return out;
  }
  if (op == "correct_key_ni.verify") {
    auto salt = unhex(req.at("salt_hex").as_str());
    static const uint8_t none = 0;
    std::vector<NiCorrectKeyProof> ps;
    std::vector<EncryptionKey> eks;
    for (auto& s : req.at("proofs").arr) ps.push_back(NiCorrectKeyProof::from_json(s.as_str()));
    for (auto& s : req.at("n").arr) eks.push_back(EncryptionKey(BigInt::from_dec(s.as_str())));
    out.set("results", outcomes(ps.size(), [&](size_t i) { ps[i].verify(eng, eks[i], salt.empty() ? &none : salt.data(), salt.size()); }));
    return out;
  }
  // CompositeDLogProof: {"items": [{"N","g","ni","secret" | "proof"}]} (decimal strings; no key)
  if (op == "dlog.prove" || op == "dlog.verify") {
    const auto& its = req.at("items").arr;
    std::vector<DLogStatement> st;
    for (auto& it : its) st.push_back({dec(it, "N"), dec(it, "g"), dec(it, "ni")});
    if (op == "dlog.prove") {
      std::vector<BigInt> secret;
      for (auto& it : its) secret.push_back(dec(it, "secret"));
      Json arr = Json::array(), sts = Json::array();
      for (auto& p : CompositeDLogProof::prove_batch(eng, st, secret, rng_from(req))) arr.push(Json::string(p.to_json()));
      for (auto& s : st) sts.push(Json::string(DLogStatement::from_json(s.to_json()).to_json()));  // serde round trip of the statement
      out.set("proofs", arr).set("statements", sts);
    } else {
      out.set("results", outcomes(its.size(), [&](size_t i) { CompositeDLogProof::from_json(its[i].at("proof").as_str()).verify(eng, st[i]); }));
    }
    return out;
  }
  // CorrectMessageProof: {"n", "valid": [...], "messages": [...]} -> proofs as plain JSON objects of decimal strings
  // (the reference derives no serde for it); cmsg.verify takes them back
  if (op == "cmsg.prove" || op == "cmsg.verify") {
    EncryptionKey cek(dec(req, "n"));
    auto decs = [](const Json& a) { std::vector<BigInt> v; for (auto& x : a.arr) v.push_back(BigInt::from_dec(x.as_str())); return v; };
    auto encs = [](const std::vector<BigInt>& v) { Json a = Json::array(); for (auto& x : v) a.push(Json::string(x.to_dec())); return a; };
    if (op == "cmsg.prove") {
      Json arr = Json::array();
      for (auto& p : CorrectMessageProof::prove_batch(eng, cek, decs(req.at("valid")), decs(req.at("messages")), rng_from(req)))
        arr.push(Json::object().set("e_vec", encs(p.e_vec)).set("z_vec", encs(p.z_vec)).set("a_vec", encs(p.a_vec))
                     .set("ciphertext", Json::string(p.ciphertext.to_dec())).set("valid_messages", encs(p.valid_messages)));
      out.set("proofs", arr);
    } else {
      const auto& ps = req.at("proofs").arr;
      out.set("results", outcomes(ps.size(), [&](size_t i) {
        CorrectMessageProof p;
        p.e_vec = decs(ps[i].at("e_vec")); p.z_vec = decs(ps[i].at("z_vec")); p.a_vec = decs(ps[i].at("a_vec"));
        p.ciphertext = dec(ps[i], "ciphertext"); p.valid_messages = decs(ps[i].at("valid_messages")); p.ek = cek;
        p.verify(eng);
      }));
    }
    return out;
  }
  if (op == "paillier.keypair") {  // {"bits", "rng_hex"?}
    DecryptionKey dk = Paillier::keypair_with_modulus_size(eng, (size_t)req.at("bits").as_num(), rng_from(req));
    out.set("p", Json::string(dk.p.to_dec())).set("q", Json::string(dk.q.to_dec()));
    return out;
  }
  if (op == "paillier.flow") {  // {"p","q","m":[...],"rng_hex"}: encrypt -> open -> verify_opening (correct_opening.rs:47-56), batched
    DecryptionKey dk{dec(req, "p"), dec(req, "q")};
    EncryptionKey fek(dk.p * dk.q);
    std::vector<BigInt> m;
    for (auto& x : req.at("m").arr) m.push_back(BigInt::from_dec(x.as_str()));
    auto enc = Paillier::encrypt_batch(eng, fek, m, rng_from(req));
    auto opened = Paillier::open_batch(eng, dk, enc.first);
    auto ok = Paillier::verify_opening_batch(eng, fek, opened.first, opened.second, enc.first);
    Json cs = Json::array(), ms = Json::array(), rs = Json::array(), oks = Json::array();
    for (size_t i = 0; i < m.size(); ++i) {
      cs.push(Json::string(enc.first[i].to_dec())); ms.push(Json::string(opened.first[i].to_dec()));
      rs.push(Json::string(opened.second[i].to_dec())); oks.push(Json::boolean(ok[i] != 0));
    }
    out.set("c", cs).set("m", ms).set("r", rs).set("ok", Json::boolean(true)).set("opening_ok", oks);
    return out;
  }
  if (op == "opening.verify") {  // {"n", "items": [{"m","r","c"}]}
    EncryptionKey oek(dec(req, "n"));
    std::vector<BigInt> m, r, c;
    for (auto& it : req.at("items").arr) { m.push_back(dec(it, "m")); r.push_back(dec(it, "r")); c.push_back(dec(it, "c")); }
    Json arr = Json::array();
    for (int v : Paillier::verify_opening_batch(eng, oek, m, r, c)) arr.push(Json::boolean(v != 0));
    out.set("results", arr);
    return out;
  }
  // sigma protocols: {"n", "items": [{witness/statement/proof fields as decimal strings}], "rng_hex"}
  EncryptionKey ek(dec(req, "n"));
  const auto& items = req.at("items").arr;
  if (op == "zero.prove") {
    std::vector<ZeroWitness> w; std::vector<ZeroStatement> st;
    for (auto& it : items) { w.push_back({dec(it, "r")}); st.push_back({ek, dec(it, "c")}); }
    Json arr = Json::array();
    for (auto& p : ZeroProof::prove_batch(eng, w, st, rng_from(req))) arr.push(Json::string(p.to_json()));
    out.set("proofs", arr);
  } else if (op == "zero.verify") {
    out.set("results", outcomes(items.size(), [&](size_t i) { ZeroProof::from_json(items[i].at("proof").as_str()).verify(eng, {ek, dec(items[i], "c")}); }));
  } else if (op == "ciphertext.prove") {
    std::vector<CiphertextWitness> w; std::vector<CiphertextStatement> st;
    for (auto& it : items) { w.push_back({dec(it, "x"), dec(it, "r")}); st.push_back({ek, dec(it, "c")}); }
    Json arr = Json::array();
    for (auto& p : CiphertextProof::prove_batch(eng, w, st, rng_from(req))) arr.push(Json::string(p.to_json()));
    out.set("proofs", arr);
  } else if (op == "ciphertext.verify") {
    out.set("results", outcomes(items.size(), [&](size_t i) { CiphertextProof::from_json(items[i].at("proof").as_str()).verify(eng, {ek, dec(items[i], "c")}); }));
  } else if (op == "mul.prove") {
    std::vector<MulWitness> w; std::vector<MulStatement> st;
    for (auto& it : items) {
      w.push_back({dec(it, "a"), dec(it, "b"), dec(it, "c"), dec(it, "r_a"), dec(it, "r_b"), dec(it, "r_c")});
      st.push_back({ek, dec(it, "e_a"), dec(it, "e_b"), dec(it, "e_c")});
    }
    Json arr = Json::array();
    for (auto& p : MulProof::prove_batch(eng, w, st, rng_from(req))) arr.push(Json::string(p.to_json()));
    out.set("proofs", arr);
  } else if (op == "mul.verify") {
    out.set("results", outcomes(items.size(), [&](size_t i) {
      MulProof::from_json(items[i].at("proof").as_str()).verify(eng, {ek, dec(items[i], "e_a"), dec(items[i], "e_b"), dec(items[i], "e_c")});
    }));
  } else if (op == "verlin.prove") {
    std::vector<VerlinWitness> w; std::vector<VerlinStatement> st;
    for (auto& it : items) {
      w.push_back({dec(it, "x"), dec(it, "x_prime"), dec(it, "x_double_prime"), dec(it, "r_x")});
      st.push_back({ek, dec(it, "c"), dec(it, "c_prime"), dec(it, "phi_x")});
    }
    Json arr = Json::array();
    for (auto& p : VerlinProof::prove_batch(eng, w, st, rng_from(req))) arr.push(Json::string(p.to_json()));
    out.set("proofs", arr);
  } else if (op == "verlin.verify") {
    out.set("results", outcomes(items.size(), [&](size_t i) {
      VerlinProof::from_json(items[i].at("proof").as_str()).verify(eng, {ek, dec(items[i], "c"), dec(items[i], "c_prime"), dec(items[i], "phi_x")});
    }));
  } else if (op == "zero.verify_batch" || op == "ciphertext.verify_batch" || op == "mul.verify_batch" || op == "verlin.verify_batch") {
    // one call over the whole list; an item may carry its own "n" (mixed-key batches).  results: 1 / 0 / -1 (panic)
    auto key_of = [&](const Json& it) { return it.find("n") ? EncryptionKey(dec(it, "n")) : ek; };
    std::vector<int> res;
    if (op == "zero.verify_batch") {
      std::vector<ZeroProof> pr; std::vector<ZeroStatement> st;
      for (auto& it : items) { pr.push_back(ZeroProof::from_json(it.at("proof").as_str())); st.push_back({key_of(it), dec(it, "c")}); }
      std::vector<const ZeroProof*> ps;
      for (auto& p : pr) ps.push_back(&p);
      res = ZeroProof::verify_batch(eng, ps, st);
    } else if (op == "ciphertext.verify_batch") {
      std::vector<CiphertextProof> pr; std::vector<CiphertextStatement> st;
      for (auto& it : items) { pr.push_back(CiphertextProof::from_json(it.at("proof").as_str())); st.push_back({key_of(it), dec(it, "c")}); }
      std::vector<const CiphertextProof*> ps;
      for (auto& p : pr) ps.push_back(&p);
      res = CiphertextProof::verify_batch(eng, ps, st);
    } else if (op == "mul.verify_batch") {
      std::vector<MulProof> pr; std::vector<MulStatement> st;
      for (auto& it : items) {
        pr.push_back(MulProof::from_json(it.at("proof").as_str()));
        st.push_back({key_of(it), dec(it, "e_a"), dec(it, "e_b"), dec(it, "e_c")});
      }
      std::vector<const MulProof*> ps;
      for (auto& p : pr) ps.push_back(&p);
      res = MulProof::verify_batch(eng, ps, st);
    } else {
      std::vector<VerlinProof> pr; std::vector<VerlinStatement> st;
      for (auto& it : items) {
        pr.push_back(VerlinProof::from_json(it.at("proof").as_str()));
        st.push_back({key_of(it), dec(it, "c"), dec(it, "c_prime"), dec(it, "phi_x")});
      }
      std::vector<const VerlinProof*> ps;
      for (auto& p : pr) ps.push_back(&p);
      res = VerlinProof::verify_batch(eng, ps, st);
    }
    Json arr = Json::array();
    for (int v : res) arr.push(Json::number(v));
    out.set("results", arr);
  } else {
    throw std::runtime_error("unknown op " + op);
  }
  return out;
}

char* dup(const std::string& s) {
  char* p = (char*)malloc(s.size() + 1);
  memcpy(p, s.c_str(), s.size() + 1);
  return p;
}

}  // namespace

extern "C" {

char* zkh_call(const char* op, const char* request_json) {
  try {
    return dup(dispatch(op, Json::parse(request_json)).dump());
  } catch (const ReferencePanic& e) {
    return dup(Json::object().set("ok", Json::boolean(false)).set("kind", Json::string("panic")).set("error", Json::string(e.what())).dump());
  } catch (const std::exception& e) {
    return dup(Json::object().set("ok", Json::boolean(false)).set("kind", Json::string("error")).set("error", Json::string(e.what())).dump());
  }
}

void zkh_free(char* p) { free(p); }

// Synthetic NiCorrectKeyProof workload with a DISTINCT modulus per proof (BASELINE configs[2]: 4096 proofs at 3072 bits),
// built on the device: `count` key pairs by Paillier::keypairs_batch (Miller-Rabin waves on K2), one honest proof per key
// by NiCorrectKeyProof::proof_batch; every `bad_every`-th proof gets sigma_0 + 1 (rejected).  Binary rows out, no JSON:
//   n_out [count][bits/32], sigma_out [count][11][bits/32], pq_out (optional) [count][2][bits/64].
// The byte stream behind every sample is SHA-256(seed || counter): the same seed gives the same keys.  Returns 0, or -1
// with the message in err (at most err_len bytes).
int zkh_correct_key_workload(int device, int bits, int count, const uint8_t* seed, int seed_len, const uint8_t* salt, int salt_len,
                             int bad_every, uint32_t* n_out, uint32_t* sigma_out, uint32_t* pq_out, char* err, int err_len) {
  try {
    Engine& eng = engine(device);
    std::vector<uint8_t> key(seed, seed + seed_len);
    auto ctr = std::make_shared<uint64_t>(0);
    auto pool = std::make_shared<std::vector<uint8_t>>();
    ByteSource rng = [key, ctr, pool](uint8_t* p, size_t n) {
      while (pool->size() < n) {
        std::vector<uint8_t> msg(key);
        for (int i = 0; i < 8; ++i) msg.push_back((uint8_t)(*ctr >> (8 * i)));
        ++*ctr;
        const std::vector<uint8_t> h = zkhost::sha256(msg.data(), msg.size());
        pool->insert(pool->end(), h.begin(), h.end());
      }
      memcpy(p, pool->data(), n);
      pool->erase(pool->begin(), pool->begin() + n);
    };
    const std::vector<DecryptionKey> dks = Paillier::keypairs_batch(eng, (size_t)bits, (size_t)count, rng);
    const std::vector<NiCorrectKeyProof> proofs = NiCorrectKeyProof::proof_batch(eng, dks, salt, (size_t)salt_len);
    const size_t nl = (size_t)bits / 32, hl = (size_t)bits / 64;
    for (size_t b = 0; b < (size_t)count; ++b) {
      const BigInt n = dks[b].p * dks[b].q;
      n.to_limbs(n_out + b * nl, nl);
      for (size_t i = 0; i < M2; ++i) {
        BigInt sgm = proofs[b].sigma_vec[i];
        if (i == 0 && bad_every > 0 && (int)(b % (size_t)bad_every) == bad_every - 1) sgm = (sgm + BigInt(1)) % n;
        sgm.to_limbs(sigma_out + (b * M2 + i) * nl, nl);
      }
      if (pq_out) {
        dks[b].p.to_limbs(pq_out + (2 * b) * hl, hl);
        dks[b].q.to_limbs(pq_out + (2 * b + 1) * hl, hl);
      }
    }
    return 0;
  } catch (const std::exception& e) {
    if (err && err_len > 0) {
      strncpy(err, e.what(), (size_t)err_len - 1);
      err[err_len - 1] = 0;
    }
    return -1;
  }
}

}  // extern "C"
