// Shared plumbing of the sigma-protocol entry points (api_sigma.cu, api_more.cu): a bump allocator over one device
// buffer and the handful of batched primitives every proof is sequenced from (Enc, mod_pow, mod_mul, Fiat-Shamir hash).
#pragma once
#include <algorithm>
#include <initializer_list>

#include "ctx.h"

namespace zkp {
// Bump allocator over one device buffer for the temporaries of a call.
struct Arena {
  zkp_ctx* c;
  size_t off = 0, cap = 0;
  uint8_t* base = nullptr;
  std::vector<size_t> wants;
  explicit Arena(zkp_ctx* ctx) : c(ctx) {}
  cudaError_t reserve(size_t bytes) {
    cudaError_t e = c->in0.ensure(bytes);
    base = c->in0.as<uint8_t>();
    cap = bytes;
    off = 0;
    return e;
  }
  template <class U>
  U* get(size_t count) {
    size_t bytes = (count * sizeof(U) + 255) & ~size_t(255);
    if (off + bytes > cap) return nullptr;
    U* p = reinterpret_cast<U*>(base + off);
    off += bytes;
    return p;
  }
};

struct Sig {
  zkp_ctx* c;
  cudaStream_t st;
  int batch, nl, nnl;
  Arena ar;
  bool bad = false;
  Sig(zkp_ctx* ctx, int b) : c(ctx), st(ctx->stream), batch(b), nl(ctx->n.limbs), nnl(ctx->nn.limbs), ar(ctx) {}

  // `jobs` < 0 means one row per proof (batch rows); the ring proof passes batch * M
  int nrows(int jobs) const { return jobs < 0 ? batch : jobs; }
  uint32_t* rows(int limbs, int jobs = -1) {
    uint32_t* p = ar.get<uint32_t>((size_t)nrows(jobs) * limbs);
    if (!p) bad = true;
    return p;
  }
  uint8_t* bytes(size_t count, bool zero = false) {
    uint8_t* p = ar.get<uint8_t>(count);
    if (!p) { bad = true; return p; }
    if (zero && cudaMemsetAsync(p, 0, count, st) != cudaSuccess) bad = true;
    return p;
  }
  uint32_t* up(const uint32_t* host, int limbs, int jobs = -1) {  // upload [rows][limbs]
    uint32_t* p = rows(limbs, jobs);
    if (p && cudaMemcpyAsync(p, host, (size_t)nrows(jobs) * limbs * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) bad = true;
    return p;
  }
  void down(uint32_t* host, const uint32_t* dev, int limbs, int jobs = -1) {
    if (host && !bad && cudaMemcpyAsync(host, dev, (size_t)nrows(jobs) * limbs * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess) bad = true;
  }
  void down8(uint8_t* host, const uint8_t* dev, size_t count) {
    if (host && !bad && cudaMemcpyAsync(host, dev, count, cudaMemcpyDeviceToHost, st) != cudaSuccess) bad = true;
  }
  void ck(cudaError_t e) {
    if (e != cudaSuccess) {
      bad = true;
      fail_cuda(c, e, "sigma-protocol launch");
    }
  }
  // Paillier::encrypt_with_chosen_randomness(ek, m, r); m == nullptr means the plaintext 0 (c = r^n mod nn)
  uint32_t* enc(const uint32_t* m, int m_limbs, const uint32_t* r, int r_limbs, int jobs = -1) {
    uint32_t* out = rows(nnl, jobs);
    if (bad) return out;
    ProfScope ps(c, KID_MODEXP_SHARED, nrows(jobs));
    ck(launch_enc(c, r, r_limbs, m, m_limbs, out, nrows(jobs)));
    return out;
  }
  // BigInt::mod_pow(base, exp, nn) / Paillier::mul, per-proof exponent
  uint32_t* powm(const uint32_t* base, int base_limbs, const uint32_t* exp, int exp_limbs, int jobs = -1) {
    uint32_t* out = rows(nnl, jobs);
    if (bad) return out;
    ProfScope ps(c, KID_MODEXP_VAR, nrows(jobs));
    ck(launch_pow_nn(c, base, base_limbs, exp, exp_limbs, 32 * exp_limbs, 1, out, nrows(jobs)));
    return out;
  }
  // BigInt::mod_mul(a, b, nn) / Paillier::add
  // row j of a is multiplied by row j / b_per of b
  uint32_t* mulm(const uint32_t* a, int a_limbs, const uint32_t* b, int b_limbs, int jobs = -1, int b_per = 1) {
    uint32_t* out = rows(nnl, jobs);
    if (bad) return out;
    ProfScope ps(c, KID_MODMUL, nrows(jobs));
    ck(launch_modmul_shared(c->nn.view(), 0, a, a_limbs, b, b_limbs, b_per, out, nnl, nrows(jobs), st));
    return out;
  }
  // the same modulo n
  uint32_t* mulm_n(const uint32_t* a, int a_limbs, const uint32_t* b, int b_limbs, int jobs = -1, int b_per = 1) {
    uint32_t* out = rows(nl, jobs);
    if (bad) return out;
    ProfScope ps(c, KID_MODMUL, nrows(jobs));
    ck(launch_modmul_shared(c->n.view(), 0, a, a_limbs, b, b_limbs, b_per, out, nl, nrows(jobs), st));
    return out;
  }
  // e = compute_digest(n, items...) as 8 limbs
  uint32_t* challenge(std::initializer_list<const uint32_t*> items) {
    uint8_t* dig = ar.get<uint8_t>((size_t)batch * 32);
    uint32_t* e = rows(8);
    if (!dig || bad) { bad = true; return e; }
    ShaSegs s;
    s.nseg = 0;
    s.seg[s.nseg++] = {c->n.mod.as<uint32_t>(), 0, 1, nl};
    for (const uint32_t* p : items) s.seg[s.nseg++] = {p, (long long)nnl, 1, nnl};
    {
      ProfScope ps(c, KID_SHA, batch);
      ck(launch_sha256_transcript(s, batch, dig, st));
    }
    ProfScope ps(c, KID_OTHER, batch);
    ck(launch_digest_to_limbs(dig, batch, e, st));
    return e;
  }
  // independent modexps of one proof run side by side: fork(k) sends what follows to auxiliary stream k (after everything
  // queued on the main stream so far), on_main() goes back without waiting, join() makes the main stream wait for all
  void fork(int k) {
    if (bad) return;
    ck(fork_stream(c, k));
    st = c->stream;
  }
  void on_main() {
    main_stream(c);
    st = c->stream;
  }
  void join() {
    cudaError_t e = join_streams(c);
    st = c->stream;
    if (e != cudaSuccess) ck(e);
  }
  // device span of the call for zkp_profile_get(5): from here (inputs uploaded) ...
  ProfScope* span = nullptr;
  void compute_begin() {
    if (!span) span = new ProfScope(c, KID_CALL, batch);
  }
  // ... to here (every kernel queued, streams joined; the downloads follow)
  void compute_end() {
    join();
    delete span;
    span = nullptr;
  }
  ~Sig() { delete span; }
  int finish(const char* what) {
    join();
    compute_end();
    if (bad) {
      cudaStreamSynchronize(st);
      if (c->err.empty()) c->err = what;
      return ZKP_E_CUDA;
    }
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail_cuda(c, e, what);
    return ZKP_OK;
  }
};

// The modular exponentiations of one proof phase, queued and then launched as ONE K2h kernel over a heterogeneous job
// list (longest exponents first).  With a key the two-digit kernels do not take, run() issues them one after the other
// through launch_enc / launch_pow_nn instead - same results.
struct PowTerm {
  const uint32_t* base;
  int base_limbs;
  const uint32_t* exp;  // nullptr: the shared exponent n (a Paillier encryption's r^n)
  int exp_limbs;
};
struct PowBatch {
  Sig& s;
  struct Item {
    PowTerm term[kMaxPowBases];
    int nterm;
    const uint32_t* plain;
    uint32_t* out;
    int plain_limbs, jobs;
  };
  std::vector<Item> items;
  explicit PowBatch(Sig& sig) : s(sig) {}
  // Paillier::encrypt_with_chosen_randomness(ek, m, r); m == nullptr: the plaintext 0
  uint32_t* enc(const uint32_t* m, int m_limbs, const uint32_t* r, int r_limbs, int jobs = -1) {
    return product({PowTerm{r, r_limbs, nullptr, 0}}, m, m_limbs, jobs);
  }
  // BigInt::mod_pow(base, exp, nn) / Paillier::mul with a per-proof exponent
  uint32_t* powm(const uint32_t* base, int base_limbs, const uint32_t* exp, int exp_limbs, int jobs = -1) {
    return product({PowTerm{base, base_limbs, exp, exp_limbs}}, nullptr, 0, jobs);
  }
  // (1 + m n) * PROD base_k^exp_k mod nn as ONE job (simultaneous exponentiation): where the reference multiplies powers
  // together and only the product is used - gen_phi (verlin_proof.rs:138-165)
  uint32_t* product(std::initializer_list<PowTerm> terms, const uint32_t* m, int m_limbs, int jobs = -1) {
    uint32_t* out = s.rows(s.nnl, jobs);
    Item it{};
    it.nterm = 0;
    for (const PowTerm& t : terms)
      if (it.nterm < kMaxPowBases) it.term[it.nterm++] = t;
    if ((size_t)it.nterm != terms.size()) s.bad = true;
    it.plain = m;
    it.plain_limbs = m ? m_limbs : 0;
    it.out = out;
    it.jobs = s.nrows(jobs);
    items.push_back(it);
    return out;
  }
  void run() {
    if (s.bad || items.empty()) return;
    zkp_ctx* c = s.c;
    if (!jobs_supported(c) || (int)items.size() > kMaxPowSegs) {
      for (const Item& it : items) run_separately(it);
      items.clear();
      return;
    }
    int n_bits = 32 * c->n.S;  // bit length of the shared exponent n
    while (n_bits > 1 && !((c->n.h_mod[(n_bits - 1) >> 5] >> ((n_bits - 1) & 31)) & 1u)) --n_bits;
    auto bits = [&](const Item& it) {
      int b = 0;
      for (int k = 0; k < it.nterm; ++k) b = std::max(b, it.term[k].exp ? 32 * it.term[k].exp_limbs : n_bits);
      return b;
    };
    std::stable_sort(items.begin(), items.end(), [&](const Item& a, const Item& b) { return bits(a) > bits(b); });
    PowJobs pj;
    for (const Item& it : items) {
      PowSeg& g = pj.seg[pj.nseg++];
      g.nbase = it.nterm;
      for (int k = 0; k < kMaxPowBases; ++k) {
        const PowTerm& t = it.term[k < it.nterm ? k : 0];
        g.base[k] = t.base;
        g.base_limbs[k] = t.base_limbs;
        g.exp[k] = t.exp ? t.exp : c->n.mod.as<uint32_t>();
        g.exp_limbs[k] = t.exp ? t.exp_limbs : c->n.S;
        g.exp_stride[k] = t.exp ? t.exp_limbs : 0;
      }
      g.exp_bits = bits(it);
      g.plain = it.plain;
      g.plain_limbs = it.plain_limbs;
      g.out = it.out;
      g.jobs = it.jobs;
      g.first = pj.total;
      pj.total += it.jobs;
    }
    items.clear();
    ProfScope ps(c, KID_MODEXP_VAR, pj.total);
    s.ck(launch_pow_jobs(c, pj));
  }

 private:
  // the same values by the single-purpose kernels: each power by launch_enc / launch_pow_nn, then the products
  void run_separately(const Item& it) {
    zkp_ctx* c = s.c;
    uint32_t* acc = nullptr;
    for (int k = 0; k < it.nterm; ++k) {
      const PowTerm& t = it.term[k];
      const bool last = k == it.nterm - 1;
      uint32_t* dst = (it.nterm == 1) ? it.out : s.rows(s.nnl, it.jobs);
      if (s.bad) return;
      {
        ProfScope ps(c, t.exp ? KID_MODEXP_VAR : KID_MODEXP_SHARED, it.jobs);
        // the plaintext factor rides on an encryption term when there is one, else on a separate Enc(m, 1)-free product below
        if (t.exp) s.ck(launch_pow_nn(c, t.base, t.base_limbs, t.exp, t.exp_limbs, 32 * t.exp_limbs, 1, dst, it.jobs));
        else s.ck(launch_enc(c, t.base, t.base_limbs, it.plain, it.plain_limbs, dst, it.jobs));
      }
      if (acc) {
        uint32_t* prod = last ? it.out : s.rows(s.nnl, it.jobs);
        if (s.bad) return;
        ProfScope ps(c, KID_MODMUL, it.jobs);
        s.ck(launch_modmul_shared(c->nn.view(), 0, acc, s.nnl, dst, s.nnl, 1, prod, s.nnl, it.jobs, s.st));
        acc = prod;
      } else {
        acc = dst;
      }
    }
    // the plaintext factor rides on THE encryption term: a plaintext without one (a separate (1 + m n) factor) or with two
    // (the factor applied twice) is not something a caller builds - refuse rather than compute something else than K2h would
    int enc_terms = 0;
    for (int k = 0; k < it.nterm; ++k) enc_terms += !it.term[k].exp;
    if (it.plain ? enc_terms != 1 : false) s.bad = true;
  }
};

inline int sigma_begin(zkp_ctx* c, int batch, int z_limbs, size_t rows_nnl, size_t rows_other_bytes, Sig& s, int pow_bases = 1) {
  if (!c->paillier) return fail(c, ZKP_E_STATE, "zkp_set_key not called");
  if (batch <= 0) return fail(c, ZKP_E_ARG, "batch must be positive");
  if (z_limbs && (z_limbs % 4 || z_limbs < c->n.limbs + 12 || z_limbs > c->nn.S))
    return fail(c, ZKP_E_ARG, "z_limbs must be a multiple of 4 in [n_limbs + 12, nn_limbs]");
  cudaError_t e = cudaSetDevice(c->device);
  if (e != cudaSuccess) return fail_cuda(c, e, "cudaSetDevice");
  c->err.clear();
  // generous: every helper result is one [batch][nn_limbs] row array
  size_t bytes = (size_t)batch * ((rows_nnl + 12) * s.nnl * 4 + rows_other_bytes + 256) + 64 * 256;
  e = s.ar.reserve(bytes);
  if (e != cudaSuccess) return fail_cuda(c, e, "arena");
  e = ensure_table(c, c->nn.S, kTableVar, 4 * batch, pow_bases);  // a K2h launch holds at most 3 modexps per proof (PowBatch)
  if (e != cudaSuccess) return fail_cuda(c, e, "table");
  return ZKP_OK;
}

}  // namespace zkp
