// C-ABI layer, part 3: NiCorrectKeyProof::verify for a batch of proofs, each with its own modulus N
// (reference src/zkproofs/correct_key_ni.rs:73-100).
#include "ctx.h"

using namespace zkp;

extern "C" {

int zkp_ck_verify_stage(zkp_ctx* c, int batch, int nl, const uint32_t* n, const uint32_t* sigma, const uint8_t* salt,
                        int salt_len) {
  if (!c) return ZKP_E_ARG;
  if (batch <= 0 || nl <= 0 || nl % 4 || !n || !sigma || salt_len < 0 || (salt_len > 0 && !salt))
    return fail(c, ZKP_E_ARG, "bad correct-key batch shape");
  const int S = pick_width(nl);
  if (S < 0) return fail(c, ZKP_E_ARG, "modulus wider than 8192 bits");
  if ((long long)batch * kCkM2 > 0x3fffffffll) return fail(c, ZKP_E_ARG, "batch too large");
  ZKP_CU(c, cudaSetDevice(c->device));
  CkState& s = c->ck;
  s.staged = s.done = false;
  const size_t bm = (size_t)batch * kCkM2;
  const int ml = nl + 8;
  ZKP_CU(c, s.n.ensure((size_t)batch * nl * 4));
  ZKP_CU(c, s.sigma.ensure(bm * nl * 4));
  ZKP_CU(c, s.salt.ensure((size_t)salt_len + 16));
  ZKP_CU(c, s.r2.ensure((size_t)batch * S * 4));
  ZKP_CU(c, s.n0inv.ensure((size_t)batch * 4));
  ZKP_CU(c, s.mask.ensure(bm * ml * 4 + (size_t)S * 4));
  ZKP_CU(c, s.rho.ensure(bm * nl * 4));
  ZKP_CU(c, s.derived.ensure(bm * nl * 4));
  ZKP_CU(c, s.accept.ensure((size_t)batch));
  ZKP_CU(c, ensure_table(c, S, kTableVar));
  cudaStream_t st = c->stream;
  if (s.nprimes == 0) {  // primes below alpha = 6370
    std::vector<uint16_t> pr;
    for (int v = 2; v < kCkAlpha; ++v) {
      bool is = true;
      for (int d = 2; d * d <= v; ++d)
        if (v % d == 0) { is = false; break; }
      if (is) pr.push_back((uint16_t)v);
    }
    ZKP_CU(c, s.primes.ensure(pr.size() * 2));
    ZKP_CU(c, cudaMemcpyAsync(s.primes.p, pr.data(), pr.size() * 2, cudaMemcpyHostToDevice, st));
    ZKP_CU(c, cudaStreamSynchronize(st));
    s.nprimes = (int)pr.size();
  }
  ZKP_CU(c, cudaMemcpyAsync(s.n.p, n, (size_t)batch * nl * 4, cudaMemcpyHostToDevice, st));
  ZKP_CU(c, cudaMemcpyAsync(s.sigma.p, sigma, bm * nl * 4, cudaMemcpyHostToDevice, st));
  if (salt_len) ZKP_CU(c, cudaMemcpyAsync(s.salt.p, salt, (size_t)salt_len, cudaMemcpyHostToDevice, st));
  ZKP_CU(c, cudaStreamSynchronize(st));
  s.batch = batch;
  s.nl = nl;
  s.S = S;
  s.salt_len = salt_len;
  s.staged = true;
  return ZKP_OK;
}

int zkp_ck_verify_run(zkp_ctx* c) {
  if (!c) return ZKP_E_ARG;
  CkState& s = c->ck;
  if (!s.staged) return fail(c, ZKP_E_STATE, "nothing staged for correct-key verify");
  ZKP_CU(c, cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const int batch = s.batch, nl = s.nl, S = s.S, ml = nl + 8;
  const int jobs = batch * kCkM2;
  const uint32_t* n = s.n.as<uint32_t>();
  {  // per-proof Montgomery constants
    ProfScope ps(c, KID_OTHER, batch);
    ZKP_CU(c, launch_mont_setup(n, nl, S, batch, s.r2.as<uint32_t>(), s.n0inv.as<uint32_t>(), st));
  }
  {  // rho_i = mask_generation(|N|, H(N, H(salt), i)) % N   (correct_key_ni.rs:75-86)
    ProfScope ps(c, KID_SHA, jobs);
    ZKP_CU(c, launch_ck_rho(n, nl, s.salt.as<uint8_t>(), s.salt_len, batch, s.mask.as<uint32_t>(), ml, st));
  }
  {
    ProfScope ps(c, KID_MODMUL, jobs);
    ZKP_CU(c, launch_ck_reduce(s.mask.as<uint32_t>(), ml, n, nl, s.r2.as<uint32_t>(), s.n0inv.as<uint32_t>(), S, batch,
                               s.rho.as<uint32_t>(), st));
  }
  {  // sigma_i^N mod N   (:90-93)
    ProfScope ps(c, KID_MODEXP_VAR, jobs);
    ZKP_CU(c, launch_modexp_var(s.sigma.as<uint32_t>(), n, nl, s.r2.as<uint32_t>(), s.n0inv.as<uint32_t>(), n, nl, 32 * nl,
                                kCkM2, kCkM2, s.derived.as<uint32_t>(), jobs, S, c->table.as<uint32_t>(), c->num_sms, st));
  }
  {  // rho == derived && gcd(P, N) == 1   (:87-88, 95)
    ProfScope ps(c, KID_OTHER, batch);
    ZKP_CU(c, launch_ck_check(n, nl, s.rho.as<uint32_t>(), s.derived.as<uint32_t>(), s.primes.as<uint16_t>(), s.nprimes,
                              batch, s.accept.as<uint8_t>(), st));
  }
  s.done = true;
  return ZKP_OK;
}

int zkp_ck_verify_fetch(zkp_ctx* c, uint8_t* accept, uint32_t* rho) {
  if (!c) return ZKP_E_ARG;
  CkState& s = c->ck;
  if (!s.done) return fail(c, ZKP_E_STATE, "correct-key verify has not run");
  ZKP_CU(c, cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  if (accept) ZKP_CU(c, cudaMemcpyAsync(accept, s.accept.p, (size_t)s.batch, cudaMemcpyDeviceToHost, st));
  if (rho) ZKP_CU(c, cudaMemcpyAsync(rho, s.rho.p, (size_t)s.batch * kCkM2 * s.nl * 4, cudaMemcpyDeviceToHost, st));
  ZKP_CU(c, cudaStreamSynchronize(st));
  return ZKP_OK;
}

int zkp_correct_key_ni_rho(zkp_ctx* c, int batch, int nl, const uint32_t* n, const uint8_t* salt, int salt_len, uint32_t* rho) {
  if (!c) return ZKP_E_ARG;
  if (batch <= 0 || nl <= 0 || nl % 4 || !n || !rho || salt_len < 0 || (salt_len > 0 && !salt))
    return fail(c, ZKP_E_ARG, "bad correct-key batch shape");
  const int S = pick_width(nl);
  if (S < 0) return fail(c, ZKP_E_ARG, "modulus wider than 8192 bits");
  for (int b = 0; b < batch; ++b)
    if (!(n[(size_t)b * nl] & 1u)) return fail(c, ZKP_E_ARG, "every modulus must be odd");
  ZKP_CU(c, cudaSetDevice(c->device));
  CkState& s = c->ck;
  s.staged = s.done = false;
  const size_t bm = (size_t)batch * kCkM2;
  const int ml = nl + 8;
  ZKP_CU(c, s.n.ensure((size_t)batch * nl * 4));
  ZKP_CU(c, s.salt.ensure((size_t)salt_len + 16));
  ZKP_CU(c, s.r2.ensure((size_t)batch * S * 4));
  ZKP_CU(c, s.n0inv.ensure((size_t)batch * 4));
  ZKP_CU(c, s.mask.ensure(bm * ml * 4 + (size_t)S * 4));
  ZKP_CU(c, s.rho.ensure(bm * nl * 4));
  cudaStream_t st = c->stream;
  ZKP_CU(c, cudaMemcpyAsync(s.n.p, n, (size_t)batch * nl * 4, cudaMemcpyHostToDevice, st));
  if (salt_len) ZKP_CU(c, cudaMemcpyAsync(s.salt.p, salt, (size_t)salt_len, cudaMemcpyHostToDevice, st));
  ZKP_CU(c, launch_mont_setup(s.n.as<uint32_t>(), nl, S, batch, s.r2.as<uint32_t>(), s.n0inv.as<uint32_t>(), st));
  ZKP_CU(c, launch_ck_rho(s.n.as<uint32_t>(), nl, s.salt.as<uint8_t>(), salt_len, batch, s.mask.as<uint32_t>(), ml, st));
  ZKP_CU(c, launch_ck_reduce(s.mask.as<uint32_t>(), ml, s.n.as<uint32_t>(), nl, s.r2.as<uint32_t>(), s.n0inv.as<uint32_t>(), S, batch,
                             s.rho.as<uint32_t>(), st));
  ZKP_CU(c, cudaMemcpyAsync(rho, s.rho.p, bm * nl * 4, cudaMemcpyDeviceToHost, st));
  ZKP_CU(c, cudaStreamSynchronize(st));
  return ZKP_OK;
}

int zkp_correct_key_ni_verify(zkp_ctx* c, int batch, int nl, const uint32_t* n, const uint32_t* sigma, const uint8_t* salt,
                              int salt_len, uint8_t* accept, uint32_t* rho) {
  int rc = zkp_ck_verify_stage(c, batch, nl, n, sigma, salt, salt_len);
  if (rc) return rc;
  rc = zkp_ck_verify_run(c);
  if (rc) return rc;
  return zkp_ck_verify_fetch(c, accept, rho);
}

}  // extern "C"
