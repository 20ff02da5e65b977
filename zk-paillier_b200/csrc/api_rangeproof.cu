// C-ABI layer, part 2: RangeProofNi::{prove, verify} for a batch of proofs under one key
// (reference src/zkproofs/range_proof_ni.rs:47-107 over range_proof.rs:128-355) and the
// stand-alone transcript hash.  Host code stages buffers and sequences kernels; all
// arithmetic, hashing and predicates run on the device.
#include "ctx.h"

using namespace zkp;

namespace {

int rp_check_shape(zkp_ctx* c, int batch, int ef, int wl) {
  if (!c->paillier) return fail(c, ZKP_E_STATE, "zkp_set_key not called");
  if (batch <= 0 || ef <= 0) return fail(c, ZKP_E_ARG, "batch and error_factor must be positive");
  if (wl <= 0 || wl % 4 || wl > 64 || wl > c->nn.S) return fail(c, ZKP_E_ARG, "w_limbs must be a multiple of 4, at most 64");
  if ((long long)batch * ef * 2 > 0x3fffffffll) return fail(c, ZKP_E_ARG, "batch * error_factor too large");
  return ZKP_OK;
}

ShaSegs rp_transcript(zkp_ctx* c, const uint32_t* cpairs, int batch, int ef) {
  const int nnl = c->nn.limbs;
  ShaSegs s;
  s.nseg = 3;
  s.seg[0] = {c->n.mod.as<uint32_t>(), 0, 1, c->n.limbs};                       // ek.n
  s.seg[1] = {cpairs, (long long)ef * nnl, ef, nnl};                             // c1[0..ef)
  s.seg[2] = {cpairs + (size_t)batch * ef * nnl, (long long)ef * nnl, ef, nnl};  // c2[0..ef)
  return s;
}

}  // namespace

extern "C" {

int zkp_sha256_transcript(zkp_ctx* c, const uint32_t* items, int limbs, int count, int batch, uint8_t* digest) {
  if (!c) return ZKP_E_ARG;
  if (!items || !digest || limbs <= 0 || count <= 0 || batch < 0) return fail(c, ZKP_E_ARG, "bad transcript shape");
  if (batch == 0) return ZKP_OK;
  ZKP_CU(c, cudaSetDevice(c->device));
  const size_t bytes = (size_t)batch * count * limbs * 4;
  ZKP_CU(c, c->in0.ensure(bytes));
  ZKP_CU(c, c->out0.ensure((size_t)batch * 32));
  ZKP_CU(c, cudaMemcpyAsync(c->in0.p, items, bytes, cudaMemcpyHostToDevice, c->stream));
  ShaSegs s;
  s.nseg = 1;
  s.seg[0] = {c->in0.as<uint32_t>(), (long long)count * limbs, count, limbs};
  {
    ProfScope ps(c, KID_SHA, batch);
    ZKP_CU(c, launch_sha256_transcript(s, batch, c->out0.as<uint8_t>(), c->stream));
  }
  ZKP_CU(c, cudaMemcpyAsync(digest, c->out0.p, (size_t)batch * 32, cudaMemcpyDeviceToHost, c->stream));
  ZKP_CU(c, cudaStreamSynchronize(c->stream));
  return ZKP_OK;
}

// ------------------------------------------------------------------ prove
int zkp_rp_prove_stage(zkp_ctx* c, int batch, int ef, int wl, const uint32_t* range, const uint32_t* x, const uint32_t* r,
                       const uint32_t* w1, const uint8_t* swap, const uint32_t* r1, const uint32_t* r2) {
  if (!c) return ZKP_E_ARG;
  int rc = rp_check_shape(c, batch, ef, wl);
  if (rc) return rc;
  if (!range || !x || !r || !w1 || !swap || !r1 || !r2) return fail(c, ZKP_E_ARG, "null input");
  ZKP_CU(c, cudaSetDevice(c->device));
  RpState& s = c->rp;
  s.prove_staged = s.prove_done = s.pairs_done = false;
  const int nl = c->n.limbs, nnl = c->nn.limbs;
  const size_t be = (size_t)batch * ef;
  ZKP_CU(c, s.range.ensure((size_t)batch * wl * 4));
  ZKP_CU(c, s.x.ensure((size_t)batch * wl * 4));
  ZKP_CU(c, s.r.ensure((size_t)batch * nl * 4));
  ZKP_CU(c, s.w1in.ensure(be * wl * 4));
  ZKP_CU(c, s.w.ensure(2 * be * wl * 4));
  ZKP_CU(c, s.swap.ensure(be));
  ZKP_CU(c, s.rr.ensure(2 * be * nl * 4));
  ZKP_CU(c, s.c.ensure(2 * be * nnl * 4));
  ZKP_CU(c, s.digest.ensure((size_t)batch * 32));
  ZKP_CU(c, s.kind.ensure(be));
  ZKP_CU(c, s.resp_w.ensure(2 * be * wl * 4));
  ZKP_CU(c, s.resp_r.ensure(2 * be * nl * 4));
  ZKP_CU(c, s.rmul.ensure(2 * be * nl * 4));
  ZKP_CU(c, s.fault.ensure((size_t)batch));
  ZKP_CU(c, ensure_table(c, c->nn.S, kTableShared));
  cudaStream_t st = c->stream;
  ZKP_CU(c, cudaMemcpyAsync(s.range.p, range, (size_t)batch * wl * 4, cudaMemcpyHostToDevice, st));
  ZKP_CU(c, cudaMemcpyAsync(s.x.p, x, (size_t)batch * wl * 4, cudaMemcpyHostToDevice, st));
  ZKP_CU(c, cudaMemcpyAsync(s.r.p, r, (size_t)batch * nl * 4, cudaMemcpyHostToDevice, st));
  ZKP_CU(c, cudaMemcpyAsync(s.w1in.p, w1, be * wl * 4, cudaMemcpyHostToDevice, st));
  ZKP_CU(c, cudaMemcpyAsync(s.swap.p, swap, be, cudaMemcpyHostToDevice, st));
  ZKP_CU(c, cudaMemcpyAsync(s.rr.p, r1, be * nl * 4, cudaMemcpyHostToDevice, st));
  ZKP_CU(c, cudaMemcpyAsync(s.rr.as<uint32_t>() + be * nl, r2, be * nl * 4, cudaMemcpyHostToDevice, st));
  ZKP_CU(c, cudaStreamSynchronize(st));
  s.batch = batch;
  s.ef = ef;
  s.wl = wl;
  s.prove_staged = true;
  return ZKP_OK;
}

// Phase 1 of prove: the encrypted pairs and their transcript hash (RangeProof::generate_encrypted_pairs,
// range_proof.rs:128-193, + range_proof_ni.rs:58-61).  Does not depend on the challenge.
int zkp_rp_prove_run_pairs(zkp_ctx* c) {
  if (!c) return ZKP_E_ARG;
  RpState& s = c->rp;
  if (!s.prove_staged || !c->paillier) return fail(c, ZKP_E_STATE, "nothing staged for prove");
  ZKP_CU(c, cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const int batch = s.batch, ef = s.ef, wl = s.wl, nl = c->n.limbs;
  const int be = batch * ef;
  s.pairs_done = s.prove_done = false;
  ZKP_CU(c, cudaMemsetAsync(s.fault.p, 0, (size_t)batch, st));
  {  // w2 = w1 - third, coin swap (range_proof.rs:141-149)
    ProfScope ps(c, KID_OTHER, be);
    ZKP_CU(c, launch_rp_prep(s.range.as<uint32_t>(), s.w1in.as<uint32_t>(), s.w.as<uint32_t>(), s.swap.as<uint8_t>(), batch, ef, wl,
                             s.fault.as<uint8_t>(), st));
  }
  {  // c1 | c2 = Enc(w1' | w2', r1 | r2)  (:161-187)
    ProfScope ps(c, KID_MODEXP_SHARED, 2.0 * be);
    ZKP_CU(c, launch_enc(c, s.rr.as<uint32_t>(), nl, s.w.as<uint32_t>(), wl, s.c.as<uint32_t>(), 2 * be));
  }
  {  // e = H(n, c1.., c2..)  (range_proof_ni.rs:58-61)
    ProfScope ps(c, KID_SHA, batch);
    ZKP_CU(c, launch_sha256_transcript(rp_transcript(c, s.c.as<uint32_t>(), batch, ef), batch, s.digest.as<uint8_t>(), st));
  }
  s.pairs_done = true;
  return ZKP_OK;
}

// Phase 2 of prove: the responses (RangeProof::generate_proof, range_proof.rs:210-252).  challenge == NULL: the
// Fiat-Shamir bits of phase 1's hash (RangeProofNi); otherwise the verifier's raw ChallengeBits bytes,
// [batch][chal_bytes] (the interactive proof).
int zkp_rp_prove_run_responses(zkp_ctx* c, const uint8_t* challenge, int chal_bytes) {
  if (!c) return ZKP_E_ARG;
  RpState& s = c->rp;
  if (!s.pairs_done || !c->paillier) return fail(c, ZKP_E_STATE, "zkp_rp_prove_run_pairs has not run");
  if (challenge && chal_bytes <= 0) return fail(c, ZKP_E_ARG, "chal_bytes must be positive");
  ZKP_CU(c, cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const int batch = s.batch, ef = s.ef, wl = s.wl, nl = c->n.limbs;
  const int be = batch * ef;
  if (challenge) {
    ZKP_CU(c, s.chal.ensure((size_t)batch * chal_bytes));
    ZKP_CU(c, cudaMemcpyAsync(s.chal.p, challenge, (size_t)batch * chal_bytes, cudaMemcpyHostToDevice, st));
  }
  s.chal_bytes = challenge ? chal_bytes : 0;
  {  // secret_r * r_j % n for both j (range_proof.rs:239,245)
    ProfScope ps(c, KID_MODMUL, 2.0 * be);
    SharedKey kn = c->n.view();
    ZKP_CU(c, launch_modmul_shared(kn, 0, s.rr.as<uint32_t>(), nl, s.r.as<uint32_t>(), nl, ef, s.rmul.as<uint32_t>(), nl, be, st));
    ZKP_CU(c, launch_modmul_shared(kn, 0, s.rr.as<uint32_t>() + (size_t)be * nl, nl, s.r.as<uint32_t>(), nl, ef,
                                   s.rmul.as<uint32_t>() + (size_t)be * nl, nl, be, st));
  }
  {  // responses (:210-252)
    ProfScope ps(c, KID_OTHER, be);
    RpProveArgs a;
    a.batch = batch; a.ef = ef; a.wl = wl; a.nl = nl;
    a.range = s.range.as<uint32_t>(); a.x = s.x.as<uint32_t>(); a.w = s.w.as<uint32_t>(); a.rr = s.rr.as<uint32_t>();
    a.rmul = s.rmul.as<uint32_t>(); a.digest = s.digest.as<uint8_t>(); a.kind = s.kind.as<uint8_t>();
    a.chal = challenge ? s.chal.as<uint8_t>() : nullptr; a.chal_bytes = s.chal_bytes;
    a.resp_w = s.resp_w.as<uint32_t>(); a.resp_r = s.resp_r.as<uint32_t>(); a.fault = s.fault.as<uint8_t>();
    ZKP_CU(c, launch_rp_respond(a, st));
  }
  s.prove_done = true;
  return ZKP_OK;
}

int zkp_rp_prove_run(zkp_ctx* c) {
  int rc = zkp_rp_prove_run_pairs(c);
  if (rc) return rc;
  return zkp_rp_prove_run_responses(c, nullptr, 0);
}

int zkp_rp_prove_fetch(zkp_ctx* c, uint32_t* c1, uint32_t* c2, uint8_t* digest, uint8_t* kind, uint32_t* resp_w,
                       uint32_t* resp_r) {
  if (!c) return ZKP_E_ARG;
  RpState& s = c->rp;
  if (!s.pairs_done) return fail(c, ZKP_E_STATE, "prove has not run");
  if (!s.prove_done && (kind || resp_w || resp_r)) return fail(c, ZKP_E_STATE, "responses have not been computed yet");
  ZKP_CU(c, cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const size_t be = (size_t)s.batch * s.ef;
  const int nl = c->n.limbs, nnl = c->nn.limbs;
  if (c1) ZKP_CU(c, cudaMemcpyAsync(c1, s.c.p, be * nnl * 4, cudaMemcpyDeviceToHost, st));
  if (c2) ZKP_CU(c, cudaMemcpyAsync(c2, s.c.as<uint32_t>() + be * nnl, be * nnl * 4, cudaMemcpyDeviceToHost, st));
  if (digest) ZKP_CU(c, cudaMemcpyAsync(digest, s.digest.p, (size_t)s.batch * 32, cudaMemcpyDeviceToHost, st));
  if (kind) ZKP_CU(c, cudaMemcpyAsync(kind, s.kind.p, be, cudaMemcpyDeviceToHost, st));
  if (resp_w) ZKP_CU(c, cudaMemcpyAsync(resp_w, s.resp_w.p, 2 * be * s.wl * 4, cudaMemcpyDeviceToHost, st));
  if (resp_r) ZKP_CU(c, cudaMemcpyAsync(resp_r, s.resp_r.p, 2 * be * nl * 4, cudaMemcpyDeviceToHost, st));
  std::vector<uint8_t> fault((size_t)s.batch);
  ZKP_CU(c, cudaMemcpyAsync(fault.data(), s.fault.p, (size_t)s.batch, cudaMemcpyDeviceToHost, st));
  ZKP_CU(c, cudaStreamSynchronize(st));
  for (uint8_t f : fault)
    if (f) return fail(c, ZKP_E_ARG, "prove input outside the engine's domain (w1 < range/3, masked_x overflows w_limbs, or degenerate digest)");
  return ZKP_OK;
}

int zkp_rangeproof_ni_prove(zkp_ctx* c, int batch, int ef, int wl, const uint32_t* range, const uint32_t* x,
                            const uint32_t* r, const uint32_t* w1, const uint8_t* swap, const uint32_t* r1,
                            const uint32_t* r2, uint32_t* c1, uint32_t* c2, uint8_t* digest, uint8_t* kind,
                            uint32_t* resp_w, uint32_t* resp_r) {
  int rc = zkp_rp_prove_stage(c, batch, ef, wl, range, x, r, w1, swap, r1, r2);
  if (rc) return rc;
  rc = zkp_rp_prove_run(c);
  if (rc) return rc;
  return zkp_rp_prove_fetch(c, c1, c2, digest, kind, resp_w, resp_r);
}

// ----------------------------------------------------------------- verify
static int rp_verify_alloc(zkp_ctx* c, int batch, int ef, int wl) {
  RpState& s = c->rp;
  const int nl = c->n.limbs, nnl = c->nn.limbs;
  const size_t be = (size_t)batch * ef;
  ZKP_CU(c, s.v_cx.ensure((size_t)batch * nnl * 4));
  ZKP_CU(c, s.v_digest.ensure((size_t)batch * 32));
  ZKP_CU(c, s.v_jobs_base.ensure(2 * be * nl * 4));
  ZKP_CU(c, s.v_jobs_plain.ensure(2 * be * wl * 4));
  ZKP_CU(c, s.v_tag.ensure(2 * be * 4));
  ZKP_CU(c, s.v_jobs_out.ensure(2 * be * nnl * 4));
  ZKP_CU(c, s.v_count.ensure(16));
  ZKP_CU(c, s.v_cmul.ensure(be * nnl * 4));
  ZKP_CU(c, s.v_sel.ensure(be));
  ZKP_CU(c, s.v_ok.ensure(be));
  ZKP_CU(c, s.v_accept.ensure((size_t)batch));
  ZKP_CU(c, s.v_fault.ensure((size_t)batch));
  ZKP_CU(c, ensure_table(c, c->nn.S, kTableShared));
  return ZKP_OK;
}

int zkp_rp_verify_stage(zkp_ctx* c, int batch, int ef, int wl, const uint32_t* range, const uint32_t* cipher_x,
                        const uint32_t* c1, const uint32_t* c2, const uint8_t* kind, const uint32_t* resp_w,
                        const uint32_t* resp_r) {
  if (!c) return ZKP_E_ARG;
  int rc = rp_check_shape(c, batch, ef, wl);
  if (rc) return rc;
  if (!range || !cipher_x || !c1 || !c2 || !kind || !resp_w || !resp_r) return fail(c, ZKP_E_ARG, "null input");
  ZKP_CU(c, cudaSetDevice(c->device));
  RpState& s = c->rp;
  s.verify_staged = s.verify_done = false;
  const int nl = c->n.limbs, nnl = c->nn.limbs;
  const size_t be = (size_t)batch * ef;
  rc = rp_verify_alloc(c, batch, ef, wl);
  if (rc) return rc;
  ZKP_CU(c, s.v_range.ensure((size_t)batch * wl * 4));
  ZKP_CU(c, s.v_c.ensure(2 * be * nnl * 4));
  ZKP_CU(c, s.v_kind.ensure(be));
  ZKP_CU(c, s.v_resp_w.ensure(2 * be * wl * 4));
  ZKP_CU(c, s.v_resp_r.ensure(2 * be * nl * 4));
  cudaStream_t st = c->stream;
  ZKP_CU(c, cudaMemcpyAsync(s.v_range.p, range, (size_t)batch * wl * 4, cudaMemcpyHostToDevice, st));
  ZKP_CU(c, cudaMemcpyAsync(s.v_cx.p, cipher_x, (size_t)batch * nnl * 4, cudaMemcpyHostToDevice, st));
  ZKP_CU(c, cudaMemcpyAsync(s.v_c.p, c1, be * nnl * 4, cudaMemcpyHostToDevice, st));
  ZKP_CU(c, cudaMemcpyAsync(s.v_c.as<uint32_t>() + be * nnl, c2, be * nnl * 4, cudaMemcpyHostToDevice, st));
  ZKP_CU(c, cudaMemcpyAsync(s.v_kind.p, kind, be, cudaMemcpyHostToDevice, st));
  ZKP_CU(c, cudaMemcpyAsync(s.v_resp_w.p, resp_w, 2 * be * wl * 4, cudaMemcpyHostToDevice, st));
  ZKP_CU(c, cudaMemcpyAsync(s.v_resp_r.p, resp_r, 2 * be * nl * 4, cudaMemcpyHostToDevice, st));
  ZKP_CU(c, cudaStreamSynchronize(st));
  s.vbatch = batch; s.vef = ef; s.vwl = wl;
  s.pv_range = s.v_range.as<uint32_t>();
  s.pv_c = s.v_c.as<uint32_t>();
  s.pv_kind = s.v_kind.as<uint8_t>();
  s.pv_resp_w = s.v_resp_w.as<uint32_t>();
  s.pv_resp_r = s.v_resp_r.as<uint32_t>();
  s.verify_staged = true;
  return ZKP_OK;
}

int zkp_rp_verify_stage_from_prove(zkp_ctx* c, const uint32_t* cipher_x) {
  if (!c) return ZKP_E_ARG;
  RpState& s = c->rp;
  if (!s.prove_done || !c->paillier) return fail(c, ZKP_E_STATE, "no proved batch on the device");
  if (!cipher_x) return fail(c, ZKP_E_ARG, "null input");
  ZKP_CU(c, cudaSetDevice(c->device));
  s.verify_staged = s.verify_done = false;
  int rc = rp_verify_alloc(c, s.batch, s.ef, s.wl);
  if (rc) return rc;
  ZKP_CU(c, cudaMemcpyAsync(s.v_cx.p, cipher_x, (size_t)s.batch * c->nn.limbs * 4, cudaMemcpyHostToDevice, c->stream));
  ZKP_CU(c, cudaStreamSynchronize(c->stream));
  s.vbatch = s.batch; s.vef = s.ef; s.vwl = s.wl;
  s.pv_range = s.range.as<uint32_t>();
  s.pv_c = s.c.as<uint32_t>();
  s.pv_kind = s.kind.as<uint8_t>();
  s.pv_resp_w = s.resp_w.as<uint32_t>();
  s.pv_resp_r = s.resp_r.as<uint32_t>();
  s.verify_staged = true;
  return ZKP_OK;
}

int zkp_rp_verify_run(zkp_ctx* c) { return zkp_rp_verify_run_with_challenge(c, nullptr, 0); }

// RangeProof::verifier_output with the verifier's own ChallengeBits (interactive proof, range_proof.rs:254-355);
// challenge == NULL recomputes the Fiat-Shamir bits (RangeProofNi::verify).
int zkp_rp_verify_run_with_challenge(zkp_ctx* c, const uint8_t* challenge, int chal_bytes) {
  if (!c) return ZKP_E_ARG;
  RpState& s = c->rp;
  if (!s.verify_staged || !c->paillier) return fail(c, ZKP_E_STATE, "nothing staged for verify");
  if (challenge && chal_bytes <= 0) return fail(c, ZKP_E_ARG, "chal_bytes must be positive");
  ZKP_CU(c, cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  if (challenge) {
    ZKP_CU(c, s.v_chal.ensure((size_t)s.vbatch * chal_bytes));
    ZKP_CU(c, cudaMemcpyAsync(s.v_chal.p, challenge, (size_t)s.vbatch * chal_bytes, cudaMemcpyHostToDevice, st));
  }
  s.v_chal_bytes = challenge ? chal_bytes : 0;
  const int batch = s.vbatch, ef = s.vef, wl = s.vwl, nl = c->n.limbs, nnl = c->nn.limbs;
  const int be = batch * ef;
  ZKP_CU(c, cudaMemsetAsync(s.v_fault.p, 0, (size_t)batch, st));
  ZKP_CU(c, cudaMemsetAsync(s.v_count.p, 0, 16, st));
  if (!challenge) {  // e = H(n, c1.., c2..)  (range_proof_ni.rs:89-92) on an auxiliary stream, beside the encryptions below
    ZKP_CU(c, fork_stream(c, 0));
    {
      ProfScope ps(c, KID_SHA, batch);
      ZKP_CU(c, launch_sha256_transcript(rp_transcript(c, s.pv_c, batch, ef), batch, s.v_digest.as<uint8_t>(), c->stream));
    }
    ZKP_CU(c, main_stream(c));
  }
  RpVerifyArgs a;
  a.batch = batch; a.ef = ef; a.wl = wl; a.nl = nl;
  a.range = s.pv_range; a.c = s.pv_c; a.kind = s.pv_kind; a.resp_w = s.pv_resp_w; a.resp_r = s.pv_resp_r;
  a.digest = s.v_digest.as<uint8_t>(); a.cmul = s.v_cmul.as<uint32_t>();
  a.chal = challenge ? s.v_chal.as<uint8_t>() : nullptr; a.chal_bytes = s.v_chal_bytes;
  a.jobs_base = s.v_jobs_base.as<uint32_t>(); a.jobs_plain = s.v_jobs_plain.as<uint32_t>();
  a.jobs_out = s.v_jobs_out.as<uint32_t>(); a.tag = s.v_tag.as<uint32_t>(); a.count = s.v_count.as<unsigned>();
  a.sel = s.v_sel.as<uint8_t>(); a.ok = s.v_ok.as<uint8_t>(); a.fault = s.v_fault.as<uint8_t>();
  {
    ProfScope ps(c, KID_OTHER, be);
    ZKP_CU(c, launch_rp_plan(a, st));
  }
  // The Enc count is data dependent (ef + #Open per proof): K1 reads it from the device,
  // the host fetches it after the launch only to report units.
  int prof_idx;
  {
    ProfScope ps(c, KID_MODEXP_SHARED, 0.0);
    prof_idx = ps.idx;
    ZKP_CU(c, launch_enc(c, a.jobs_base, nl, a.jobs_plain, wl, a.jobs_out, 2 * be, a.count));
  }
  {  // c_j * cipher_x mod n^2 for the Mask rows (range_proof.rs:321-327)
    ProfScope ps(c, KID_MODMUL, be);
    ZKP_CU(c, launch_modmul_select(c->nn.view(), a.sel, s.pv_c, s.pv_c + (size_t)be * nnl, nnl, s.v_cx.as<uint32_t>(), nnl,
                                   ef, s.v_cmul.as<uint32_t>(), nnl, be, st));
  }
  ZKP_CU(c, join_streams(c));  // the digest is needed from here on
  {
    ProfScope ps(c, KID_OTHER, be);
    ZKP_CU(c, launch_rp_bits(a, st));
    ZKP_CU(c, launch_rp_check(a, st));
    ZKP_CU(c, launch_rp_accept(a.ok, a.fault, batch, ef, s.v_accept.as<uint8_t>(), st));
  }
  unsigned count = 0;
  ZKP_CU(c, cudaMemcpyAsync(&count, s.v_count.p, 4, cudaMemcpyDeviceToHost, st));
  ZKP_CU(c, cudaStreamSynchronize(st));
  s.enc_count = count;
  if (prof_idx >= 0) c->prof[prof_idx].units = (double)count;
  s.verify_done = true;
  return ZKP_OK;
}

int zkp_rp_verify_fetch(zkp_ctx* c, uint8_t* accept, uint8_t* fault, uint8_t* digest) {
  if (!c) return ZKP_E_ARG;
  RpState& s = c->rp;
  if (!s.verify_done) return fail(c, ZKP_E_STATE, "verify has not run");
  ZKP_CU(c, cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  if (accept) ZKP_CU(c, cudaMemcpyAsync(accept, s.v_accept.p, (size_t)s.vbatch, cudaMemcpyDeviceToHost, st));
  if (fault) ZKP_CU(c, cudaMemcpyAsync(fault, s.v_fault.p, (size_t)s.vbatch, cudaMemcpyDeviceToHost, st));
  if (digest) ZKP_CU(c, cudaMemcpyAsync(digest, s.v_digest.p, (size_t)s.vbatch * 32, cudaMemcpyDeviceToHost, st));
  ZKP_CU(c, cudaStreamSynchronize(st));
  return ZKP_OK;
}

long long zkp_rp_verify_enc_count(zkp_ctx* c) { return c ? c->rp.enc_count : 0; }

int zkp_rangeproof_ni_verify(zkp_ctx* c, int batch, int ef, int wl, const uint32_t* range, const uint32_t* cipher_x,
                             const uint32_t* c1, const uint32_t* c2, const uint8_t* kind, const uint32_t* resp_w,
                             const uint32_t* resp_r, uint8_t* accept, uint8_t* fault, uint8_t* digest) {
  int rc = zkp_rp_verify_stage(c, batch, ef, wl, range, cipher_x, c1, c2, kind, resp_w, resp_r);
  if (rc) return rc;
  rc = zkp_rp_verify_run(c);
  if (rc) return rc;
  return zkp_rp_verify_fetch(c, accept, fault, digest);
}

}  // extern "C"
