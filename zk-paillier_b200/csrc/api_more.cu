// C-ABI layer, part 5: the remaining public proofs of the reference, batched:
//   CorrectOpening::verify_opening  reference src/zkproofs/correct_opening.rs:17-30
//   CompositeDLogProof              reference src/zkproofs/wi_dlog_proof.rs:46-91      (one modulus N per statement)
//   CorrectMessageProof             reference src/zkproofs/correct_message.rs:35-162   (ring proof, M valid messages)
// Same conventions as api_sigma.cu: randomness is an input, Err(IncorrectProof) comes back in accept[], the inputs on
// which the reference panics (assert!, unwrap(), index out of range) in fault[].
#include "sig.h"

using namespace zkp;

namespace {
struct DLog {
  zkp_ctx* c;
  cudaStream_t st;
  int batch, nl, S;
  Arena ar;
  bool bad = false;
  uint32_t *N = nullptr, *g = nullptr, *ni = nullptr, *r2 = nullptr, *n0 = nullptr;
  DLog(zkp_ctx* ctx, int b, int nl_) : c(ctx), st(ctx->stream), batch(b), nl(nl_), S(pick_width(nl_)), ar(ctx) {}
  template <class U>
  U* get(size_t count) {
    U* p = ar.get<U>(count);
    if (!p) bad = true;
    return p;
  }
  uint32_t* up(const uint32_t* host, int limbs) {
    uint32_t* p = get<uint32_t>((size_t)batch * limbs);
    if (p && cudaMemcpyAsync(p, host, (size_t)batch * limbs * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) bad = true;
    return p;
  }
  void ck(cudaError_t e) {
    if (e != cudaSuccess) {
      bad = true;
      fail_cuda(c, e, "CompositeDLogProof launch");
    }
  }
  int begin(const uint32_t* hN, const uint32_t* hg, const uint32_t* hni, size_t extra_bytes) {
    if (batch <= 0 || nl <= 0 || nl % 4) return fail(c, ZKP_E_ARG, "bad batch / n_limbs");
    if (S < 0) return fail(c, ZKP_E_ARG, "modulus wider than 8192 bits");
    for (int b = 0; b < batch; ++b)
      if (!(hN[(size_t)b * nl] & 1u)) return fail(c, ZKP_E_ARG, "every modulus must be odd");
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return fail_cuda(c, e, "cudaSetDevice");
    c->err.clear();
    e = ar.reserve((size_t)batch * ((size_t)12 * nl * 4 + (size_t)S * 4 + extra_bytes + 1024) + 64 * 256);
    if (e != cudaSuccess) return fail_cuda(c, e, "arena");
    e = ensure_table(c, S, kTableVar);
    if (e != cudaSuccess) return fail_cuda(c, e, "table");
    N = up(hN, nl);
    g = up(hg, nl);
    ni = up(hni, nl);
    r2 = get<uint32_t>((size_t)batch * S);
    n0 = get<uint32_t>((size_t)batch);
    if (!bad) {
      ProfScope ps(c, KID_OTHER, batch);
      ck(launch_mont_setup(N, nl, S, batch, r2, n0, st));
    }
    return ZKP_OK;
  }
  // BigInt::mod_pow(base_b, exp_b, N_b)
  uint32_t* powm(const uint32_t* base, const uint32_t* exp, int exp_limbs) {
    uint32_t* out = get<uint32_t>((size_t)batch * nl);
    if (bad) return out;
    ProfScope ps(c, KID_MODEXP_VAR, batch);
    ck(launch_modexp_var(base, N, nl, r2, n0, exp, exp_limbs, 32 * exp_limbs, 1, 1, out, batch, S, c->table.as<uint32_t>(),
                         c->num_sms, st));
    return out;
  }
  // e = compute_digest(x, g, N, ni)      wi_dlog_proof.rs:55-60,74-79
  uint32_t* challenge(const uint32_t* x) {
    uint8_t* dig = get<uint8_t>((size_t)batch * 32);
    uint32_t* e = get<uint32_t>((size_t)batch * 8);
    if (bad) return e;
    ShaSegs s;
    s.nseg = 0;
    for (const uint32_t* p : {x, (const uint32_t*)g, (const uint32_t*)N, (const uint32_t*)ni}) s.seg[s.nseg++] = {p, (long long)nl, 1, nl};
    {
      ProfScope ps(c, KID_SHA, batch);
      ck(launch_sha256_transcript(s, batch, dig, st));
    }
    ProfScope ps(c, KID_OTHER, batch);
    ck(launch_digest_to_limbs(dig, batch, e, st));
    return e;
  }
  int finish(const char* what) {
    cudaError_t e = cudaStreamSynchronize(st);
    if (bad) {
      if (c->err.empty()) c->err = what;
      return ZKP_E_CUDA;
    }
    if (e != cudaSuccess) return fail_cuda(c, e, what);
    return ZKP_OK;
  }
};
}  // namespace

namespace {
// u_i = ciphertext * (1 + m_i n)^-1 mod nn for the M valid messages of every proof   (correct_message.rs:50-56,134-142)
uint32_t* cm_u(Sig& s, const uint32_t* d_valid, int ml, const uint32_t* d_cipher, int M) {
  const int rows = s.batch * M;
  uint32_t* one = s.rows(s.nl, 1);
  if (!s.bad) {
    cudaMemsetAsync(one, 0, (size_t)s.nl * 4, s.st);
    const uint32_t v = 1u;
    cudaMemcpyAsync(one, &v, 4, cudaMemcpyHostToDevice, s.st);
  }
  uint32_t* vred = s.mulm_n(d_valid, ml, one, s.nl, rows, rows);   // m_i mod n
  uint32_t* gminv = s.rows(s.nnl, rows);
  if (!s.bad) s.ck(launch_gm_inv(vred, s.c->n.mod.as<uint32_t>(), s.nl, rows, gminv, s.st));
  return s.mulm(gminv, s.nnl, d_cipher, s.nnl, rows, M);           // row (b, i) times ciphertext_b
}
// chal = compute_digest(a_vec) mod 2^256: the digest itself     (:85-87,128-129)
uint32_t* cm_challenge(Sig& s, const uint32_t* d_a, int M) {
  uint8_t* dig = s.bytes((size_t)s.batch * 32);
  uint32_t* e = s.rows(8);
  if (s.bad) return e;
  ShaSegs sg;
  sg.nseg = 1;
  sg.seg[0] = {d_a, (long long)M * s.nnl, M, s.nnl};
  {
    ProfScope ps(s.c, KID_SHA, s.batch);
    s.ck(launch_sha256_transcript(sg, s.batch, dig, s.st));
  }
  ProfScope ps(s.c, KID_OTHER, s.batch);
  s.ck(launch_digest_to_limbs(dig, s.batch, e, s.st));
  return e;
}
}  // namespace

extern "C" {

// ------------------------------------------------------------- CorrectOpening
// ok[b] = (c[b] == Paillier::encrypt_with_chosen_randomness(ek, m[b], r[b]))      correct_opening.rs:26-29
int zkp_verify_opening(zkp_ctx* c, int batch, int m_limbs, const uint32_t* m, const uint32_t* r, const uint32_t* cc, uint8_t* ok) {
  if (!c || !m || !r || !cc || !ok) return c ? fail(c, ZKP_E_ARG, "null buffer") : ZKP_E_ARG;
  Sig s(c, batch);
  if (c->paillier && (m_limbs <= 0 || m_limbs % 4 || m_limbs > c->nn.S)) return fail(c, ZKP_E_ARG, "bad m_limbs");
  int rc = sigma_begin(c, batch, 0, 3, 4 * (size_t)(m_limbs + s.nl) + 64, s);
  if (rc) return rc;
  uint32_t* d_m = s.up(m, m_limbs);
  uint32_t* d_r = s.up(r, s.nl);
  uint32_t* d_c = s.up(cc, s.nnl);
  uint8_t* d_ok = s.bytes((size_t)batch);
  uint32_t* d = s.enc(d_m, m_limbs, d_r, s.nl);
  if (!s.bad) s.ck(launch_rows_equal(d_c, d, s.nnl, batch, 0, d_ok, s.st));
  s.down8(ok, d_ok, (size_t)batch);
  return s.finish("zkp_verify_opening");
}

// --------------------------------------------------------- CompositeDLogProof

// prove (wi_dlog_proof.rs:46-65): x = g^r mod N, e = H(x, g, N, ni), y = r + e * secret (unreduced).
// r is the caller's sample below 2^(K + K' + S) = 2^512 (r_limbs >= 16 for the reference's constants);
// fault[b] = 1 if y does not fit y_limbs.
int zkp_dlog_prove(zkp_ctx* c, int batch, int n_limbs, const uint32_t* N, const uint32_t* g, const uint32_t* ni,
                   const uint32_t* secret, int secret_limbs, const uint32_t* r, int r_limbs, int y_limbs, uint32_t* x, uint32_t* y,
                   uint8_t* fault) {
  if (!c) return ZKP_E_ARG;
  if (!N || !g || !ni || !secret || !r || !x || !y || !fault) return fail(c, ZKP_E_ARG, "null buffer");
  if (secret_limbs <= 0 || r_limbs <= 0 || r_limbs % 2 || y_limbs < r_limbs || y_limbs % 2) return fail(c, ZKP_E_ARG, "bad widths");
  DLog d(c, batch, n_limbs);
  int rc = d.begin(N, g, ni, 4 * (size_t)(secret_limbs + r_limbs + y_limbs) + 64);
  if (rc) return rc;
  uint32_t* d_s = d.up(secret, secret_limbs);
  uint32_t* d_r = d.up(r, r_limbs);
  uint8_t* d_fault = d.get<uint8_t>((size_t)batch);
  if (d_fault) cudaMemsetAsync(d_fault, 0, (size_t)batch, d.st);
  uint32_t* d_x = d.powm(d.g, d_r, r_limbs);                                           // :54
  uint32_t* e = d.challenge(d_x);                                                      // :55-60
  uint32_t* d_y = d.get<uint32_t>((size_t)batch * y_limbs);
  if (!d.bad) d.ck(launch_muladd(d_r, r_limbs, d_s, secret_limbs, e, 8, batch, d_y, y_limbs, d_fault, d.st));  // :61
  if (!d.bad) {
    if (cudaMemcpyAsync(x, d_x, (size_t)batch * n_limbs * 4, cudaMemcpyDeviceToHost, d.st) != cudaSuccess) d.bad = true;
    if (cudaMemcpyAsync(y, d_y, (size_t)batch * y_limbs * 4, cudaMemcpyDeviceToHost, d.st) != cudaSuccess) d.bad = true;
    if (cudaMemcpyAsync(fault, d_fault, (size_t)batch, cudaMemcpyDeviceToHost, d.st) != cudaSuccess) d.bad = true;
  }
  return d.finish("zkp_dlog_prove");
}

// verify (wi_dlog_proof.rs:66-91): fault[b] where the reference's asserts fire (N <= 2^128, gcd(g, N) != 1,
// gcd(ni, N) != 1); accept[b] = (x == g^y * ni^e mod N) otherwise.
int zkp_dlog_verify(zkp_ctx* c, int batch, int n_limbs, const uint32_t* N, const uint32_t* g, const uint32_t* ni, const uint32_t* x,
                    const uint32_t* y, int y_limbs, uint8_t* accept, uint8_t* fault) {
  if (!c) return ZKP_E_ARG;
  if (!N || !g || !ni || !x || !y || !accept || !fault) return fail(c, ZKP_E_ARG, "null buffer");
  if (y_limbs <= 0 || y_limbs % 2) return fail(c, ZKP_E_ARG, "bad y_limbs");
  DLog d(c, batch, n_limbs);
  int rc = d.begin(N, g, ni, 4 * (size_t)y_limbs + 64);
  if (rc) return rc;
  const int nl = n_limbs;
  uint32_t* d_x = d.up(x, nl);
  uint32_t* d_y = d.up(y, y_limbs);
  uint8_t* d_fault = d.get<uint8_t>((size_t)batch);
  uint8_t* d_acc = d.get<uint8_t>((size_t)batch);
  uint32_t* scratch = d.get<uint32_t>((size_t)batch * nl);
  if (d_fault) cudaMemsetAsync(d_fault, 0, (size_t)batch, d.st);
  if (!d.bad) d.ck(launch_gt_pow2(d.N, nl, 128, batch, d_fault, d.st));                                       // :68
  if (!d.bad) d.ck(launch_modinv(d.g, d.N, nl, batch, nullptr, scratch, d_fault, d.st, nl));                  // :71  gcd(g, N) == 1
  if (!d.bad) d.ck(launch_modinv(d.ni, d.N, nl, batch, nullptr, scratch, d_fault, d.st, nl));                 // :72  gcd(ni, N) == 1
  uint32_t* e = d.challenge(d_x);                                                                            // :74-79
  uint32_t* ni_e = d.powm(d.ni, e, 8);                                                                       // :80
  uint32_t* g_y = d.powm(d.g, d_y, y_limbs);                                                                 // :81
  uint32_t* prod = d.get<uint32_t>((size_t)batch * nl);
  if (!d.bad) {
    ProfScope ps(c, KID_MODMUL, batch);
    d.ck(launch_modmul_var(g_y, ni_e, nl, d.N, d.r2, d.n0, 1, d.S, batch, prod, d.st));                      // :82
  }
  if (!d.bad) d.ck(launch_rows_equal(d_x, prod, nl, batch, 0, d_acc, d.st));                                 // :85
  if (!d.bad) {
    if (cudaMemcpyAsync(accept, d_acc, (size_t)batch, cudaMemcpyDeviceToHost, d.st) != cudaSuccess) d.bad = true;
    if (cudaMemcpyAsync(fault, d_fault, (size_t)batch, cudaMemcpyDeviceToHost, d.st) != cudaSuccess) d.bad = true;
  }
  rc = d.finish("zkp_dlog_verify");
  if (rc == ZKP_OK)
    for (int i = 0; i < batch; ++i)
      if (fault[i]) accept[i] = 0;
  return rc;
}

// -------------------------------------------------------- CorrectMessageProof

// prove (correct_message.rs:35-125).  Per proof: M valid messages [M][m_limbs], the message to encrypt, and the prover's
// randomness r, e_rand [M-1][8] (256-bit), z_rand [M-1][n_limbs], w.  Out: ciphertext, e_vec [M][8], z_vec [M][n_limbs],
// a_vec [M][nn_limbs].  fault[b] = 1 when the message is not among the valid ones (the reference indexes past its
// random vectors and panics) or an inverse does not exist (unwrap()).
int zkp_correct_message_prove(zkp_ctx* c, int batch, int M, int m_limbs, const uint32_t* valid, const uint32_t* msg, const uint32_t* r,
                              const uint32_t* e_rand, const uint32_t* z_rand, const uint32_t* w, uint32_t* ciphertext,
                              uint32_t* e_vec, uint32_t* z_vec, uint32_t* a_vec, uint8_t* fault) {
  if (!c) return ZKP_E_ARG;
  if (!valid || !msg || !r || (M > 1 && (!e_rand || !z_rand)) || !w || !ciphertext || !e_vec || !z_vec || !a_vec || !fault)
    return fail(c, ZKP_E_ARG, "null buffer");
  if (M <= 0 || M > 4096) return fail(c, ZKP_E_ARG, "bad number of messages");
  if (c->paillier && (m_limbs <= 0 || m_limbs % 4 || m_limbs > c->n.limbs)) return fail(c, ZKP_E_ARG, "m_limbs must be a multiple of 4, at most n_limbs");
  if ((long long)batch * M > 0x3fffffffll) return fail(c, ZKP_E_ARG, "batch too large");
  Sig s(c, batch);
  const int el = 8;
  int rc = sigma_begin(c, batch, 0, (size_t)10 * M + 4, (size_t)M * (4 * (size_t)(m_limbs + 3 * s.nl + 3 * el) + 16) + 256, s);
  if (rc) return rc;
  const int nl = s.nl, nnl = s.nnl, rows = batch * M, rnd = batch * (M - 1);
  uint32_t* d_valid = s.up(valid, m_limbs, rows);
  uint32_t* d_msg = s.up(msg, m_limbs);
  uint32_t* d_r = s.up(r, nl);
  uint32_t* d_er = rnd ? s.up(e_rand, el, rnd) : s.rows(el, 1);
  uint32_t* d_zr = rnd ? s.up(z_rand, nl, rnd) : s.rows(nl, 1);
  uint32_t* d_w = s.up(w, nl);
  uint8_t* d_fault = s.bytes((size_t)batch, true);
  uint8_t* d_rowfault = s.bytes((size_t)rows, true);
  uint8_t* d_match = s.bytes((size_t)rows);
  uint32_t *esel = s.rows(el, rows), *zsel = s.rows(nl, rows), *esum = s.rows(el);
  if (!s.bad) s.ck(launch_cm_layout(d_valid, d_msg, m_limbs, d_er, el, d_zr, d_w, nl, batch, M, d_match, esel, zsel, esum, d_fault, s.st));
  uint32_t* d_c = s.enc(d_msg, m_limbs, d_r, nl);                                   // ciphertext = Enc(message, r)     :43-49
  uint32_t* u = cm_u(s, d_valid, m_limbs, d_c, M);                                  // u_i                              :50-56
  s.fork(0);
  uint32_t* zn = s.enc(nullptr, 0, zsel, nl, rows);                                 // z_i^n (w^n in the true slot)      :69,71
  s.on_main();
  uint32_t* ue = s.powm(u, nnl, esel, el, rows);                                    // u_i^e_i (u^0 = 1 in the true slot) :72
  uint32_t* ueinv = s.rows(nnl, rows);
  if (!s.bad) s.ck(launch_modinv(ue, c->nn.mod.as<uint32_t>(), nnl, rows, nullptr, ueinv, d_rowfault, s.st));  // :73 unwrap()
  s.join();
  uint32_t* d_a = s.mulm(zn, nnl, ueinv, nnl, rows);                                // a_i                              :75
  uint32_t* chal = cm_challenge(s, d_a, M);                                         // :85-87
  uint32_t* ei = s.rows(el);
  if (!s.bad) s.ck(launch_sub_pow2(chal, esum, el, batch, ei, s.st));               // e = chal - sum e_j mod 2^256     :88-91
  uint32_t* r_ei = s.rows(nl);
  if (!s.bad) {                                                                     // r^e mod n                        :92
    ProfScope ps(c, KID_MODEXP_VAR, batch);
    s.ck(launch_modexp_var(d_r, c->n.mod.as<uint32_t>(), nl, c->n.r2.as<uint32_t>(), c->n.n0.as<uint32_t>(), ei, el, 32 * el, 1,
                           0x7fffffff, r_ei, batch, c->n.S, c->table.as<uint32_t>(), c->num_sms, s.st));
  }
  uint32_t* zi = s.mulm_n(d_w, nl, r_ei, nl);                                       // z = w r^e mod n                  :93
  if (!s.bad) s.ck(launch_cm_finish(d_match, ei, el, zi, nl, batch, M, esel, zsel, s.st));  // :95-121
  if (!s.bad) s.ck(launch_rows_reduce(d_rowfault, batch, M, 1, 1, d_fault, s.st));
  s.down(ciphertext, d_c, nnl);
  s.down(e_vec, esel, el, rows);
  s.down(z_vec, zsel, nl, rows);
  s.down(a_vec, d_a, nnl, rows);
  s.down8(fault, d_fault, (size_t)batch);
  return s.finish("zkp_correct_message_prove");
}

// verify (correct_message.rs:126-161): fault[b] = 1 where assert_eq!(chal, ei_sum) fires; otherwise
// accept[b] = AND_i ( u_i^e_i * a_i == z_i^n  mod nn ).  e_vec rows are e_limbs wide (8 for honest proofs).
int zkp_correct_message_verify(zkp_ctx* c, int batch, int M, int m_limbs, int e_limbs, const uint32_t* ciphertext, const uint32_t* valid,
                               const uint32_t* e_vec, const uint32_t* z_vec, const uint32_t* a_vec, uint8_t* accept, uint8_t* fault) {
  if (!c) return ZKP_E_ARG;
  if (!ciphertext || !valid || !e_vec || !z_vec || !a_vec || !accept || !fault) return fail(c, ZKP_E_ARG, "null buffer");
  if (M <= 0 || M > 4096 || e_limbs < 8 || e_limbs % 2) return fail(c, ZKP_E_ARG, "bad number of messages / e_limbs");
  if (c->paillier && (m_limbs <= 0 || m_limbs % 4 || m_limbs > c->n.limbs)) return fail(c, ZKP_E_ARG, "m_limbs must be a multiple of 4, at most n_limbs");
  if ((long long)batch * M > 0x3fffffffll) return fail(c, ZKP_E_ARG, "batch too large");
  Sig s(c, batch);
  int rc = sigma_begin(c, batch, 0, (size_t)9 * M + 4, (size_t)M * (4 * (size_t)(m_limbs + 2 * s.nl + 2 * e_limbs) + 16) + 256, s);
  if (rc) return rc;
  const int nl = s.nl, nnl = s.nnl, rows = batch * M;
  uint32_t* d_c = s.up(ciphertext, nnl);
  uint32_t* d_valid = s.up(valid, m_limbs, rows);
  uint32_t* d_e = s.up(e_vec, e_limbs, rows);
  uint32_t* d_z = s.up(z_vec, nl, rows);
  uint32_t* d_a = s.up(a_vec, nnl, rows);
  uint8_t* d_fault = s.bytes((size_t)batch, true);
  uint8_t* d_rowok = s.bytes((size_t)rows);
  uint8_t* d_acc = s.bytes((size_t)batch);
  uint32_t* chal = cm_challenge(s, d_a, M);                                         // :128-129
  uint32_t* esum = s.rows(8);
  if (!s.bad) s.ck(launch_sum_pow2(d_e, e_limbs, 8, batch, M, esum, s.st));         // sum e_i mod 2^256                :130-131
  if (!s.bad) s.ck(launch_rows_differ_fault(chal, esum, 8, batch, d_fault, s.st));  // assert_eq!(chal, ei_sum)         :133
  uint32_t* u = cm_u(s, d_valid, m_limbs, d_c, M);                                  // :134-142
  s.fork(0);
  uint32_t* zn = s.enc(nullptr, 0, d_z, nl, rows);                                  // z_i^n                            :145
  s.on_main();
  uint32_t* ue = s.powm(u, nnl, d_e, e_limbs, rows);                                // u_i^e_i                          :146
  uint32_t* lhs = s.mulm(ue, nnl, d_a, nnl, rows);                                  // :147
  s.join();
  if (!s.bad) s.ck(launch_rows_equal(lhs, zn, nnl, rows, 0, d_rowok, s.st));        // :148
  if (!s.bad) s.ck(launch_rows_reduce(d_rowok, batch, M, 0, 0, d_acc, s.st));       // :151
  s.down8(accept, d_acc, (size_t)batch);
  s.down8(fault, d_fault, (size_t)batch);
  rc = s.finish("zkp_correct_message_verify");
  if (rc == ZKP_OK)
    for (int i = 0; i < batch; ++i)
      if (fault[i]) accept[i] = 0;
  return rc;
}

}  // extern "C"
