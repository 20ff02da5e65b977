// C-ABI layer, part 4: the sigma-protocol proofs for a batch of statements under one key:
//   ZeroProof        reference src/zkproofs/zero_enc_proof.rs:44-94
//   CiphertextProof  reference src/zkproofs/correct_ciphertext.rs:42-98
//   MulProof         reference src/zkproofs/multiplication_proof.rs:60-145
//   VerlinProof      reference src/zkproofs/verlin_proof.rs:60-165
// Each call uploads the statement/witness/randomness rows, sequences K1 (Enc), K2 (mod_pow with the
// per-proof challenge / response exponents), K3 (mod_mul), K4 (the Fiat-Shamir hash) and the
// per-proof helpers of sigma.cu on the device, and downloads the proof rows or verdicts.
#include "sig.h"

using namespace zkp;

extern "C" {

// ------------------------------------------------------------------ ZeroProof
int zkp_zero_prove(zkp_ctx* c, int batch, const uint32_t* r, const uint32_t* cc, const uint32_t* r_prime, uint32_t* z,
                   uint32_t* a) {
  if (!c || !r || !cc || !r_prime || !z || !a) return c ? fail(c, ZKP_E_ARG, "null buffer") : ZKP_E_ARG;
  Sig s(c, batch);
  int rc = sigma_begin(c, batch, 0, 6, 4 * 4 * (size_t)s.nl, s);
  if (rc) return rc;
  uint32_t* d_r = s.up(r, s.nl);
  uint32_t* d_c = s.up(cc, s.nnl);
  uint32_t* d_rp = s.up(r_prime, s.nl);
  uint32_t* d_a = s.enc(nullptr, 0, d_rp, s.nl);                 // a = Enc(0, r')            :46-52
  uint32_t* e = s.challenge({d_c, d_a});                         // e = H(n, c, a)            :54-58
  uint32_t* r_e = s.powm(d_r, s.nl, e, 8);                       // r^e mod nn                :60
  uint32_t* d_z = s.mulm(d_rp, s.nl, r_e, s.nnl);                // z = r' * r^e mod nn       :61
  s.down(z, d_z, s.nnl);
  s.down(a, d_a, s.nnl);
  return s.finish("zkp_zero_prove");
}

int zkp_zero_verify(zkp_ctx* c, int batch, const uint32_t* cc, const uint32_t* z, const uint32_t* a, uint8_t* accept) {
  if (!c || !cc || !z || !a || !accept) return c ? fail(c, ZKP_E_ARG, "null buffer") : ZKP_E_ARG;
  Sig s(c, batch);
  int rc = sigma_begin(c, batch, 0, 7, 64, s);
  if (rc) return rc;
  uint32_t* d_c = s.up(cc, s.nnl);
  uint32_t* d_z = s.up(z, s.nnl);
  uint32_t* d_a = s.up(a, s.nnl);
  uint8_t* d_acc = s.ar.get<uint8_t>((size_t)batch);
  uint32_t* e = s.challenge({d_c, d_a});                         // :67-71
  PowBatch pb(s);                                                // both modexps of the proof in one launch
  uint32_t* c_z = pb.enc(nullptr, 0, d_z, s.nnl);                // Enc(0, z)                 :73-79
  uint32_t* c_e = pb.powm(d_c, s.nnl, e, 8);                     // Paillier::mul(c, e)       :81-85
  pb.run();
  uint32_t* c_z_test = s.mulm(c_e, s.nnl, d_a, s.nnl);           // Paillier::add(c_e, a)     :86-88
  if (!s.bad) s.ck(launch_rows_equal(c_z, c_z_test, s.nnl, batch, 0, d_acc, s.st));
  if (cudaMemcpyAsync(accept, d_acc, (size_t)batch, cudaMemcpyDeviceToHost, s.st) != cudaSuccess) s.bad = true;
  return s.finish("zkp_zero_verify");
}

// ------------------------------------------------------------ CiphertextProof
int zkp_ciphertext_prove(zkp_ctx* c, int batch, int z_limbs, const uint32_t* x, const uint32_t* r, const uint32_t* cc,
                         const uint32_t* x_prime, const uint32_t* r_prime, uint32_t* z1, uint32_t* z2, uint32_t* c_prime) {
  if (!c || !x || !r || !cc || !x_prime || !r_prime || !z1 || !z2 || !c_prime) return c ? fail(c, ZKP_E_ARG, "null buffer") : ZKP_E_ARG;
  Sig s(c, batch);
  int rc = sigma_begin(c, batch, z_limbs, 6, 4 * (4 * (size_t)s.nl + z_limbs) + 64, s);
  if (rc) return rc;
  uint32_t* d_x = s.up(x, s.nl);
  uint32_t* d_r = s.up(r, s.nl);
  uint32_t* d_c = s.up(cc, s.nnl);
  uint32_t* d_xp = s.up(x_prime, s.nl);
  uint32_t* d_rp = s.up(r_prime, s.nl);
  uint8_t* d_fault = s.ar.get<uint8_t>((size_t)batch);
  cudaMemsetAsync(d_fault, 0, (size_t)batch, s.st);
  uint32_t* d_cp = s.enc(d_xp, s.nl, d_rp, s.nl);                // c' = Enc(x', r')          :45-51
  uint32_t* e = s.challenge({d_c, d_cp});                        // :53-57
  uint32_t* d_z1 = s.rows(z_limbs);
  if (!s.bad) s.ck(launch_muladd(d_xp, s.nl, d_x, s.nl, e, 8, batch, d_z1, z_limbs, d_fault, s.st));  // z1 = x' + x*e (unreduced) :59
  uint32_t* r_e = s.powm(d_r, s.nl, e, 8);                       // :60
  uint32_t* d_z2 = s.mulm(d_rp, s.nl, r_e, s.nnl);               // :61
  s.down(z1, d_z1, z_limbs);
  s.down(z2, d_z2, s.nnl);
  s.down(c_prime, d_cp, s.nnl);
  return s.finish("zkp_ciphertext_prove");
}

int zkp_ciphertext_verify(zkp_ctx* c, int batch, int z_limbs, const uint32_t* cc, const uint32_t* z1, const uint32_t* z2,
                          const uint32_t* c_prime, uint8_t* accept) {
  if (!c || !cc || !z1 || !z2 || !c_prime || !accept) return c ? fail(c, ZKP_E_ARG, "null buffer") : ZKP_E_ARG;
  Sig s(c, batch);
  int rc = sigma_begin(c, batch, z_limbs, 7, 4 * (size_t)z_limbs + 64, s);
  if (rc) return rc;
  uint32_t* d_c = s.up(cc, s.nnl);
  uint32_t* d_z1 = s.up(z1, z_limbs);
  uint32_t* d_z2 = s.up(z2, s.nnl);
  uint32_t* d_cp = s.up(c_prime, s.nnl);
  uint8_t* d_acc = s.ar.get<uint8_t>((size_t)batch);
  uint32_t* e = s.challenge({d_c, d_cp});                        // :67-71
  PowBatch pb(s);
  uint32_t* c_z = pb.enc(d_z1, z_limbs, d_z2, s.nnl);            // Enc(z1, z2)               :73-79
  uint32_t* c_e = pb.powm(d_c, s.nnl, e, 8);                     // :81-85
  pb.run();
  uint32_t* c_z_test = s.mulm(c_e, s.nnl, d_cp, s.nnl);          // :86-92
  if (!s.bad) s.ck(launch_rows_equal(c_z, c_z_test, s.nnl, batch, 0, d_acc, s.st));
  if (cudaMemcpyAsync(accept, d_acc, (size_t)batch, cudaMemcpyDeviceToHost, s.st) != cudaSuccess) s.bad = true;
  return s.finish("zkp_ciphertext_verify");
}

// -------------------------------------------------------------------- MulProof
int zkp_mul_prove(zkp_ctx* c, int batch, const uint32_t* a, const uint32_t* b, const uint32_t* r_a, const uint32_t* r_b,
                  const uint32_t* r_c, const uint32_t* e_a, const uint32_t* e_b, const uint32_t* e_c, const uint32_t* d,
                  const uint32_t* r_d, uint32_t* f, uint32_t* z1, uint32_t* z2, uint32_t* e_d, uint32_t* e_db, uint8_t* fault) {
  if (!c || !a || !b || !r_a || !r_b || !r_c || !e_a || !e_b || !e_c || !d || !r_d || !f || !z1 || !z2 || !e_d || !e_db || !fault)
    return c ? fail(c, ZKP_E_ARG, "null buffer") : ZKP_E_ARG;
  Sig s(c, batch);
  int rc = sigma_begin(c, batch, 0, 22, 4 * 12 * (size_t)s.nl + 64, s);
  if (rc) return rc;
  const int nl = s.nl, nnl = s.nnl;
  uint32_t *d_a = s.up(a, nl), *d_b = s.up(b, nl), *d_ra = s.up(r_a, nl), *d_rb = s.up(r_b, nl), *d_rc = s.up(r_c, nl);
  uint32_t *d_ea = s.up(e_a, nnl), *d_eb = s.up(e_b, nnl), *d_ec = s.up(e_c, nnl), *d_d = s.up(d, nl), *d_rd = s.up(r_d, nl);
  uint8_t* d_fault = s.ar.get<uint8_t>((size_t)batch);
  cudaMemsetAsync(d_fault, 0, (size_t)batch, s.st);
  uint32_t* r_db = s.rows(nnl);
  uint32_t* db = s.rows(nnl);
  if (!s.bad) s.ck(launch_muladd(nullptr, 0, d_rd, nl, d_rb, nl, batch, r_db, nnl, d_fault, s.st));  // r_db = r_d * r_b (unreduced) :70
  if (!s.bad) s.ck(launch_muladd(nullptr, 0, d_d, nl, d_b, nl, batch, db, nnl, d_fault, s.st));      // db = d * b (unreduced)       :71
  PowBatch p1(s);                                                              // the two encryptions in one launch
  uint32_t* d_ed = p1.enc(d_d, nl, d_rd, nl);                                  // e_d = Enc(d, r_d)             :63-69
  uint32_t* d_edb = p1.enc(db, nnl, r_db, nnl);                                // e_db = Enc(db, r_db)          :72-78
  p1.run();
  uint32_t* e = s.challenge({d_ea, d_eb, d_ec, d_ed, d_edb});                  // :80-87
  uint32_t* ea = s.rows(nl);
  {
    ProfScope ps(c, KID_MODMUL, batch);
    s.ck(launch_modmul_shared(c->n.view(), 0, e, 8, d_a, nl, 1, ea, nl, batch, s.st));  // ea = e*a mod n       :89
  }
  uint32_t* d_f = s.rows(nl);
  if (!s.bad) s.ck(launch_modadd(ea, d_d, c->n.mod.as<uint32_t>(), nl, batch, d_f, s.st));           // f = ea + d mod n     :90
  // the short-exponent modexps and the inversion that follows them on an auxiliary stream, next to the long one
  s.fork(0);
  PowBatch p2(s);
  uint32_t* r_a_e = p2.powm(d_ra, nl, e, 8);                                   // :91
  uint32_t* r_c_e = p2.powm(d_rc, nl, e, 8);                                   // :94
  p2.run();
  uint32_t* d_z1 = s.mulm(r_a_e, nnl, d_rd, nl);                               // :92
  uint32_t* v = s.mulm(r_db, nnl, r_c_e, nnl);                                 // :95
  uint32_t* vinv = s.rows(nnl);
  uint32_t* scratch = s.rows(4 * nnl);
  if (!s.bad) s.ck(launch_modinv(v, c->nn.mod.as<uint32_t>(), nnl, batch, scratch, vinv, d_fault, s.st));  // mod_inv(..).unwrap() :96
  s.on_main();
  PowBatch p3(s);
  uint32_t* r_b_f = p3.powm(d_rb, nl, d_f, nl);                                // :93
  p3.run();
  s.join();
  uint32_t* d_z2 = s.mulm(r_b_f, nnl, vinv, nnl);                              // :97
  s.down(f, d_f, nl);
  s.down(z1, d_z1, nnl);
  s.down(z2, d_z2, nnl);
  s.down(e_d, d_ed, nnl);
  s.down(e_db, d_edb, nnl);
  if (cudaMemcpyAsync(fault, d_fault, (size_t)batch, cudaMemcpyDeviceToHost, s.st) != cudaSuccess) s.bad = true;
  return s.finish("zkp_mul_prove");
}

int zkp_mul_verify(zkp_ctx* c, int batch, const uint32_t* e_a, const uint32_t* e_b, const uint32_t* e_c, const uint32_t* f,
                   const uint32_t* z1, const uint32_t* z2, const uint32_t* e_d, const uint32_t* e_db, uint8_t* accept,
                   uint8_t* fault) {
  if (!c || !e_a || !e_b || !e_c || !f || !z1 || !z2 || !e_d || !e_db || !accept || !fault)
    return c ? fail(c, ZKP_E_ARG, "null buffer") : ZKP_E_ARG;
  Sig s(c, batch);
  int rc = sigma_begin(c, batch, 0, 22, 4 * (size_t)s.nl + 128, s);
  if (rc) return rc;
  const int nl = s.nl, nnl = s.nnl;
  uint32_t *d_ea = s.up(e_a, nnl), *d_eb = s.up(e_b, nnl), *d_ec = s.up(e_c, nnl), *d_f = s.up(f, nl);
  uint32_t *d_z1 = s.up(z1, nnl), *d_z2 = s.up(z2, nnl), *d_ed = s.up(e_d, nnl), *d_edb = s.up(e_db, nnl);
  uint8_t* d_fault = s.ar.get<uint8_t>((size_t)batch);
  uint8_t* d_acc = s.ar.get<uint8_t>((size_t)batch);
  cudaMemsetAsync(d_fault, 0, (size_t)batch, s.st);
  s.compute_begin();
  // five independent modexps per proof in two K2h launches: the two with the 256-bit challenge as exponent (and the
  // products and the inversion that consume them) on an auxiliary stream, the three |n|-bit ones on the main stream
  uint32_t* e = s.challenge({d_ea, d_eb, d_ec, d_ed, d_edb});                  // :109-116
  s.fork(0);
  PowBatch ps(s);
  uint32_t* e_a_e = ps.powm(d_ea, nnl, e, 8);                                  // :133
  uint32_t* e_c_e = ps.powm(d_ec, nnl, e, 8);                                  // :135
  ps.run();
  uint32_t* lhs1 = s.mulm(e_a_e, nnl, d_ed, nnl);                              // :134
  uint32_t* v = s.mulm(d_edb, nnl, e_c_e, nnl);                                // :136
  uint32_t* vinv = s.rows(nnl);
  uint32_t* scratch = s.rows(4 * nnl);
  if (!s.bad) s.ck(launch_modinv(v, c->nn.mod.as<uint32_t>(), nnl, batch, scratch, vinv, d_fault, s.st));  // :137 (unwrap -> fault)
  s.on_main();
  PowBatch pl(s);
  uint32_t* enc_f_z1 = pl.enc(d_f, nl, d_z1, nnl);                             // :118-124
  uint32_t* enc_0_z2 = pl.enc(nullptr, 0, d_z2, nnl);                          // :125-131
  uint32_t* e_b_f = pl.powm(d_eb, nnl, d_f, nl);                               // :138
  pl.run();
  s.join();
  uint32_t* lhs2 = s.mulm(e_b_f, nnl, vinv, nnl);                              // :139
  if (!s.bad) s.ck(launch_rows_equal(lhs1, enc_f_z1, nnl, batch, 0, d_acc, s.st));         // :141
  if (!s.bad) s.ck(launch_rows_equal(lhs2, enc_0_z2, nnl, batch, 1, d_acc, s.st));
  s.compute_end();
  if (cudaMemcpyAsync(accept, d_acc, (size_t)batch, cudaMemcpyDeviceToHost, s.st) != cudaSuccess) s.bad = true;
  if (cudaMemcpyAsync(fault, d_fault, (size_t)batch, cudaMemcpyDeviceToHost, s.st) != cudaSuccess) s.bad = true;
  rc = s.finish("zkp_mul_verify");
  if (rc == ZKP_OK)
    for (int i = 0; i < batch; ++i)
      if (fault[i]) accept[i] = 0;
  return rc;
}

// ----------------------------------------------------------------- VerlinProof
namespace {
// gen_phi (verlin_proof.rs:138-165): c^y * c'^y' * Enc(y'', r_y) mod nn.  Only the product is ever used (it is phi_a in the
// proof, and the left side of the verifier's comparison), so the three powers are ONE simultaneous exponentiation job:
// one squaring chain for all three bases, the factor (1 + y'' n) as the job's final multiplier.
uint32_t* gen_phi(Sig& s, const uint32_t* cc, const uint32_t* cp, const uint32_t* y, const uint32_t* yp, const uint32_t* ydp,
                  int y_limbs, const uint32_t* r_y, int r_limbs) {
  PowBatch pb(s);
  uint32_t* phi = pb.product({PowTerm{cc, s.nnl, y, y_limbs}, PowTerm{cp, s.nnl, yp, y_limbs}, PowTerm{r_y, r_limbs, nullptr, 0}}, ydp, y_limbs);
  pb.run();
  s.join();
  return phi;
}
}  // namespace

int zkp_verlin_prove(zkp_ctx* c, int batch, int z_limbs, const uint32_t* x, const uint32_t* x_prime, const uint32_t* x_dp,
                     const uint32_t* r_x, const uint32_t* cc, const uint32_t* c_prime, const uint32_t* phi_x,
                     const uint32_t* a, const uint32_t* a_prime, const uint32_t* a_dp, const uint32_t* r_a, uint32_t* phi_a,
                     uint32_t* z, uint32_t* z_prime, uint32_t* z_dp, uint32_t* r_z) {
  if (!c || !x || !x_prime || !x_dp || !r_x || !cc || !c_prime || !phi_x || !a || !a_prime || !a_dp || !r_a || !phi_a || !z ||
      !z_prime || !z_dp || !r_z)
    return c ? fail(c, ZKP_E_ARG, "null buffer") : ZKP_E_ARG;
  Sig s(c, batch);
  int rc = sigma_begin(c, batch, z_limbs, 12, 4 * (8 * (size_t)s.nl + 3 * (size_t)z_limbs) + 128, s, 3);
  if (rc) return rc;
  const int nl = s.nl, nnl = s.nnl;
  uint32_t *d_x = s.up(x, nl), *d_xp = s.up(x_prime, nl), *d_xdp = s.up(x_dp, nl), *d_rx = s.up(r_x, nl);
  uint32_t *d_c = s.up(cc, nnl), *d_cp = s.up(c_prime, nnl), *d_phix = s.up(phi_x, nnl);
  uint32_t *d_a = s.up(a, nl), *d_ap = s.up(a_prime, nl), *d_adp = s.up(a_dp, nl), *d_ra = s.up(r_a, nl);
  uint8_t* d_fault = s.ar.get<uint8_t>((size_t)batch);
  cudaMemsetAsync(d_fault, 0, (size_t)batch, s.st);
  uint32_t* d_phia = gen_phi(s, d_c, d_cp, d_a, d_ap, d_adp, nl, d_ra, nl);    // :70-78
  uint32_t* e = s.challenge({d_c, d_cp, d_phix, d_phia});                      // :80-86
  uint32_t *d_z = s.rows(z_limbs), *d_zp = s.rows(z_limbs), *d_zdp = s.rows(z_limbs);
  if (!s.bad) s.ck(launch_muladd(d_a, nl, d_x, nl, e, 8, batch, d_z, z_limbs, d_fault, s.st));       // z = x*e + a      :87
  if (!s.bad) s.ck(launch_muladd(d_ap, nl, d_xp, nl, e, 8, batch, d_zp, z_limbs, d_fault, s.st));    // :88
  if (!s.bad) s.ck(launch_muladd(d_adp, nl, d_xdp, nl, e, 8, batch, d_zdp, z_limbs, d_fault, s.st)); // :89
  uint32_t* r_x_e = s.powm(d_rx, nl, e, 8);                                    // :90
  uint32_t* d_rz = s.mulm(r_x_e, nnl, d_ra, nl);                               // :91
  s.down(phi_a, d_phia, nnl);
  s.down(z, d_z, z_limbs);
  s.down(z_prime, d_zp, z_limbs);
  s.down(z_dp, d_zdp, z_limbs);
  s.down(r_z, d_rz, nnl);
  return s.finish("zkp_verlin_prove");
}

int zkp_verlin_verify(zkp_ctx* c, int batch, int z_limbs, const uint32_t* cc, const uint32_t* c_prime, const uint32_t* phi_x,
                      const uint32_t* phi_a, const uint32_t* z, const uint32_t* z_prime, const uint32_t* z_dp,
                      const uint32_t* r_z, uint8_t* accept) {
  if (!c || !cc || !c_prime || !phi_x || !phi_a || !z || !z_prime || !z_dp || !r_z || !accept)
    return c ? fail(c, ZKP_E_ARG, "null buffer") : ZKP_E_ARG;
  Sig s(c, batch);
  int rc = sigma_begin(c, batch, z_limbs, 13, 4 * 3 * (size_t)z_limbs + 128, s, 3);
  if (rc) return rc;
  const int nnl = s.nnl;
  uint32_t *d_c = s.up(cc, nnl), *d_cp = s.up(c_prime, nnl), *d_phix = s.up(phi_x, nnl), *d_phia = s.up(phi_a, nnl);
  uint32_t *d_z = s.up(z, z_limbs), *d_zp = s.up(z_prime, z_limbs), *d_zdp = s.up(z_dp, z_limbs), *d_rz = s.up(r_z, nnl);
  uint8_t* d_acc = s.ar.get<uint8_t>((size_t)batch);
  s.compute_begin();
  uint32_t* e = s.challenge({d_c, d_cp, d_phix, d_phia});                      // :102-108
  s.fork(2);
  PowBatch ps(s);
  uint32_t* phi_x_e = ps.powm(d_phix, nnl, e, 8);                              // :109-113
  ps.run();
  uint32_t* rhs = s.mulm(phi_x_e, nnl, d_phia, nnl);                           // :114-118
  s.on_main();
  uint32_t* phi_z = gen_phi(s, d_c, d_cp, d_z, d_zp, d_zdp, z_limbs, d_rz, nnl);  // :120-128 (joins)
  if (!s.bad) s.ck(launch_rows_equal(phi_z, rhs, nnl, batch, 0, d_acc, s.st));
  s.compute_end();
  if (cudaMemcpyAsync(accept, d_acc, (size_t)batch, cudaMemcpyDeviceToHost, s.st) != cudaSuccess) s.bad = true;
  return s.finish("zkp_verlin_verify");
}

}  // extern "C"
