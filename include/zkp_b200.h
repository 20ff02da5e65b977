/*
 * zkp_b200.h - C ABI of the B200-native batched Paillier zero-knowledge engine.
 *
 * This is the drop-in boundary for the hot path of ZenGo-X/zk-paillier: the
 * security-parameter loops of big-integer modular exponentiation inside
 * zkproofs::RangeProofNi::{prove,verify} and NiCorrectKeyProof::verify (plus
 * the modexps of ZeroProof / CiphertextProof / MulProof / VerlinProof).  The
 * reference has no FFI of its own: its seam is the Rust call surface between the
 * proof protocols (src/zkproofs/) and curv-kzen / kzen-paillier (GMP).  Each
 * entry point below names the reference code it replaces; the Rust `extern "C"`
 * block a maintainer would add is in INTEGRATION.md and rust/src/ffi.rs.
 *
 * Conventions
 *  - Integers cross as little-endian arrays of uint32_t limbs, fixed width per
 *    call, zero padded; batches are dense row-major.  (On little-endian hosts the
 *    same bytes are little-endian uint64_t limbs when the limb count is even.)
 *    Every limb count passed in must be a multiple of 4 (16-byte rows: rows are
 *    staged with 1-D TMA bulk copies).
 *  - All buffers are HOST memory owned by the caller; nothing is retained after
 *    the call returns.  Device memory lives in the opaque zkp_ctx (one per GPU).
 *    The *_stage / *_run / *_fetch triples split one call into host->device
 *    copy, kernels only, device->host copy (for measurement and pipelining).
 *  - Every function returns 0 on success, a negative ZKP_E_* code otherwise;
 *    zkp_last_error() gives text.  Nothing throws across the boundary.  There is
 *    NO CPU fallback: without a CUDA device zkp_ctx_create fails.
 *  - Soundness failures are reported per item in accept[] (1 = Ok(()), 0 =
 *    Err(IncorrectProof)).  Inputs on which the reference would PANIC instead
 *    (response/bit-vector shorter than error_factor, non-invertible value in
 *    MulProof) set fault[] = 1 so the shim can re-raise.
 *  - Randomness is always an input; the library never samples.
 *  - A context is single-threaded; use one context per thread / per GPU.
 */
#ifndef ZKP_B200_H
#define ZKP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZKP_OK 0
#define ZKP_E_ARG (-1)     /* bad argument (width, null, range) */
#define ZKP_E_CUDA (-2)    /* CUDA runtime error, see zkp_last_error */
#define ZKP_E_STATE (-3)   /* call order (no key set, nothing staged) */
#define ZKP_E_NOMEM (-4)

#define ZKP_RP_OPEN 0      /* Response::Open  {w1,r1,w2,r2}            (range_proof.rs:54-66) */
#define ZKP_RP_MASK1 1     /* Response::Mask  {j:1, masked_x, masked_r} (range_proof.rs:68-77) */
#define ZKP_RP_MASK2 2     /* Response::Mask  {j:2, ...} */

#define ZKP_CK_M2 11       /* correct_key_ni.rs:29 */

typedef struct zkp_ctx zkp_ctx;

/* ---- context ------------------------------------------------------------ */
/* stream: a cudaStream_t to launch on (e.g. torch's current stream), or NULL
 * for a private stream. */
int zkp_ctx_create(int device, void* stream, zkp_ctx** out);
void zkp_ctx_destroy(zkp_ctx* ctx);
const char* zkp_last_error(const zkp_ctx* ctx);
int zkp_version(void);
int zkp_sm_count(const zkp_ctx* ctx);
int zkp_sync(zkp_ctx* ctx);

/* Per-kernel device-time accounting (CUDA events on the launch stream).
 * kernel ids: 0 = modexp_shared (K1 / K1m), 1 = modexp_var (K2 / K2m / K2h), 2 = modmul (K3),
 * 3 = sha256_transcript (K4), 4 = everything else, 5 = the device span of whole zkp_mul_verify /
 * zkp_verlin_verify calls (first kernel to last kernel; the host<->device copies excluded). */
int zkp_profile_enable(zkp_ctx* ctx, int on);
int zkp_profile_reset(zkp_ctx* ctx);
int zkp_profile_get(zkp_ctx* ctx, int kernel, double* ms_total, long long* launches, double* units);

/* ---- key ---------------------------------------------------------------- */
/* Paillier EncryptionKey{n, nn} (kzen-paillier): derives nn = n^2, Montgomery
 * constants for nn and n, and the sliding-window schedule of the exponent n. */
int zkp_set_key(zkp_ctx* ctx, const uint32_t* n, int n_limbs);
/* Generic shared (modulus, exponent) for zkp_modexp_shared. */
int zkp_set_modulus(zkp_ctx* ctx, const uint32_t* mod, int mod_limbs, const uint32_t* exp, int exp_limbs);
int zkp_nn_limbs(const zkp_ctx* ctx); /* 2*n_limbs after zkp_set_key */

/* ---- K1: BigInt::mod_pow(base, E, M) with one (M,E) per call ---------------
 * out[j] = bases[j]^E mod M.  bases: [batch][base_limbs] (< 2^(32*mod_limbs)),
 * out: [batch][mod_limbs].  After zkp_set_key: M = nn, E = n. */
int zkp_modexp_shared(zkp_ctx* ctx, const uint32_t* bases, int base_limbs, int batch, uint32_t* out);

/* Paillier::encrypt_with_chosen_randomness(ek, m, r) = (1 + m*n) * r^n mod nn
 * (kzen-paillier; called at range_proof.rs:165,179,280,286,330).
 * m: [batch][m_limbs], r: [batch][r_limbs], out: [batch][nn_limbs]. */
int zkp_paillier_enc(zkp_ctx* ctx, const uint32_t* m, int m_limbs, const uint32_t* r, int r_limbs, int batch,
                     uint32_t* out);

/* ---- K2: BigInt::mod_pow with per-instance modulus / exponent --------------
 * out[j] = bases[j]^exps[j/exp_per] mod mods[j/mod_per]
 * (correct_key_ni.rs:90-93 with exp_per = mod_per = 11; the Paillier::mul / mod_pow
 * sites of the sigma protocols -- zero_enc_proof.rs:60,81, correct_ciphertext.rs:60,81,
 * multiplication_proof.rs:91-94,133-138, verlin_proof.rs:89,109,147,152 -- with
 * exp_per = 1 and mod_per = batch for one shared key).
 * bases, out: [batch][mod_limbs]; mods: [ceil(batch/mod_per)][mod_limbs] (odd);
 * exps: [ceil(batch/exp_per)][exp_limbs]; exp_bits: bits scanned (every exponent < 2^exp_bits). */
int zkp_modexp_var(zkp_ctx* ctx, const uint32_t* bases, const uint32_t* exps, int exp_limbs, int exp_bits, int exp_per,
                   const uint32_t* mods, int mod_limbs, int mod_per, int batch, uint32_t* out);

/* ---- K3: BigInt::mod_mul / Paillier::add under the key ----------------------
 * out[j] = a[j] * b[j/b_per] mod (which ? nn : n).  Widths = that modulus. */
int zkp_modmul(zkp_ctx* ctx, int which_nn, const uint32_t* a, const uint32_t* b, int b_per, int batch,
               uint32_t* out);

/* ---- K4: compute_digest (utils.rs:9-22) -----------------------------------
 * digest[b] = SHA-256( to_bytes(items[b][0]) || ... || to_bytes(items[b][count-1]) )
 * where to_bytes is the minimal-length big-endian magnitude and zero -> 0x00
 * (curv-kzen BigInt::to_bytes over GMP).  items: [batch][count][limbs]. */
int zkp_sha256_transcript(zkp_ctx* ctx, const uint32_t* items, int limbs, int count, int batch,
                          uint8_t* digest /* [batch][32] */);

/* ---- RangeProofNi (range_proof_ni.rs:47-107, range_proof.rs:128-355) --------
 * All proofs of a batch are under the key set by zkp_set_key.
 * ef = error_factor (SECURITY_PARAMETER = 128, range_proof_ni.rs:23).
 * w_limbs: row width of range / x / w / masked_x values (>= bits(10000*range)/32+1
 * is always enough), n_limbs / nn_limbs: widths of n / n^2.
 *
 * prove inputs  range[b][w], x[b][w], r[b][n]             secret_x, secret_r
 *               w1[b][ef][w]   samples of [third, 2*third) (range_proof.rs:136-139)
 *               swap[b][ef]    the coin flips            (range_proof.rs:144-149)
 *               r1[b][ef][n], r2[b][ef][n]                (range_proof.rs:151-159)
 * prove outputs c1[b][ef][nn], c2[b][ef][nn]              EncryptedPairs
 *               digest[b][32]                             compute_digest(n, c1.., c2..)
 *               kind[b][ef]                               ZKP_RP_*
 *               resp_w[b][ef][2][w]  Open: (w1,w2)  Mask: (masked_x, 0)
 *               resp_r[b][ef][2][n]  Open: (r1,r2)  Mask: (masked_r, 0)
 * Any output pointer may be NULL (not fetched). */
int zkp_rangeproof_ni_prove(zkp_ctx* ctx, int batch, int ef, int w_limbs, const uint32_t* range, const uint32_t* x,
                            const uint32_t* r, const uint32_t* w1, const uint8_t* swap, const uint32_t* r1,
                            const uint32_t* r2, uint32_t* c1, uint32_t* c2, uint8_t* digest, uint8_t* kind,
                            uint32_t* resp_w, uint32_t* resp_r);
int zkp_rp_prove_stage(zkp_ctx* ctx, int batch, int ef, int w_limbs, const uint32_t* range, const uint32_t* x,
                       const uint32_t* r, const uint32_t* w1, const uint8_t* swap, const uint32_t* r1,
                       const uint32_t* r2);
int zkp_rp_prove_run(zkp_ctx* ctx);
/* The same run in the two phases of the INTERACTIVE RangeProof (range_proof.rs): the encrypted pairs first
 * (generate_encrypted_pairs, :128-193; fetch c1/c2 with zkp_rp_prove_fetch and NULL response pointers), then,
 * once the verifier has opened its commitment, the responses to ITS ChallengeBits bytes (generate_proof,
 * :210-252): challenge[b][chal_bytes], bit i = (byte[i/8] >> (7 - i%8)) & 1 as BitVec::from_bytes reads them.
 * challenge == NULL uses the Fiat-Shamir bits of the transcript hash (= zkp_rp_prove_run). */
int zkp_rp_prove_run_pairs(zkp_ctx* ctx);
int zkp_rp_prove_run_responses(zkp_ctx* ctx, const uint8_t* challenge, int chal_bytes);
int zkp_rp_prove_fetch(zkp_ctx* ctx, uint32_t* c1, uint32_t* c2, uint8_t* digest, uint8_t* kind, uint32_t* resp_w,
                       uint32_t* resp_r);

/* verify inputs: range[b][w], cipher_x[b][nn] and the proof arrays as produced
 * by prove.  accept[b] = 1 iff RangeProofNi::verify returns Ok(()); fault[b] = 1
 * where the reference would panic (kind byte out of range).  digest (optional
 * out) receives the recomputed challenge hash. */
int zkp_rangeproof_ni_verify(zkp_ctx* ctx, int batch, int ef, int w_limbs, const uint32_t* range,
                             const uint32_t* cipher_x, const uint32_t* c1, const uint32_t* c2, const uint8_t* kind,
                             const uint32_t* resp_w, const uint32_t* resp_r, uint8_t* accept, uint8_t* fault,
                             uint8_t* digest);
int zkp_rp_verify_stage(zkp_ctx* ctx, int batch, int ef, int w_limbs, const uint32_t* range,
                        const uint32_t* cipher_x, const uint32_t* c1, const uint32_t* c2, const uint8_t* kind,
                        const uint32_t* resp_w, const uint32_t* resp_r);
/* Chain on the device: verify the batch most recently produced by
 * zkp_rp_prove_run (no host round trip); only cipher_x comes from the host. */
int zkp_rp_verify_stage_from_prove(zkp_ctx* ctx, const uint32_t* cipher_x);
int zkp_rp_verify_run(zkp_ctx* ctx);
/* RangeProof::verifier_output against the verifier's own ChallengeBits (interactive proof, range_proof.rs:254-355). */
int zkp_rp_verify_run_with_challenge(zkp_ctx* ctx, const uint8_t* challenge, int chal_bytes);
int zkp_rp_verify_fetch(zkp_ctx* ctx, uint8_t* accept, uint8_t* fault, uint8_t* digest);
/* Number of Paillier encryptions the last verify_run performed (ef + #Open per proof: the plan follows the variant of each
 * response, so that the encryptions run beside the transcript hash; a response whose variant contradicts its challenge bit is
 * encrypted too - the reference skips it, the verdict is the same). */
long long zkp_rp_verify_enc_count(zkp_ctx* ctx);

/* ---- NiCorrectKeyProof::verify (correct_key_ni.rs:73-100) -------------------
 * n[b][n_limbs] (one modulus per proof), sigma[b][11][n_limbs], salt shared.
 * accept[b] = 1 iff all 11 sigma_i^N mod N equal the derived rho_i and
 * gcd(primorial(6370), N) == 1.  rho (optional out): [b][11][n_limbs]. */
int zkp_correct_key_ni_verify(zkp_ctx* ctx, int batch, int n_limbs, const uint32_t* n, const uint32_t* sigma,
                              const uint8_t* salt, int salt_len, uint8_t* accept, uint32_t* rho);
/* The prover's half of the same derivation (correct_key_ni.rs:44-63): rho[b][11][n_limbs] =
 * mask_generation(|N_b|, H(N_b, H(salt), i)) % N_b.  NiCorrectKeyProof::proof then takes N-th roots of
 * these with the decryption key (extract_nroot: two half-width zkp_modexp_var calls + CRT on the host). */
int zkp_correct_key_ni_rho(zkp_ctx* ctx, int batch, int n_limbs, const uint32_t* n, const uint8_t* salt, int salt_len,
                           uint32_t* rho);
int zkp_ck_verify_stage(zkp_ctx* ctx, int batch, int n_limbs, const uint32_t* n, const uint32_t* sigma,
                        const uint8_t* salt, int salt_len);
int zkp_ck_verify_run(zkp_ctx* ctx);
int zkp_ck_verify_fetch(zkp_ctx* ctx, uint8_t* accept, uint32_t* rho);

/* ---- sigma protocols under the key of zkp_set_key -----------------------------
 * One row per proof; n-wide rows are [batch][n_limbs], ciphertext-wide rows are
 * [batch][nn_limbs].  The Fiat-Shamir challenge e = compute_digest(n, ...) is
 * computed on the device.  Randomness (r_prime, x_prime, d, r_d, a.., r_a) is an
 * input.  z_limbs: row width of the UNREDUCED responses x' + x*e (correct_ciphertext.rs:59,
 * verlin_proof.rs:87-89), a multiple of 4 with n_limbs + 12 <= z_limbs <= nn_limbs.
 * accept[b] = 1 iff verify returns Ok(()).
 *
 * ZeroProof (zero_enc_proof.rs:44-94): witness r, statement c; proof (z, a). */
int zkp_zero_prove(zkp_ctx* ctx, int batch, const uint32_t* r, const uint32_t* c, const uint32_t* r_prime, uint32_t* z,
                   uint32_t* a);
int zkp_zero_verify(zkp_ctx* ctx, int batch, const uint32_t* c, const uint32_t* z, const uint32_t* a, uint8_t* accept);
/* CiphertextProof (correct_ciphertext.rs:42-98): witness (x, r), statement c; proof (z1, z2, c_prime). */
int zkp_ciphertext_prove(zkp_ctx* ctx, int batch, int z_limbs, const uint32_t* x, const uint32_t* r, const uint32_t* c,
                         const uint32_t* x_prime, const uint32_t* r_prime, uint32_t* z1, uint32_t* z2, uint32_t* c_prime);
int zkp_ciphertext_verify(zkp_ctx* ctx, int batch, int z_limbs, const uint32_t* c, const uint32_t* z1, const uint32_t* z2,
                          const uint32_t* c_prime, uint8_t* accept);
/* MulProof (multiplication_proof.rs:60-145): witness (a, b, r_a, r_b, r_c), statement (e_a, e_b, e_c),
 * randomness (d, r_d); proof (f, z1, z2, e_d, e_db).  fault[b] = 1 where BigInt::mod_inv(..).unwrap()
 * (:96, :137) would panic (value not invertible mod n^2); such proofs are not accepted. */
int zkp_mul_prove(zkp_ctx* ctx, int batch, const uint32_t* a, const uint32_t* b, const uint32_t* r_a, const uint32_t* r_b,
                  const uint32_t* r_c, const uint32_t* e_a, const uint32_t* e_b, const uint32_t* e_c, const uint32_t* d,
                  const uint32_t* r_d, uint32_t* f, uint32_t* z1, uint32_t* z2, uint32_t* e_d, uint32_t* e_db, uint8_t* fault);
int zkp_mul_verify(zkp_ctx* ctx, int batch, const uint32_t* e_a, const uint32_t* e_b, const uint32_t* e_c, const uint32_t* f,
                   const uint32_t* z1, const uint32_t* z2, const uint32_t* e_d, const uint32_t* e_db, uint8_t* accept,
                   uint8_t* fault);
/* VerlinProof (verlin_proof.rs:60-165): witness (x, x', x'', r_x), statement (c, c', phi_x),
 * randomness (a, a', a'', r_a); proof (phi_a, z, z', z'', r_z). */
int zkp_verlin_prove(zkp_ctx* ctx, int batch, int z_limbs, const uint32_t* x, const uint32_t* x_prime, const uint32_t* x_dp,
                     const uint32_t* r_x, const uint32_t* c, const uint32_t* c_prime, const uint32_t* phi_x,
                     const uint32_t* a, const uint32_t* a_prime, const uint32_t* a_dp, const uint32_t* r_a, uint32_t* phi_a,
                     uint32_t* z, uint32_t* z_prime, uint32_t* z_dp, uint32_t* r_z);
int zkp_verlin_verify(zkp_ctx* ctx, int batch, int z_limbs, const uint32_t* c, const uint32_t* c_prime, const uint32_t* phi_x,
                      const uint32_t* phi_a, const uint32_t* z, const uint32_t* z_prime, const uint32_t* z_dp,
                      const uint32_t* r_z, uint8_t* accept);

/* ---- the remaining public proofs ------------------------------
 * CorrectOpening::verify_opening (correct_opening.rs:17-30): ok[b] = (c[b] == Enc(m[b], r[b])) under the current key.
 * m: [batch][m_limbs], r: [batch][n_limbs], c: [batch][nn_limbs]. */
int zkp_verify_opening(zkp_ctx* ctx, int batch, int m_limbs, const uint32_t* m, const uint32_t* r, const uint32_t* c, uint8_t* ok);
/* CompositeDLogProof (wi_dlog_proof.rs:46-91): one DLogStatement {N, g, ni} per proof (rows of n_limbs limbs, every N odd;
 * no zkp_set_key needed).  prove: secret [batch][secret_limbs], the prover's sample r < 2^(K + K' + S) [batch][r_limbs];
 * out x = g^r mod N [batch][n_limbs], y = r + e * secret (unreduced) [batch][y_limbs], e = H(x, g, N, ni);
 * fault[b] = 1 if y overflows y_limbs.
 * verify: fault[b] = 1 where the reference's asserts fire (N <= 2^128, gcd(g, N) != 1, gcd(ni, N) != 1: panics);
 * accept[b] = (x == g^y * ni^e mod N). */
int zkp_dlog_prove(zkp_ctx* ctx, int batch, int n_limbs, const uint32_t* N, const uint32_t* g, const uint32_t* ni,
                   const uint32_t* secret, int secret_limbs, const uint32_t* r, int r_limbs, int y_limbs, uint32_t* x, uint32_t* y,
                   uint8_t* fault);
int zkp_dlog_verify(zkp_ctx* ctx, int batch, int n_limbs, const uint32_t* N, const uint32_t* g, const uint32_t* ni, const uint32_t* x,
                    const uint32_t* y, int y_limbs, uint8_t* accept, uint8_t* fault);
/* CorrectMessageProof (correct_message.rs:35-162) under the current key: M valid messages per proof.
 * prove: valid [batch][M][m_limbs], msg [batch][m_limbs], randomness r [batch][n_limbs], e_rand [batch][M-1][8] (B = 256 bits),
 * z_rand [batch][M-1][n_limbs], w [batch][n_limbs]; out ciphertext [batch][nn_limbs], e_vec [batch][M][8],
 * z_vec [batch][M][n_limbs], a_vec [batch][M][nn_limbs]; fault[b] = 1 where the reference panics (message not among the
 * valid ones: index past the random vectors; a non-invertible value under unwrap()).
 * verify: fault[b] = 1 where assert_eq!(chal, ei_sum) fires; accept[b] = AND_i (u_i^e_i * a_i == z_i^n mod nn). */
int zkp_correct_message_prove(zkp_ctx* ctx, int batch, int M, int m_limbs, const uint32_t* valid, const uint32_t* msg,
                              const uint32_t* r, const uint32_t* e_rand, const uint32_t* z_rand, const uint32_t* w,
                              uint32_t* ciphertext, uint32_t* e_vec, uint32_t* z_vec, uint32_t* a_vec, uint8_t* fault);
int zkp_correct_message_verify(zkp_ctx* ctx, int batch, int M, int m_limbs, int e_limbs, const uint32_t* ciphertext,
                               const uint32_t* valid, const uint32_t* e_vec, const uint32_t* z_vec, const uint32_t* a_vec,
                               uint8_t* accept, uint8_t* fault);

/* ---- A/B hooks (per context; results are bit-identical whatever the setting) --------------------------------
 * The library reads no environment variables.  (The measured-and-rejected kernel variants of DESIGN.md section 3.7
 * and their ZKP_B200_* environment knobs exist only in the lab build, `make lab` -> libzkp_b200_lab.so.)
 *   ZKP_TUNE_ENC_KERNEL  0 = by launch size (default): K1m, the two-digit Montgomery form, whenever the key qualifies, and
 *                            K2h's one-job-per-warp layout for launches of a few hundred encryptions (one proof's latency);
 *                        1 = K1, Montgomery modulo n^2;   2 = K1m whatever the launch size
 *                        (the parity tests run all three and compare)
 *   ZKP_TUNE_JOBS_SHAPE  lane layout of K2h, the one-launch heterogeneous modexp list of the sigma protocols:
 *                        0 = by job count (default), 1 = wide lanes (as K1m / K2m), 2 = narrow lanes (one job over
 *                        twice the lanes: fills the GPU at a few hundred proofs of 4096-bit n)
 *   ZKP_TUNE_JOBS_ROWS   row form of K2h's narrow-lane and latency layouts: 0 = by layout (pair rows in the one-job-per-warp
 *                        latency layout, single rows in the narrow-lane layout), 1 = single rows (one quotient digit per step,
 *                        as K1m / K2m), 2 = pair rows (two multiplier limbs and a two-limb quotient per step) */
#define ZKP_TUNE_ENC_KERNEL 0
#define ZKP_TUNE_JOBS_SHAPE 1
#define ZKP_TUNE_JOBS_ROWS 2
int zkp_tune(zkp_ctx* ctx, int knob, int value);

/* ---- measurement ----------------------------------------------------------
 * Register-only multiply-add issue-rate microbenchmark (the roofline denominator
 * for the modexp kernels).  variant 0: independent IMAD.WIDE.U32, 1: the
 * carry-chained IMAD.WIDE.U32.X rows the Montgomery loop is made of, 2: 32-bit
 * IMAD.  Returns multiply-adds per second in *mads_per_s. */
int zkp_imad_peak(zkp_ctx* ctx, int variant, double* mads_per_s);
/* Which kernel served Paillier::encrypt_with_chosen_randomness so far on this context: launches of K1m (two-digit
 * Montgomery form, the default) and of K1 (Montgomery modulo n^2: rows wider than n, keys K1m does not take, or
 * zkp_tune(ctx, ZKP_TUNE_ENC_KERNEL, 1)).  Either pointer may be NULL. */
int zkp_enc_kernel_launches(const zkp_ctx* ctx, long long* k1m, long long* k1);
/* IMAD.WIDE.U32 (32x32+64 multiply-adds) one encryption EXECUTES under the current key, counted from the kernels'
 * own op lists: K1m = (4 S^2 per squaring, 5 S^2 per multiplication, S^2 for the final X0 + X1 n, S = limbs of n);
 * K1 = 2 (2S)^2 per Montgomery multiplication modulo n^2.  The roofline's "executed" view; the algorithmic figure
 * of SURVEY.md section 8d (fixed 5-bit window, schoolbook CIOS modulo n^2) is larger than either. */
int zkp_enc_executed_mads(const zkp_ctx* ctx, double* k1m, double* k1);

#ifdef __cplusplus
}
#endif
#endif /* ZKP_B200_H */
