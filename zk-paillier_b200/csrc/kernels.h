// Internal launch interface between the C-ABI layer (zkp_api.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

namespace zkp {

constexpr int kCtaThreads = 128;       // 4 warps per CTA
constexpr int kWindowShared = 5;       // sliding window, shared exponent (K1)
constexpr int kTableShared = 16;       // odd powers x^1..x^31
constexpr int kWindow2m = 6;           // sliding window of K1m (32 odd powers + x^2 per encryption in flight)
constexpr int kWindowVar = 5;          // fixed window, per-instance exponent (K2)
constexpr int kTableVar = 32;          // x^0..x^31
constexpr int kMaxSchedSteps = 2048;

// Widths the kernels are instantiated for (limbs of 32 bits).
// A modulus of any other width runs zero-extended on the next size up.
int pick_width(int limbs);             // returns S in {32,64,96,128,192,256} or -1
int group_threads(int S);              // T for width S
int resident_groups(int S, int num_sms);  // groups a full persistent grid holds

// Per-key constants for the shared-modulus kernels (device pointers, S limbs each).
struct SharedKey {
  const uint32_t* mod;    // M
  const uint32_t* r2;     // R^2 mod M, R = 2^(32 S)
  const uint32_t* nR;     // (Paillier n) * R mod M  (Montgomery form of n); Enc epilogue
  const uint32_t* sched;  // sliding-window schedule of the shared exponent
  int nsteps;
  uint32_t n0inv;         // -M^{-1} mod 2^32
  int S;
};

// All row widths below are in 32-bit limbs and must be even; rows narrower than
// the kernel width S are zero-extended on load, and only the low `out_limbs`
// limbs of a result are stored (the caller's width of the modulus).

// K1: out[j] = bases[j]^E mod M for a shared (M, E); optional Paillier epilogue
//     out[j] = (1 + plain[j]*n) * bases[j]^n mod n^2   when plain != nullptr.
// bases: [jobs][base_limbs], plain: [jobs][plain_limbs], out: [jobs][out_limbs].
// base_limbs / plain_limbs must be multiples of 4 (16-byte TMA rows).
// table: scratch of resident_groups(S) * kTableShared * S limbs.
// jobs_dev (optional): device pointer to the actual job count (<= jobs), read by the kernel.
cudaError_t launch_modexp_shared(const SharedKey& key, const uint32_t* bases, int base_limbs,
                                 const uint32_t* plain, int plain_limbs, uint32_t* out, int out_limbs, int jobs,
                                 uint32_t* table, int num_sms, cudaStream_t st, const unsigned* jobs_dev = nullptr);

// K1m (modexp2m.cu): the same encryption by Montgomery arithmetic in two-digit base-n form (half the limb products
// of K1).  S = kernel width of n in limbs (32, 64, 96 or 128); n odd, 1 < n <= 2^(32 S) - 4.
struct Enc2mKey {
  const uint32_t* mod;     // n, S limbs
  const uint32_t* consts;  // K_lo = -W mod n | pair(W^2 mod n^2) | pair(W mod n^2): 5 S limbs (W = 2^(32 S); pair(v) = v mod n | v div n)
  const uint32_t* ops;     // op list (enc2m_ops)
  int nops;
  uint32_t n0inv;          // -n^{-1} mod 2^32
  uint32_t n0inv_hi;       // bits 32..63 of -n^{-1} mod 2^64 (pair rows: two quotient digits per step)
  int S;
};
bool enc2m_supported(const uint32_t* n_host, int S);
void enc2m_host_constants(const uint32_t* n_host, int S, uint32_t* consts /* [5 S] */);
double enc2m_sqr_products();  // limb products of one two-digit squaring in units of S^2: 4, or 3 + (T/2 + 1)/T with the symmetric variant
int enc2m_window();  // sliding-window width of K1m (5, or 6 with ZKP_B200_K1M_WINDOW=6)
std::vector<uint32_t> enc2m_ops(const uint32_t* sched, int nsteps);  // from the schedule of the exponent n recoded with enc2m_window()
int enc2m_resident_groups(int S, int num_sms);
size_t enc2m_table_limbs(int S, int num_sms);
// bases: [jobs][base_limbs] (any value below 2^(32 base_limbs), base_limbs <= 2 S), plain: [jobs][plain_limbs] or null
// (plain_limbs <= 2 S), out: [jobs][out_limbs] (out_limbs <= 2 S); base_limbs / plain_limbs multiples of 4.
cudaError_t launch_enc2m(const Enc2mKey& key, const uint32_t* bases, int base_limbs, const uint32_t* plain, int plain_limbs,
                         uint32_t* out, int out_limbs, int jobs, uint32_t* table, int num_sms, cudaStream_t st,
                         const unsigned* jobs_dev = nullptr);

// K2m (modexp2m.cu): out[j] = bases[j]^exps[j / exp_per] mod n^2 in the same form, fixed 5-bit window.
// bases: [jobs][base_limbs] (base_limbs <= 2 S, even), exps: [ceil(jobs/exp_per)][exp_limbs], out: [jobs][out_limbs].
// table: var2m_table_limbs() limbs of scratch.  key.ops / key.nops are not used.
size_t var2m_table_limbs(int S, int num_sms);
cudaError_t launch_modexp2m_var(const Enc2mKey& key, const uint32_t* bases, int base_limbs, const uint32_t* exps, int exp_limbs,
                                int exp_bits, int exp_per, uint32_t* out, int out_limbs, int jobs, uint32_t* table, int num_sms,
                                cudaStream_t st);

// K2h (modexp2m.cu): every modexp of a batch of sigma-protocol proofs in ONE launch.  A job list is up to kMaxPowSegs
// homogeneous segments; job j of segment s computes
//     out[j] = (1 + plain[j] n) * PROD_k base_k[j]^exp_k[j] mod n^2      (k < nbase <= 3; plain == nullptr: the factor is 1)
// i.e. BigInt::mod_pow / Paillier::mul with a per-job exponent (nbase = 1), Paillier::encrypt_with_chosen_randomness when the
// exponent is the shared n (exp = n on the device, exp_stride = 0), or a whole product of powers by simultaneous (Straus)
// exponentiation - one squaring chain for all bases - where the reference multiplies the powers together anyway:
// gen_phi = c^y c'^y' Enc(y'', r_y) (verlin_proof.rs:138-165) is one job of three bases.
// Segments must be listed longest exponent scan first.
constexpr int kMaxPowSegs = 8;
constexpr int kMaxPowBases = 3;
struct PowSeg {
  const uint32_t* base[kMaxPowBases];   // [jobs][base_limbs[k]], base_limbs <= 2 S, even
  const uint32_t* exp[kMaxPowBases];    // row j at exp[k] + j * exp_stride[k] (limbs); every exponent < 2^exp_bits
  long long exp_stride[kMaxPowBases];
  int base_limbs[kMaxPowBases], exp_limbs[kMaxPowBases];
  int nbase;
  const uint32_t* plain;  // [jobs][plain_limbs] or nullptr; plain_limbs <= 2 S, even
  uint32_t* out;          // [jobs][out_limbs]
  int exp_bits, plain_limbs, jobs, first;  // exp_bits: bits scanned (all bases); first: index of the segment's first job in the launch
};
struct PowJobs {
  PowSeg seg[kMaxPowSegs];
  int nseg = 0, total = 0;
};
// shape: 0 = pick by job count, 1 = the wide-lane layout of K2m (few lanes per job), 2 = the narrow-lane layout (one job
// over twice the lanes, half the limbs per lane: fills the machine at small batches of wide moduli)
// Scratch: jobs2m_scratch_limbs(S, num_sms, total jobs) limbs at `table` (window tables; for a launch that under-fills
// the GPU also the per-job accumulators and phase counters of the phased schedule, see modexp2m.cu).
size_t jobs2m_scratch_limbs(int S, int num_sms, int total_jobs, int max_bases = 1);
// shape 3 = one job per warp (the latency layout, chosen when a launch has at most one long job per SM sub-partition).
// jobs_dev (optional, single-segment launches): device pointer to the actual job count (<= jobs.total).
cudaError_t launch_modexp2m_jobs(const Enc2mKey& key, const PowJobs& jobs, int out_limbs, uint32_t* table, size_t table_limbs,
                                 unsigned* cursor, int num_sms, cudaStream_t st, int shape = 0, const unsigned* jobs_dev = nullptr,
                                 int rows = 0 /* narrow / latency layouts: 0 = default, 1 = single rows, 2 = pair rows */);

// Montgomery setup for per-instance moduli: r2[i] = R^2 mod mods[i] ([count][S]), n0inv[i].
// mods: [count][mod_limbs].
cudaError_t launch_mont_setup(const uint32_t* mods, int mod_limbs, int S, int count, uint32_t* r2, uint32_t* n0inv,
                              cudaStream_t st);

// K2: out[j] = bases[j]^exps[j / exp_per] mod mods[j / mod_per], fixed 5-bit window.
// bases/out: [jobs][mod_limbs]; mods: [ceil(jobs/mod_per)][mod_limbs]; r2: [..][S];
// exps: [ceil(jobs/exp_per)][exp_limbs]; exp_bits: number of exponent bits scanned (uniform).
cudaError_t launch_modexp_var(const uint32_t* bases, const uint32_t* mods, int mod_limbs, const uint32_t* r2,
                              const uint32_t* n0inv, const uint32_t* exps, int exp_limbs, int exp_bits,
                              int exp_per, int mod_per, uint32_t* out, int jobs, int S, uint32_t* table,
                              int num_sms, cudaStream_t st, int base_limbs = 0 /* 0: bases are mod_limbs wide */);

// K3: shared modulus.  mode 0: out[j] = a[j] * b[j / b_per] mod M
//                      mode 1: out[j] = a[j] * R mod M (to Montgomery form; b unused; out rows are S limbs)
// a: [jobs][a_limbs], b: [ceil(jobs/b_per)][b_limbs], out: [jobs][out_limbs].
cudaError_t launch_modmul_shared(const SharedKey& key, int mode, const uint32_t* a, int a_limbs, const uint32_t* b,
                                 int b_limbs, int b_per, uint32_t* out, int out_limbs, int jobs, cudaStream_t st);

// K3 with a per-job operand selector (RangeProofNi verify, range_proof.rs:321-327):
// out[t] = (sel[t] == 2 ? a1[t] : a0[t]) * b[t / b_per] mod M; rows with sel[t] == 0 are not stored.
cudaError_t launch_modmul_select(const SharedKey& key, const uint8_t* sel, const uint32_t* a0, const uint32_t* a1,
                                 int a_limbs, const uint32_t* b, int b_limbs, int b_per, uint32_t* out, int out_limbs,
                                 int jobs, cudaStream_t st);

// ---- K4: Fiat-Shamir transcript hash (utils.rs:9-22) -------------------------------------------
constexpr int kShaThreads = 64;
struct ShaSeg {
  const uint32_t* base;    // items of proof b start at base + b * batch_stride
  long long batch_stride;  // in limbs (0 = the same items for every proof, e.g. n)
  int count;               // items per proof in this segment
  int limbs;               // limbs per item
};
struct ShaSegs {
  ShaSeg seg[8];
  int nseg;
};
cudaError_t launch_sha256_transcript(const ShaSegs& segs, int batch, uint8_t* digest, cudaStream_t st);

// ---- K5: RangeProofNi glue (rangeproof.cu) -----------------------------------------------------
constexpr uint8_t ZKP_RP_OPEN_ = 0, ZKP_RP_MASK1_ = 1, ZKP_RP_MASK2_ = 2;

struct RpProveArgs {
  int batch, ef, wl, nl;
  const uint32_t* range;   // [batch][wl]
  const uint32_t* x;       // [batch][wl]
  const uint32_t* w;       // [2][batch*ef][wl]  w1' | w2'
  const uint32_t* rr;      // [2][batch*ef][nl]  r1 | r2
  const uint32_t* rmul;    // [2][batch*ef][nl]  r*r1 mod n | r*r2 mod n
  const uint8_t* digest;   // [batch][32]
  const uint8_t* chal;     // interactive proof: raw ChallengeBits bytes [batch][chal_bytes] instead of the digest (or null)
  int chal_bytes;
  uint8_t* kind;           // [batch*ef]
  uint32_t* resp_w;        // [batch*ef][2][wl]
  uint32_t* resp_r;        // [batch*ef][2][nl]
  uint8_t* fault;          // [batch]
};
struct RpVerifyArgs {
  int batch, ef, wl, nl;
  const uint32_t* range;    // [batch][wl]
  const uint32_t* c;        // [2][batch*ef][2nl]  c1 | c2
  const uint8_t* kind;      // [batch*ef]
  const uint32_t* resp_w;   // [batch*ef][2][wl]
  const uint32_t* resp_r;   // [batch*ef][2][nl]
  const uint8_t* digest;    // [batch][32]
  const uint8_t* chal;      // interactive proof: raw ChallengeBits bytes [batch][chal_bytes] (or null)
  int chal_bytes;
  const uint32_t* cmul;     // [batch*ef][2nl]     c_j * cipher_x mod n^2 (Mask rows)
  uint32_t* jobs_base;      // [2*batch*ef][nl]
  uint32_t* jobs_plain;     // [2*batch*ef][wl]
  uint32_t* jobs_out;       // [2*batch*ef][2nl]
  uint32_t* tag;            // [2*batch*ef]  (t << 1) | which
  unsigned* count;          // number of jobs
  uint8_t* sel;             // [batch*ef]
  uint8_t* ok;              // [batch*ef]
  uint8_t* fault;           // [batch]
};
cudaError_t launch_rp_prep(const uint32_t* range, const uint32_t* w1in, uint32_t* w, const uint8_t* swap, int batch,
                           int ef, int wl, uint8_t* fault, cudaStream_t st);
cudaError_t launch_rp_respond(const RpProveArgs& a, cudaStream_t st);
cudaError_t launch_rp_plan(const RpVerifyArgs& a, cudaStream_t st);
cudaError_t launch_rp_bits(const RpVerifyArgs& a, cudaStream_t st);
cudaError_t launch_rp_check(const RpVerifyArgs& a, cudaStream_t st);
cudaError_t launch_rp_accept(const uint8_t* ok, const uint8_t* fault, int batch, int ef, uint8_t* accept,
                             cudaStream_t st);

// ---- NiCorrectKeyProof glue (correctkey.cu) ----------------------------------------------------
constexpr int kCkM2 = 11;      // correct_key_ni.rs:29
constexpr int kCkAlpha = 6370; // primes below alpha make up P (correct_key_ni.rs:25-26)
// mask[b][i][ml] = mask_generation(bit_length(n_b), H(n_b || H(salt) || i)), ml = nl + 8
cudaError_t launch_ck_rho(const uint32_t* n, int nl, const uint8_t* salt, int salt_len, int batch, uint32_t* mask, int ml,
                          cudaStream_t st);
// rho[b][i][nl] = mask[b][i] % n_b   (r2: [batch][S], n0inv: [batch] from launch_mont_setup)
cudaError_t launch_ck_reduce(const uint32_t* mask, int ml, const uint32_t* mods, int nl, const uint32_t* r2,
                             const uint32_t* n0inv, int S, int batch, uint32_t* rho, cudaStream_t st);
// accept[b] = (rho[b] == derived[b]) && no prime in primes[] divides n_b
cudaError_t launch_ck_check(const uint32_t* n, int nl, const uint32_t* rho, const uint32_t* derived,
                            const uint16_t* primes, int nprimes, int batch, uint8_t* accept, cudaStream_t st);

// ---- sigma-protocol helpers (sigma.cu), one thread per proof ------------------------------------
cudaError_t launch_digest_to_limbs(const uint8_t* digest, int batch, uint32_t* out /* [batch][8] */, cudaStream_t st);
// out[b] = a[b] + x[b] * e[b] as plain integers (a may be null); sets fault[b] on overflow of out_limbs
cudaError_t launch_muladd(const uint32_t* a, int a_limbs, const uint32_t* x, int x_limbs, const uint32_t* e, int e_limbs,
                          int batch, uint32_t* out, int out_limbs, uint8_t* fault, cudaStream_t st);
// out[b] = (a[b] + c[b]) mod m, a and c already below m (m: one shared modulus of `limbs` limbs)
cudaError_t launch_modadd(const uint32_t* a, const uint32_t* c, const uint32_t* m, int limbs, int batch, uint32_t* out,
                          cudaStream_t st);
// accept[b] = (and_in ? accept[b] : 1) && x[b] == y[b]
cudaError_t launch_rows_equal(const uint32_t* x, const uint32_t* y, int limbs, int batch, int and_in, uint8_t* accept,
                              cudaStream_t st);
// out[b] = v[b]^-1 mod m (m odd, shared); fault[b] = 1 when not invertible.  scratch: [batch][4*limbs]
// m_stride = 0: one shared modulus; otherwise row b uses m + b * m_stride (per-statement moduli; gcd(v, m) == 1 checks).
cudaError_t launch_modinv(const uint32_t* v, const uint32_t* m, int limbs, int batch, uint32_t* scratch, uint32_t* out,
                          uint8_t* fault, cudaStream_t st, long long m_stride = 0);

// ---- helpers of the remaining proofs (more.cu): CompositeDLogProof, CorrectMessageProof ----------
// out[j] = a[j] * b[j] mod mods[j / mod_per] (rows of `limbs` limbs; r2 / n0inv from launch_mont_setup, S = kernel width)
cudaError_t launch_modmul_var(const uint32_t* a, const uint32_t* b, int limbs, const uint32_t* mods, const uint32_t* r2,
                              const uint32_t* n0inv, int mod_per, int S, int jobs, uint32_t* out, cudaStream_t st);
// fault[b] = 1 unless N_b > 2^bits
cudaError_t launch_gt_pow2(const uint32_t* n, int limbs, int bits, int batch, uint8_t* fault, cudaStream_t st);
// out[t] = (1 + m_t n)^-1 mod n^2 = 1 + ((n - m_t) mod n) n, m_t < n: [rows][nl] -> [rows][2 nl]
cudaError_t launch_gm_inv(const uint32_t* m, const uint32_t* n, int nl, int rows, uint32_t* out, cudaStream_t st);
cudaError_t launch_cm_layout(const uint32_t* valid, const uint32_t* msg, int ml, const uint32_t* e_rand, int el,
                             const uint32_t* z_rand, const uint32_t* w, int nl, int batch, int M, uint8_t* match, uint32_t* esel,
                             uint32_t* zsel, uint32_t* esum, uint8_t* fault, cudaStream_t st);
cudaError_t launch_sub_pow2(const uint32_t* a, const uint32_t* c, int limbs, int batch, uint32_t* out, cudaStream_t st);
cudaError_t launch_cm_finish(const uint8_t* match, const uint32_t* e, int el, const uint32_t* z, int nl, int batch, int M,
                             uint32_t* esel, uint32_t* zsel, cudaStream_t st);
cudaError_t launch_sum_pow2(const uint32_t* e, int el, int ol, int batch, int M, uint32_t* esum, cudaStream_t st);
// out[b] = AND (mode 0) / OR (mode 1) of rows[b][0..M); merge: combined with the value already in out[b]
cudaError_t launch_rows_reduce(const uint8_t* rows, int batch, int M, int mode, int merge, uint8_t* out, cudaStream_t st);
cudaError_t launch_rows_differ_fault(const uint32_t* x, const uint32_t* y, int limbs, int batch, uint8_t* fault, cudaStream_t st);

// IMAD.WIDE.U32 peak microbenchmark (register-only).  variant 0: independent
// IMAD.WIDE.U32; 1: carry-chained IMAD.WIDE.U32.X rows; 2: plain IMAD (32-bit).
// Returns multiply-adds issued in *ops; caller times it.
cudaError_t launch_imad_peak(int variant, int blocks, int iters, uint32_t* sink, double* ops, cudaStream_t st);

}  // namespace zkp
