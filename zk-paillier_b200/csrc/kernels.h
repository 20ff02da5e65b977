// Internal launch interface between the C-ABI layer (zkp_api.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace zkp {

constexpr int kCtaThreads = 128;       // 4 warps per CTA
constexpr int kWindowShared = 5;       // sliding window, shared exponent (K1)
constexpr int kTableShared = 16;       // odd powers x^1..x^31
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
cudaError_t launch_modexp_shared(const SharedKey& key, const uint32_t* bases, int base_limbs,
                                 const uint32_t* plain, int plain_limbs, uint32_t* out, int out_limbs, int jobs,
                                 uint32_t* table, int num_sms, cudaStream_t st);

// Montgomery setup for per-instance moduli: r2[i] = R^2 mod mods[i] ([count][S]), n0inv[i].
// mods: [count][mod_limbs].
cudaError_t launch_mont_setup(const uint32_t* mods, int mod_limbs, int S, int count, uint32_t* r2, uint32_t* n0inv,
                              cudaStream_t st);

// K2: out[j] = bases[j]^exps[j / per] mod mods[j / per], fixed 5-bit window.
// bases/out: [jobs][mod_limbs]; mods: [count][mod_limbs]; r2: [count][S];
// exps: [count][exp_limbs]; exp_bits: number of exponent bits scanned (uniform).
cudaError_t launch_modexp_var(const uint32_t* bases, const uint32_t* mods, int mod_limbs, const uint32_t* r2,
                              const uint32_t* n0inv, const uint32_t* exps, int exp_limbs, int exp_bits,
                              int per, uint32_t* out, int jobs, int S, uint32_t* table, int num_sms,
                              cudaStream_t st);

// K3: shared modulus.  mode 0: out[j] = a[j] * b[j / b_per] mod M
//                      mode 1: out[j] = a[j] * R mod M (to Montgomery form; b unused; out rows are S limbs)
// a: [jobs][a_limbs], b: [ceil(jobs/b_per)][b_limbs], out: [jobs][out_limbs].
cudaError_t launch_modmul_shared(const SharedKey& key, int mode, const uint32_t* a, int a_limbs, const uint32_t* b,
                                 int b_limbs, int b_per, uint32_t* out, int out_limbs, int jobs, cudaStream_t st);

// IMAD.WIDE.U32 peak microbenchmark (register-only).  variant 0: independent
// IMAD.WIDE.U32; 1: carry-chained IMAD.WIDE.U32.X rows; 2: plain IMAD (32-bit).
// Returns multiply-adds issued in *ops; caller times it.
cudaError_t launch_imad_peak(int variant, int blocks, int iters, uint32_t* sink, double* ops, cudaStream_t st);

}  // namespace zkp
