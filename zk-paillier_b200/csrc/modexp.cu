// K1/K2/K3: batched Montgomery modular exponentiation / multiplication kernels (sm_100a).
//
// K1 modexp_shared_kernel : base[j]^E mod M, one (M,E) per launch, sliding window
//    over a host-recoded schedule of the shared exponent; optional Paillier
//    epilogue  c = (1 + m*n) * r^n mod n^2.
//    Replaces Paillier::encrypt_with_chosen_randomness as called from
//    reference src/zkproofs/range_proof.rs:161-187 (prove) and :280-291,330-334
//    (verify); zero_enc_proof.rs:46,73; correct_ciphertext.rs:45,73;
//    multiplication_proof.rs:63,72,118,125; verlin_proof.rs:157.
// K2 modexp_var_kernel : per-instance modulus and exponent, fixed 5-bit window.
//    Replaces BigInt::mod_pow at reference src/zkproofs/correct_key_ni.rs:90-93
//    and the per-proof-exponent mod_pow / Paillier::mul sites of the sigma protocols.
// K3 modmul_shared_kernel : a*b mod M  (range_proof.rs:239,245,325,327 and the
//    Paillier::add / mod_mul sites).
//
// Bases / plaintexts of one CTA's jobs are staged into shared memory with one
// 1-D TMA bulk copy (cp.async.bulk -> UBLKCP) per array and an mbarrier.
#include <stdlib.h>

#include "kernels.h"
#include "mp_coop.cuh"
#include "blockmul.cuh"
#include "tma.cuh"

namespace zkp {

template <int T, int L>
struct Occ {
  // registers/thread target: <=128 up to L=16 (16 warps/SM), <=168 for L=24, <=255 beyond
  static constexpr int kMinBlocks = (L <= 16) ? 4 : (L <= 24 ? 3 : 2);
};

// x += 1 over the whole group
template <int T, int L>
__device__ __forceinline__ void add_one(uint32_t (&x)[L], int lane) {
  using M = Mp<T, L>;
  const int g = lane & (T - 1);
  add_cc(x[0], g == 0 ? 1u : 0u);
#pragma unroll
  for (int j = 1; j < L; ++j) addc_cc(x[j], 0);
  uint32_t co = addc_out();
  uint32_t top;
  uint32_t cin = M::resolve(co != 0, M::all_ones(x), lane, top);
  M::add_small(x, cin);
}

// ------------------------------------------------------------------------ K1
struct SharedParams {
  SharedKey key;
  const uint32_t* bases;
  const uint32_t* plain;
  uint32_t* out;
  uint32_t* table;
  int base_limbs;
  int plain_limbs;
  int out_limbs;
  int jobs;
  const unsigned* jobs_dev;
  int sched_pad;  // schedule entries padded to a multiple of 4
};

template <int T, int L>
__global__ void __launch_bounds__(kCtaThreads, Occ<T, L>::kMinBlocks) modexp_shared_kernel(const SharedParams p) {
  using M = Mp<T, L>;
  constexpr int S = T * L;
  constexpr int G = kCtaThreads / T;  // big integers (jobs) per CTA pass
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* s_sched = reinterpret_cast<uint32_t*>(smem_raw + 16);
  uint32_t* s_bases = s_sched + p.sched_pad;
  uint32_t* s_plain = s_bases + G * p.base_limbs;

  const int lane = threadIdx.x & 31;
  const int g = lane & (T - 1);
  const int grp = threadIdx.x / T;
  const uint32_t n0inv = p.key.n0inv;
  uint32_t n[L];
  M::load(n, p.key.mod + g * L);

  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  uint32_t phase = 0;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, (uint32_t)p.sched_pad * 4u);
    bulk_g2s(s_sched, p.key.sched, (uint32_t)p.sched_pad * 4u, bar);
  }
  mbar_wait(bar, phase);
  phase ^= 1;

  uint32_t* tab = p.table + ((size_t)(blockIdx.x * G + grp) * kTableShared) * S + g * L;
  const int jobs = p.jobs_dev ? min((int)*p.jobs_dev, p.jobs) : p.jobs;
  const int npass = (jobs + G - 1) / G;
  for (int cj = blockIdx.x; cj < npass; cj += gridDim.x) {
    const int job0 = cj * G;
    const int nvalid = min(G, jobs - job0);
    if (threadIdx.x == 0) {
      uint32_t bb = (uint32_t)(nvalid * p.base_limbs) * 4u;
      uint32_t pb = p.plain ? (uint32_t)(nvalid * p.plain_limbs) * 4u : 0u;
      mbar_expect_tx(bar, bb + pb);
      bulk_g2s(s_bases, p.bases + (size_t)job0 * p.base_limbs, bb, bar);
      if (p.plain) bulk_g2s(s_plain, p.plain + (size_t)job0 * p.plain_limbs, pb, bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    const bool valid = grp < nvalid;
    const int src = valid ? grp : 0;

    uint32_t acc[L], y[L];
    M::load_ext(acc, s_bases + src * p.base_limbs, p.base_limbs, g);
    M::load(y, p.key.r2 + g * L);
    M::mont_mul(acc, acc, y, n, n0inv, lane);  // base * R mod M
    M::store(tab, acc);
    M::mont_sqr(y, acc, n, n0inv, lane);       // base^2 * R
#pragma unroll 1
    for (int e = 1; e < kTableShared; ++e) {   // odd powers base^(2e+1)
      M::mont_mul(acc, acc, y, n, n0inv, lane);
      M::store(tab + e * S, acc);
    }
    uint32_t st = s_sched[0];
    M::load(acc, tab + (st & 0xffu) * S);
#pragma unroll 1
    for (int k = 1; k < p.key.nsteps; ++k) {
      st = s_sched[k];
      const uint32_t idx = st & 0xffu;
      const int nsq = (int)(st >> 8);
      if (idx != 0xffu) M::load(y, tab + idx * S);  // prefetch multiplier ahead of the squarings
#pragma unroll 1
      for (int q = 0; q < nsq; ++q) M::mont_sqr(acc, acc, n, n0inv, lane);
      if (idx != 0xffu) M::mont_mul(acc, acc, y, n, n0inv, lane);
    }
    // leave Montgomery form, fusing the Paillier factor (1 + m*n) when asked
    if (p.plain) {
      uint32_t nr[L];
      M::load_ext(y, s_plain + src * p.plain_limbs, p.plain_limbs, g);
      M::load(nr, p.key.nR + g * L);
      M::mont_mul(y, y, nr, n, n0inv, lane);  // m*n, exact (m < n  =>  m*n < n^2)
      add_one<T, L>(y, lane);
    } else {
#pragma unroll
      for (int j = 0; j < L; ++j) y[j] = 0;
      if (g == 0) y[0] = 1;
    }
    M::mont_mul(acc, acc, y, n, n0inv, lane);
    if (valid) M::store_ext(p.out + (size_t)(job0 + grp) * p.out_limbs, acc, p.out_limbs, g);
    __syncthreads();  // everyone is done with the staged inputs
    fence_proxy_async();
  }
}

// ------------------------------------------------------------ Montgomery setup
template <int T, int L>
__global__ void __launch_bounds__(kCtaThreads) mont_setup_kernel(const uint32_t* mods, int mod_limbs, int count,
                                                               uint32_t* r2, uint32_t* n0inv) {
  using M = Mp<T, L>;
  constexpr int S = T * L;
  constexpr int G = kCtaThreads / T;
  const int lane = threadIdx.x & 31;
  const int g = lane & (T - 1);
  const int i = blockIdx.x * G + threadIdx.x / T;
  const bool valid = i < count;
  const int src = valid ? i : 0;
  uint32_t n[L], x[L];
  M::load_ext(n, mods + (size_t)src * mod_limbs, mod_limbs, g);
  uint32_t n0 = __shfl_sync(ZKP_FULL, n[0], 0, T);
  uint32_t inv = n0;  // Newton: inv = n0^-1 mod 2^32 (n0 odd)
#pragma unroll
  for (int k = 0; k < 5; ++k) inv *= 2u - n0 * inv;
#pragma unroll
  for (int j = 0; j < L; ++j) x[j] = 0;
  if (g == 0) x[0] = 1;
#pragma unroll 1
  for (int k = 0; k < 64 * S; ++k) M::mod_double(x, n, lane);  // 2^(64 S) mod n = R^2 mod n
  if (valid) {
    M::store(r2 + (size_t)i * S + g * L, x);
    if (g == 0) n0inv[i] = 0u - inv;
  }
}

// ------------------------------------------------------------------------ K2
struct VarParams {
  const uint32_t* bases;
  const uint32_t* mods;
  const uint32_t* r2;
  const uint32_t* n0inv;
  const uint32_t* exps;
  uint32_t* out;
  uint32_t* table;
  int mod_limbs;
  int base_limbs;
  int exp_limbs;
  int exp_bits;
  int exp_per;
  int mod_per;
  int jobs;
};

__device__ __forceinline__ uint32_t exp_window(const uint32_t* e, int exp_limbs, int bit) {
  int limb = bit >> 5, sh = bit & 31;
  uint64_t v = e[limb];
  if (limb + 1 < exp_limbs) v |= (uint64_t)e[limb + 1] << 32;
  return (uint32_t)(v >> sh) & ((1u << kWindowVar) - 1u);
}

template <int T, int L>
__global__ void __launch_bounds__(kCtaThreads, Occ<T, L>::kMinBlocks) modexp_var_kernel(const VarParams p) {
  using M = Mp<T, L>;
  constexpr int S = T * L;
  constexpr int G = kCtaThreads / T;
  const int lane = threadIdx.x & 31;
  const int g = lane & (T - 1);
  const int grp = threadIdx.x / T;
  uint32_t* tab = p.table + ((size_t)(blockIdx.x * G + grp) * kTableVar) * S + g * L;
  const int npass = (p.jobs + G - 1) / G;
  const int nwin = (p.exp_bits + kWindowVar - 1) / kWindowVar;
  for (int cj = blockIdx.x; cj < npass; cj += gridDim.x) {
    const int job = cj * G + grp;
    const bool valid = job < p.jobs;
    const int src = valid ? job : 0;
    const int mi = src / p.mod_per;
    const uint32_t n0inv = p.n0inv[mi];
    const uint32_t* e = p.exps + (size_t)(src / p.exp_per) * p.exp_limbs;
    uint32_t n[L], acc[L], y[L];
    M::load_ext(n, p.mods + (size_t)mi * p.mod_limbs, p.mod_limbs, g);
    M::load(y, p.r2 + (size_t)mi * S + g * L);
    M::load_ext(acc, p.bases + (size_t)src * p.base_limbs, p.base_limbs, g);
    M::mont_mul(acc, acc, y, n, n0inv, lane);  // x R
    {
      uint32_t one[L];
#pragma unroll
      for (int j = 0; j < L; ++j) one[j] = 0;
      if (g == 0) one[0] = 1;
      M::mont_mul(y, y, one, n, n0inv, lane);  // R mod n  (= x^0 in Montgomery form)
    }
    M::store(tab, y);
    M::store(tab + S, acc);
#pragma unroll
    for (int j = 0; j < L; ++j) y[j] = acc[j];
#pragma unroll 1
    for (int k = 2; k < kTableVar; ++k) {
      M::mont_mul(acc, acc, y, n, n0inv, lane);
      M::store(tab + (size_t)k * S, acc);
    }
    M::load(acc, tab + (size_t)exp_window(e, p.exp_limbs, (nwin - 1) * kWindowVar) * S);
#pragma unroll 1
    for (int w = nwin - 2; w >= 0; --w) {
      M::load(y, tab + (size_t)exp_window(e, p.exp_limbs, w * kWindowVar) * S);
#pragma unroll 1
      for (int q = 0; q < kWindowVar; ++q) M::mont_sqr(acc, acc, n, n0inv, lane);
      M::mont_mul(acc, acc, y, n, n0inv, lane);
    }
#pragma unroll
    for (int j = 0; j < L; ++j) y[j] = 0;
    if (g == 0) y[0] = 1;
    M::mont_mul(acc, acc, y, n, n0inv, lane);
    if (valid) M::store_ext(p.out + (size_t)job * p.mod_limbs, acc, p.mod_limbs, g);
  }
}

// ------------------------------------------------------------------------ K3
struct MulParams {
  SharedKey key;
  const uint32_t* a;
  const uint32_t* b;
  uint32_t* out;
  int a_limbs, b_limbs, out_limbs, b_per, jobs, mode;
};

template <int T, int L>
__global__ void __launch_bounds__(kCtaThreads) modmul_shared_kernel(const MulParams p) {
  using M = Mp<T, L>;
  constexpr int G = kCtaThreads / T;
  const int lane = threadIdx.x & 31;
  const int g = lane & (T - 1);
  const int job = blockIdx.x * G + threadIdx.x / T;
  const bool valid = job < p.jobs;
  const int src = valid ? job : 0;
  uint32_t n[L], x[L], y[L];
  M::load(n, p.key.mod + g * L);
  M::load_ext(x, p.a + (size_t)src * p.a_limbs, p.a_limbs, g);
  M::load(y, p.key.r2 + g * L);
  M::mont_mul(x, x, y, n, p.key.n0inv, lane);  // a R
  if (p.mode == 0) {
    M::load_ext(y, p.b + (size_t)(src / p.b_per) * p.b_limbs, p.b_limbs, g);
    M::mont_mul(x, x, y, n, p.key.n0inv, lane);  // a b
  }
  if (valid) M::store_ext(p.out + (size_t)job * p.out_limbs, x, p.out_limbs, g);
}

struct MulSelParams {
  SharedKey key;
  const uint8_t* sel;
  const uint32_t* a0;
  const uint32_t* a1;
  const uint32_t* b;
  uint32_t* out;
  int a_limbs, b_limbs, out_limbs, b_per, jobs;
};

template <int T, int L>
__global__ void __launch_bounds__(kCtaThreads) modmul_select_kernel(const MulSelParams p) {
  using M = Mp<T, L>;
  constexpr int G = kCtaThreads / T;
  const int lane = threadIdx.x & 31;
  const int g = lane & (T - 1);
  const int job = blockIdx.x * G + threadIdx.x / T;
  const bool valid = job < p.jobs;
  const int src = valid ? job : 0;
  const uint8_t sel = p.sel[src];
  // a warp is skipped only as a whole: its groups share shuffles
  if (__ballot_sync(ZKP_FULL, valid && sel != 0) == 0u) return;
  uint32_t n[L], x[L], y[L];
  M::load(n, p.key.mod + g * L);
  M::load_ext(x, (sel == 2 ? p.a1 : p.a0) + (size_t)src * p.a_limbs, p.a_limbs, g);
  M::load(y, p.key.r2 + g * L);
  M::mont_mul(x, x, y, n, p.key.n0inv, lane);
  M::load_ext(y, p.b + (size_t)(src / p.b_per) * p.b_limbs, p.b_limbs, g);
  M::mont_mul(x, x, y, n, p.key.n0inv, lane);
  if (valid && sel != 0) M::store_ext(p.out + (size_t)job * p.out_limbs, x, p.out_limbs, g);
}

// ----------------------------------------------------------------- dispatch
int pick_width(int limbs) {
  static const int w[] = {32, 64, 96, 128, 192, 256};
  for (int s : w)
    if (limbs <= s) return s;
  return -1;
}
int group_threads(int S) { return S <= 32 ? 4 : (S <= 128 ? 8 : 16); }

template <int T, int L>
static int blocks_per_sm_shared() {
  return Occ<T, L>::kMinBlocks;
}
// Persistent CTAs per SM of K2 (modexp_var_kernel).  Its register count would allow more than the 4 of K1 at the narrow
// shapes (72 registers at L = 8, 100 at L = 12), but measured on B200 at 3072-bit N (NiCorrectKeyProof verify, batch 4096):
// 4 -> 9 786/s, 5 -> 9 709/s, 6 -> 9 900/s: within noise, the multiplier pipe is already 81 % busy at 16 warps per SM.
// ZKP_B200_K2_CTAS overrides (tuning).
int k2_ctas_per_sm(int L) {
#ifdef ZKP_B200_LAB
  static const int env = [] {
    const char* e = getenv("ZKP_B200_K2_CTAS");
    return e ? atoi(e) : 0;
  }();
  if (env > 0) return env;
#endif
  return (L <= 16) ? 4 : (L <= 24 ? 3 : 2);
}
int resident_groups(int S, int num_sms) {  // sized for the larger of K1's and K2's persistent grids
  int T = group_threads(S);
  int L = S / T;
  int mb = (L <= 16) ? 4 : (L <= 24 ? 3 : 2);
  mb = mb > k2_ctas_per_sm(L) ? mb : k2_ctas_per_sm(L);
  return num_sms * mb * (kCtaThreads / T);
}

#define ZKP_DISPATCH(S, CALL)                 \
  switch (S) {                                \
    case 32:  { CALL(4, 8);  } break;         \
    case 64:  { CALL(8, 8);  } break;         \
    case 96:  { CALL(8, 12); } break;         \
    case 128: { CALL(8, 16); } break;         \
    case 192: { CALL(16, 12); } break;        \
    case 256: { CALL(16, 16); } break;        \
    default: return cudaErrorInvalidValue;    \
  }

cudaError_t launch_modexp_shared(const SharedKey& key, const uint32_t* bases, int base_limbs, const uint32_t* plain,
                                 int plain_limbs, uint32_t* out, int out_limbs, int jobs, uint32_t* table,
                                 int num_sms, cudaStream_t st, const unsigned* jobs_dev) {
  if (jobs <= 0) return cudaSuccess;
  if (base_limbs % 4 || (plain && plain_limbs % 4) || base_limbs > key.S || (plain && plain_limbs > key.S) ||
      out_limbs % 2 || out_limbs > key.S || key.nsteps <= 0)
    return cudaErrorInvalidValue;
  SharedParams p;
  p.key = key;
  p.bases = bases;
  p.plain = plain;
  p.out = out;
  p.table = table;
  p.base_limbs = base_limbs;
  p.plain_limbs = plain ? plain_limbs : 0;
  p.out_limbs = out_limbs;
  p.jobs = jobs;
  p.jobs_dev = jobs_dev;
  p.sched_pad = (key.nsteps + 3) & ~3;
#define CALL(T_, L_)                                                                                          \
  {                                                                                                           \
    constexpr int G = kCtaThreads / T_;                                                                       \
    size_t smem = 16 + (size_t)p.sched_pad * 4 + (size_t)G * (p.base_limbs + p.plain_limbs) * 4;              \
    int grid = num_sms * Occ<T_, L_>::kMinBlocks;                                                             \
    int npass = (jobs + G - 1) / G;                                                                           \
    if (grid > npass) grid = npass;                                                                           \
    cudaError_t e = cudaFuncSetAttribute(modexp_shared_kernel<T_, L_>,                                        \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);             \
    if (e != cudaSuccess) return e;                                                                           \
    modexp_shared_kernel<T_, L_><<<grid, kCtaThreads, smem, st>>>(p);                                         \
  }
  ZKP_DISPATCH(key.S, CALL)
#undef CALL
  return cudaGetLastError();
}

cudaError_t launch_mont_setup(const uint32_t* mods, int mod_limbs, int S, int count, uint32_t* r2,
                              uint32_t* n0inv, cudaStream_t st) {
  if (count <= 0) return cudaSuccess;
  if (mod_limbs % 2 || mod_limbs > S) return cudaErrorInvalidValue;
#define CALL(T_, L_)                                                                    \
  {                                                                                     \
    constexpr int G = kCtaThreads / T_;                                                 \
    mont_setup_kernel<T_, L_><<<(count + G - 1) / G, kCtaThreads, 0, st>>>(mods, mod_limbs, count, r2, n0inv); \
  }
  ZKP_DISPATCH(S, CALL)
#undef CALL
  return cudaGetLastError();
}

cudaError_t launch_modexp_var(const uint32_t* bases, const uint32_t* mods, int mod_limbs, const uint32_t* r2,
                              const uint32_t* n0inv, const uint32_t* exps, int exp_limbs, int exp_bits, int exp_per,
                              int mod_per, uint32_t* out, int jobs, int S, uint32_t* table, int num_sms,
                              cudaStream_t st, int base_limbs) {
  if (base_limbs <= 0) base_limbs = mod_limbs;
  if (base_limbs % 2 || base_limbs > S) return cudaErrorInvalidValue;
  if (jobs <= 0) return cudaSuccess;
  if (exp_per <= 0 || mod_per <= 0 || exp_bits <= 0 || exp_bits > 32 * exp_limbs || mod_limbs % 2 || mod_limbs > S)
    return cudaErrorInvalidValue;
  VarParams p;
  p.bases = bases;
  p.mods = mods;
  p.r2 = r2;
  p.n0inv = n0inv;
  p.exps = exps;
  p.out = out;
  p.table = table;
  p.mod_limbs = mod_limbs;
  p.base_limbs = base_limbs;
  p.exp_limbs = exp_limbs;
  p.exp_bits = exp_bits;
  p.exp_per = exp_per;
  p.mod_per = mod_per;
  p.jobs = jobs;
#define CALL(T_, L_)                                                   \
  {                                                                    \
    constexpr int G = kCtaThreads / T_;                                \
    int grid = num_sms * k2_ctas_per_sm(L_);                           \
    int npass = (jobs + G - 1) / G;                                    \
    if (grid > npass) grid = npass;                                    \
    modexp_var_kernel<T_, L_><<<grid, kCtaThreads, 0, st>>>(p);        \
  }
  ZKP_DISPATCH(S, CALL)
#undef CALL
  return cudaGetLastError();
}

cudaError_t launch_modmul_shared(const SharedKey& key, int mode, const uint32_t* a, int a_limbs, const uint32_t* b,
                                 int b_limbs, int b_per, uint32_t* out, int out_limbs, int jobs, cudaStream_t st) {
  if (jobs <= 0) return cudaSuccess;
  if (b_per <= 0 || a_limbs % 2 || a_limbs > key.S || out_limbs % 2 || out_limbs > key.S ||
      (mode == 0 && (b_limbs % 2 || b_limbs > key.S)))
    return cudaErrorInvalidValue;
  MulParams p;
  p.key = key;
  p.a = a;
  p.b = b;
  p.out = out;
  p.a_limbs = a_limbs;
  p.b_limbs = b_limbs;
  p.out_limbs = out_limbs;
  p.b_per = b_per;
  p.jobs = jobs;
  p.mode = mode;
#define CALL(T_, L_)                                                                 \
  {                                                                                  \
    constexpr int G = kCtaThreads / T_;                                              \
    modmul_shared_kernel<T_, L_><<<(jobs + G - 1) / G, kCtaThreads, 0, st>>>(p);     \
  }
  ZKP_DISPATCH(key.S, CALL)
#undef CALL
  return cudaGetLastError();
}

cudaError_t launch_modmul_select(const SharedKey& key, const uint8_t* sel, const uint32_t* a0, const uint32_t* a1,
                                 int a_limbs, const uint32_t* b, int b_limbs, int b_per, uint32_t* out, int out_limbs,
                                 int jobs, cudaStream_t st) {
  if (jobs <= 0) return cudaSuccess;
  if (b_per <= 0 || a_limbs % 2 || a_limbs > key.S || out_limbs % 2 || out_limbs > key.S || b_limbs % 2 ||
      b_limbs > key.S)
    return cudaErrorInvalidValue;
  MulSelParams p;
  p.key = key;
  p.sel = sel;
  p.a0 = a0;
  p.a1 = a1;
  p.b = b;
  p.out = out;
  p.a_limbs = a_limbs;
  p.b_limbs = b_limbs;
  p.out_limbs = out_limbs;
  p.b_per = b_per;
  p.jobs = jobs;
#define CALL(T_, L_)                                                                 \
  {                                                                                  \
    constexpr int G = kCtaThreads / T_;                                              \
    modmul_select_kernel<T_, L_><<<(jobs + G - 1) / G, kCtaThreads, 0, st>>>(p);     \
  }
  ZKP_DISPATCH(key.S, CALL)
#undef CALL
  return cudaGetLastError();
}

// ------------------------------------------------- IMAD peak microbenchmark
template <int VARIANT>
__global__ void __launch_bounds__(256) imad_peak_kernel(int iters, uint32_t* sink) {
  uint32_t a[16], u[18], z[18];
  uint32_t seed = threadIdx.x * 2654435761u + blockIdx.x;
#pragma unroll
  for (int j = 0; j < 16; ++j) a[j] = seed = seed * 1664525u + 1013904223u;
#pragma unroll
  for (int j = 0; j < 18; ++j) {
    u[j] = seed = seed * 1664525u + 1013904223u;
    z[j] = ~seed;
  }
  uint32_t b = seed | 1u;
  for (int it = 0; it < iters; ++it) {
    if (VARIANT == 0 || VARIANT == 3) {  // 16 independent 64-bit accumulators: IMAD.WIDE.U32
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        uint64_t acc = ((uint64_t)u[j + 1] << 32) | u[j];
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a[j]), "r"(b));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a[j + 1]), "r"(b));
        u[j] = (uint32_t)acc;
        u[j + 1] = (uint32_t)(acc >> 32);
      }
    } else if (VARIANT == 1) {  // the row update the Montgomery loop is made of
      Mp<8, 16>::mad_even(u, a, b);
      Mp<8, 16>::mad_odd(z, a, b);
    } else {  // 32-bit IMAD
#pragma unroll
      for (int j = 0; j < 16; ++j) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(u[j]) : "r"(a[j]), "r"(b));
    }
    b += u[0] & 2u;
  }
  uint32_t x = 0;
#pragma unroll
  for (int j = 0; j < 18; ++j) x ^= u[j] ^ z[j];
  if (x == 0x12345678u) sink[0] = x;
}

#ifdef ZKP_B200_LAB
// Probes of the multiplier pipe (lab build only; scripts/imad_probes.py, profiles/r02_imad_probes.json): why does a
// register-only IMAD.WIDE.U32 loop stop at 86 % of one warp instruction per 4 cycles per sub-partition?
//   7  IMAD.HI.U32, independent            8  IMAD.WIDE.U32 with a zero addend (mul.wide)
//   9  IMAD.WIDE.U32 : IADD3 = 1 : 1       10 IMAD.WIDE.U32 : LOP3 = 1 : 2
//   11 32 independent accumulators         12 both multiplicands loop-invariant
//   13 IMAD.WIDE.U32 : IMAD (32-bit) = 1 : 1 on independent registers
template <int VARIANT>
__global__ void __launch_bounds__(256) imad_probe_kernel(int iters, uint32_t* sink) {
  constexpr int NA = VARIANT == 11 ? 32 : 16;
  uint32_t a[16], u[NA], v[16];
  uint32_t seed = threadIdx.x * 2654435761u + blockIdx.x;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    a[j] = seed = seed * 1664525u + 1013904223u;
    v[j] = ~seed;
  }
#pragma unroll
  for (int j = 0; j < NA; ++j) u[j] = seed = seed * 1664525u + 1013904223u;
  uint32_t b = seed | 1u;
  for (int it = 0; it < iters; ++it) {
    if (VARIANT == 7) {
#pragma unroll
      for (int j = 0; j < 16; ++j) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(u[j]) : "r"(a[j]), "r"(b));
    } else if (VARIANT == 8) {
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        uint64_t acc;
        asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(acc) : "r"(a[j] ^ u[j]), "r"(b));
        uint64_t acc2;
        asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(acc2) : "r"(a[j + 1] ^ u[j + 1]), "r"(b));
        u[j] = (uint32_t)acc ^ (uint32_t)(acc2 >> 32);
        u[j + 1] = (uint32_t)(acc >> 32) ^ (uint32_t)acc2;
      }
    } else if (VARIANT == 9 || VARIANT == 10 || VARIANT == 13) {
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        uint64_t acc = ((uint64_t)u[j + 1] << 32) | u[j];
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a[j]), "r"(b));
        if (VARIANT == 9) asm volatile("add.u32 %0, %0, %1;" : "+r"(v[j]) : "r"(a[j]));
        if (VARIANT == 10) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;\n\tlop3.b32 %3, %3, %1, %2, 0x96;" : "+r"(v[j]), "+r"(v[j + 1]) : "r"(a[j]), "r"(b));
        if (VARIANT == 13) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(v[j]) : "r"(a[j]), "r"(b));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a[j + 1]), "r"(b));
        if (VARIANT == 9) asm volatile("add.u32 %0, %0, %1;" : "+r"(v[j + 1]) : "r"(a[j + 1]));
        if (VARIANT == 13) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(v[j + 1]) : "r"(a[j + 1]), "r"(b));
        u[j] = (uint32_t)acc;
        u[j + 1] = (uint32_t)(acc >> 32);
      }
    } else if (VARIANT == 11) {
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        uint64_t acc = ((uint64_t)u[j + 1] << 32) | u[j];
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a[j >> 1]), "r"(b));
        u[j] = (uint32_t)acc;
        u[j + 1] = (uint32_t)(acc >> 32);
      }
    } else {  // 12
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        uint64_t acc = ((uint64_t)u[j + 1] << 32) | u[j];
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a[0]), "r"(b));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a[0]), "r"(b));
        u[j] = (uint32_t)acc;
        u[j + 1] = (uint32_t)(acc >> 32);
      }
    }
    b += u[0] & 2u;
  }
  uint32_t x = 0;
#pragma unroll
  for (int j = 0; j < NA; ++j) x ^= u[j];
#pragma unroll
  for (int j = 0; j < 16; ++j) x ^= v[j];
  if (x == 0x12345678u) sink[0] = x;
}

// Probes of the FP64 pipe ("what comes next": 52-bit limbs on DFMA, Emmart et al., ARITH 2018).
//   17 DFMA, 16 independent accumulators
//   18 one exact 52x52 -> 104-bit product per step: hi = fma_rz(a, b, 2^104), lo = fma_rz(a, b, (2^104 + 2^52) - hi), both words
//      added into 64-bit integer accumulators (2 DFMA + 1 DADD + 2 IADD3 pairs); the rate counts PRODUCTS
//   19 DFMA : IMAD.WIDE.U32 = 1 : 1 on independent registers (do the two pipes overlap?); the rate counts the DFMA
template <int VARIANT>
__global__ void __launch_bounds__(256) dfma_probe_kernel(int iters, uint32_t* sink) {
  double a[16], u[16];
  uint64_t lo[8], hi[8];
  uint32_t w[16];
  uint32_t seed = threadIdx.x * 2654435761u + blockIdx.x;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    seed = seed * 1664525u + 1013904223u;
    a[j] = (double)(seed >> 6) * 67108864.0 + (double)(seed & 0x3ffffffu);  // an integer below 2^52
    u[j] = (double)(seed >> 8);
    w[j] = seed;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) lo[j] = hi[j] = 0;
  double b = (double)(seed | 1u) * 1048576.0;
  const double c1 = 20282409603651670423947251286016.0;  // 2^104
  const double c2 = 20282409603651674927546878656512.0;  // 2^104 + 2^52
  for (int it = 0; it < iters; ++it) {
    if (VARIANT == 17) {
#pragma unroll
      for (int j = 0; j < 16; ++j) asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(u[j]) : "d"(a[j]), "d"(b));
    } else if (VARIANT == 18) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        double h, l, sub;
        asm volatile("fma.rz.f64 %0, %1, %2, %3;" : "=d"(h) : "d"(a[j]), "d"(b), "d"(c1));
        asm volatile("sub.rz.f64 %0, %1, %2;" : "=d"(sub) : "d"(c2), "d"(h));
        asm volatile("fma.rz.f64 %0, %1, %2, %3;" : "=d"(l) : "d"(a[j]), "d"(b), "d"(sub));
        hi[j & 7] += (uint64_t)__double_as_longlong(h);
        lo[j & 7] += (uint64_t)__double_as_longlong(l);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(u[j]) : "d"(a[j]), "d"(b));
        uint64_t acc = ((uint64_t)w[j + 1] << 32) | w[j];
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(w[(j + 2) & 15]), "r"(seed));
        asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(u[j + 1]) : "d"(a[j + 1]), "d"(b));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(w[(j + 3) & 15]), "r"(seed));
        w[j] = (uint32_t)acc;
        w[j + 1] = (uint32_t)(acc >> 32);
      }
    }
    b += (double)(it & 1);
  }
  uint64_t x = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) x ^= (uint64_t)__double_as_longlong(u[j]) ^ w[j];
#pragma unroll
  for (int j = 0; j < 8; ++j) x ^= lo[j] ^ hi[j];
  if (x == 0x12345678u) sink[0] = (uint32_t)x;
}
#endif  // ZKP_B200_LAB

#ifdef ZKP_B200_LAB
// In-lane 32x32 block products by product scanning (blockmul.cuh): the building block of a
// two-digit base-n engine.  MINB = resident CTAs of 128 threads per SM (register budget).
template <int MINB, class Shape>
__global__ void __launch_bounds__(128, MINB) blockmul_peak_kernel(int iters, uint32_t* sink) {
  uint32_t a[32], b[32], out[64];
  uint32_t seed = threadIdx.x * 2654435761u + blockIdx.x;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    a[j] = seed = seed * 1664525u + 1013904223u;
    b[j] = seed = seed * 1664525u + 1013904223u;
  }
  for (int it = 0; it < iters; ++it) {
    block_mul<32, 32, Shape>(out, a, b);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      a[j] ^= out[j];
      b[j] += out[32 + j];
    }
  }
  uint32_t x = 0;
#pragma unroll
  for (int j = 0; j < 32; ++j) x ^= a[j] ^ b[j];
  if (x == 0x12345678u) sink[0] = x;
}

#endif  // ZKP_B200_LAB

cudaError_t launch_imad_peak(int variant, int blocks, int iters, uint32_t* sink, double* ops, cudaStream_t st) {
#ifdef ZKP_B200_LAB
  if (variant >= 7 && variant <= 13) {
    switch (variant) {
      case 7: imad_probe_kernel<7><<<blocks, 256, 0, st>>>(iters, sink); break;
      case 8: imad_probe_kernel<8><<<blocks, 256, 0, st>>>(iters, sink); break;
      case 9: imad_probe_kernel<9><<<blocks, 256, 0, st>>>(iters, sink); break;
      case 10: imad_probe_kernel<10><<<blocks, 256, 0, st>>>(iters, sink); break;
      case 11: imad_probe_kernel<11><<<blocks, 256, 0, st>>>(iters, sink); break;
      case 12: imad_probe_kernel<12><<<blocks, 256, 0, st>>>(iters, sink); break;
      default: imad_probe_kernel<13><<<blocks, 256, 0, st>>>(iters, sink); break;
    }
    *ops = (double)blocks * 256.0 * (double)iters * 16.0;  // multiply-adds of the probed kind (the filler instructions are not counted)
    return cudaGetLastError();
  }
  if (variant >= 17 && variant <= 19) {
    if (variant == 17) dfma_probe_kernel<17><<<blocks, 256, 0, st>>>(iters, sink);
    if (variant == 18) dfma_probe_kernel<18><<<blocks, 256, 0, st>>>(iters, sink);
    if (variant == 19) dfma_probe_kernel<19><<<blocks, 256, 0, st>>>(iters, sink);
    *ops = (double)blocks * 256.0 * (double)iters * 16.0;  // DFMA (17, 19) or whole 52x52 products (18)
    return cudaGetLastError();
  }
  if (variant >= 14 && variant <= 16) {  // variant 0 at 1 / 2 / 4 warps per sub-partition (4 / 8 / 16 warps per SM)
    const int per_sm = variant == 14 ? 1 : (variant == 15 ? 2 : 4);
    const int grid = blocks / 8 * per_sm;
    imad_peak_kernel<3><<<grid, 128, 0, st>>>(iters, sink);
    *ops = (double)grid * 128.0 * (double)iters * 16.0;
    return cudaGetLastError();
  }
  if (variant >= 3 && variant <= 6) {
    // 3: full block, 4 CTAs/SM (16 warps); 4: full, 2 CTAs/SM (8 warps); 5: lower triangle, 4 CTAs/SM; 6: full, 3 CTAs/SM
    const int it2 = iters / 64 > 0 ? iters / 64 : 1;
    const int per_sm = variant == 4 ? 2 : (variant == 6 ? 3 : 4);
    const int grid = blocks / 8 * per_sm;
    if (variant == 3) blockmul_peak_kernel<4, ShapeFull><<<grid, 128, 0, st>>>(it2, sink);
    if (variant == 4) blockmul_peak_kernel<2, ShapeFull><<<grid, 128, 0, st>>>(it2, sink);
    if (variant == 5) blockmul_peak_kernel<4, ShapeLowTri><<<grid, 128, 0, st>>>(it2, sink);
    if (variant == 6) blockmul_peak_kernel<3, ShapeFull><<<grid, 128, 0, st>>>(it2, sink);
    *ops = (double)grid * 128.0 * (double)it2 * (variant == 5 ? 528.0 : 1024.0);
    return cudaGetLastError();
  }
#endif
  switch (variant) {
    case 0: imad_peak_kernel<0><<<blocks, 256, 0, st>>>(iters, sink); break;
    case 1: imad_peak_kernel<1><<<blocks, 256, 0, st>>>(iters, sink); break;
    case 2: imad_peak_kernel<2><<<blocks, 256, 0, st>>>(iters, sink); break;
    default: return cudaErrorInvalidValue;
  }
  *ops = (double)blocks * 256.0 * (double)iters * 16.0;
  return cudaGetLastError();
}

}  // namespace zkp
