// K1v2: Paillier encryption c = (1 + m n) r^n mod n^2 in TWO-DIGIT BASE-n ARITHMETIC (|n| = 2048 exactly).
//
// Replaces the same reference lines as K1 (Paillier::encrypt_with_chosen_randomness at
// range_proof.rs:165,179,280,286,330 ...) with fewer multiplies.  An element x of Z_{n^2} is held as
// x = X0 + X1 n with 0 <= X0, X1 < n, so
//     x y = X0 Y0 + (X0 Y1 + X1 Y0) n             (mod n^2: the X1 Y1 n^2 term vanishes)
//         = R + (Q + X0 Y1 + X1 Y0 mod n) n        with (Q, R) = divmod(X0 Y0, n)
// i.e. three (two for a squaring) 2048 x 2048-bit products, each followed by a Barrett reduction modulo n
// (two more 2048 x 2048 products by the constants mu' and n), instead of one 4096 x 4096-bit Montgomery
// multiplication: 24.6 k limb products per squaring against 32.8 k in K1.
//
// Layout: one modexp per PAIR of lanes (Mp<2, 32>): a 64-limb digit is spread over the pair, 32 limbs per lane.
// Every product runs on the same split-accumulator rows as K1's Montgomery loop (mp_coop.cuh: mul_step), one
// row per limb of the streamed operand, in a ROLLED loop (a few hundred instructions in total: the whole
// engine stays in the instruction cache).  The multiplicand sits in registers; the streamed operand, the
// captured low half of each product and the digits that live across products sit in per-lane slots of
// shared memory (hot) or of an L2-resident scratch (cold), word-interleaved across threads.
//
// Barrett (W = 2^2048, mu' = floor(W^2 / n) - W, P = H W + P_lo < n W):
//     q^ = H + floor(H mu' / W) >= q - 2,   r = (P - q^ n) mod 2^2080 in [0, 3n): up to three corrections.
#include <cstring>

#include "kernels.h"
#include "mp_coop.cuh"

namespace zkp {
namespace v2 {

constexpr int BL = 32;        // limbs per lane per digit
constexpr int DL = 64;        // limbs per digit
constexpr int kThreads = 128;
constexpr int kCtasPerSm = 3;
using M2 = Mp<2, BL>;

struct KeyConst {
  uint32_t n[DL];   // modulus, exactly 2048 bits
  uint32_t mu[DL];  // mu' = floor(2^4096 / n) - 2^2048
};
// Per-key constants in the constant bank (one key per device at a time: launches of contexts holding
// different keys must not overlap).
__constant__ KeyConst c_key;

struct Enc2dParams {
  const uint32_t* sched;
  const uint32_t* bases;   // [jobs][64]
  const uint32_t* plain;   // [jobs][plain_limbs] or null
  uint32_t* out;           // [jobs][128]
  uint32_t* table;         // [groups][17][128]: odd powers x^1..x^31 and x^2, as (X0 | X1) digit pairs
  uint32_t* cold;          // [grid][2][32][kThreads]: the cold slots X1 and ACC
  const unsigned* jobs_dev;
  int nsteps, plain_limbs, jobs;
};

// Digit slots.  Slot s, limb i of the lane in column `col` (a thread index within the CTA).
enum Slot { S_X0 = 0, S_LO = 1, S_HI = 2, S_Q = 3, S_X1 = 4, S_ACC = 5 };
struct Slots {
  uint32_t smem;       // shared-window byte address of the CTA's hot slots [4][32][kThreads]
  uint32_t* cold;      // this CTA's cold slots [2][32][kThreads] (global, L1/L2 resident)
  int tid;
  __device__ __forceinline__ uint32_t ld(int s, int i, int col) const {
    if (s < 4) {
      uint32_t v;
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem + 4u * (uint32_t)((s * BL + i) * kThreads + col)) : "memory");
      return v;
    }
    return cold[((s - 4) * BL + i) * kThreads + col];
  }
  __device__ __forceinline__ void st(int s, int i, int col, uint32_t v) const {
    if (s < 4) {
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(smem + 4u * (uint32_t)((s * BL + i) * kThreads + col)), "r"(v) : "memory");
    } else {
      cold[((s - 4) * BL + i) * kThreads + col] = v;
    }
  }
  __device__ __forceinline__ void load(uint32_t (&x)[BL], int s) const {
#pragma unroll
    for (int i = 0; i < BL; ++i) x[i] = ld(s, i, tid);
  }
  __device__ __forceinline__ void store(int s, const uint32_t (&x)[BL]) const {
#pragma unroll
    for (int i = 0; i < BL; ++i) st(s, i, tid, x[i]);
  }
};

enum BSrc { B_SLOT = 0, B_N = 1, B_MU = 2 };

// hi = floor(a * B / W) exactly; the low 64 limbs of the product are captured into slot `cap` (cap < 0: dropped).
// a: this lane's half of the multiplicand digit (registers); B: slot `bslot` (must be a hot slot), n or mu'.
template <int BSRC>
__device__ __forceinline__ void mulw(uint32_t (&hi)[BL], const uint32_t (&a)[BL], int bslot, int cap, const Slots& sl, int lane) {
  const int g = lane & 1;
  const int col0 = sl.tid & ~1;
  uint32_t E[BL + 2], O[BL + 2];
#pragma unroll
  for (int j = 0; j < BL + 2; ++j) E[j] = O[j] = 0;
  // 16 rows per iteration: inside the unrolled body the accumulator arrays are renamed, not copied
#pragma unroll 8
  for (int k = 0; k < DL; k += 2) {
    uint32_t b0, b1;
    if (BSRC == B_SLOT) {
      b0 = sl.ld(bslot, k & (BL - 1), col0 | (k >> 5));
      b1 = sl.ld(bslot, (k + 1) & (BL - 1), col0 | (k >> 5));
    } else if (BSRC == B_N) {
      b0 = c_key.n[k];
      b1 = c_key.n[k + 1];
    } else {
      b0 = c_key.mu[k];
      b1 = c_key.mu[k + 1];
    }
    M2::mul_step(E, O, a, b0, g);
    if (cap >= 0 && g == 0) sl.st(cap, k & (BL - 1), col0 | (k >> 5), E[0]);
    M2::mul_step(O, E, a, b1, g);
    if (cap >= 0 && g == 0) sl.st(cap, (k + 1) & (BL - 1), col0 | (k >> 5), O[0]);
  }
  M2::mul_finish(hi, E, O, lane);
}

// x += c (small) over the pair; returns the carry out of the digit
__device__ __forceinline__ uint32_t add_small_digit(uint32_t (&x)[BL], uint32_t c, int lane) {
  const int g = lane & 1;
  add_cc(x[0], g == 0 ? c : 0u);
#pragma unroll
  for (int j = 1; j < BL; ++j) addc_cc(x[j], 0);
  uint32_t co = addc_out();
  uint32_t top;
  uint32_t cin = M2::resolve(co != 0, M2::all_ones(x), lane, top);
  M2::add_small(x, cin);
  return top;
}

// x = (x + y) reduced by at most one subtraction of n (x + y < 2n)
__device__ __forceinline__ void add_mod(uint32_t (&x)[BL], const uint32_t (&y)[BL], const uint32_t (&nreg)[BL], int lane) {
  const uint32_t ovf = M2::add_full(x, y, lane);
  uint32_t d[BL];
  const uint32_t borrow = M2::sub_full(d, x, nreg, lane);
  const bool take = (ovf != 0u) || (borrow == 0u);
#pragma unroll
  for (int j = 0; j < BL; ++j) x[j] = take ? d[j] : x[j];
}

// (LO, HI) slots hold P < n W.  On exit slot Q = floor(P / n), slot LO = P mod n (also returned in r).
__device__ __forceinline__ void barrett(uint32_t (&r)[BL], const uint32_t (&nreg)[BL], const Slots& sl, int lane) {
  const int l0 = lane & ~1;
  uint32_t h[BL], t[BL];
  sl.load(h, S_HI);
  const uint32_t p64 = __shfl_sync(ZKP_FULL, h[0], l0);
  mulw<B_MU>(t, h, 0, -1, sl, lane);             // t = floor(H mu' / W)
  M2::add_full(h, t, lane);                      // q^ = H + t   (<= q < W)
  mulw<B_N>(t, h, 0, S_HI, sl, lane);            // low 64 limbs of q^ n -> slot HI; t = high part
  const uint32_t L64 = __shfl_sync(ZKP_FULL, t[0], l0);
  __syncwarp();
  uint32_t d[BL];
  sl.load(r, S_LO);
  sl.load(t, S_HI);
  uint32_t borrow = M2::sub_full(d, r, t, lane);
  uint32_t r64 = p64 - L64 - borrow;
#pragma unroll
  for (int j = 0; j < BL; ++j) r[j] = d[j];
  uint32_t corr = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    borrow = M2::sub_full(d, r, nreg, lane);
    const bool take = (r64 != 0u) || (borrow == 0u);
#pragma unroll
    for (int j = 0; j < BL; ++j) r[j] = take ? d[j] : r[j];
    r64 = take ? r64 - borrow : r64;
    corr += take ? 1u : 0u;
  }
  add_small_digit(h, corr, lane);
  sl.store(S_Q, h);
  sl.store(S_LO, r);
  __syncwarp();
}

// One round (the only out-of-line routine: three product loops live here, once):
//   a = slot `aslot`, or the 32 limbs at `aptr` (global) when aptr != nullptr;  P = a * B(slot bslot);
//   (Q, R) = divmod(P, n)  ->  slot Q, slot LO.
__device__ __noinline__ void round_mul(int aslot, const uint32_t* __restrict__ aptr, int alimbs, int bslot, Slots sl, int lane) {
  const int g = lane & 1;
  uint32_t a[BL], hi[BL], r[BL], nreg[BL];
#pragma unroll
  for (int j = 0; j < BL; ++j) nreg[j] = c_key.n[BL * g + j];
  if (aptr) M2::load_ext(a, aptr, alimbs, g);
  else sl.load(a, aslot);
  mulw<B_SLOT>(hi, a, bslot, S_LO, sl, lane);
  sl.store(S_HI, hi);
  __syncwarp();
  barrett(r, nreg, sl, lane);
}

// slot x = (slot x [+ slot x again if DOUBLE] + slot y) mod n ... small glue between rounds
__device__ __forceinline__ void slot_add_mod(int dst, int xs, int ys, const Slots& sl, int lane) {
  const int g = lane & 1;
  uint32_t x[BL], y[BL], nreg[BL];
#pragma unroll
  for (int j = 0; j < BL; ++j) nreg[j] = c_key.n[BL * g + j];
  sl.load(x, xs);
  sl.load(y, ys);
  add_mod(x, y, nreg, lane);
  sl.store(dst, x);
}
__device__ __forceinline__ void slot_copy(int dst, int src, const Slots& sl) {
  uint32_t x[BL];
  sl.load(x, src);
  sl.store(dst, x);
}

// (X0, X1) <- (X0, X1)^2      [slots X0 (hot), X1 (cold)]
__device__ __forceinline__ void sqr2(const Slots& sl, int lane) {
  round_mul(S_X1, nullptr, 0, S_X0, sl, lane);   // U = X0 X1 mod n          -> LO
  slot_add_mod(S_ACC, S_LO, S_LO, sl, lane);     // ACC = 2U mod n
  round_mul(S_X0, nullptr, 0, S_X0, sl, lane);   // (Q, R) = divmod(X0^2, n) -> Q, LO
  slot_add_mod(S_X1, S_ACC, S_Q, sl, lane);      // X1 = 2U + Q mod n
  slot_copy(S_X0, S_LO, sl);                     // X0 = R
  __syncwarp();
}

// (X0, X1) <- (X0, X1) * (Y0, Y1), the table entry at `y` (this lane's halves: Y0 at y, Y1 at y + 32)
__device__ __forceinline__ void mul2(const uint32_t* y, const Slots& sl, int lane) {
  const int g = lane & 1;
  round_mul(0, y + BL - g * BL, DL, S_X0, sl, lane);  // U = X0 Y1 mod n  (load_ext indexes by g: pass the digit base)
  slot_copy(S_ACC, S_LO, sl);
  {  // V = X1 Y0 mod n: stream Y0 from a hot slot (Q is free here), multiplicand X1
    uint32_t t[BL];
    M2::load(t, y);
    sl.store(S_Q, t);
    __syncwarp();
  }
  round_mul(S_X1, nullptr, 0, S_Q, sl, lane);
  slot_add_mod(S_ACC, S_ACC, S_LO, sl, lane);    // U + V mod n
  round_mul(0, y - g * BL, DL, S_X0, sl, lane);  // (Q, R) = divmod(X0 Y0, n)
  slot_add_mod(S_X1, S_ACC, S_Q, sl, lane);      // X1 = U + V + Q mod n
  slot_copy(S_X0, S_LO, sl);
  __syncwarp();
}

__global__ void __launch_bounds__(kThreads, kCtasPerSm) enc2d_kernel(const __grid_constant__ Enc2dParams p) {
  extern __shared__ __align__(16) uint32_t smem[];
  uint32_t* s_sched = smem;
  const int sched_pad = (p.nsteps + 3) & ~3;
  const int lane = threadIdx.x & 31;
  const int g = lane & 1;
  Slots sl;
  sl.smem = (uint32_t)__cvta_generic_to_shared(smem + sched_pad);
  sl.cold = p.cold + (size_t)blockIdx.x * (2 * BL * kThreads);
  sl.tid = threadIdx.x;
  for (int i = threadIdx.x; i < p.nsteps; i += blockDim.x) s_sched[i] = p.sched[i];
  __syncthreads();

  constexpr int G = kThreads / 2;  // modexps per CTA pass
  constexpr int kEntry = 2 * DL;   // limbs per table entry
  const int grp = threadIdx.x >> 1;
  const int jobs = p.jobs_dev ? min((int)*p.jobs_dev, p.jobs) : p.jobs;
  const int npass = (jobs + G - 1) / G;
  uint32_t* tab = p.table + ((size_t)(blockIdx.x * G + grp) * (kTableShared + 1)) * kEntry + g * DL;
  for (int cj = blockIdx.x; cj < npass; cj += gridDim.x) {
    const int job = cj * G + grp;
    const bool valid = job < jobs;
    const int src = valid ? job : 0;
    {
      uint32_t x0[BL], x1[BL], d[BL], nreg[BL];
#pragma unroll
      for (int j = 0; j < BL; ++j) nreg[j] = c_key.n[BL * g + j];
      M2::load(x0, p.bases + (size_t)src * DL + g * BL);
      // r may exceed n (r < W < 2n): X0 = r mod n, X1 = floor(r / n)
      const uint32_t borrow = M2::sub_full(d, x0, nreg, lane);
      const bool take = borrow == 0u;
#pragma unroll
      for (int j = 0; j < BL; ++j) {
        x0[j] = take ? d[j] : x0[j];
        x1[j] = 0;
      }
      if (g == 0) x1[0] = take ? 1u : 0u;
      sl.store(S_X0, x0);
      sl.store(S_X1, x1);
      M2::store(tab, x0);
      M2::store(tab + BL, x1);
      __syncwarp();
    }
    // x^2 -> table entry 16, then the odd powers x^3 .. x^31
    sqr2(sl, lane);
    {
      uint32_t t[BL];
      sl.load(t, S_X0);
      M2::store(tab + kTableShared * kEntry, t);
      sl.load(t, S_X1);
      M2::store(tab + kTableShared * kEntry + BL, t);
      M2::load(t, tab);
      sl.store(S_X0, t);
      M2::load(t, tab + BL);
      sl.store(S_X1, t);
      __syncwarp();
    }
#pragma unroll 1
    for (int e = 1; e < kTableShared; ++e) {
      mul2(tab + kTableShared * kEntry, sl, lane);
      uint32_t t[BL];
      sl.load(t, S_X0);
      M2::store(tab + e * kEntry, t);
      sl.load(t, S_X1);
      M2::store(tab + e * kEntry + BL, t);
    }
    uint32_t st = s_sched[0];
    {
      uint32_t t[BL];
      M2::load(t, tab + (st & 0xffu) * kEntry);
      sl.store(S_X0, t);
      M2::load(t, tab + (st & 0xffu) * kEntry + BL);
      sl.store(S_X1, t);
      __syncwarp();
    }
#pragma unroll 1
    for (int k = 1; k < p.nsteps; ++k) {
      st = s_sched[k];
      const uint32_t idx = st & 0xffu;
      const int nsq = (int)(st >> 8);
#pragma unroll 1
      for (int q = 0; q < nsq; ++q) sqr2(sl, lane);
      if (idx != 0xffu) mul2(tab + idx * kEntry, sl, lane);
    }
    // c = (1 + m n) x = X0 + ((X1 + m X0) mod n) n   (mod n^2), written out as the 128-limb integer X0 + c1 n
    {
      uint32_t c1[BL], nreg[BL];
#pragma unroll
      for (int j = 0; j < BL; ++j) nreg[j] = c_key.n[BL * g + j];
      sl.load(c1, S_X1);
      if (p.plain) {
        uint32_t r[BL];
        round_mul(0, p.plain + (size_t)src * p.plain_limbs, p.plain_limbs, S_X0, sl, lane);  // m X0 mod n -> LO
        sl.load(r, S_LO);
        add_mod(c1, r, nreg, lane);
      }
      uint32_t hi[BL], lo[BL], x0[BL];
      mulw<B_N>(hi, c1, 0, S_LO, sl, lane);        // c1 * n: low half -> slot LO
      __syncwarp();
      sl.load(lo, S_LO);
      sl.load(x0, S_X0);
      const uint32_t c = M2::add_full(lo, x0, lane);
      add_small_digit(hi, c, lane);
      if (valid) {
        M2::store(p.out + (size_t)job * (2 * DL) + g * BL, lo);
        M2::store(p.out + (size_t)job * (2 * DL) + DL + g * BL, hi);
      }
      __syncwarp();
    }
  }
}

}  // namespace v2

// Host side: mu' = floor(2^4096 / n) - 2^2048 by bitwise restoring division (once per launch: 4097 x 66 limb ops).
static void barrett_mu(const uint32_t* n, uint32_t* mu) {
  uint32_t rem[66] = {0};
  uint32_t quo[130] = {0};
  for (int bit = 4096; bit >= 0; --bit) {
    uint32_t carry = (bit == 4096) ? 1u : 0u;  // the numerator 2^4096 has a single set bit
    for (int i = 0; i < 66; ++i) {
      uint32_t nc = rem[i] >> 31;
      rem[i] = (rem[i] << 1) | carry;
      carry = nc;
    }
    bool ge = rem[65] != 0 || rem[64] != 0;
    if (!ge) {
      ge = true;
      for (int i = 63; i >= 0; --i) {
        if (rem[i] != n[i]) {
          ge = rem[i] > n[i];
          break;
        }
      }
    }
    if (ge) {
      uint64_t br = 0;
      for (int i = 0; i < 66; ++i) {
        uint64_t t = (uint64_t)rem[i] - (i < 64 ? n[i] : 0u) - br;
        rem[i] = (uint32_t)t;
        br = (t >> 63) & 1u;
      }
      quo[bit >> 5] |= 1u << (bit & 31);
    }
  }
  for (int i = 0; i < 64; ++i) mu[i] = quo[i];  // quo = 2^2048 + mu'
}

bool enc2d_supported(const uint32_t* n_host, int n_limbs_exact) {
  return n_limbs_exact == v2::DL && (n_host[v2::DL - 1] >> 31) == 1u && (n_host[0] & 1u);
}

int enc2d_resident_groups(int num_sms) { return num_sms * v2::kCtasPerSm * (v2::kThreads / 2); }

// scratch limbs: the window tables (17 entries of 128 limbs per group) + the cold slots (64 limbs per thread)
size_t enc2d_scratch_limbs(int num_sms) {
  return (size_t)enc2d_resident_groups(num_sms) * ((kTableShared + 1) * 2 * v2::DL + 2 * 2 * v2::BL);
}

cudaError_t launch_enc2d(const uint32_t* n_host, const uint32_t* sched_dev, int nsteps, const uint32_t* bases, const uint32_t* plain,
                         int plain_limbs, uint32_t* out, int jobs, uint32_t* scratch, int num_sms, cudaStream_t st,
                         const unsigned* jobs_dev) {
  if (jobs <= 0) return cudaSuccess;
  if (plain && (plain_limbs % 2 || plain_limbs > v2::DL)) return cudaErrorInvalidValue;
  v2::KeyConst hk;
  memcpy(hk.n, n_host, sizeof(hk.n));
  barrett_mu(n_host, hk.mu);
  cudaError_t e = cudaMemcpyToSymbolAsync(v2::c_key, &hk, sizeof(hk), 0, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return e;
  e = cudaStreamSynchronize(st);  // hk lives on this stack frame
  if (e != cudaSuccess) return e;
  v2::Enc2dParams p;
  p.sched = sched_dev;
  p.nsteps = nsteps;
  p.bases = bases;
  p.plain = plain;
  p.plain_limbs = plain ? plain_limbs : 0;
  p.out = out;
  p.table = scratch;
  p.cold = scratch + (size_t)enc2d_resident_groups(num_sms) * (kTableShared + 1) * 2 * v2::DL;
  p.jobs = jobs;
  p.jobs_dev = jobs_dev;
  constexpr int G = v2::kThreads / 2;
  int grid = num_sms * v2::kCtasPerSm;
  const int npass = (jobs + G - 1) / G;
  if (grid > npass) grid = npass;
  const size_t smem = ((size_t)((nsteps + 3) & ~3) + 4 * v2::BL * v2::kThreads) * 4;
  e = cudaFuncSetAttribute(v2::enc2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  v2::enc2d_kernel<<<grid, v2::kThreads, smem, st>>>(p);
  return cudaGetLastError();
}

}  // namespace zkp
