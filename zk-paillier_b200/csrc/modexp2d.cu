// K1v2: Paillier encryption c = (1 + m n) r^n mod n^2 in TWO-DIGIT BASE-n ARITHMETIC (|n| = 2048 exactly).
//
// Replaces the same reference lines as K1 (Paillier::encrypt_with_chosen_randomness at
// range_proof.rs:165,179,280,286,330 ...) with fewer multiplies: an element x of Z_{n^2} is held as
// x = X0 + X1 n with 0 <= X0, X1 < n, so
//     x y = X0 Y0 + (X0 Y1 + X1 Y0) n            (mod n^2; the X1 Y1 n^2 term vanishes)
//         = R + (Q + X0 Y1 + X1 Y0 mod n) n       with (Q, R) = divmod(X0 Y0, n)
// i.e. three (two for a squaring) 2048 x 2048-bit products and Barrett reductions modulo n instead of a
// 4096 x 4096-bit Montgomery multiplication: ~21 k / ~31 k limb products per squaring / multiplication
// against 32.8 k in K1.
//
// Layout: one modexp per PAIR of lanes (T = 2); a 64-limb digit is two 32-limb blocks, lane g of the pair
// holds block g in registers.  Every big product is built from in-lane 32x32 block products by product
// scanning (blockmul.cuh: pure IMAD.WIDE streams, no shuffles); the lane's two block results go to a
// limb-interleaved shared-memory scratch, and each lane then sums the pieces that fall into its own two
// 32-limb slices of the 128-limb product.  All lanes of a warp execute the same shapes (SIMT-uniform).
//
// Barrett (W = 2^2048, mu' = floor(W^2 / n) - W, P < n W):
//     q^ = H + floor(H mu' / W)   with H = floor(P / W), the product truncated below limb 62
//     r  = (P - q^ n) mod 2^2080  in [0, 4n);  up to three corrections  r -= n, q^ += 1.
#include <cstring>

#include "kernels.h"
#include "mp_coop.cuh"
#include "blockmul.cuh"

namespace zkp {
namespace v2 {

constexpr int BL = 32;            // limbs per block (= per lane per digit)
constexpr int DL = 64;            // limbs per digit
constexpr int kPiece = 66;        // limbs per block-product piece (64 + 2 accumulator words)
constexpr int kThreads = 128;
constexpr int kWarpWords = 2 * kPiece * 32;  // scratch words per warp: 2 pieces per lane, limb-interleaved

using M2 = Mp<2, BL>;

struct KeyConst {
  uint32_t n[DL];    // modulus, exactly 2048 bits
  uint32_t mu[DL];   // mu' = floor(2^4096 / n) - 2^2048
};

// Per-key constants in the constant bank: every limb of n / mu' is a static c[][] operand of an IMAD.WIDE.
// (One key per device at a time: launches of contexts with different keys must not overlap.)
__constant__ KeyConst c_key;

struct Enc2dParams {
  const uint32_t* sched;
  const uint32_t* bases;   // [jobs][64]
  const uint32_t* plain;   // [jobs][plain_limbs] or null
  uint32_t* out;           // [jobs][128]
  uint32_t* table;         // window table scratch: [groups][16][128]
  const unsigned* jobs_dev;
  int nsteps, plain_limbs, jobs;
};

// limb k of piece p of lane ln, limb-interleaved across the warp (conflict-free for same-k accesses)
struct Scratch {
  uint32_t base;  // shared-window byte address of this warp's region
  __device__ __forceinline__ uint32_t addr(int p, int k, int ln) const { return base + 4u * (uint32_t)((p * kPiece + k) * 32 + ln); }
  __device__ __forceinline__ uint32_t ld(int p, int k, int ln) const {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr(p, k, ln)) : "memory");
    return v;
  }
  __device__ __forceinline__ void st(int p, int k, int ln, uint32_t v) const {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr(p, k, ln)), "r"(v) : "memory");
  }
};

// x += c (small), returns the carry out of the 32 limbs
__device__ __forceinline__ uint32_t add_word(uint32_t (&x)[BL], uint32_t c) {
  add_cc(x[0], c);
#pragma unroll
  for (int j = 1; j < BL; ++j) addc_cc(x[j], 0);
  return addc_out();
}

// (lo, hi) = A * B for two digits: lane g passes its blocks a = A_g and b = B_g.
//   piece 0 = A_g x B_g (own x own), piece 1 = A_g x B_{1-g} (own x partner's), 64 limbs each.
//   With Z_uv = A_u x B_v at limb offset 32 (u + v):
//     lane 0 owns P[0,32)  = Z00[0,32)                              and P[64,96)  = Z11[0,32) + Z01[32,64) + Z10[32,64)
//     lane 1 owns P[32,64) = Z00[32,64) + Z01[0,32) + Z10[0,32)     and P[96,128) = Z11[32,64)
__device__ __forceinline__ void assemble_full(uint32_t (&lo)[BL], uint32_t (&hi)[BL], const Scratch& sc, int lane) {
  const int g = lane & 1, l0 = lane & ~1, l1 = lane | 1;
  __syncwarp();
  // Z00 = piece0 of lane0, Z11 = piece0 of lane1, Z01 = piece1 of lane0, Z10 = piece1 of lane1.
  // All loads of a slice are issued back to back before the (serial) carry chain consumes them.
  uint32_t x[BL], y[BL], z[BL];
#pragma unroll
  for (int t = 0; t < BL; ++t) {
    x[t] = sc.ld(0, 32 * g + t, l0);
    y[t] = sc.ld(1, t, l0);
    z[t] = sc.ld(1, t, l1);
  }
  unsigned long long c = 0;
#pragma unroll
  for (int t = 0; t < BL; ++t) {
    unsigned long long v = (unsigned long long)x[t] + c;
    if (g) v += (unsigned long long)y[t] + z[t];
    lo[t] = (uint32_t)v;
    c = v >> 32;
  }
  uint32_t cA = (uint32_t)c;
#pragma unroll
  for (int t = 0; t < BL; ++t) {
    x[t] = sc.ld(0, 32 * g + t, l1);
    y[t] = sc.ld(1, 32 + t, l0);
    z[t] = sc.ld(1, 32 + t, l1);
  }
  c = 0;
#pragma unroll
  for (int t = 0; t < BL; ++t) {
    unsigned long long v = (unsigned long long)x[t] + c;
    if (!g) v += (unsigned long long)y[t] + z[t];
    hi[t] = (uint32_t)v;
    c = v >> 32;
  }
  uint32_t cB = (uint32_t)c;
  // carries: lane1.sliceA -> lane0.sliceB -> lane1.sliceB
  const uint32_t cA_p = __shfl_xor_sync(ZKP_FULL, cA, 1);
  cB += add_word(hi, g ? 0u : cA_p);
  const uint32_t cB_p = __shfl_xor_sync(ZKP_FULL, cB, 1);
  add_word(hi, g ? cB_p : 0u);
  __syncwarp();
}

__device__ __forceinline__ void prod_full_body(uint32_t (&lo)[BL], uint32_t (&hi)[BL], const uint32_t (&a)[BL], const uint32_t (&b)[BL],
                                               const Scratch& sc, int lane) {
  uint32_t bp[BL];
#pragma unroll
  for (int j = 0; j < BL; ++j) bp[j] = __shfl_xor_sync(ZKP_FULL, b[j], 1);
  block_mul_cols<BL, BL, ShapeFull, 0, 2 * BL - 1>(a, [&](int j) { return b[j]; }, [&](int k, uint32_t v) { sc.st(0, k, lane, v); });
  block_mul_cols<BL, BL, ShapeFull, 0, 2 * BL - 1>(a, [&](int j) { return bp[j]; }, [&](int k, uint32_t v) { sc.st(1, k, lane, v); });
  assemble_full(lo, hi, sc, lane);
}

// One out-of-line copy of each big routine (they are ~5-8 k instructions each); operands cross the call in
// local memory and are pulled into registers on entry.
__device__ __noinline__ void prod_full(uint32_t* __restrict__ lo_, uint32_t* __restrict__ hi_, const uint32_t* __restrict__ a_,
                                       const uint32_t* __restrict__ b_, Scratch sc, int lane) {
  uint32_t a[BL], b[BL], lo[BL], hi[BL];
#pragma unroll
  for (int j = 0; j < BL; ++j) {
    a[j] = a_[j];
    b[j] = b_[j];
  }
  prod_full_body(lo, hi, a, b, sc, lane);
#pragma unroll
  for (int j = 0; j < BL; ++j) {
    lo_[j] = lo[j];
    hi_[j] = hi[j];
  }
}

// Exact (q, r) = divmod(P, n) for P = hi * W + lo < n W.  q may be discarded by the caller.
template <bool WANT_Q>
__device__ __forceinline__ void barrett_body(uint32_t (&q)[BL], uint32_t (&r)[BL], const uint32_t (&lo)[BL], const uint32_t (&hi)[BL],
                                        const uint32_t (&nreg)[BL], const Scratch& sc, int lane) {
  const int g = lane & 1, l0 = lane & ~1, l1 = lane | 1;
  // ---- q^ = H + floor(H mu' / W): pieces p = H_g x mu'_p, global limb offset 32 (g + p); limbs >= 62 wanted
  block_mul_cols<BL, BL, ShapeFull, 30, 2 * BL - 1>(hi, [&](int j) { return c_key.mu[j]; }, [&](int k, uint32_t v) { sc.st(0, k, lane, v); });
  block_mul_cols<BL, BL, ShapeFull, 0, 2 * BL - 1>(hi, [&](int j) { return c_key.mu[BL + j]; }, [&](int k, uint32_t v) { sc.st(1, k, lane, v); });
  __syncwarp();
  uint32_t qh[BL];
  {
    // guard limbs 62, 63: Z00[62..63] + Z01[30..31] + Z10[30..31]  (Z_gp = piece p of lane g)
    unsigned long long g0 = (unsigned long long)sc.ld(0, 62, l0) + sc.ld(1, 30, l0) + sc.ld(0, 30, l1);
    unsigned long long g1 = (unsigned long long)sc.ld(0, 63, l0) + sc.ld(1, 31, l0) + sc.ld(0, 31, l1) + (g0 >> 32);
    unsigned long long c = g ? 0ull : (g1 >> 32);
    uint32_t x[BL], y[BL], z[BL];
#pragma unroll
    for (int t = 0; t < BL; ++t) {
      x[t] = sc.ld(1, 32 * g + t, l1);
      y[t] = sc.ld(1, 32 + t, l0);
      z[t] = sc.ld(0, 32 + t, l1);
    }
#pragma unroll
    for (int t = 0; t < BL; ++t) {
      // limb 64 + 32 g + t:  H[.] + Z11[32 g + t] + (g == 0 ? Z01[32 + t] + Z10[32 + t] : 0)
      unsigned long long v = (unsigned long long)hi[t] + x[t] + c;
      if (!g) v += (unsigned long long)y[t] + z[t];
      qh[t] = (uint32_t)v;
      c = v >> 32;
    }
    const uint32_t c_p = __shfl_xor_sync(ZKP_FULL, (uint32_t)c, 1);
    add_word(qh, g ? c_p : 0u);  // q^ <= q < W: nothing leaves lane 1
  }
  __syncwarp();
  // ---- L = q^ n mod 2^(32*65): pieces p = Q_g x n_p; global limbs <= 64 wanted
  block_mul_cols<BL, BL, ShapeFull, 0, 2 * BL - 1>(qh, [&](int j) { return c_key.n[j]; }, [&](int k, uint32_t v) { sc.st(0, k, lane, v); });
  block_mul_cols<BL, BL, ShapeFull, 0, 33>(qh, [&](int j) { return c_key.n[BL + j]; }, [&](int k, uint32_t v) { sc.st(1, k, lane, v); });
  __syncwarp();
  uint32_t L[BL];
  uint32_t L64;
  {
    unsigned long long c = 0;
    uint32_t x[BL], y[BL], z[BL];
#pragma unroll
    for (int t = 0; t < BL; ++t) {
      x[t] = sc.ld(0, 32 * g + t, l0);
      y[t] = sc.ld(1, t, l0);
      z[t] = sc.ld(0, t, l1);
    }
#pragma unroll
    for (int t = 0; t < BL; ++t) {
      // limb 32 g + t: Z00[32 g + t] + (g ? Z01[t] + Z10[t] : 0)
      unsigned long long v = (unsigned long long)x[t] + c;
      if (g) v += (unsigned long long)y[t] + z[t];
      L[t] = (uint32_t)v;
      c = v >> 32;
    }
    // limb 64 = Z01[32] + Z10[32] + Z11[0] + carry out of lane 1's slice (lane 0's slice carries nothing out)
    unsigned long long v64 = (unsigned long long)sc.ld(1, 32, l0) + sc.ld(0, 32, l1) + sc.ld(1, 0, l1) + c;
    L64 = __shfl_sync(ZKP_FULL, (uint32_t)v64, l1);
  }
  __syncwarp();
  // ---- r = (P - L) mod 2^(32*65), limb 64 kept as a scalar (the same in both lanes)
  const uint32_t p64 = __shfl_sync(ZKP_FULL, hi[0], l0);
  uint32_t d[BL];
  uint32_t borrow = M2::sub_full(d, lo, L, lane);
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
  if (WANT_Q) {
    // q = q^ + corr
    uint32_t co = add_word(qh, g ? 0u : corr);
    const uint32_t co_p = __shfl_xor_sync(ZKP_FULL, co, 1);
    add_word(qh, g ? co_p : 0u);
#pragma unroll
    for (int j = 0; j < BL; ++j) q[j] = qh[j];
  }
}

template <bool WANT_Q>
__device__ __noinline__ void barrett(uint32_t* __restrict__ q_, uint32_t* __restrict__ r_, const uint32_t* __restrict__ lo_,
                                     const uint32_t* __restrict__ hi_, const uint32_t* __restrict__ n_, Scratch sc,
                                     int lane) {
  uint32_t q[BL], r[BL], lo[BL], hi[BL], nreg[BL];
#pragma unroll
  for (int j = 0; j < BL; ++j) {
    lo[j] = lo_[j];
    hi[j] = hi_[j];
    nreg[j] = n_[j];
  }
  barrett_body<WANT_Q>(q, r, lo, hi, nreg, sc, lane);
#pragma unroll
  for (int j = 0; j < BL; ++j) {
    r_[j] = r[j];
    if (WANT_Q) q_[j] = q[j];
  }
}

// x = (x + y) mod n for x + y < 2n... generic: returns x + y reduced by up to `rounds` subtractions of n
template <int ROUNDS>
__device__ __forceinline__ void add_mod(uint32_t (&x)[BL], const uint32_t (&y)[BL], const uint32_t (&nreg)[BL], uint32_t& ovf, int lane) {
  ovf += M2::add_full(x, y, lane);
  uint32_t d[BL];
#pragma unroll
  for (int k = 0; k < ROUNDS; ++k) {
    const uint32_t borrow = M2::sub_full(d, x, nreg, lane);
    const bool take = (ovf != 0u) || (borrow == 0u);
#pragma unroll
    for (int j = 0; j < BL; ++j) x[j] = take ? d[j] : x[j];
    ovf = take ? ovf - borrow : ovf;
  }
}

// (X0, X1) <- (X0, X1)^2
__device__ __forceinline__ void sqr2(uint32_t (&X0)[BL], uint32_t (&X1)[BL], const uint32_t (&nreg)[BL],
                                     const Scratch& sc, int lane) {
  uint32_t lo[BL], hi[BL], Q[BL], R[BL], U[BL], dummy[BL];
  prod_full(lo, hi, X0, X1, sc, lane);          // U = X0 X1 first (needs the old X0)
  barrett<false>(dummy, U, lo, hi, nreg, sc, lane);
  prod_full(lo, hi, X0, X0, sc, lane);          // P = X0^2
  barrett<true>(Q, R, lo, hi, nreg, sc, lane);
  // X1 = (2 U + Q) mod n
  uint32_t ovf = 0;
#pragma unroll
  for (int j = 0; j < BL; ++j) X1[j] = U[j];
  add_mod<1>(X1, U, nreg, ovf, lane);           // 2U < 2n
  add_mod<1>(X1, Q, nreg, ovf, lane);           // < 2n
#pragma unroll
  for (int j = 0; j < BL; ++j) X0[j] = R[j];
}

// (X0, X1) <- (X0, X1) * (Y0, Y1)
__device__ __forceinline__ void mul2(uint32_t (&X0)[BL], uint32_t (&X1)[BL], const uint32_t (&Y0)[BL], const uint32_t (&Y1)[BL],
                                     const uint32_t (&nreg)[BL], const Scratch& sc, int lane) {
  uint32_t lo[BL], hi[BL], Q[BL], R[BL], U[BL], V[BL], dummy[BL];
  prod_full(lo, hi, X0, Y1, sc, lane);
  barrett<false>(dummy, U, lo, hi, nreg, sc, lane);
  prod_full(lo, hi, X1, Y0, sc, lane);
  barrett<false>(dummy, V, lo, hi, nreg, sc, lane);
  prod_full(lo, hi, X0, Y0, sc, lane);
  barrett<true>(Q, R, lo, hi, nreg, sc, lane);
  uint32_t ovf = 0;
#pragma unroll
  for (int j = 0; j < BL; ++j) X1[j] = U[j];
  add_mod<1>(X1, V, nreg, ovf, lane);
  add_mod<1>(X1, Q, nreg, ovf, lane);
#pragma unroll
  for (int j = 0; j < BL; ++j) X0[j] = R[j];
}

__global__ void __launch_bounds__(kThreads, 2) enc2d_kernel(const __grid_constant__ Enc2dParams p) {
  extern __shared__ __align__(16) uint32_t smem[];
  uint32_t* s_sched = smem;
  const int sched_pad = (p.nsteps + 3) & ~3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane & 1;
  Scratch sc{(uint32_t)__cvta_generic_to_shared(smem + sched_pad + warp * kWarpWords)};
  for (int i = threadIdx.x; i < p.nsteps; i += blockDim.x) s_sched[i] = p.sched[i];
  __syncthreads();

  uint32_t nreg[BL];
#pragma unroll
  for (int j = 0; j < BL; ++j) nreg[j] = c_key.n[BL * g + j];

  constexpr int G = kThreads / 2;  // modexps per CTA pass
  const int grp = threadIdx.x >> 1;
  const int jobs = p.jobs_dev ? min((int)*p.jobs_dev, p.jobs) : p.jobs;
  const int npass = (jobs + G - 1) / G;
  uint32_t* tab = p.table + ((size_t)(blockIdx.x * G + grp) * kTableShared) * (2 * DL) + g * DL;
  for (int cj = blockIdx.x; cj < npass; cj += gridDim.x) {
    const int job = cj * G + grp;
    const bool valid = job < jobs;
    const int src = valid ? job : 0;
    uint32_t X0[BL], X1[BL], Y0[BL], Y1[BL];
    M2::load(X0, p.bases + (size_t)src * DL + g * BL);
    {  // r may exceed n (r < W < 2n): X0 = r mod n, X1 = floor(r / n)
      uint32_t d[BL];
      const uint32_t borrow = M2::sub_full(d, X0, nreg, lane);
      const bool take = borrow == 0u;
#pragma unroll
      for (int j = 0; j < BL; ++j) {
        X0[j] = take ? d[j] : X0[j];
        X1[j] = 0;
      }
      if (g == 0) X1[0] = take ? 1u : 0u;
    }
    // odd powers x, x^3, ..., x^31
    M2::store(tab, X0);
    M2::store(tab + BL, X1);
#pragma unroll
    for (int j = 0; j < BL; ++j) {
      Y0[j] = X0[j];
      Y1[j] = X1[j];
    }
    sqr2(Y0, Y1, nreg, sc, lane);  // x^2
#pragma unroll 1
    for (int e = 1; e < kTableShared; ++e) {
      mul2(X0, X1, Y0, Y1, nreg, sc, lane);
      M2::store(tab + e * (2 * DL), X0);
      M2::store(tab + e * (2 * DL) + BL, X1);
    }
    uint32_t st = s_sched[0];
    M2::load(X0, tab + (st & 0xffu) * (2 * DL));
    M2::load(X1, tab + (st & 0xffu) * (2 * DL) + BL);
#pragma unroll 1
    for (int k = 1; k < p.nsteps; ++k) {
      st = s_sched[k];
      const uint32_t idx = st & 0xffu;
      const int nsq = (int)(st >> 8);
#pragma unroll 1
      for (int q = 0; q < nsq; ++q) sqr2(X0, X1, nreg, sc, lane);
      if (idx != 0xffu) {
        M2::load(Y0, tab + idx * (2 * DL));
        M2::load(Y1, tab + idx * (2 * DL) + BL);
        mul2(X0, X1, Y0, Y1, nreg, sc, lane);
      }
    }
    // c = (1 + m n) x = X0 + (X1 + m X0 mod n) n   (mod n^2)
    if (p.plain) {
      uint32_t m[BL], lo[BL], hi[BL], T[BL], dummy[BL];
      M2::load_ext(m, p.plain + (size_t)src * p.plain_limbs, p.plain_limbs, g);
      prod_full(lo, hi, m, X0, sc, lane);
      barrett<false>(dummy, T, lo, hi, nreg, sc, lane);
      uint32_t ovf = 0;
      add_mod<1>(X1, T, nreg, ovf, lane);
    }
    {
      uint32_t lo[BL], hi[BL];
      prod_full(lo, hi, X1, nreg, sc, lane);  // X1 * n
      const uint32_t c = M2::add_full(lo, X0, lane);
      uint32_t co = add_word(hi, g ? 0u : c);
      const uint32_t co_p = __shfl_xor_sync(ZKP_FULL, co, 1);
      add_word(hi, g ? co_p : 0u);
      if (valid) {
        M2::store(p.out + (size_t)job * (2 * DL) + g * BL, lo);
        M2::store(p.out + (size_t)job * (2 * DL) + DL + g * BL, hi);
      }
    }
  }
}

}  // namespace v2

// Host side: mu' = floor(2^4096 / n) - 2^2048 by schoolbook long division (once per key).
static void barrett_mu(const uint32_t* n, uint32_t* mu) {
  // numerator 2^4096 as 129 limbs; divide by the 64-limb n (top bit set) with 64-bit partial remainders:
  // simple bitwise restoring division is fast enough once per key (4097 iterations x 65 limbs).
  uint32_t rem[66] = {0};
  uint32_t quo[130] = {0};
  for (int bit = 4096; bit >= 0; --bit) {
    // rem = rem * 2 + numerator bit (only bit 4096 is set)
    uint32_t carry = (bit == 4096) ? 1u : 0u;
    for (int i = 0; i < 66; ++i) {
      uint32_t nc = rem[i] >> 31;
      rem[i] = (rem[i] << 1) | carry;
      carry = nc;
    }
    // if rem >= n: rem -= n, quotient bit = 1
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
  // quo = floor(2^4096 / n) = 2^2048 + mu'
  for (int i = 0; i < 64; ++i) mu[i] = quo[i];
}

bool enc2d_supported(const uint32_t* n_host, int n_limbs_exact) {
  return n_limbs_exact == v2::DL && (n_host[v2::DL - 1] >> 31) == 1u && (n_host[0] & 1u);
}

constexpr int kEnc2dCtasPerSm = 2;
int enc2d_resident_groups(int num_sms) { return num_sms * kEnc2dCtasPerSm * (v2::kThreads / 2); }

cudaError_t launch_enc2d(const uint32_t* n_host, const uint32_t* sched_dev, int nsteps, const uint32_t* bases, const uint32_t* plain,
                         int plain_limbs, uint32_t* out, int jobs, uint32_t* table, int num_sms, cudaStream_t st,
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
  p.table = table;
  p.jobs = jobs;
  p.jobs_dev = jobs_dev;
  constexpr int G = v2::kThreads / 2;
  int grid = num_sms * kEnc2dCtasPerSm;
  const int npass = (jobs + G - 1) / G;
  if (grid > npass) grid = npass;
  const size_t smem = ((size_t)((nsteps + 3) & ~3) + 4 * v2::kWarpWords) * 4;
  e = cudaFuncSetAttribute(v2::enc2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  v2::enc2d_kernel<<<grid, v2::kThreads, smem, st>>>(p);
  return cudaGetLastError();
}

}  // namespace zkp
