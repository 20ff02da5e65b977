// Cooperative fixed-width multiword arithmetic for sm_100a.
//
// One big integer of S = T*L 32-bit limbs is spread over a group of T adjacent
// lanes of a warp (T in {4,8,16,32}); lane g of the group holds limbs
// [g*L, (g+1)*L) in registers.  All groups of a warp run in lock-step, so every
// shuffle/ballot is issued with the full-warp mask and a sub-warp width of T.
//
// The arithmetic this replaces in the reference is GMP's mpz_powm / mpz_mul /
// mpz_mod underneath curv-kzen's BigInt::{mod_pow, mod_mul} and kzen-paillier's
// Paillier::encrypt_with_chosen_randomness (call sites: reference
// src/zkproofs/range_proof.rs:161-187,280-291,325-334; correct_key_ni.rs:90-93).
//
// Machine unit: the 32x32+64 -> 64 multiply-add.  A mad.lo.cc.u32 / madc.hi.cc.u32
// pair on the same operands is fused by ptxas into ONE IMAD.WIDE.U32(.X) with
// carry-in/out in a predicate.  IMAD.WIDE needs 64-bit aligned register pairs, so
// the accumulator is kept as TWO arrays: one takes the even-limb products (pairs
// at limbs 2m,2m+1), the other the odd-limb products (pairs at 2m+1,2m+2).  The
// per-step division by 2^32 swaps their roles, and the stale array is moved down
// one aligned pair through the 3-operand form d = a*b + c, so a CIOS step is
// 2L IMAD.WIDE + 3 SHFL + a handful of adds, with no register moves.
#pragma once
#include <stdint.h>

#include "blockmul.cuh"

namespace zkp {

#define ZKP_FULL 0xffffffffu

// ---- single-instruction carry-chain helpers (PTX) -------------------------
__device__ __forceinline__ void add_cc(uint32_t& r, uint32_t a) {
  asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(r) : "r"(a));
}
__device__ __forceinline__ void addc_cc(uint32_t& r, uint32_t a) {
  asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(r) : "r"(a));
}
__device__ __forceinline__ void addc(uint32_t& r, uint32_t a) {
  asm volatile("addc.u32 %0, %0, %1;" : "+r"(r) : "r"(a));
}
__device__ __forceinline__ void sub_cc(uint32_t& r, uint32_t a, uint32_t b) {
  asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
}
__device__ __forceinline__ void subc_cc(uint32_t& r, uint32_t a, uint32_t b) {
  asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
}
__device__ __forceinline__ uint32_t subc_out() {  // 0 if no borrow, 0xffffffff if borrow
  uint32_t r;
  asm volatile("subc.u32 %0, 0, 0;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t addc_out() {  // carry flag -> 0/1
  uint32_t r;
  asm volatile("addc.u32 %0, 0, 0;" : "=r"(r));
  return r;
}
// (hi:lo) += a*b, starting a carry chain
__device__ __forceinline__ void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;"
               : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
// (hi:lo) += a*b + carry, continuing a carry chain
__device__ __forceinline__ void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;"
               : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}

// (d1:d0) = a*b + (c1:c0) + carry, continuing a carry chain; the destination pair
// may differ from the addend pair, which is how the accumulator is shifted down
// by one aligned limb pair for free.
__device__ __forceinline__ void madc_wide3_cc(uint32_t& d0, uint32_t& d1, uint32_t a, uint32_t b, uint32_t c0,
                                              uint32_t c1) {
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %4;\n\tmadc.hi.cc.u32 %1, %2, %3, %5;"
               : "=&r"(d0), "=r"(d1) : "r"(a), "r"(b), "r"(c0), "r"(c1));
}

// (d1:d0) = a*b + (c1:c0), STARTING a carry chain, destination pair different from the addend pair
__device__ __forceinline__ void mad_wide3_start_cc(uint32_t& d0, uint32_t& d1, uint32_t a, uint32_t b, uint32_t c0, uint32_t c1) {
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %4;\n\tmadc.hi.cc.u32 %1, %2, %3, %5;"
               : "=&r"(d0), "=r"(d1) : "r"(a), "r"(b), "r"(c0), "r"(c1));
}
// r = a + b + carry, continuing a carry chain
__device__ __forceinline__ void addc_cc3(uint32_t& r, uint32_t a, uint32_t b) {
  asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
}

template <int T, int L>
struct Mp {
  static_assert(T == 2 || T == 4 || T == 8 || T == 16 || T == 32, "group width");
  static_assert(L >= 2 && (L % 2) == 0, "limbs per lane must be even");
  static constexpr int S = T * L;  // limbs per integer
  static constexpr uint32_t GM = (T == 32) ? 0xffffffffu : ((1u << T) - 1u);

  // Group-wide resolution of per-lane carries (or borrows).  `gen` = this lane
  // produced a carry out, `prop` = this lane would pass an incoming carry on
  // (gen and prop are never both set).  Returns the carry INTO this lane and,
  // in `top`, the carry out of the most significant lane of the group.
  static __device__ __forceinline__ uint32_t resolve(bool gen, bool prop, int lane, uint32_t& top) {
    uint32_t gb = __ballot_sync(ZKP_FULL, gen);
    uint32_t pb = __ballot_sync(ZKP_FULL, prop);
    int base = lane & ~(T - 1);
    int g = lane & (T - 1);
    if constexpr (T == 32) {
      uint64_t a = (uint64_t)(gb | pb), b = (uint64_t)gb;
      uint64_t s = a + b;
      uint64_t cin = s ^ a ^ b;
      top = (uint32_t)(s >> 32) & 1u;
      return (uint32_t)(cin >> g) & 1u;
    } else {
      uint32_t G = (gb >> base) & GM, P = (pb >> base) & GM;
      uint32_t a = G | P;
      uint32_t s = a + G;
      uint32_t cin = s ^ a ^ G;
      top = (s >> T) & 1u;
      return (cin >> g) & 1u;
    }
  }

  static __device__ __forceinline__ bool all_ones(const uint32_t (&x)[L]) {
    uint32_t m = x[0];
#pragma unroll
    for (int j = 1; j < L; ++j) m &= x[j];
    return m == 0xffffffffu;
  }
  static __device__ __forceinline__ bool all_zero(const uint32_t (&x)[L]) {
    uint32_t m = x[0];
#pragma unroll
    for (int j = 1; j < L; ++j) m |= x[j];
    return m == 0u;
  }

  // x += c (c in {0,1}) rippled through this lane's limbs only.
  static __device__ __forceinline__ void add_small(uint32_t (&x)[L], uint32_t c) {
    add_cc(x[0], c);
#pragma unroll
    for (int j = 1; j < L - 1; ++j) addc_cc(x[j], 0);
    addc(x[L - 1], 0);
  }
  static __device__ __forceinline__ void sub_small(uint32_t (&x)[L], uint32_t c) {
    uint32_t t;
    sub_cc(t, x[0], c);
    x[0] = t;
#pragma unroll
    for (int j = 1; j < L; ++j) {
      subc_cc(t, x[j], 0);
      x[j] = t;
    }
  }

  // d = a - b over the whole group. Returns the final borrow (1 iff a < b).
  static __device__ __forceinline__ uint32_t sub_full(uint32_t (&d)[L], const uint32_t (&a)[L], const uint32_t (&b)[L], int lane) {
    sub_cc(d[0], a[0], b[0]);
#pragma unroll
    for (int j = 1; j < L; ++j) subc_cc(d[j], a[j], b[j]);
    uint32_t bo = subc_out() & 1u;
    uint32_t top;
    uint32_t bin = resolve(bo != 0, all_zero(d), lane, top);
    sub_small(d, bin);
    return top;
  }

  // a += b over the whole group. Returns the carry out of the top lane.
  static __device__ __forceinline__ uint32_t add_full(uint32_t (&a)[L], const uint32_t (&b)[L], int lane) {
    add_cc(a[0], b[0]);
#pragma unroll
    for (int j = 1; j < L; ++j) addc_cc(a[j], b[j]);
    uint32_t co = addc_out();
    uint32_t top;
    uint32_t cin = resolve(co != 0, all_ones(a), lane, top);
    add_small(a, cin);
    return top;
  }

  // 1 iff a >= b (group-wide unsigned compare)
  static __device__ __forceinline__ uint32_t geq(const uint32_t (&a)[L], const uint32_t (&b)[L], int lane) {
    uint32_t d[L];
    return sub_full(d, a, b, lane) ^ 1u;
  }

  // Group-wide equality; result uniform in the group.
  static __device__ __forceinline__ bool equal(const uint32_t (&a)[L], const uint32_t (&b)[L], int lane) {
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < L; ++j) m |= a[j] ^ b[j];
    uint32_t nb = __ballot_sync(ZKP_FULL, m != 0);
    int base = lane & ~(T - 1);
    return ((nb >> base) & GM) == 0u;
  }

  // Turn the redundant accumulator (L+1 significant limbs per lane: u[L] is the
  // lane's overflow, weight 2^(32L)) into the canonical residue r in [0, n):
  // push overflows one lane up, resolve carries, subtract n once if needed.
  // Pre-condition: the represented value is < 2n.
  static __device__ __forceinline__ void finish(uint32_t (&r)[L], uint32_t (&u)[L + 2], const uint32_t (&n)[L], int lane) {
    const int g = lane & (T - 1);
    uint32_t ov = __shfl_up_sync(ZKP_FULL, u[L], 1, T);
    if (g == 0) ov = 0;
    uint32_t x[L];
#pragma unroll
    for (int j = 0; j < L; ++j) x[j] = u[j];
    add_cc(x[0], ov);
#pragma unroll
    for (int j = 1; j < L; ++j) addc_cc(x[j], 0);
    uint32_t co = addc_out();
    uint32_t topc;
    uint32_t cin = resolve(co != 0, all_ones(x), lane, topc);
    add_small(x, cin);
    // value >= 2^(32 S) ?  (top lane's own overflow limb, plus the resolved carry)
    uint32_t hi = __shfl_sync(ZKP_FULL, u[L], T - 1, T) + topc;
    uint32_t d[L];
    uint32_t borrow = sub_full(d, x, n, lane);
    bool take = (hi != 0) || (borrow == 0);
#pragma unroll
    for (int j = 0; j < L; ++j) r[j] = take ? d[j] : x[j];
  }

  // X[0..L+1] += (a[0], a[2], ...) * b : the even-limb products, one carry chain
  static __device__ __forceinline__ void mad_even(uint32_t (&X)[L + 2], const uint32_t (&a)[L], uint32_t b) {
    mad_wide_cc(X[0], X[1], a[0], b);
#pragma unroll
    for (int j = 2; j < L; j += 2) madc_wide_cc(X[j], X[j + 1], a[j], b);
    addc_cc(X[L], 0);
    addc(X[L + 1], 0);
  }
  // Z[0..L+1] += (a[1], a[3], ...) * b : the odd-limb products (Z[w] has weight 2^(32(w+1)))
  static __device__ __forceinline__ void mad_odd(uint32_t (&Z)[L + 2], const uint32_t (&a)[L], uint32_t b) {
    mad_wide_cc(Z[0], Z[1], a[1], b);
#pragma unroll
    for (int j = 2; j < L; j += 2) madc_wide_cc(Z[j], Z[j + 1], a[j + 1], b);
    addc_cc(Z[L], 0);
    addc(Z[L + 1], 0);
  }

  // One CIOS step  acc = (acc + a*b + q*n) / 2^32  on the split accumulator.
  // On entry X is the even-aligned array (X[w] at limb w) and Y is the array
  // that was even-aligned one step ago (Y[w] at limb w-1, Y[0] belongs to the
  // lane below).  On exit Y is the even-aligned array and X the stale one, so
  // callers alternate step(X,Y) / step(Y,X).  Both operand-carrying IMAD.WIDE
  // chains read/write 64-bit aligned register pairs only: no MOVs.
  static __device__ __forceinline__ void cios_step(uint32_t (&X)[L + 2], uint32_t (&Y)[L + 2], const uint32_t (&a)[L],
                                                   const uint32_t (&n)[L], uint32_t b, uint32_t n0inv, int g) {
    uint32_t q;
    cios_step(X, Y, a, n, b, n0inv, g, q);
  }
  // Same step, handing back the quotient digit q it added (uniform across the group).
  static __device__ __forceinline__ void cios_step(uint32_t (&X)[L + 2], uint32_t (&Y)[L + 2], const uint32_t (&a)[L],
                                                   const uint32_t (&n)[L], uint32_t b, uint32_t n0inv, int g, uint32_t& q,
                                                   uint32_t zr = 0u) {
    // zr: a zero the compiler cannot see through.  With a literal 0 ptxas turns the carry add below into IMAD.X,
    // which competes with the IMAD.WIDE rows for the multiplier pipe; with a register it stays an IADD3.X.
    uint32_t in = __shfl_down_sync(ZKP_FULL, Y[0], 1, T);
    if (g == T - 1) in = 0;
    add_cc(Y[L], in);
    addc(Y[L + 1], zr);
    uint32_t Z[L + 2];
    add_cc(X[0], Y[1]);  // limb 0; the carry enters the odd chain at limb 1
#pragma unroll
    for (int j = 0; j < L; j += 2) madc_wide3_cc(Z[j], Z[j + 1], a[j + 1], b, Y[j + 2], Y[j + 3]);
    Z[L] = addc_out();
    Z[L + 1] = 0;
    mad_even(X, a, b);
    q = __shfl_sync(ZKP_FULL, X[0], 0, T) * n0inv;
    mad_even(X, n, q);
    mad_odd(Z, n, q);
#pragma unroll
    for (int j = 0; j < L + 2; ++j) Y[j] = Z[j];
  }

  // r = a * b * 2^(-32 S) mod n     (CIOS Montgomery; a, b < n; n odd;
  // n0inv = -n^{-1} mod 2^32).  r may alias a or b.
  static __device__ __forceinline__ void mont_mul(uint32_t (&r)[L], const uint32_t (&a)[L], const uint32_t (&b)[L],
                                                  const uint32_t (&n)[L], uint32_t n0inv, int lane) {
    const int g = lane & (T - 1);
    uint32_t E[L + 2], O[L + 2];
#pragma unroll
    for (int j = 0; j < L + 2; ++j) E[j] = O[j] = 0;
#pragma unroll 1
    for (int owner = 0; owner < T; ++owner) {
#pragma unroll
      for (int j = 0; j < L; j += 2) {
        uint32_t b0 = __shfl_sync(ZKP_FULL, b[j], owner, T);
        uint32_t b1 = __shfl_sync(ZKP_FULL, b[j + 1], owner, T);
        cios_step(E, O, a, n, b0, n0inv, g);
        cios_step(O, E, a, n, b1, n0inv, g);
      }
    }
    // merge: value = E + (O >> 32), O[0] going to the lane below
    uint32_t in = __shfl_down_sync(ZKP_FULL, O[0], 1, T);
    if (g == T - 1) in = 0;
    add_cc(O[L], in);
    addc(O[L + 1], 0);
    add_cc(E[0], O[1]);
#pragma unroll
    for (int j = 1; j <= L; ++j) addc_cc(E[j], O[j + 1]);
    addc(E[L + 1], 0);
    finish(r, E, n, lane);
  }

  // finish() for a value < NSUB*n + 2 (NSUB = 1: the usual < 2n): NSUB conditional subtractions, the overflow above
  // 2^(32 S) tracked across them.  Returns 1 iff the first subtraction was taken (uniform across the group).
  template <int NSUB>
  static __device__ __forceinline__ uint32_t finish_x(uint32_t (&r)[L], uint32_t (&u)[L + 2], const uint32_t (&n)[L], int lane) {
    const int g = lane & (T - 1);
    uint32_t ov = __shfl_up_sync(ZKP_FULL, u[L], 1, T);
    if (g == 0) ov = 0;
#pragma unroll
    for (int j = 0; j < L; ++j) r[j] = u[j];
    add_cc(r[0], ov);
#pragma unroll
    for (int j = 1; j < L; ++j) addc_cc(r[j], 0);
    uint32_t co = addc_out();
    uint32_t topc;
    uint32_t cin = resolve(co != 0, all_ones(r), lane, topc);
    add_small(r, cin);
    uint32_t hi = __shfl_sync(ZKP_FULL, u[L], T - 1, T) + topc;
    uint32_t first = 0;
#ifdef ZKP_B200_LAB_NOSUB
    // lab probe only (make lab LABTAG=_nosub LABFLAGS=-DZKP_B200_LAB_NOSUB): WRONG RESULTS - no conditional subtraction at
    // all, to time the upper bound of what a lazy (redundant-residue) reduction could save
    (void)hi;
    return first;
#endif
#pragma unroll
    for (int s = 0; s < NSUB; ++s) {
      uint32_t d[L];
      const uint32_t borrow = sub_full(d, r, n, lane);
      const bool take = (hi != 0) || (borrow == 0);
#pragma unroll
      for (int j = 0; j < L; ++j) r[j] = take ? d[j] : r[j];
      hi = take ? hi - borrow : hi;
      if (s == 0) first = take ? 1u : 0u;
    }
    return first;
  }

  // Montgomery multiplication with the two hooks the two-digit base-n form needs (modexp2m.cu):
  //   INIT : the accumulator starts at init + init_top * 2^(32 S) instead of 0
  //          (r = (init + a b) / 2^(32 S) mod n; init < 2^(32 S) + n, so r needs NSUB = 2);
  //   CAPQ : the quotient digits q_i of the reduction are kept, digit i in the lane that owns limb i, so that
  //          a b = r' 2^(32 S) - q n holds exactly for the unreduced r' (r = r' - n iff the return value is 1).
  // Only a b < 2^(32 S) n is needed (a may be any S-limb value when b < n).  r may alias a or b.
  template <bool INIT, bool CAPQ, int NSUB, int U = 1>
  static __device__ __forceinline__ uint32_t mont_mul_x(uint32_t (&r)[L], const uint32_t (&a)[L], const uint32_t (&b)[L],
                                                        const uint32_t (&n)[L], uint32_t n0inv, int lane,
                                                        const uint32_t (&init)[L], uint32_t init_top, uint32_t (&qcap)[L],
                                                        uint32_t zr = 0u) {
    const int g = lane & (T - 1);
    uint32_t E[L + 2], O[L + 2];
#pragma unroll
    for (int j = 0; j < L + 2; ++j) E[j] = O[j] = 0;
    if (INIT) {
#pragma unroll
      for (int j = 0; j < L; ++j) E[j] = init[j];
      E[L] = (g == T - 1) ? init_top : 0u;
    }
#pragma unroll U
    for (int owner = 0; owner < T; ++owner) {
      const bool mine = CAPQ && (g == owner);
#pragma unroll
      for (int j = 0; j < L; j += 2) {
        uint32_t b0 = __shfl_sync(ZKP_FULL, b[j], owner, T);
        uint32_t b1 = __shfl_sync(ZKP_FULL, b[j + 1], owner, T);
        uint32_t q0, q1;
        cios_step(E, O, a, n, b0, n0inv, g, q0, zr);
        cios_step(O, E, a, n, b1, n0inv, g, q1, zr);
        if (CAPQ) {
          qcap[j] = mine ? q0 : qcap[j];
          qcap[j + 1] = mine ? q1 : qcap[j + 1];
        }
      }
    }
    uint32_t in = __shfl_down_sync(ZKP_FULL, O[0], 1, T);
    if (g == T - 1) in = 0;
    add_cc(O[L], in);
    addc(O[L + 1], 0);
    add_cc(E[0], O[1]);
#pragma unroll
    for (int j = 1; j <= L; ++j) addc_cc(E[j], O[j + 1]);
    addc(E[L + 1], 0);
    return finish_x<NSUB>(r, E, n, lane);
  }

  // One step of the two-product form  acc = (acc + a1*b1 + a2*b2 + q*n) / 2^32  (the second digit of a two-digit
  // multiplication needs X0 Y1 + X1 Y0 under ONE reduction).  Same array roles as cios_step.
  static __device__ __forceinline__ void cios_step2(uint32_t (&X)[L + 2], uint32_t (&Y)[L + 2], const uint32_t (&a1)[L],
                                                    const uint32_t (&a2)[L], const uint32_t (&n)[L], uint32_t b1, uint32_t b2,
                                                    uint32_t n0inv, int g, uint32_t zr = 0u) {
    uint32_t in = __shfl_down_sync(ZKP_FULL, Y[0], 1, T);
    if (g == T - 1) in = 0;
    add_cc(Y[L], in);
    addc(Y[L + 1], zr);
    uint32_t Z[L + 2];
    add_cc(X[0], Y[1]);
#pragma unroll
    for (int j = 0; j < L; j += 2) madc_wide3_cc(Z[j], Z[j + 1], a1[j + 1], b1, Y[j + 2], Y[j + 3]);
    Z[L] = addc_out();
    Z[L + 1] = 0;
    mad_odd(Z, a2, b2);
    mad_even(X, a1, b1);
    mad_even(X, a2, b2);
    const uint32_t q = __shfl_sync(ZKP_FULL, X[0], 0, T) * n0inv;
    mad_even(X, n, q);
    mad_odd(Z, n, q);
#pragma unroll
    for (int j = 0; j < L + 2; ++j) Y[j] = Z[j];
  }

  // r = (init + init_top 2^(32 S) + a1 b1 + a2 b2) / 2^(32 S) mod n   (each product < 2^(32 S) n, init < 2^(32 S) + n:
  // the unreduced value is < 3n + 2).  r may alias any operand.
  template <int U = 1>
  static __device__ __forceinline__ void mont_mul2_x(uint32_t (&r)[L], const uint32_t (&a1)[L], const uint32_t (&b1)[L],
                                                     const uint32_t (&a2)[L], const uint32_t (&b2)[L], const uint32_t (&n)[L],
                                                     uint32_t n0inv, int lane, const uint32_t (&init)[L], uint32_t init_top,
                                                     uint32_t zr = 0u) {
    const int g = lane & (T - 1);
    uint32_t E[L + 2], O[L + 2];
#pragma unroll
    for (int j = 0; j < L; ++j) {
      E[j] = init[j];
      O[j] = 0;
    }
    E[L] = (g == T - 1) ? init_top : 0u;
    E[L + 1] = O[L] = O[L + 1] = 0;
#pragma unroll U
    for (int owner = 0; owner < T; ++owner) {
#pragma unroll
      for (int j = 0; j < L; j += 2) {
        const uint32_t p0 = __shfl_sync(ZKP_FULL, b1[j], owner, T);
        const uint32_t s0 = __shfl_sync(ZKP_FULL, b2[j], owner, T);
        const uint32_t p1 = __shfl_sync(ZKP_FULL, b1[j + 1], owner, T);
        const uint32_t s1 = __shfl_sync(ZKP_FULL, b2[j + 1], owner, T);
        cios_step2(E, O, a1, a2, n, p0, s0, n0inv, g, zr);
        cios_step2(O, E, a1, a2, n, p1, s1, n0inv, g, zr);
      }
    }
    uint32_t in = __shfl_down_sync(ZKP_FULL, O[0], 1, T);
    if (g == T - 1) in = 0;
    add_cc(O[L], in);
    addc(O[L + 1], 0);
    add_cc(E[0], O[1]);
#pragma unroll
    for (int j = 1; j <= L; ++j) addc_cc(E[j], O[j + 1]);
    addc(E[L + 1], 0);
    finish_x<3>(r, E, n, lane);
  }

  // ---- pair rows (model: tests/models/cios_pair_model.py) ------------------------------------------------------------
  // Two multiplier limbs and a TWO-limb quotient per step: acc = (acc + a (b0 + b1 B) + (q0 + q1 B) n) / B^2, B = 2^32, with
  // q0 + q1 B = (low 64 bits of acc + a (b0 + b1 B)) n' mod B^2 (n' = -n^-1 mod B^2): the same quotient digits as two single
  // rows, but the serial chain through the quotient - broadcast, multiply, first products - runs once per TWO rows.  That is
  // what a launch with few warps per sub-partition waits for (one proof's latency; the narrow-lane layouts of K2h); for the
  // pipe-saturated wide rows of K1m it would only add multiplies.
  // Arrays: E[w] at lane-local limb w (w <= L + 2), O[w] at limb w + 1 (w <= L + 1).  a[2m] b0 -> E pair m, a[2m+1] b0 -> O pair
  // m, a[2m] b1 -> O pair m, a[2m+1] b1 -> E pair m + 1.  Dividing by B^2 shifts BOTH arrays by one aligned pair - done by the
  // first chain of each array writing to the shifted destination - plus one bridge: O[1] (limb 2) lands on limb 0 and is added
  // into E there with the carry of the vanishing limb 1, and the bridge's carry enters the first odd chain.  The two lowest
  // limbs of every lane but lane 0 travel to the lane below.  The shift of a step is folded into the NEXT step; pair_finish
  // does the last one.
  //
  // four in-place product chains of one (x, y0, y1): the n q products, and the second operand pair of the two-product form
  static __device__ __forceinline__ void pair_mad(uint32_t (&E)[L + 3], uint32_t (&O)[L + 2], const uint32_t (&x)[L], uint32_t y0, uint32_t y1,
                                                  uint32_t zr) {
    mad_wide_cc(E[0], E[1], x[0], y0);  // even A
#pragma unroll
    for (int j = 2; j < L; j += 2) madc_wide_cc(E[j], E[j + 1], x[j], y0);
    addc_cc(E[L], zr);
    addc_cc(E[L + 1], zr);
    addc(E[L + 2], zr);
    mad_wide_cc(O[0], O[1], x[1], y0);  // odd A
#pragma unroll
    for (int j = 2; j < L; j += 2) madc_wide_cc(O[j], O[j + 1], x[j + 1], y0);
    addc_cc(O[L], zr);
    addc(O[L + 1], zr);
    mad_wide_cc(O[0], O[1], x[0], y1);  // odd B
#pragma unroll
    for (int j = 2; j < L; j += 2) madc_wide_cc(O[j], O[j + 1], x[j], y1);
    addc_cc(O[L], zr);
    addc(O[L + 1], zr);
    mad_wide_cc(E[2], E[3], x[1], y1);  // even B: one pair up
#pragma unroll
    for (int j = 2; j < L; j += 2) madc_wide_cc(E[j + 2], E[j + 3], x[j + 1], y1);
    addc(E[L + 2], zr);
  }
  // One pair step.  TWO: a second product pair (a2, c0, c1) under the same reduction (mont_mul2_x).
  template <bool TWO>
  static __device__ __forceinline__ void cios_pair(uint32_t (&E)[L + 3], uint32_t (&O)[L + 2], const uint32_t (&a)[L], const uint32_t (&a2)[L],
                                                   const uint32_t (&n)[L], uint32_t b0, uint32_t b1, uint32_t c0, uint32_t c1,
                                                   uint32_t np0, uint32_t np1, int g, uint32_t& q0, uint32_t& q1, uint32_t zr) {
    // the shift of the previous step: limb 1 vanishes (s1 goes to the lane below), the bridge, and its carry into odd chain A
    uint32_t s1 = E[1];
    add_cc(s1, O[0]);
    addc_cc(E[2], O[1]);
    uint32_t NO[L + 2], NE[L + 3];
#pragma unroll
    for (int j = 0; j < L; j += 2) madc_wide3_cc(NO[j], NO[j + 1], a[j + 1], b0, O[j + 2], O[j + 3]);  // odd A, shifted
    NO[L] = addc_out();
    NO[L + 1] = 0;
    uint32_t r0 = __shfl_down_sync(ZKP_FULL, E[0], 1, T), r1 = __shfl_down_sync(ZKP_FULL, s1, 1, T);
    if (g == T - 1) r0 = r1 = 0;
    add_cc(E[L], r0);
    addc_cc(E[L + 1], r1);
    addc(E[L + 2], zr);
    mad_wide3_start_cc(NE[0], NE[1], a[0], b0, E[2], E[3]);  // even A, shifted
#pragma unroll
    for (int j = 2; j < L; j += 2) madc_wide3_cc(NE[j], NE[j + 1], a[j], b0, E[j + 2], E[j + 3]);
    addc_cc3(NE[L], E[L + 2], zr);
    NE[L + 1] = addc_out();
    NE[L + 2] = 0;
    mad_wide_cc(NO[0], NO[1], a[0], b1);  // odd B
#pragma unroll
    for (int j = 2; j < L; j += 2) madc_wide_cc(NO[j], NO[j + 1], a[j], b1);
    addc_cc(NO[L], zr);
    addc(NO[L + 1], zr);
    mad_wide_cc(NE[2], NE[3], a[1], b1);  // even B
#pragma unroll
    for (int j = 2; j < L; j += 2) madc_wide_cc(NE[j + 2], NE[j + 3], a[j + 1], b1);
    addc(NE[L + 2], zr);
    if (TWO) pair_mad(NE, NO, a2, c0, c1, zr);
    // the two quotient digits from lane 0's low 64 bits
    const uint32_t t0 = __shfl_sync(ZKP_FULL, NE[0], 0, T);
    const uint32_t t1 = __shfl_sync(ZKP_FULL, NE[1] + NO[0], 0, T);
    q0 = t0 * np0;
    q1 = __umulhi(t0, np0) + t0 * np1 + t1 * np0;
    pair_mad(NE, NO, n, q0, q1, zr);
#pragma unroll
    for (int j = 0; j < L + 3; ++j) E[j] = NE[j];
#pragma unroll
    for (int j = 0; j < L + 2; ++j) O[j] = NO[j];
  }
  // The shift of the last step, then the two arrays folded into the redundant accumulator finish_x takes (u[L] = the lane's overflow).
  static __device__ __forceinline__ void pair_finish(uint32_t (&u)[L + 2], uint32_t (&E)[L + 3], uint32_t (&O)[L + 2], int g) {
    uint32_t s1 = E[1];
    add_cc(s1, O[0]);
    addc_cc(E[2], O[1]);
#pragma unroll
    for (int j = 2; j < L + 1; ++j) addc_cc(O[j], 0);
    addc(O[L + 1], 0);
    uint32_t r0 = __shfl_down_sync(ZKP_FULL, E[0], 1, T), r1 = __shfl_down_sync(ZKP_FULL, s1, 1, T);
    if (g == T - 1) r0 = r1 = 0;
    add_cc(E[L], r0);
    addc_cc(E[L + 1], r1);
    addc(E[L + 2], 0);
    u[0] = E[2];  // new limb w = old limb w + 2 = E[w + 2] + O[w + 1]
    u[1] = E[3];
    add_cc(u[1], O[2]);
#pragma unroll
    for (int j = 2; j <= L; ++j) {
      u[j] = E[j + 2];
      addc_cc(u[j], O[j + 1]);
    }
    u[L + 1] = addc_out();
  }
  // mont_mul_x by pair rows (same contract: INIT, CAPQ, NSUB; returns 1 iff the first subtraction was taken)
  template <bool INIT, bool CAPQ, int NSUB, int U = 1>
  static __device__ __forceinline__ uint32_t mont_mul_p(uint32_t (&r)[L], const uint32_t (&a)[L], const uint32_t (&b)[L], const uint32_t (&n)[L],
                                                        uint32_t np0, uint32_t np1, int lane, const uint32_t (&init)[L], uint32_t init_top,
                                                        uint32_t (&qcap)[L], uint32_t zr = 0u) {
    const int g = lane & (T - 1);
    uint32_t E[L + 3], O[L + 2];
#pragma unroll
    for (int j = 0; j < L + 3; ++j) E[j] = 0;
#pragma unroll
    for (int j = 0; j < L + 2; ++j) O[j] = 0;
    // the first step shifts "the previous step": start two limbs up (E[w + 2] = limb w), as if a step had just ended
    if (INIT) {
#pragma unroll
      for (int j = 0; j < L; ++j) E[j + 2] = init[j];
      E[L + 2] = (g == T - 1) ? init_top : 0u;
    }
#pragma unroll U
    for (int owner = 0; owner < T; ++owner) {
      const bool mine = CAPQ && (g == owner);
#pragma unroll
      for (int j = 0; j < L; j += 2) {
        const uint32_t b0 = __shfl_sync(ZKP_FULL, b[j], owner, T);
        const uint32_t b1 = __shfl_sync(ZKP_FULL, b[j + 1], owner, T);
        uint32_t q0, q1;
        cios_pair<false>(E, O, a, a, n, b0, b1, 0u, 0u, np0, np1, g, q0, q1, zr);
        if (CAPQ) {
          qcap[j] = mine ? q0 : qcap[j];
          qcap[j + 1] = mine ? q1 : qcap[j + 1];
        }
      }
    }
    uint32_t u[L + 2];
    pair_finish(u, E, O, g);
    return finish_x<NSUB>(r, u, n, lane);
  }
  // mont_mul2_x by pair rows
  template <int U = 1>
  static __device__ __forceinline__ void mont_mul2_p(uint32_t (&r)[L], const uint32_t (&a1)[L], const uint32_t (&b1)[L], const uint32_t (&a2)[L],
                                                     const uint32_t (&b2)[L], const uint32_t (&n)[L], uint32_t np0, uint32_t np1, int lane,
                                                     const uint32_t (&init)[L], uint32_t init_top, uint32_t zr = 0u) {
    const int g = lane & (T - 1);
    uint32_t E[L + 3], O[L + 2];
    E[0] = E[1] = 0;
#pragma unroll
    for (int j = 0; j < L; ++j) E[j + 2] = init[j];
    E[L + 2] = (g == T - 1) ? init_top : 0u;
#pragma unroll
    for (int j = 0; j < L + 2; ++j) O[j] = 0;
#pragma unroll U
    for (int owner = 0; owner < T; ++owner) {
#pragma unroll
      for (int j = 0; j < L; j += 2) {
        const uint32_t p0 = __shfl_sync(ZKP_FULL, b1[j], owner, T);
        const uint32_t p1 = __shfl_sync(ZKP_FULL, b1[j + 1], owner, T);
        const uint32_t s0 = __shfl_sync(ZKP_FULL, b2[j], owner, T);
        const uint32_t s1 = __shfl_sync(ZKP_FULL, b2[j + 1], owner, T);
        uint32_t q0, q1;
        cios_pair<true>(E, O, a1, a2, n, p0, p1, s0, s1, np0, np1, g, q0, q1, zr);
      }
    }
    uint32_t u[L + 2];
    pair_finish(u, E, O, g);
    finish_x<3>(r, u, n, lane);
  }

  // ---- variants of the two-digit rows with a smaller code footprint (modexp2m.cu, TwoDigit MODE 1 / 2) -------------
  // mont_mul_x<INIT, CAPQ, 2> with the multiplier picked at run time (b = alt ? b1 : b0), so that both halves of a
  // two-digit squaring run through ONE copy of the row loop (instruction-cache footprint).
  template <int U = 1>
  static __device__ __forceinline__ uint32_t mont_mul_sel(uint32_t (&r)[L], const uint32_t (&a)[L], const uint32_t (&b0)[L],
                                                          const uint32_t (&b1)[L], bool alt, const uint32_t (&n)[L],
                                                          uint32_t n0inv, int lane, const uint32_t (&init)[L],
                                                          uint32_t init_top, uint32_t (&qcap)[L], uint32_t zr = 0u) {
    const int g = lane & (T - 1);
    uint32_t E[L + 2], O[L + 2];
#pragma unroll
    for (int j = 0; j < L; ++j) {
      E[j] = init[j];
      O[j] = 0;
    }
    E[L] = (g == T - 1) ? init_top : 0u;
    E[L + 1] = O[L] = O[L + 1] = 0;
#pragma unroll U
    for (int owner = 0; owner < T; ++owner) {
      const bool mine = g == owner;
#pragma unroll
      for (int j = 0; j < L; j += 2) {
        uint32_t b0v = __shfl_sync(ZKP_FULL, alt ? b1[j] : b0[j], owner, T);
        uint32_t b1v = __shfl_sync(ZKP_FULL, alt ? b1[j + 1] : b0[j + 1], owner, T);
        uint32_t q0, q1;
        cios_step(E, O, a, n, b0v, n0inv, g, q0, zr);
        cios_step(O, E, a, n, b1v, n0inv, g, q1, zr);
        qcap[j] = mine ? q0 : qcap[j];
        qcap[j + 1] = mine ? q1 : qcap[j + 1];
      }
    }
    uint32_t in = __shfl_down_sync(ZKP_FULL, O[0], 1, T);
    if (g == T - 1) in = 0;
    add_cc(O[L], in);
    addc(O[L + 1], 0);
    add_cc(E[0], O[1]);
#pragma unroll
    for (int j = 1; j <= L; ++j) addc_cc(E[j], O[j + 1]);
    addc(E[L + 1], 0);
    return finish_x<2>(r, E, n, lane);
  }

  // The same product with the multiplier limbs read from the group's shared-memory row sb[0..S) (LDS broadcast instead
  // of SHFL) and the quotient digits written to sq[0..S): the row loop can then be rolled at any depth (UNR pairs of
  // steps per iteration) because nothing in it indexes registers by the row number.
  template <int UNR>
  static __device__ __forceinline__ uint32_t mont_mul_s(uint32_t (&r)[L], const uint32_t (&a)[L], const uint32_t* sb, uint32_t* sq,
                                                        const uint32_t (&n)[L], uint32_t n0inv, int lane,
                                                        const uint32_t (&init)[L], uint32_t init_top, uint32_t zr = 0u) {
    // any depth: the tail runs pair by pair
    const int g = lane & (T - 1);
    uint32_t E[L + 2], O[L + 2];
#pragma unroll
    for (int j = 0; j < L; ++j) {
      E[j] = init[j];
      O[j] = 0;
    }
    E[L] = (g == T - 1) ? init_top : 0u;
    E[L + 1] = O[L] = O[L + 1] = 0;
    constexpr int kMain = S / (2 * UNR) * (2 * UNR);  // rows in whole iterations; the rest runs pair by pair
#pragma unroll 1
    for (int i = 0; i < kMain; i += 2 * UNR) {
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const uint2 bb = *reinterpret_cast<const uint2*>(sb + i + 2 * u);
        uint32_t q0, q1;
        cios_step(E, O, a, n, bb.x, n0inv, g, q0, zr);
        cios_step(O, E, a, n, bb.y, n0inv, g, q1, zr);
        if (g == 0) *reinterpret_cast<uint2*>(sq + i + 2 * u) = make_uint2(q0, q1);
      }
    }
#pragma unroll 1
    for (int i = kMain; i < S; i += 2) {
      const uint2 bb = *reinterpret_cast<const uint2*>(sb + i);
      uint32_t q0, q1;
      cios_step(E, O, a, n, bb.x, n0inv, g, q0, zr);
      cios_step(O, E, a, n, bb.y, n0inv, g, q1, zr);
      if (g == 0) *reinterpret_cast<uint2*>(sq + i) = make_uint2(q0, q1);
    }
    uint32_t in = __shfl_down_sync(ZKP_FULL, O[0], 1, T);
    if (g == T - 1) in = 0;
    add_cc(O[L], in);
    addc(O[L + 1], 0);
    add_cc(E[0], O[1]);
#pragma unroll
    for (int j = 1; j <= L; ++j) addc_cc(E[j], O[j + 1]);
    addc(E[L + 1], 0);
    return finish_x<2>(r, E, n, lane);
  }

  // mont_mul2_x with both multipliers in shared memory (sb1 pairs with a1, sb2 with a2).
  template <int UNR>
  static __device__ __forceinline__ void mont_mul2_s(uint32_t (&r)[L], const uint32_t (&a1)[L], const uint32_t* sb1,
                                                     const uint32_t (&a2)[L], const uint32_t* sb2, const uint32_t (&n)[L],
                                                     uint32_t n0inv, int lane, const uint32_t (&init)[L], uint32_t init_top,
                                                     uint32_t zr = 0u) {
    const int g = lane & (T - 1);
    uint32_t E[L + 2], O[L + 2];
#pragma unroll
    for (int j = 0; j < L; ++j) {
      E[j] = init[j];
      O[j] = 0;
    }
    E[L] = (g == T - 1) ? init_top : 0u;
    E[L + 1] = O[L] = O[L + 1] = 0;
    constexpr int kMain = S / (2 * UNR) * (2 * UNR);
#pragma unroll 1
    for (int i = 0; i < kMain; i += 2 * UNR) {
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const uint2 p = *reinterpret_cast<const uint2*>(sb1 + i + 2 * u);
        const uint2 s = *reinterpret_cast<const uint2*>(sb2 + i + 2 * u);
        cios_step2(E, O, a1, a2, n, p.x, s.x, n0inv, g, zr);
        cios_step2(O, E, a1, a2, n, p.y, s.y, n0inv, g, zr);
      }
    }
#pragma unroll 1
    for (int i = kMain; i < S; i += 2) {
      const uint2 p = *reinterpret_cast<const uint2*>(sb1 + i);
      const uint2 s = *reinterpret_cast<const uint2*>(sb2 + i);
      cios_step2(E, O, a1, a2, n, p.x, s.x, n0inv, g, zr);
      cios_step2(O, E, a1, a2, n, p.y, s.y, n0inv, g, zr);
    }
    uint32_t in = __shfl_down_sync(ZKP_FULL, O[0], 1, T);
    if (g == T - 1) in = 0;
    add_cc(O[L], in);
    addc(O[L + 1], 0);
    add_cc(E[0], O[1]);
#pragma unroll
    for (int j = 1; j <= L; ++j) addc_cc(E[j], O[j + 1]);
    addc(E[L + 1], 0);
    finish_x<3>(r, E, n, lane);
  }

  // ---- symmetric squaring (modexp2m.cu, TwoDigit MODE 3; model: tests/models/sqr_sym_model.py) -------------------------
  // x^2 = sum_g A_g^2 B^(2g) + 2 sum_{g<h} A_g A_h B^(g+h) over the lane blocks A_g (B = 2^(32 L)), each unordered pair of
  // blocks multiplied ONCE.  Round R = 0 .. T/2: lane g multiplies its block by the block of lane (g + R) mod T, an in-lane
  // L x L product (blockmul.cuh).  Without wrap-around that is the pair of difference R at block position p = 2g + R, with
  // wrap-around the pair of difference T - R at p = 2g + R - T; R = T/2 yields every pair twice (weight 1, not 2).
  // Position p belongs to lane p >> 1, slot p & 1 = R & 1 (static per round), so lane d adds what lanes d - h and
  // d - h + T/2 (h = R >> 1) made, when those are valid sources.  (T/2 + 1) L^2 limb products per lane instead of T L^2.
  // one round: acc += weight * (A_src A_(src + R)) for the valid sources of this lane (R, h = R >> 1 and the doubling
  // flag are run-time values, so that the rounds can run as a rolled loop: the code has to stay inside the instruction cache)
  static __device__ __forceinline__ void sqr_round(uint32_t (&acc)[2 * L + 1], const uint32_t (&x)[L], int g, int R, int h, uint32_t dbl) {
    uint32_t b[L], prod[2 * L], pw[2 * L + 1];
#pragma unroll
    for (int j = 0; j < L; ++j) b[j] = __shfl_sync(ZKP_FULL, x[j], (g + R) & (T - 1), T);
    block_mul<L, L, ShapeFull>(prod, x, b);
    pw[0] = prod[0] << dbl;
#pragma unroll
    for (int i = 1; i < 2 * L; ++i) pw[i] = __funnelshift_l(prod[i - 1], prod[i], dbl);
    pw[2 * L] = dbl ? prod[2 * L - 1] >> 31 : 0u;
    const int s1 = g - h, s2 = g - h + T / 2;
    const bool v1 = s1 >= 0 && s1 + R <= T - 1;  // s1's product did not wrap: it sits at 2 s1 + R, owner s1 + h = g
    const bool v2 = s2 <= T - 1 && s2 + R >= T;  // s2's product wrapped: 2 s2 + R - T, owner s2 + h - T/2 = g
    {
      uint32_t t = __shfl_sync(ZKP_FULL, pw[0], s1 & (T - 1), T);
      add_cc(acc[0], v1 ? t : 0u);
#pragma unroll
      for (int i = 1; i < 2 * L; ++i) {
        t = __shfl_sync(ZKP_FULL, pw[i], s1 & (T - 1), T);
        addc_cc(acc[i], v1 ? t : 0u);
      }
      t = __shfl_sync(ZKP_FULL, pw[2 * L], s1 & (T - 1), T);
      addc(acc[2 * L], v1 ? t : 0u);
    }
    {
      uint32_t t = __shfl_sync(ZKP_FULL, pw[0], s2 & (T - 1), T);
      add_cc(acc[0], v2 ? t : 0u);
#pragma unroll
      for (int i = 1; i < 2 * L; ++i) {
        t = __shfl_sync(ZKP_FULL, pw[i], s2 & (T - 1), T);
        addc_cc(acc[i], v2 ? t : 0u);
      }
      t = __shfl_sync(ZKP_FULL, pw[2 * L], s2 & (T - 1), T);
      addc(acc[2 * L], v2 ? t : 0u);
    }
  }

  // plo / phi = block g of the low / high half of x^2 (canonical limbs), x any S-limb value
  static __device__ __forceinline__ void sqr_product(uint32_t (&plo)[L], uint32_t (&phi)[L], const uint32_t (&x)[L], int lane) {
    static_assert(T == 4 || T == 8 || T == 16, "symmetric squaring: rounds 1 .. T/2 run in (odd, even) pairs");
    using M2 = Mp<T, 2 * L>;
    const int g = lane & (T - 1);
    uint32_t acc0[2 * L + 1], acc1[2 * L + 1];
    {
      uint32_t prod[2 * L];
      block_mul<L, L, ShapeFull>(prod, x, x);  // round 0: the diagonal block, position 2g: it stays in this lane
#pragma unroll
      for (int i = 0; i < 2 * L; ++i) {
        acc0[i] = prod[i];
        acc1[i] = 0;
      }
      acc0[2 * L] = acc1[2 * L] = 0;
    }
#pragma unroll 1
    for (int k = 0; k < T / 4; ++k) {  // rounds 2k + 1 (odd positions: slot 1) and 2k + 2 (slot 0; weight 1 when it is round T/2)
      sqr_round(acc1, x, g, 2 * k + 1, k, 1u);
      sqr_round(acc0, x, g, 2 * k + 2, k + 1, (2 * k + 2 == T / 2) ? 0u : 1u);
    }
    // this lane's two positions as one value: limbs [0, 2L) stay here, the L + 1 limbs above go one lane up
    uint32_t V[2 * L], OV[L + 1];
#pragma unroll
    for (int j = 0; j < L; ++j) V[j] = acc0[j];
    V[L] = acc0[L];
    add_cc(V[L], acc1[0]);
#pragma unroll
    for (int j = 1; j < L; ++j) {
      V[L + j] = acc0[L + j];
      addc_cc(V[L + j], acc1[j]);
    }
    OV[0] = acc0[2 * L];
    addc_cc(OV[0], acc1[L]);
#pragma unroll
    for (int j = 1; j < L; ++j) {
      OV[j] = acc1[L + j];
      addc_cc(OV[j], 0u);
    }
    OV[L] = acc1[2 * L];
    addc(OV[L], 0u);
    {
      uint32_t o = __shfl_up_sync(ZKP_FULL, OV[0], 1, T);
      add_cc(V[0], g == 0 ? 0u : o);
#pragma unroll
      for (int j = 1; j <= L; ++j) {
        o = __shfl_up_sync(ZKP_FULL, OV[j], 1, T);
        addc_cc(V[j], g == 0 ? 0u : o);
      }
#pragma unroll
      for (int j = L + 1; j < 2 * L; ++j) addc_cc(V[j], 0u);
      const uint32_t co = addc_out();
      uint32_t top;
      const uint32_t cin = M2::resolve(co != 0, M2::all_ones(V), lane, top);
      M2::add_small(V, cin);
    }
    // two blocks per lane -> block g of each half in lane g
#pragma unroll
    for (int j = 0; j < L; ++j) {
      const uint32_t a0 = __shfl_sync(ZKP_FULL, V[j], g >> 1, T), a1 = __shfl_sync(ZKP_FULL, V[L + j], g >> 1, T);
      const uint32_t c0 = __shfl_sync(ZKP_FULL, V[j], T / 2 + (g >> 1), T), c1 = __shfl_sync(ZKP_FULL, V[L + j], T / 2 + (g >> 1), T);
      plo[j] = (g & 1) ? a1 : a0;
      phi[j] = (g & 1) ? c1 : c0;
    }
  }

  // One reduction-only row  acc = (acc + q n) / 2^32 + feed 2^(32 (S - 1))  on the split accumulator (cios_step without the
  // a b products): `feed` is the limb of the high half that must sit at limb S - 1 after the previous row's shift.
  static __device__ __forceinline__ void redc_step(uint32_t (&X)[L + 2], uint32_t (&Y)[L + 2], const uint32_t (&n)[L], uint32_t feed,
                                                   uint32_t n0inv, int g, uint32_t& q, uint32_t zr) {
    uint32_t in = __shfl_down_sync(ZKP_FULL, Y[0], 1, T);
    if (g == T - 1) in = feed;
    add_cc(Y[L], in);
    addc(Y[L + 1], zr);
    q = __shfl_sync(ZKP_FULL, X[0] + Y[1], 0, T) * n0inv;
    uint32_t Z[L + 2];
    add_cc(X[0], Y[1]);  // limb 0; the carry enters the odd chain at limb 1
#pragma unroll
    for (int j = 0; j < L; j += 2) madc_wide3_cc(Z[j], Z[j + 1], n[j + 1], q, Y[j + 2], Y[j + 3]);
    Z[L] = addc_out();
    Z[L + 1] = 0;
    mad_even(X, n, q);
#pragma unroll
    for (int j = 0; j < L + 2; ++j) Y[j] = Z[j];
  }

  // r = (plo + phi W) / W mod n for a 2 S-limb value below n W given as its two halves (block g of each in lane g): S
  // reduction-only rows.  The quotient digits are kept as in mont_mul_x<.., CAPQ = true, ..>; returns 1 iff n was subtracted.
  static __device__ __forceinline__ uint32_t mont_redc_x(uint32_t (&r)[L], const uint32_t (&plo)[L], const uint32_t (&phi)[L],
                                                         const uint32_t (&n)[L], uint32_t n0inv, int lane, uint32_t (&qcap)[L],
                                                         uint32_t zr = 0u) {
    const int g = lane & (T - 1);
    uint32_t E[L + 2], O[L + 2];
#pragma unroll
    for (int j = 0; j < L; ++j) {
      E[j] = plo[j];
      O[j] = 0;
    }
    E[L] = E[L + 1] = O[L] = O[L + 1] = 0;
    uint32_t prev = 0;
#pragma unroll 1
    for (int owner = 0; owner < T; ++owner) {
      const bool mine = g == owner;
#pragma unroll
      for (int j = 0; j < L; j += 2) {
        const uint32_t f0 = __shfl_sync(ZKP_FULL, phi[j], owner, T);
        const uint32_t f1 = __shfl_sync(ZKP_FULL, phi[j + 1], owner, T);
        uint32_t q0, q1;
        redc_step(E, O, n, prev, n0inv, g, q0, zr);
        redc_step(O, E, n, f0, n0inv, g, q1, zr);
        prev = f1;
        qcap[j] = mine ? q0 : qcap[j];
        qcap[j + 1] = mine ? q1 : qcap[j + 1];
      }
    }
    uint32_t in = __shfl_down_sync(ZKP_FULL, O[0], 1, T);
    if (g == T - 1) in = prev;
    add_cc(O[L], in);
    addc(O[L + 1], 0);
    add_cc(E[0], O[1]);
#pragma unroll
    for (int j = 1; j <= L; ++j) addc_cc(E[j], O[j + 1]);
    addc(E[L + 1], 0);
    return finish_x<1>(r, E, n, lane);
  }

  // x = (x + y) mod n for x, y < n
  static __device__ __forceinline__ void add_mod(uint32_t (&x)[L], const uint32_t (&y)[L], const uint32_t (&n)[L], int lane) {
    const uint32_t ovf = add_full(x, y, lane);
    uint32_t d[L];
    const uint32_t borrow = sub_full(d, x, n, lane);
    const bool take = (ovf != 0u) || (borrow == 0u);
#pragma unroll
    for (int j = 0; j < L; ++j) x[j] = take ? d[j] : x[j];
  }

  // ---- plain (non-modular) product on the same split accumulator -----------------------------------
  // One step  acc = (acc + a*b) / 2^32 : cios_step without the q*n rows.  The limb shifted out of lane 0 is
  // the next limb of the low half of the product: it is left in lane 0's X[0] (callers capture it there).
  static __device__ __forceinline__ void mul_step(uint32_t (&X)[L + 2], uint32_t (&Y)[L + 2], const uint32_t (&a)[L], uint32_t b, int g) {
    uint32_t in = __shfl_down_sync(ZKP_FULL, Y[0], 1, T);
    if (g == T - 1) in = 0;
    add_cc(Y[L], in);
    addc(Y[L + 1], 0);
    uint32_t Z[L + 2];
    add_cc(X[0], Y[1]);
#pragma unroll
    for (int j = 0; j < L; j += 2) madc_wide3_cc(Z[j], Z[j + 1], a[j + 1], b, Y[j + 2], Y[j + 3]);
    Z[L] = addc_out();
    Z[L + 1] = 0;
    mad_even(X, a, b);
#pragma unroll
    for (int j = 0; j < L + 2; ++j) Y[j] = Z[j];
  }

  // After the last step: fold the two accumulator arrays and resolve the lane overflows into the exact
  // S-limb value (the high half of a product; nothing can be left above the top lane).
  static __device__ __forceinline__ void mul_finish(uint32_t (&r)[L], uint32_t (&E)[L + 2], uint32_t (&O)[L + 2], int lane) {
    const int g = lane & (T - 1);
    uint32_t in = __shfl_down_sync(ZKP_FULL, O[0], 1, T);
    if (g == T - 1) in = 0;
    add_cc(O[L], in);
    addc(O[L + 1], 0);
    add_cc(E[0], O[1]);
#pragma unroll
    for (int j = 1; j <= L; ++j) addc_cc(E[j], O[j + 1]);
    addc(E[L + 1], 0);
    uint32_t ov = __shfl_up_sync(ZKP_FULL, E[L], 1, T);
    if (g == 0) ov = 0;
#pragma unroll
    for (int j = 0; j < L; ++j) r[j] = E[j];
    add_cc(r[0], ov);
#pragma unroll
    for (int j = 1; j < L; ++j) addc_cc(r[j], 0);
    uint32_t co = addc_out();
    uint32_t topc;
    uint32_t cin = resolve(co != 0, all_ones(r), lane, topc);
    add_small(r, cin);
  }

  static __device__ __forceinline__ void mont_sqr(uint32_t (&r)[L], const uint32_t (&a)[L], const uint32_t (&n)[L],
                                                  uint32_t n0inv, int lane) {
    mont_mul(r, a, a, n, n0inv, lane);
  }

  // x = 2x mod n  (x < n)
  static __device__ __forceinline__ void mod_double(uint32_t (&x)[L], const uint32_t (&n)[L], int lane) {
    const int g = lane & (T - 1);
    uint32_t topbit = x[L - 1] >> 31;
    uint32_t in = __shfl_up_sync(ZKP_FULL, topbit, 1, T);
    if (g == 0) in = 0;
    uint32_t hi = __shfl_sync(ZKP_FULL, topbit, T - 1, T);
#pragma unroll
    for (int j = L - 1; j > 0; --j) x[j] = (x[j] << 1) | (x[j - 1] >> 31);
    x[0] = (x[0] << 1) | in;
    uint32_t d[L];
    uint32_t borrow = sub_full(d, x, n, lane);
    bool take = (hi != 0) || (borrow == 0);
#pragma unroll
    for (int j = 0; j < L; ++j) x[j] = take ? d[j] : x[j];
  }

  // ---- coalesced, vectorised limb moves (this lane's L limbs) -------------
  static __device__ __forceinline__ void load(uint32_t (&x)[L], const uint32_t* p) {
    if (L % 4 == 0) {
      const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
      for (int j = 0; j < L / 4; ++j) {
        uint4 v = q[j];
        x[4 * j] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w;
      }
    } else {
      const uint2* q = reinterpret_cast<const uint2*>(p);
#pragma unroll
      for (int j = 0; j < L / 2; ++j) {
        uint2 v = q[j];
        x[2 * j] = v.x; x[2 * j + 1] = v.y;
      }
    }
  }
  // the same through L2 only (ld.global.cg): for rows another SM may have rewritten since this SM last read them
  static __device__ __forceinline__ void load_cg(uint32_t (&x)[L], const uint32_t* p) {
    const uint2* q = reinterpret_cast<const uint2*>(p);
#pragma unroll
    for (int j = 0; j < L / 2; ++j) {
      uint2 v = __ldcg(q + j);
      x[2 * j] = v.x; x[2 * j + 1] = v.y;
    }
  }
  static __device__ __forceinline__ void store(uint32_t* p, const uint32_t (&x)[L]) {
    if (L % 4 == 0) {
      uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
      for (int j = 0; j < L / 4; ++j) q[j] = make_uint4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
    } else {
      uint2* q = reinterpret_cast<uint2*>(p);
#pragma unroll
      for (int j = 0; j < L / 2; ++j) q[j] = make_uint2(x[2 * j], x[2 * j + 1]);
    }
  }
  // Load a narrower integer (`limbs` <= S limbs, multiple of 2) zero-extended.
  static __device__ __forceinline__ void load_ext(uint32_t (&x)[L], const uint32_t* p, int limbs, int g) {
#pragma unroll
    for (int j = 0; j < L; j += 2) {
      int idx = g * L + j;
      uint2 v = make_uint2(0u, 0u);
      if (idx < limbs) v = *reinterpret_cast<const uint2*>(p + idx);
      x[j] = v.x; x[j + 1] = v.y;
    }
  }
  // Store only the low `limbs` limbs (multiple of 2) of the integer.
  static __device__ __forceinline__ void store_ext(uint32_t* p, const uint32_t (&x)[L], int limbs, int g) {
#pragma unroll
    for (int j = 0; j < L; j += 2) {
      int idx = g * L + j;
      if (idx < limbs) *reinterpret_cast<uint2*>(p + idx) = make_uint2(x[j], x[j + 1]);
    }
  }
  static __device__ __forceinline__ void set_small(uint32_t (&x)[L], uint32_t v, int g) {
#pragma unroll
    for (int j = 0; j < L; ++j) x[j] = 0;
    if (g == 0) x[0] = v;
  }
};

}  // namespace zkp
