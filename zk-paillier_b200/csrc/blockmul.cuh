// In-lane block products by PRODUCT SCANNING (column sums in a 3-word accumulator), fully unrolled.
// Every limb product is one IMAD.WIDE.U32 (+ a carry add into the third accumulator word on the otherwise
// idle ALU pipe; ptxas pairs two of them per IADD3.X); both operands are register arrays (or constant-bank
// values) with static indices, so there are no shuffles inside a block product and any shape -- full,
// triangular, a column range -- costs exactly its number of limb products.  Two columns are kept in flight
// in independent accumulators, so one warp already has two dependency chains to overlap.
// Measured on B200 (profiles/r01_integer_pipe_probes.json): 8.0 T multiply-adds/s = the IMAD.WIDE pipe limit,
// at 16 and at 8 warps per SM.
#pragma once
#include <stdint.h>

namespace zkp {

// (c2:c1:c0) += a * b
__device__ __forceinline__ void mac3(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t a, uint32_t b) {
  asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
               : "+r"(c0), "+r"(c1), "+r"(c2)
               : "r"(a), "r"(b));
}
// (d2:d1:d0) += (c2:c1): carry the upper two words of a finished column into the next one
__device__ __forceinline__ void carry3(uint32_t& d0, uint32_t& d1, uint32_t& d2, uint32_t c1, uint32_t c2) {
  asm volatile("add.cc.u32 %0, %0, %3;\n\taddc.cc.u32 %1, %1, %4;\n\taddc.u32 %2, %2, 0;"
               : "+r"(d0), "+r"(d1), "+r"(d2)
               : "r"(c1), "r"(c2));
}

// Which (i, j) cells of the LA x LB grid are multiplied.
struct ShapeFull { static __device__ __forceinline__ constexpr bool on(int, int) { return true; } };
struct ShapeLowTri { static __device__ __forceinline__ constexpr bool on(int i, int j) { return i >= j; } };
struct ShapeStrictLow { static __device__ __forceinline__ constexpr bool on(int i, int j) { return i > j; } };

// Columns [CLO, CHI) of sum_{enabled (i,j)} a[i] * b(j) * 2^(32 (i + j)), the accumulator starting at zero at
// column CLO (lower columns are dropped: callers use that only where a truncated product is wanted).
// emit(k, limb) receives limbs k = CLO .. CHI + 1 (the last two are the accumulator's remaining words;
// with CHI = LA + LB - 1, limb CHI is the top limb of the product and limb CHI + 1 is zero).
// `b` is any callable j -> uint32_t with a compile-time-constant argument after unrolling.
template <int LA, int LB, class Shape, int CLO, int CHI, class BFn, class Emit>
__device__ __forceinline__ void block_mul_cols(const uint32_t (&a)[LA], BFn b, Emit emit) {
  static_assert(CLO >= 0 && CHI <= LA + LB - 1 && CLO < CHI, "column range");
  uint32_t c0 = 0, c1 = 0, c2 = 0, d0 = 0, d1 = 0, d2 = 0;
#pragma unroll
  for (int k = CLO; k < CHI; k += 2) {
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      const int j = k - i;
      if (j >= 0 && j < LB && Shape::on(i, j)) mac3(c0, c1, c2, a[i], b(j));
      const int j1 = k + 1 - i;
      if (k + 1 < CHI && j1 >= 0 && j1 < LB && Shape::on(i, j1)) mac3(d0, d1, d2, a[i], b(j1));
    }
    emit(k, c0);
    carry3(d0, d1, d2, c1, c2);
    emit(k + 1, d0);
    c0 = d1;
    c1 = d2;
    c2 = 0;
    d0 = d1 = d2 = 0;
  }
  if (((CHI - CLO) & 1) == 0) {  // even column count: two words are still pending
    emit(CHI, c0);
    emit(CHI + 1, c1);
  } else {                       // odd: limb CHI went out as the last `d0`; one word pending
    emit(CHI + 1, c0);
  }
}

// out[0 .. LA+LB) = full (or shaped) product of two register blocks
template <int LA, int LB, class Shape>
__device__ __forceinline__ void block_mul(uint32_t (&out)[LA + LB], const uint32_t (&a)[LA], const uint32_t (&b)[LB]) {
  block_mul_cols<LA, LB, Shape, 0, LA + LB - 1>(
      a, [&](int j) { return b[j]; }, [&](int k, uint32_t v) { if (k < LA + LB) out[k] = v; });
}

}  // namespace zkp
