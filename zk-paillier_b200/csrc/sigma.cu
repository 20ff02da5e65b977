// Small per-thread helpers for the sigma-protocol proofs (ZeroProof, CiphertextProof, MulProof,
// VerlinProof; reference src/zkproofs/{zero_enc_proof,correct_ciphertext,multiplication_proof,
// verlin_proof}.rs).  The modexps of those proofs run in K1/K2 and the mulmods in K3; what is left is
// one-per-proof bookkeeping on full-width integers, done here one thread per proof:
//   digest -> exponent limbs, z = a + x*e (unreduced), (a + b) mod n, row equality, and
//   BigInt::mod_inv (multiplication_proof.rs:96,137) by the binary extended Euclid.
#include "kernels.h"
#include "mp_coop.cuh"

namespace zkp {

constexpr int kMaxLimbs = 256;  // 8192-bit modulus

// e = compute_digest(...) as a BigInt (utils.rs:21): 32 big-endian bytes -> 8 little-endian limbs
__global__ void digest_to_limbs_kernel(const uint8_t* digest, int batch, uint32_t* out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= batch * 8) return;
  const int b = t >> 3, k = t & 7;  // limb k = bytes [28-4k, 32-4k)
  const uint8_t* d = digest + (size_t)b * 32 + 28 - 4 * k;
  out[t] = ((uint32_t)d[0] << 24) | ((uint32_t)d[1] << 16) | ((uint32_t)d[2] << 8) | d[3];
}

// out[b] = a[b] + x[b] * e[b]   (plain integers; a may be null).  out_limbs >= x_limbs + e_limbs.
__global__ void muladd_kernel(const uint32_t* a, int a_limbs, const uint32_t* x, int x_limbs, const uint32_t* e,
                              int e_limbs, int batch, uint32_t* out, int out_limbs, uint8_t* fault) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  uint32_t* o = out + (size_t)b * out_limbs;
  const uint32_t* xx = x + (size_t)b * x_limbs;
  const uint32_t* ee = e + (size_t)b * e_limbs;
  for (int i = 0; i < out_limbs; ++i) o[i] = (a && i < a_limbs) ? a[(size_t)b * a_limbs + i] : 0u;
  uint32_t over = 0;
  for (int j = 0; j < e_limbs; ++j) {
    const uint32_t ej = ee[j];
    unsigned long long carry = 0;
    for (int i = 0; i < x_limbs; ++i) {
      if (i + j >= out_limbs) { over |= (ej && xx[i]) ? 1u : 0u; continue; }
      unsigned long long t = (unsigned long long)xx[i] * ej + o[i + j] + carry;
      o[i + j] = (uint32_t)t;
      carry = t >> 32;
    }
    for (int k = x_limbs + j; carry; ++k) {
      if (k >= out_limbs) { over = 1; break; }
      unsigned long long t = (unsigned long long)o[k] + carry;
      o[k] = (uint32_t)t;
      carry = t >> 32;
    }
  }
  if (over && fault) fault[b] = 1;
}

// out[b] = (a[b] + c[b]) mod m   with a, c < m  (BigInt::mod_add, multiplication_proof.rs:90)
__global__ void modadd_kernel(const uint32_t* a, const uint32_t* c, const uint32_t* m, int limbs, int batch,
                              uint32_t* out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const uint32_t* aa = a + (size_t)b * limbs;
  const uint32_t* cc = c + (size_t)b * limbs;
  uint32_t* o = out + (size_t)b * limbs;
  uint32_t carry = 0;
  for (int i = 0; i < limbs; ++i) {
    unsigned long long t = (unsigned long long)aa[i] + cc[i] + carry;
    o[i] = (uint32_t)t;
    carry = (uint32_t)(t >> 32);
  }
  bool ge = carry != 0;
  if (!ge) {
    ge = true;
    for (int i = limbs - 1; i >= 0; --i)
      if (o[i] != m[i]) { ge = o[i] > m[i]; break; }
  }
  if (ge) {
    uint32_t br = 0;
    for (int i = 0; i < limbs; ++i) {
      unsigned long long t = (unsigned long long)o[i] - m[i] - br;
      o[i] = (uint32_t)t;
      br = (uint32_t)(t >> 63);
    }
  }
}

// accept[b] = (and_in ? accept[b] : 1) && x[b] == y[b]      one warp per row
__global__ void rows_equal_kernel(const uint32_t* x, const uint32_t* y, int limbs, int batch, int and_in,
                                  uint8_t* accept) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= batch) return;
  uint32_t diff = 0;
  for (int i = lane; i < limbs; i += 32) diff |= x[(size_t)warp * limbs + i] ^ y[(size_t)warp * limbs + i];
  diff = __reduce_or_sync(0xffffffffu, diff);
  if (lane == 0) accept[warp] = (uint8_t)((and_in ? accept[warp] : 1) && diff == 0);
}

// ---- BigInt::mod_inv(v, m) for odd m: binary extended Euclid, ONE WARP per instance.
// The four working integers (u, w and the cofactors x1, x2) are spread over the 32 lanes, L limbs per lane in
// registers (Mp<32, L>): shifts borrow one bit from the neighbouring lane, add / sub / compare resolve
// their carries with one ballot.  Control flow depends only on the instance, so it is uniform per warp.
// Invariants: u = x1 * v (mod m), w = x2 * v (mod m).  Ends with u == 1 (inverse x1), or a common factor
// (u == 0 or u == w): not invertible -> fault, where the reference's unwrap() panics.
template <int L>
struct Inv {
  using M = Mp<32, L>;
  static __device__ __forceinline__ void shr1(uint32_t (&x)[L], uint32_t top, int lane) {
    uint32_t nb = __shfl_down_sync(ZKP_FULL, x[0], 1);
    if (lane == 31) nb = top;
#pragma unroll
    for (int j = 0; j < L - 1; ++j) x[j] = (x[j] >> 1) | (x[j + 1] << 31);
    x[L - 1] = (x[L - 1] >> 1) | (nb << 31);
  }
  static __device__ __forceinline__ bool is_odd(const uint32_t (&x)[L]) { return (__shfl_sync(ZKP_FULL, x[0], 0) & 1u) != 0; }
  static __device__ __forceinline__ bool is_small(const uint32_t (&x)[L], uint32_t v, int lane) {  // x == v (v < 2^32)
    uint32_t m = (lane == 0) ? (x[0] ^ v) : x[0];
#pragma unroll
    for (int j = 1; j < L; ++j) m |= x[j];
    return __ballot_sync(ZKP_FULL, m != 0) == 0u;
  }
  // x = (x + m) >> 1 if x is odd else x >> 1   (halving modulo the odd m)
  static __device__ __forceinline__ void half_mod(uint32_t (&x)[L], const uint32_t (&m)[L], int lane) {
    uint32_t top = 0;
    if (is_odd(x)) top = M::add_full(x, m, lane);
    shr1(x, top, lane);
  }
  // x = (x - y) mod m
  static __device__ __forceinline__ void sub_mod(uint32_t (&x)[L], const uint32_t (&y)[L], const uint32_t (&m)[L], int lane) {
    uint32_t d[L];
    uint32_t borrow = M::sub_full(d, x, y, lane);
#pragma unroll
    for (int j = 0; j < L; ++j) x[j] = d[j];
    if (borrow) M::add_full(x, m, lane);
  }
};

template <int L>
__global__ void __launch_bounds__(128) modinv_kernel(const uint32_t* v, const uint32_t* m, long long m_stride, int limbs, int batch,
                                                     uint32_t* out, uint8_t* fault) {
  using M = Mp<32, L>;
  using I = Inv<L>;
  const int lane = threadIdx.x & 31;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= batch) return;
  uint32_t u[L], w[L], x1[L], x2[L], mm[L], d[L];
  M::load_ext(u, v + (size_t)b * limbs, limbs, lane);
  M::load_ext(mm, m + (size_t)b * m_stride, limbs, lane);  // m_stride = 0: one shared modulus
#pragma unroll
  for (int j = 0; j < L; ++j) {
    w[j] = mm[j];
    x1[j] = 0;
    x2[j] = 0;
  }
  if (lane == 0) x1[0] = 1;
  // v < 2^(32 limbs) may exceed m by a small factor when m's top limb is set: reduce by subtraction
  bool ok = true;
  for (int k = 0; k < 64; ++k) {
    if (M::sub_full(d, u, mm, lane)) break;  // u < m
#pragma unroll
    for (int j = 0; j < L; ++j) u[j] = d[j];
    if (k == 63) ok = false;
  }
  if (I::is_small(u, 0u, lane)) ok = false;
  while (ok && !I::is_small(u, 1u, lane)) {
    while (!I::is_odd(u)) {
      I::shr1(u, 0, lane);
      I::half_mod(x1, mm, lane);
    }
    while (!I::is_odd(w)) {
      I::shr1(w, 0, lane);
      I::half_mod(x2, mm, lane);
    }
    if (I::is_small(u, 1u, lane)) break;
    const uint32_t borrow = M::sub_full(d, u, w, lane);  // d = u - w
    if (!borrow) {
      if (I::is_small(d, 0u, lane)) { ok = false; break; }  // u == w != 1: common factor
#pragma unroll
      for (int j = 0; j < L; ++j) u[j] = d[j];
      I::sub_mod(x1, x2, mm, lane);
    } else {
      M::sub_full(d, w, u, lane);
#pragma unroll
      for (int j = 0; j < L; ++j) w[j] = d[j];
      I::sub_mod(x2, x1, mm, lane);
    }
  }
  if (!ok) {
#pragma unroll
    for (int j = 0; j < L; ++j) x1[j] = 0;
    if (lane == 0) fault[b] = 1;
  }
  M::store_ext(out + (size_t)b * limbs, x1, limbs, lane);
}

static inline unsigned blocks_for(long long n, int t) { return (unsigned)((n + t - 1) / t); }

cudaError_t launch_digest_to_limbs(const uint8_t* digest, int batch, uint32_t* out, cudaStream_t st) {
  if (batch <= 0) return cudaSuccess;
  digest_to_limbs_kernel<<<blocks_for(batch * 8ll, 128), 128, 0, st>>>(digest, batch, out);
  return cudaGetLastError();
}
cudaError_t launch_muladd(const uint32_t* a, int a_limbs, const uint32_t* x, int x_limbs, const uint32_t* e, int e_limbs,
                          int batch, uint32_t* out, int out_limbs, uint8_t* fault, cudaStream_t st) {
  if (batch <= 0) return cudaSuccess;
  muladd_kernel<<<blocks_for(batch, 64), 64, 0, st>>>(a, a_limbs, x, x_limbs, e, e_limbs, batch, out, out_limbs, fault);
  return cudaGetLastError();
}
cudaError_t launch_modadd(const uint32_t* a, const uint32_t* c, const uint32_t* m, int limbs, int batch, uint32_t* out,
                          cudaStream_t st) {
  if (batch <= 0) return cudaSuccess;
  modadd_kernel<<<blocks_for(batch, 64), 64, 0, st>>>(a, c, m, limbs, batch, out);
  return cudaGetLastError();
}
cudaError_t launch_rows_equal(const uint32_t* x, const uint32_t* y, int limbs, int batch, int and_in, uint8_t* accept,
                              cudaStream_t st) {
  if (batch <= 0) return cudaSuccess;
  rows_equal_kernel<<<blocks_for(batch * 32ll, 256), 256, 0, st>>>(x, y, limbs, batch, and_in, accept);
  return cudaGetLastError();
}
cudaError_t launch_modinv(const uint32_t* v, const uint32_t* m, int limbs, int batch, uint32_t* scratch, uint32_t* out,
                          uint8_t* fault, cudaStream_t st, long long m_stride) {
  (void)scratch;
  if (batch <= 0) return cudaSuccess;
  if (limbs > kMaxLimbs || limbs % 2) return cudaErrorInvalidValue;
  const unsigned grid = blocks_for(batch * 32ll, 128);
  const int per_lane = (limbs + 31) / 32;
  if (per_lane <= 2) modinv_kernel<2><<<grid, 128, 0, st>>>(v, m, m_stride, limbs, batch, out, fault);
  else if (per_lane <= 4) modinv_kernel<4><<<grid, 128, 0, st>>>(v, m, m_stride, limbs, batch, out, fault);
  else if (per_lane <= 6) modinv_kernel<6><<<grid, 128, 0, st>>>(v, m, m_stride, limbs, batch, out, fault);
  else modinv_kernel<8><<<grid, 128, 0, st>>>(v, m, m_stride, limbs, batch, out, fault);
  return cudaGetLastError();
}

}  // namespace zkp
