// Small per-thread helpers for the sigma-protocol proofs (ZeroProof, CiphertextProof, MulProof,
// VerlinProof; reference src/zkproofs/{zero_enc_proof,correct_ciphertext,multiplication_proof,
// verlin_proof}.rs).  The modexps of those proofs run in K1/K2 and the mulmods in K3; what is left is
// one-per-proof bookkeeping on full-width integers, done here one thread per proof:
//   digest -> exponent limbs, z = a + x*e (unreduced), (a + b) mod n, row equality, and
//   BigInt::mod_inv (multiplication_proof.rs:96,137) by the binary extended Euclid.
#include "kernels.h"

namespace zkp {

constexpr int kMaxLimbs = 272;  // 8192-bit modulus + slack

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

// ---- BigInt::mod_inv(v, m) for odd m: binary extended Euclid, one thread per instance.
// Invariants: u = x1 * v (mod m), w = x2 * v (mod m).  Ends with u == 1 (inverse x1), or u == 0
// (gcd(v, m) = w != 1: not invertible -> fault, where the reference's unwrap() panics).
struct Big {
  uint32_t* d;
  int n;
  __device__ bool is_zero() const { for (int i = 0; i < n; ++i) if (d[i]) return false; return true; }
  __device__ bool is_one() const { if (d[0] != 1u) return false; for (int i = 1; i < n; ++i) if (d[i]) return false; return true; }
  __device__ bool even() const { return !(d[0] & 1u); }
  __device__ void shr1(uint32_t top) { for (int i = 0; i < n - 1; ++i) d[i] = (d[i] >> 1) | (d[i + 1] << 31); d[n - 1] = (d[n - 1] >> 1) | (top << 31); }
  __device__ uint32_t add(const uint32_t* o) { uint32_t c = 0; for (int i = 0; i < n; ++i) { unsigned long long t = (unsigned long long)d[i] + o[i] + c; d[i] = (uint32_t)t; c = (uint32_t)(t >> 32); } return c; }
  __device__ uint32_t sub(const uint32_t* o) { uint32_t b = 0; for (int i = 0; i < n; ++i) { unsigned long long t = (unsigned long long)d[i] - o[i] - b; d[i] = (uint32_t)t; b = (uint32_t)(t >> 63); } return b; }
  __device__ int cmp(const uint32_t* o) const { for (int i = n - 1; i >= 0; --i) if (d[i] != o[i]) return d[i] < o[i] ? -1 : 1; return 0; }
};

__global__ void modinv_kernel(const uint32_t* v, const uint32_t* m, int limbs, int batch, uint32_t* scratch,
                              uint32_t* out, uint8_t* fault) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  uint32_t* base = scratch + (size_t)b * 4 * limbs;
  Big u{base, limbs}, w{base + limbs, limbs}, x1{base + 2 * limbs, limbs}, x2{base + 3 * limbs, limbs};
  for (int i = 0; i < limbs; ++i) {
    u.d[i] = v[(size_t)b * limbs + i];
    w.d[i] = m[i];
    x1.d[i] = i == 0 ? 1u : 0u;
    x2.d[i] = 0u;
  }
  // reduce v below m first (v < 2^(32 limbs); m has its top limb set in practice, a few subtractions at most)
  while (u.cmp(m) >= 0) u.sub(m);
  bool ok = !u.is_zero();
  while (ok && !u.is_one()) {
    while (u.even()) {
      u.shr1(0);
      uint32_t c = x1.even() ? 0u : x1.add(m);
      x1.shr1(c);
    }
    while (w.even()) {
      w.shr1(0);
      uint32_t c = x2.even() ? 0u : x2.add(m);
      x2.shr1(c);
    }
    if (u.is_one()) break;
    int c = u.cmp(w.d);
    if (c == 0) { ok = false; break; }  // u == w != 1: common factor
    if (c > 0) {
      u.sub(w.d);
      if (x1.sub(x2.d)) x1.add(m);
    } else {
      w.sub(u.d);
      if (x2.sub(x1.d)) x2.add(m);
    }
  }
  for (int i = 0; i < limbs; ++i) out[(size_t)b * limbs + i] = ok ? x1.d[i] : 0u;
  if (!ok) fault[b] = 1;
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
                          uint8_t* fault, cudaStream_t st) {
  if (batch <= 0) return cudaSuccess;
  if (limbs > kMaxLimbs) return cudaErrorInvalidValue;
  modinv_kernel<<<blocks_for(batch, 32), 32, 0, st>>>(v, m, limbs, batch, scratch, out, fault);
  return cudaGetLastError();
}

}  // namespace zkp
