// NiCorrectKeyProof::verify glue (reference src/zkproofs/correct_key_ni.rs:73-117): everything but the
// eleven sigma_i^N mod N (K2, modexp.cu):
//   ck_rho_kernel     salt_bn, seed_i = H(N || salt_bn || i), mask_generation (MGF1-like, :105-117)
//   ck_reduce_kernel  rho_i = mask_i % N on the cooperative Montgomery arithmetic
//   ck_check_kernel   rho == derived  and  gcd(P, N) == 1 with P the primorial of 6370 (:87-88, 95)
#include "kernels.h"
#include "mp_coop.cuh"
#include "sha256.cuh"

namespace zkp {

constexpr int kCkThreads = 64;

// One thread per (proof b, index i).  mask row: ml = nl + 8 limbs, little endian, zero padded.
__global__ void __launch_bounds__(kCkThreads) ck_rho_kernel(const uint32_t* n, int nl, const uint8_t* salt, int salt_len,
                                                            int batch, uint32_t* mask, int ml) {
  __shared__ uint32_t wbuf[16 * kCkThreads];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= batch * kCkM2) return;
  const int b = t / kCkM2, i = t % kCkM2;
  const uint32_t* N = n + (size_t)b * nl;
  Sha256 s;
  uint32_t salt_bn[8], seed[8], d[8];
  // salt_bn = compute_digest(once(BigInt::from_bytes(salt)))  (:75): to_bytes drops the salt's leading zero bytes
  s.init(wbuf + threadIdx.x, kCkThreads);
  {
    int k = 0;
    while (k < salt_len && salt[k] == 0) ++k;
    if (k == salt_len) s.push(0u, 1);
    for (; k < salt_len; ++k) s.push(salt[k], 1);
  }
  s.finish(salt_bn);
  // seed_bn = compute_digest(n, salt_bn, BigInt::from(i))  (:79-83)
  s.init(wbuf + threadIdx.x, kCkThreads);
  s.push_bigint(nl, [&](int k) { return __ldg(N + k); });
  s.push_bigint(8, [&](int k) { return salt_bn[7 - k]; });
  s.push((uint32_t)i, 1);
  s.finish(seed);
  // key_length = n.bit_length(); msklen = key_length / 256 + 1  (:74, :106)
  int top = nl - 1;
  while (top >= 0 && N[top] == 0u) --top;
  const int key_length = top < 0 ? 0 : 32 * top + (32 - __clz(N[top]));
  const int msklen = key_length / 256 + 1;
  uint32_t* row = mask + (size_t)t * ml;
  for (int j = 0; j < msklen; ++j) {  // H(seed || j) << (256 j)  (:107-116)
    s.init(wbuf + threadIdx.x, kCkThreads);
    s.push_bigint(8, [&](int k) { return seed[7 - k]; });
    s.push((uint32_t)j, 1);
    s.finish(d);
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (8 * j + k < ml) row[8 * j + k] = d[7 - k];
  }
  for (int k = 8 * msklen; k < ml; ++k) row[k] = 0u;
}

// rho = mask % N.  mask = lo + hi * 2^(32 S) with lo < R = 2^(32 S), hi < 2^256 <= N:
//   lo mod N = montmul(montmul(lo, R^2), 1),  hi * R mod N = montmul(hi, R^2).
template <int T, int L>
__global__ void __launch_bounds__(kCtaThreads) ck_reduce_kernel(const uint32_t* mask, int ml, const uint32_t* mods, int nl,
                                                              const uint32_t* r2, const uint32_t* n0inv, int jobs,
                                                              uint32_t* rho) {
  using M = Mp<T, L>;
  constexpr int S = T * L;
  constexpr int G = kCtaThreads / T;
  const int lane = threadIdx.x & 31;
  const int g = lane & (T - 1);
  const int job = blockIdx.x * G + threadIdx.x / T;
  const bool valid = job < jobs;
  const int src = valid ? job : 0;
  const int mi = src / kCkM2;
  const uint32_t ni = n0inv[mi];
  uint32_t n[L], rr[L], lo[L], hi[L], one[L];
  M::load_ext(n, mods + (size_t)mi * nl, nl, g);
  M::load(rr, r2 + (size_t)mi * S + g * L);
  const uint32_t* row = mask + (size_t)src * ml;
  M::load_ext(lo, row, ml < S ? ml : S, g);
  M::load_ext(hi, row + S, ml > S ? ml - S : 0, g);
  M::set_small(one, 1u, g);
  M::mont_mul(lo, lo, rr, n, ni, lane);
  M::mont_mul(lo, lo, one, n, ni, lane);
  M::mont_mul(hi, hi, rr, n, ni, lane);
  uint32_t carry = M::add_full(lo, hi, lane);
  uint32_t d[L];
  uint32_t borrow = M::sub_full(d, lo, n, lane);
  const bool take = carry != 0 || borrow == 0;
#pragma unroll
  for (int j = 0; j < L; ++j) lo[j] = take ? d[j] : lo[j];
  if (valid) M::store_ext(rho + (size_t)job * nl, lo, nl, g);
}

// One warp per proof: all eleven rho_i == sigma_i^N mod N, and no prime < 6370 divides N
// (P is the square-free product of exactly those primes, so gcd(P, N) == 1 iff none divides N).
__global__ void ck_check_kernel(const uint32_t* n, int nl, const uint32_t* rho, const uint32_t* derived,
                                const uint16_t* primes, int nprimes, int batch, uint8_t* accept) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= batch) return;
  const size_t row = (size_t)warp * kCkM2 * nl;
  uint32_t bad = 0;
  for (int k = lane; k < kCkM2 * nl; k += 32) bad |= rho[row + k] ^ derived[row + k];
  const uint32_t* N = n + (size_t)warp * nl;
  for (int pi = lane; pi < nprimes; pi += 32) {
    const uint32_t p = primes[pi];
    uint32_t rem = 0;
    for (int k = nl - 1; k >= 0; --k) {
      const uint32_t v = __ldg(N + k);
      rem = ((rem << 16) | (v >> 16)) % p;
      rem = ((rem << 16) | (v & 0xffffu)) % p;
    }
    if (rem == 0) bad |= 1u;
  }
  bad = __reduce_or_sync(0xffffffffu, bad);
  if (lane == 0) accept[warp] = bad ? 0 : 1;
}

cudaError_t launch_ck_rho(const uint32_t* n, int nl, const uint8_t* salt, int salt_len, int batch, uint32_t* mask, int ml,
                          cudaStream_t st) {
  const int total = batch * kCkM2;
  if (total <= 0) return cudaSuccess;
  ck_rho_kernel<<<(total + kCkThreads - 1) / kCkThreads, kCkThreads, 0, st>>>(n, nl, salt, salt_len, batch, mask, ml);
  return cudaGetLastError();
}

cudaError_t launch_ck_reduce(const uint32_t* mask, int ml, const uint32_t* mods, int nl, const uint32_t* r2,
                             const uint32_t* n0inv, int S, int batch, uint32_t* rho, cudaStream_t st) {
  const int jobs = batch * kCkM2;
  if (jobs <= 0) return cudaSuccess;
  if (ml % 2 || nl % 2 || nl > S || ml > 2 * S) return cudaErrorInvalidValue;
#define CALL(T_, L_)                                                                                              \
  {                                                                                                               \
    constexpr int G = kCtaThreads / T_;                                                                           \
    ck_reduce_kernel<T_, L_><<<(jobs + G - 1) / G, kCtaThreads, 0, st>>>(mask, ml, mods, nl, r2, n0inv, jobs, rho); \
  }
  switch (S) {
    case 32:  { CALL(4, 8);  } break;
    case 64:  { CALL(8, 8);  } break;
    case 96:  { CALL(8, 12); } break;
    case 128: { CALL(8, 16); } break;
    case 192: { CALL(16, 12); } break;
    case 256: { CALL(16, 16); } break;
    default: return cudaErrorInvalidValue;
  }
#undef CALL
  return cudaGetLastError();
}

cudaError_t launch_ck_check(const uint32_t* n, int nl, const uint32_t* rho, const uint32_t* derived,
                            const uint16_t* primes, int nprimes, int batch, uint8_t* accept, cudaStream_t st) {
  if (batch <= 0) return cudaSuccess;
  ck_check_kernel<<<(batch * 32 + 255) / 256, 256, 0, st>>>(n, nl, rho, derived, primes, nprimes, batch, accept);
  return cudaGetLastError();
}

}  // namespace zkp
