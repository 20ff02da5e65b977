// Helpers for the remaining public proofs of the reference:
//   CompositeDLogProof   reference src/zkproofs/wi_dlog_proof.rs:46-91   (per-statement modulus N)
//   CorrectMessageProof  reference src/zkproofs/correct_message.rs:35-162 (ring proof over a small message space)
// Their modexps run in K1m / K2m / K2 and the mulmods in K3; this file holds what is left: the product modulo a
// per-statement modulus, the inverse of g^m = 1 + m n, the 256-bit challenge bookkeeping and the ring layout.
#include "kernels.h"
#include "mp_coop.cuh"

namespace zkp {

// ---- out[j] = a[j] * b[j] mod mods[j / mod_per]   (BigInt::mod_mul with a per-statement modulus, wi_dlog_proof.rs:80)
// Two Montgomery multiplications: (a b / R) (R^2) / R.  r2 / n0inv from launch_mont_setup.
template <int T, int L>
__global__ void __launch_bounds__(kCtaThreads) modmul_var_kernel(const uint32_t* a, const uint32_t* b, int limbs, const uint32_t* mods,
                                                               const uint32_t* r2, const uint32_t* n0inv, int mod_per, int jobs,
                                                               uint32_t* out) {
  using M = Mp<T, L>;
  constexpr int S = T * L;
  constexpr int G = kCtaThreads / T;
  const int lane = threadIdx.x & 31;
  const int g = lane & (T - 1);
  const int job = blockIdx.x * G + threadIdx.x / T;
  const bool valid = job < jobs;
  const int src = valid ? job : 0;
  const int mi = src / mod_per;
  const uint32_t ni = n0inv[mi];
  uint32_t n[L], rr[L], x[L], y[L];
  M::load_ext(n, mods + (size_t)mi * limbs, limbs, g);
  M::load(rr, r2 + (size_t)mi * S + g * L);
  M::load_ext(x, a + (size_t)src * limbs, limbs, g);
  M::load_ext(y, b + (size_t)src * limbs, limbs, g);
  // operands may exceed the modulus (they are only below 2^(32 limbs)): bring x below n first, x R / R = x mod n
  M::mont_mul(x, x, rr, n, ni, lane);
  uint32_t one[L];
  M::set_small(one, 1u, g);
  M::mont_mul(x, x, one, n, ni, lane);
  M::mont_mul(x, x, y, n, ni, lane);   // x y / R  (x < n, y < R)
  M::mont_mul(x, x, rr, n, ni, lane);  // x y mod n
  if (valid) M::store_ext(out + (size_t)job * limbs, x, limbs, g);
}

cudaError_t launch_modmul_var(const uint32_t* a, const uint32_t* b, int limbs, const uint32_t* mods, const uint32_t* r2,
                              const uint32_t* n0inv, int mod_per, int S, int jobs, uint32_t* out, cudaStream_t st) {
  if (jobs <= 0) return cudaSuccess;
  if (limbs % 2 || limbs > S || mod_per <= 0) return cudaErrorInvalidValue;
#define CALL(T_, L_)                                                                                                   \
  {                                                                                                                    \
    constexpr int G = kCtaThreads / T_;                                                                                \
    modmul_var_kernel<T_, L_><<<(jobs + G - 1) / G, kCtaThreads, 0, st>>>(a, b, limbs, mods, r2, n0inv, mod_per, jobs, out); \
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

// ---- fault[b] |= !(N_b > 2^bits)      (assert!(statement.N > 2^K), wi_dlog_proof.rs:68)
__global__ void gt_pow2_kernel(const uint32_t* n, int limbs, int bits, int batch, uint8_t* fault) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const uint32_t* N = n + (size_t)b * limbs;
  // N > 2^bits  <=>  some bit above `bits` is set, or bit `bits` is set and some lower bit is set
  const int wl = bits >> 5, wb = bits & 31;
  bool above = false, at = false, below = false;
  for (int i = 0; i < limbs; ++i) {
    const uint32_t v = N[i];
    if (i > wl) above |= v != 0u;
    else if (i == wl) {
      above |= wb < 31 && (v >> (wb + 1)) != 0u;
      at = (v >> wb) & 1u;
      below |= (v & ((1u << wb) - 1u)) != 0u;
    } else below |= v != 0u;
  }
  if (!(above || (at && below))) fault[b] = 1;
}
cudaError_t launch_gt_pow2(const uint32_t* n, int limbs, int bits, int batch, uint8_t* fault, cudaStream_t st) {
  if (batch <= 0) return cudaSuccess;
  gt_pow2_kernel<<<(batch + 63) / 64, 64, 0, st>>>(n, limbs, bits, batch, fault);
  return cudaGetLastError();
}

// ---- out[t] = (1 + m_t n)^(-1) mod n^2 = 1 + ((n - m_t) mod n) n     (gm_inv, correct_message.rs:52-53,136-139)
// m_t < n (already reduced), nl limbs; out rows are 2 nl limbs.  One thread per row.
__global__ void gm_inv_kernel(const uint32_t* m, const uint32_t* n, int nl, int rows, uint32_t* out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows) return;
  const uint32_t* mm = m + (size_t)t * nl;
  uint32_t* o = out + (size_t)t * 2 * nl;
  bool zero = true;
  for (int i = 0; i < nl; ++i) zero &= mm[i] == 0u;
  for (int i = 0; i < 2 * nl; ++i) o[i] = 0u;
  o[0] = 1u;
  if (zero) return;
  // neg = n - m, accumulated limb by limb into the product neg * n (schoolbook, row i of neg)
  uint32_t br = 0;
  for (int i = 0; i < nl; ++i) {
    const unsigned long long d = (unsigned long long)n[i] - mm[i] - br;
    const uint32_t ni = (uint32_t)d;
    br = (uint32_t)(d >> 63);
    unsigned long long carry = 0;
    for (int j = 0; j < nl; ++j) {
      const unsigned long long v = (unsigned long long)ni * n[j] + o[i + j] + carry;
      o[i + j] = (uint32_t)v;
      carry = v >> 32;
    }
    for (int k = i + nl; carry && k < 2 * nl; ++k) {
      const unsigned long long v = (unsigned long long)o[k] + carry;
      o[k] = (uint32_t)v;
      carry = v >> 32;
    }
  }
}
cudaError_t launch_gm_inv(const uint32_t* m, const uint32_t* n, int nl, int rows, uint32_t* out, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  gm_inv_kernel<<<(rows + 63) / 64, 64, 0, st>>>(m, n, nl, rows, out);
  return cudaGetLastError();
}

// ---- CorrectMessageProof::prove ring layout (correct_message.rs:62-83,98-121): one thread per proof.
// match[b][i] = valid_messages[b][i] == message[b]; the non-matching slots take the prover's random (e_j, z_j) in order,
// the matching slots take (0, w) for now (u^0 = 1, so a_i = w^n there) and the real (e, z) in cm_finish_kernel.
// fault[b] = 1 when the random vectors run out (no slot matches: the reference indexes past ei_vec and panics).
__global__ void cm_layout_kernel(const uint32_t* valid, const uint32_t* msg, int ml, const uint32_t* e_rand, int el,
                                 const uint32_t* z_rand, const uint32_t* w, int nl, int batch, int M, uint8_t* match,
                                 uint32_t* esel, uint32_t* zsel, uint32_t* esum, uint8_t* fault) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const uint32_t* mg = msg + (size_t)b * ml;
  int j = 0;
  bool bad = false;
  for (int i = 0; i < M; ++i) {
    const uint32_t* v = valid + ((size_t)b * M + i) * ml;
    bool eq = true;
    for (int k = 0; k < ml; ++k) eq &= v[k] == mg[k];
    match[(size_t)b * M + i] = eq ? 1 : 0;
    uint32_t* eo = esel + ((size_t)b * M + i) * el;
    uint32_t* zo = zsel + ((size_t)b * M + i) * nl;
    if (eq || j >= M - 1) {
      if (!eq) bad = true;
      for (int k = 0; k < el; ++k) eo[k] = 0u;
      for (int k = 0; k < nl; ++k) zo[k] = w[(size_t)b * nl + k];
    } else {
      for (int k = 0; k < el; ++k) eo[k] = e_rand[((size_t)b * (M - 1) + j) * el + k];
      for (int k = 0; k < nl; ++k) zo[k] = z_rand[((size_t)b * (M - 1) + j) * nl + k];
      ++j;
    }
  }
  // ei_sum = sum of the M - 1 random challenges mod 2^(32 el)   (:88-89)
  uint32_t* s = esum + (size_t)b * el;
  for (int k = 0; k < el; ++k) s[k] = 0u;
  for (int r = 0; r < M - 1; ++r) {
    uint32_t carry = 0;
    for (int k = 0; k < el; ++k) {
      const unsigned long long t = (unsigned long long)s[k] + e_rand[((size_t)b * (M - 1) + r) * el + k] + carry;
      s[k] = (uint32_t)t;
      carry = (uint32_t)(t >> 32);
    }
  }
  if (bad) fault[b] = 1;
}
cudaError_t launch_cm_layout(const uint32_t* valid, const uint32_t* msg, int ml, const uint32_t* e_rand, int el,
                             const uint32_t* z_rand, const uint32_t* w, int nl, int batch, int M, uint8_t* match, uint32_t* esel,
                             uint32_t* zsel, uint32_t* esum, uint8_t* fault, cudaStream_t st) {
  if (batch <= 0) return cudaSuccess;
  cm_layout_kernel<<<(batch + 63) / 64, 64, 0, st>>>(valid, msg, ml, e_rand, el, z_rand, w, nl, batch, M, match, esel, zsel, esum,
                                                      fault);
  return cudaGetLastError();
}

// out[b] = (a[b] - c[b]) mod 2^(32 limbs)     (BigInt::mod_sub(chal, ei_sum, 2^256), :91)
__global__ void sub_pow2_kernel(const uint32_t* a, const uint32_t* c, int limbs, int batch, uint32_t* out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  uint32_t br = 0;
  for (int k = 0; k < limbs; ++k) {
    const unsigned long long t = (unsigned long long)a[(size_t)b * limbs + k] - c[(size_t)b * limbs + k] - br;
    out[(size_t)b * limbs + k] = (uint32_t)t;
    br = (uint32_t)(t >> 63);
  }
}
cudaError_t launch_sub_pow2(const uint32_t* a, const uint32_t* c, int limbs, int batch, uint32_t* out, cudaStream_t st) {
  if (batch <= 0) return cudaSuccess;
  sub_pow2_kernel<<<(batch + 63) / 64, 64, 0, st>>>(a, c, limbs, batch, out);
  return cudaGetLastError();
}

// The matching slots get the real response (e, z)   (:95-121); one thread per (proof, slot).
__global__ void cm_finish_kernel(const uint8_t* match, const uint32_t* e, int el, const uint32_t* z, int nl, int batch, int M,
                                 uint32_t* esel, uint32_t* zsel) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= batch * M) return;
  if (!match[t]) return;
  const int b = t / M;
  for (int k = 0; k < el; ++k) esel[(size_t)t * el + k] = e[(size_t)b * el + k];
  for (int k = 0; k < nl; ++k) zsel[(size_t)t * nl + k] = z[(size_t)b * nl + k];
}
cudaError_t launch_cm_finish(const uint8_t* match, const uint32_t* e, int el, const uint32_t* z, int nl, int batch, int M,
                             uint32_t* esel, uint32_t* zsel, cudaStream_t st) {
  if (batch <= 0) return cudaSuccess;
  cm_finish_kernel<<<(batch * M + 63) / 64, 64, 0, st>>>(match, e, el, z, nl, batch, M, esel, zsel);
  return cudaGetLastError();
}

// CorrectMessageProof::verify bookkeeping: esum[b] = sum_i e[b][i] mod 2^(32 ol), rows of e are el >= ol limbs wide
// (:130-131); one thread per proof
__global__ void sum_pow2_kernel(const uint32_t* e, int el, int ol, int batch, int M, uint32_t* esum) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  uint32_t* s = esum + (size_t)b * ol;
  for (int k = 0; k < ol; ++k) s[k] = 0u;
  for (int i = 0; i < M; ++i) {
    uint32_t carry = 0;
    for (int k = 0; k < ol; ++k) {
      const unsigned long long t = (unsigned long long)s[k] + e[((size_t)b * M + i) * el + k] + carry;
      s[k] = (uint32_t)t;
      carry = (uint32_t)(t >> 32);
    }
  }
}
cudaError_t launch_sum_pow2(const uint32_t* e, int el, int ol, int batch, int M, uint32_t* esum, cudaStream_t st) {
  if (batch <= 0) return cudaSuccess;
  if (ol > el) return cudaErrorInvalidValue;
  sum_pow2_kernel<<<(batch + 63) / 64, 64, 0, st>>>(e, el, ol, batch, M, esum);
  return cudaGetLastError();
}

// out[b] = reduce over the M rows of proof b: mode 0: AND of ok rows (accept), mode 1: OR (fault); |= into out when or_in
__global__ void rows_reduce_kernel(const uint8_t* rows, int batch, int M, int mode, int merge, uint8_t* out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  uint8_t acc = mode == 0 ? 1 : 0;
  for (int i = 0; i < M; ++i) {
    const uint8_t v = rows[(size_t)b * M + i] ? 1 : 0;
    acc = mode == 0 ? (acc & v) : (acc | v);
  }
  if (merge) acc = mode == 0 ? (acc & (out[b] ? 1 : 0)) : (acc | (out[b] ? 1 : 0));
  out[b] = acc;
}
cudaError_t launch_rows_reduce(const uint8_t* rows, int batch, int M, int mode, int merge, uint8_t* out, cudaStream_t st) {
  if (batch <= 0) return cudaSuccess;
  rows_reduce_kernel<<<(batch + 63) / 64, 64, 0, st>>>(rows, batch, M, mode, merge, out);
  return cudaGetLastError();
}

// fault[b] |= !(x[b] == y[b])     (assert_eq!(chal, ei_sum), correct_message.rs:133)
__global__ void rows_differ_fault_kernel(const uint32_t* x, const uint32_t* y, int limbs, int batch, uint8_t* fault) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  uint32_t d = 0;
  for (int k = 0; k < limbs; ++k) d |= x[(size_t)b * limbs + k] ^ y[(size_t)b * limbs + k];
  if (d) fault[b] = 1;
}
cudaError_t launch_rows_differ_fault(const uint32_t* x, const uint32_t* y, int limbs, int batch, uint8_t* fault, cudaStream_t st) {
  if (batch <= 0) return cudaSuccess;
  rows_differ_fault_kernel<<<(batch + 63) / 64, 64, 0, st>>>(x, y, limbs, batch, fault);
  return cudaGetLastError();
}

}  // namespace zkp
