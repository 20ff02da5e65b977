// K4 (Fiat-Shamir transcript hash) and the RangeProofNi glue kernels K5: everything of
// RangeProof::{generate_encrypted_pairs, generate_proof, verifier_output}
// (reference src/zkproofs/range_proof.rs:128-193, 210-252, 254-355) that is not a
// Paillier encryption -- thirds of the range, w2 = w1 - third, the coin swap,
// challenge bits, response selection, interval predicates, ciphertext equality
// and the AND over the security parameter.  The encryptions themselves run in
// K1 (modexp.cu); the two mulmods (r*r_j mod n, c_j*cipher_x mod n^2) in K3.
#include "kernels.h"
#include "sha256.cuh"

namespace zkp {

// ------------------------------------------------------------------------ K4
// digest[b] = SHA-256( to_bytes(seg0 items of b) || to_bytes(seg1 items) || ... )
__global__ void __launch_bounds__(kShaThreads) sha256_transcript_kernel(const ShaSegs segs, int batch, uint8_t* digest) {
  __shared__ uint32_t wbuf[32 * kShaThreads];
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  Sha256Ring s;
  s.init(wbuf + threadIdx.x, kShaThreads);
  for (int k = 0; k < segs.nseg; ++k) {
    const ShaSeg sg = segs.seg[k];
    const uint32_t* base = sg.base + (size_t)b * sg.batch_stride;
    for (int it = 0; it < sg.count; ++it) {
      const uint32_t* p = base + (size_t)it * sg.limbs;
      // BigInt::to_bytes(): minimal big-endian magnitude, zero -> one 0x00 byte.  Every lane walks all limbs of the item
      // (leading zero limbs contribute 0 bytes), so that the lanes of a warp stay in the same iteration.
      bool started = false;
      for (int i = sg.limbs - 1; i >= 0; --i) {
        const uint32_t v = __ldg(p + i);
        int nb = 4;
        if (!started) {
          nb = v ? 4 - (__clz(v) >> 3) : (i == 0 ? 1 : 0);
          started = v != 0u;
        }
        s.push(v, nb);
      }
    }
  }
  uint32_t out[8];
  s.finish(out);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint32_t v = out[i];
    digest[(size_t)b * 32 + 4 * i + 0] = (uint8_t)(v >> 24);
    digest[(size_t)b * 32 + 4 * i + 1] = (uint8_t)(v >> 16);
    digest[(size_t)b * 32 + 4 * i + 2] = (uint8_t)(v >> 8);
    digest[(size_t)b * 32 + 4 * i + 3] = (uint8_t)v;
  }
}

// ------------------------------------------------------------------------ K4w
// The same digest with ONE WARP per transcript, for batches too small to fill the GPU with one thread each (a batch of 1024
// RangeProofNi transcripts is 32 warps of K4: 3 % of the SMs busy for 10 ms; one proof alone waits the same 10 ms twice).
// The byte stream of a transcript is laid out first - every lane measures items (minimal big-endian length of to_bytes(),
// zero -> 1 byte) and a warp scan turns the lengths into byte offsets in shared memory - after which any block can be
// fetched independently (binary search of the item, then its limbs): each lane fetches ONE block of the next 32 and expands
// its message schedule, and the 32 blocks are then compressed in order with the expanded words broadcast by shuffle.
constexpr int kShaWarpThreads = 128;
__device__ __forceinline__ const uint32_t* sha_item(const ShaSegs& segs, int b, int item, int& limbs) {
  int k = 0, first = 0;
  while (k + 1 < segs.nseg && item >= first + segs.seg[k].count) {
    first += segs.seg[k].count;
    ++k;
  }
  limbs = segs.seg[k].limbs;
  return segs.seg[k].base + (size_t)b * segs.seg[k].batch_stride + (size_t)(item - first) * limbs;
}
__global__ void __launch_bounds__(kShaWarpThreads) sha256_transcript_warp_kernel(const __grid_constant__ ShaSegs segs, int nitems, int batch,
                                                                                uint8_t* digest) {
  extern __shared__ uint32_t s_off_all[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x * (kShaWarpThreads / 32) + warp;
  if (b >= batch) return;
  uint32_t* off = s_off_all + (size_t)warp * (nitems + 1);  // off[i] = first byte of item i in the message; off[nitems] = length
  // ---- lengths and offsets
  uint32_t run = 0;
  for (int base = 0; base < nitems; base += 32) {
    const int it = base + lane;
    uint32_t len = 0;
    if (it < nitems) {
      int limbs;
      const uint32_t* p = sha_item(segs, b, it, limbs);
      int top = limbs - 1;
      uint32_t v = 0;
      while (top >= 0 && (v = __ldg(p + top)) == 0u) --top;
      len = top < 0 ? 1u : (uint32_t)(4 * top + 4 - (__clz(v) >> 3));
    }
    uint32_t inc = len;  // inclusive scan over the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (it < nitems) off[it] = run + inc - len;
    run += __shfl_sync(0xffffffffu, inc, 31);
  }
  if (lane == 0) off[nitems] = run;
  __syncwarp();
  const uint32_t total = run;
  const uint32_t nblocks = (total + 9 + 63) / 64;
  const unsigned long long bits = (unsigned long long)total * 8ull;
  uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  // the 16 message words of block blk (padding and length included), by the lane that owns the block
  auto fetch = [&](uint32_t blk, uint32_t (&m)[16]) {
    const uint32_t p0 = blk * 64u;
    int it = 0, limbs = 0;
    const uint32_t* ptr = nullptr;
    uint32_t end = 0;
    if (blk < nblocks && p0 < total) {
      int lo = 0, hi = nitems - 1;  // the item that holds byte p0: largest i with off[i] <= p0
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (off[mid] <= p0) lo = mid;
        else hi = mid - 1;
      }
      it = lo;
      ptr = sha_item(segs, b, it, limbs);
      end = off[it + 1];
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const uint32_t pos = p0 + 4u * (uint32_t)k;
      uint32_t word = 0;
      if (blk < nblocks) {
        if (pos < total) {
          while (pos >= end) {  // next item (an item is at least one byte long)
            ++it;
            ptr = sha_item(segs, b, it, limbs);
            end = off[it + 1];
          }
          if (pos + 4u <= end) {
            // four bytes of one item: bits [sh, sh + 32) of its little-endian limbs, sh = 8 (bytes below the word)
            const uint32_t sh = 8u * (end - pos - 4u);
            const uint32_t lo32 = __ldg(ptr + (sh >> 5));
            const uint32_t hi32 = (sh & 31u) ? __ldg(ptr + (sh >> 5) + 1) : 0u;
            word = __funnelshift_r(lo32, hi32, sh & 31u);
          } else {
#pragma unroll
            for (int bq = 0; bq < 4; ++bq) {  // the word straddles items or the end of the message
              const uint32_t pb = pos + (uint32_t)bq;
              uint32_t byte;
              if (pb < total) {
                while (pb >= end) {
                  ++it;
                  ptr = sha_item(segs, b, it, limbs);
                        end = off[it + 1];
                }
                const uint32_t le = end - 1u - pb;  // position from the least significant byte
                byte = (__ldg(ptr + (le >> 2)) >> (8u * (le & 3u))) & 0xffu;
              } else {
                byte = pb == total ? 0x80u : 0u;
              }
              word = (word << 8) | byte;
            }
          }
        } else if (pos == total) {
          word = 0x80000000u;
        }
      }
      m[k] = word;
    }
    if (blk == nblocks - 1) {
      m[14] = (uint32_t)(bits >> 32);
      m[15] = (uint32_t)bits;
    }
  };
  // 32 blocks per pass: lane j fetches block base + j and expands its message schedule (W_i + K_i, 64 registers) - the part
  // of SHA-256 that is parallel across blocks - then the blocks are compressed in order, every lane running the rounds on its
  // own copy of the state with the expanded words broadcast from the owning lane.  What is left per block is the serial
  // round chain and one shuffle per round.
  for (uint32_t base = 0; base < nblocks; base += 32u) {
    uint32_t m[16], kw[64];
    fetch(base + (uint32_t)lane, m);
#pragma unroll
    for (int i = 0; i < 16; ++i) kw[i] = m[i] + kSha256K[i];
#pragma unroll
    for (int i = 16; i < 64; ++i) {
      const uint32_t w15 = m[(i + 1) & 15], w2 = m[(i + 14) & 15];
      const uint32_t s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3);
      const uint32_t s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
      const uint32_t wi = m[i & 15] + s0 + m[(i + 9) & 15] + s1;
      m[i & 15] = wi;
      kw[i] = wi + kSha256K[i];
    }
    const int nb = (int)min(32u, nblocks - base);
#pragma unroll 1
    for (int j = 0; j < nb; ++j) {
      uint32_t a = h[0], bb = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        const uint32_t x = __shfl_sync(0xffffffffu, kw[i], j);
        const uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
        const uint32_t ch = (e & f) ^ (~e & g);
        const uint32_t t1 = (hh + x) + S1 + ch;
        const uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
        const uint32_t mj = (a & bb) ^ (a & c) ^ (bb & c);
        const uint32_t t2 = S0 + mj;
        hh = g; g = f; f = e; e = d + t1; d = c; c = bb; bb = a; a = t1 + t2;
      }
      h[0] += a; h[1] += bb; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
  }
  if (lane < 8) {
    const uint32_t v = h[lane];  // h is the same in every lane; lane i writes word i
    uint8_t* o = digest + (size_t)b * 32 + 4 * lane;
    o[0] = (uint8_t)(v >> 24);
    o[1] = (uint8_t)(v >> 16);
    o[2] = (uint8_t)(v >> 8);
    o[3] = (uint8_t)v;
  }
}

// below this many transcripts one warp each (K4w), above it one thread each (K4): 148 SMs x 64 resident warps
constexpr int kShaWarpBatchMax = 148 * 64;

cudaError_t launch_sha256_transcript(const ShaSegs& segs, int batch, uint8_t* digest, cudaStream_t st) {
  if (batch <= 0) return cudaSuccess;
  int nitems = 0;
  for (int k = 0; k < segs.nseg; ++k) nitems += segs.seg[k].count;
  const size_t smem = (size_t)(kShaWarpThreads / 32) * (nitems + 1) * sizeof(uint32_t);
  if (batch < kShaWarpBatchMax && nitems > 0 && smem <= 200 * 1024) {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(sha256_transcript_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
    }
    const int per = kShaWarpThreads / 32;
    sha256_transcript_warp_kernel<<<(batch + per - 1) / per, kShaWarpThreads, smem, st>>>(segs, nitems, batch, digest);
    return cudaGetLastError();
  }
  sha256_transcript_kernel<<<(batch + kShaThreads - 1) / kShaThreads, kShaThreads, 0, st>>>(segs, batch, digest);
  return cudaGetLastError();
}

// ------------------------------------------------ small fixed-width helpers
// (values of w limbs, w <= kMaxW, one thread each; these are ~256-bit numbers)
constexpr int kMaxW = 64;

// q = floor(a / 3)     range.div_floor(3)  (range_proof.rs:133, 218, 264)
__device__ __forceinline__ void div3(uint32_t* q, const uint32_t* a, int w) {
  uint32_t rem = 0;
  for (int i = w - 1; i >= 0; --i) {
    unsigned long long cur = ((unsigned long long)rem << 32) | a[i];
    q[i] = (uint32_t)(cur / 3ull);
    rem = (uint32_t)(cur % 3ull);
  }
}
// -1, 0, 1
__device__ __forceinline__ int cmpw(const uint32_t* a, const uint32_t* b, int w) {
  for (int i = w - 1; i >= 0; --i) {
    if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
  }
  return 0;
}
// d = a + b, returns carry out
__device__ __forceinline__ uint32_t addw(uint32_t* d, const uint32_t* a, const uint32_t* b, int w) {
  uint32_t c = 0;
  for (int i = 0; i < w; ++i) {
    unsigned long long t = (unsigned long long)a[i] + b[i] + c;
    d[i] = (uint32_t)t;
    c = (uint32_t)(t >> 32);
  }
  return c;
}
// d = a - b, returns borrow out
__device__ __forceinline__ uint32_t subw(uint32_t* d, const uint32_t* a, const uint32_t* b, int w) {
  uint32_t br = 0;
  for (int i = 0; i < w; ++i) {
    unsigned long long t = (unsigned long long)a[i] - b[i] - br;
    d[i] = (uint32_t)t;
    br = (uint32_t)(t >> 63);
  }
  return br;
}

// ------------------------------------------------------------- prove: prep
// w1in: [batch*ef][wl] the w1 samples; w: [2][batch*ef][wl] receives w1' (half 0) and
// w2' (half 1) after the swap.  range_proof.rs:141-149.
__global__ void rp_prep_kernel(const uint32_t* range, const uint32_t* w1in, uint32_t* w, const uint8_t* swap, int batch,
                               int ef, int wl, uint8_t* fault) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= batch * ef) return;
  const int b = t / ef;
  uint32_t third[kMaxW], w1[kMaxW], w2[kMaxW];
  div3(third, range + (size_t)b * wl, wl);
  uint32_t* p1 = w + (size_t)t * wl;
  uint32_t* p2 = w + ((size_t)batch * ef + t) * wl;
  for (int i = 0; i < wl; ++i) w1[i] = w1in[(size_t)t * wl + i];
  uint32_t borrow = subw(w2, w1, third, wl);
  if (borrow) fault[b] = 1;  // w1 < third: negative plaintext, outside the engine's domain
  const bool sw = swap[t] != 0;
  for (int i = 0; i < wl; ++i) {
    p1[i] = sw ? w2[i] : w1[i];
    p2[i] = sw ? w1[i] : w2[i];
  }
}

// ---------------------------------------------------------- prove: respond
// range_proof.rs:210-252 for one (proof, i).
__global__ void rp_respond_kernel(RpProveArgs a) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.batch * a.ef) return;
  const int b = t / a.ef, i = t % a.ef;
  const int wl = a.wl, nl = a.nl;
  const size_t half = (size_t)a.batch * a.ef;
  uint32_t e = a.chal ? challenge_bit_raw(a.chal + (size_t)b * a.chal_bytes, a.chal_bytes, i)
                      : challenge_bit(a.digest + (size_t)b * 32, i);
  if (e == 2u) {
    a.fault[b] = 1;
    e = 0;
  }
  const uint32_t* w1 = a.w + (size_t)t * wl;
  const uint32_t* w2 = a.w + (half + t) * wl;
  uint32_t* rw = a.resp_w + (size_t)t * 2 * wl;
  uint32_t* rr = a.resp_r + (size_t)t * 2 * nl;
  if (!e) {
    a.kind[t] = ZKP_RP_OPEN_;
    for (int k = 0; k < wl; ++k) {
      rw[k] = w1[k];
      rw[wl + k] = w2[k];
    }
    const uint32_t* r1 = a.rr + (size_t)t * nl;
    const uint32_t* r2 = a.rr + (half + t) * nl;
    for (int k = 0; k < nl; ++k) {
      rr[k] = r1[k];
      rr[nl + k] = r2[k];
    }
    return;
  }
  uint32_t third[kMaxW], two[kMaxW], s[kMaxW];
  div3(third, a.range + (size_t)b * wl, wl);
  addw(two, third, third, wl);
  const uint32_t* x = a.x + (size_t)b * wl;
  uint32_t carry = addw(s, x, w1, wl);
  const bool first = !carry && cmpw(s, third, wl) > 0 && cmpw(s, two, wl) < 0;
  if (!first) carry = addw(s, x, w2, wl);
  if (carry) a.fault[b] = 1;  // masked_x does not fit w_limbs
  a.kind[t] = first ? ZKP_RP_MASK1_ : ZKP_RP_MASK2_;
  const uint32_t* mr = a.rmul + ((first ? 0 : half) + t) * nl;
  for (int k = 0; k < wl; ++k) {
    rw[k] = s[k];
    rw[wl + k] = 0;
  }
  for (int k = 0; k < nl; ++k) {
    rr[k] = mr[k];
    rr[nl + k] = 0;
  }
}

cudaError_t launch_rp_prep(const uint32_t* range, const uint32_t* w1in, uint32_t* w, const uint8_t* swap, int batch,
                           int ef, int wl, uint8_t* fault, cudaStream_t st) {
  if (wl > kMaxW) return cudaErrorInvalidValue;
  const int total = batch * ef;
  if (total <= 0) return cudaSuccess;
  rp_prep_kernel<<<(total + 127) / 128, 128, 0, st>>>(range, w1in, w, swap, batch, ef, wl, fault);
  return cudaGetLastError();
}

cudaError_t launch_rp_respond(const RpProveArgs& a, cudaStream_t st) {
  if (a.wl > kMaxW) return cudaErrorInvalidValue;
  const int total = a.batch * a.ef;
  if (total <= 0) return cudaSuccess;
  rp_respond_kernel<<<(total + 127) / 128, 128, 0, st>>>(a);
  return cudaGetLastError();
}

// ------------------------------------------------------------ verify: plan
// For each (proof, i): the interval predicates (range_proof.rs:300-309, :338-340) and the Paillier encryptions to run,
// appended to a compact job list (2 for Open, 1 for Mask).  sel[t] tells K3 which ciphertext to multiply by cipher_x
// (0 none, 1 c1, 2 c2).  The plan follows the VARIANT of each response, not the challenge bit: what the verifier has
// to encrypt is written in the proof, so the encryptions can start while the transcript is still being hashed on another
// stream; rp_bits_kernel applies the (bit, variant) match of :276/314/345 afterwards.  A response whose variant
// contradicts its bit is therefore encrypted although the reference would skip it - the verdict is the same `false`.
__global__ void rp_plan_kernel(RpVerifyArgs a) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.batch * a.ef) return;
  const int b = t / a.ef;
  const int wl = a.wl, nl = a.nl;
  const uint8_t k = a.kind[t];
  a.sel[t] = 0;
  if (k > ZKP_RP_MASK2_) {  // the reference has no such variant
    a.fault[b] = 1;
    a.ok[t] = 0;
    return;
  }
  uint32_t third[kMaxW], two[kMaxW];
  div3(third, a.range + (size_t)b * wl, wl);
  addw(two, third, third, wl);
  const uint32_t* rw = a.resp_w + (size_t)t * 2 * wl;
  const uint32_t* rr = a.resp_r + (size_t)t * 2 * nl;
  bool ok;
  int njobs;
  if (k == ZKP_RP_OPEN_) {
    const uint32_t* w1 = rw;
    const uint32_t* w2 = rw + wl;
    const bool f1 = cmpw(w2, third, wl) < 0 && cmpw(w1, third, wl) > 0 && cmpw(w1, two, wl) < 0;
    const bool f2 = cmpw(w1, third, wl) < 0 && cmpw(w2, third, wl) > 0 && cmpw(w2, two, wl) < 0;
    ok = f1 || f2;
    njobs = 2;
  } else {
    ok = !(cmpw(rw, third, wl) < 0 || cmpw(rw, two, wl) > 0);
    njobs = 1;
    a.sel[t] = k;  // 1 -> c1, 2 -> c2 (j == 1 ? c1 : c2; only 1 and 2 are representable here)
  }
  a.ok[t] = ok ? 1 : 0;
  // The encryptions are still performed when a predicate already failed, as in the reference.
  const unsigned slot = atomicAdd(a.count, (unsigned)njobs);
  for (int j = 0; j < njobs; ++j) {
    a.tag[slot + j] = ((uint32_t)t << 1) | (uint32_t)j;
    uint32_t* jb = a.jobs_base + (size_t)(slot + j) * nl;
    uint32_t* jp = a.jobs_plain + (size_t)(slot + j) * wl;
    for (int q = 0; q < nl; ++q) jb[q] = rr[(size_t)j * nl + q];
    for (int q = 0; q < wl; ++q) jp[q] = rw[(size_t)j * wl + q];
  }
}

// ------------------------------------------------------------ verify: bits
// The challenge bit of every (proof, i) against the variant of its response (range_proof.rs:276/314/345: `_ => false`);
// a bit index past the digest (leading zero bytes stripped) is where the reference panics (:273-274).
__global__ void rp_bits_kernel(RpVerifyArgs a) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.batch * a.ef) return;
  const int b = t / a.ef, i = t % a.ef;
  const uint32_t e = a.chal ? challenge_bit_raw(a.chal + (size_t)b * a.chal_bytes, a.chal_bytes, i)
                            : challenge_bit(a.digest + (size_t)b * 32, i);
  const uint8_t k = a.kind[t];
  if (e == 2u) {
    a.fault[b] = 1;
    a.ok[t] = 0;
  } else if (k <= ZKP_RP_MASK2_ && (e == 0u) != (k == ZKP_RP_OPEN_)) {
    a.ok[t] = 0;
  }
}

// ----------------------------------------------------------- verify: check
// One warp per executed encryption: compare with the committed ciphertext
// (Open: c1[i] / c2[i], range_proof.rs:293-298) or with c_j * cipher_x mod n^2
// (Mask, :321-337); clear ok[t] on mismatch.
__global__ void rp_check_kernel(RpVerifyArgs a) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const unsigned njobs = *a.count;
  if ((unsigned)warp >= njobs) return;
  const uint32_t tg = a.tag[warp];
  const uint32_t t = tg >> 1, j = tg & 1u;
  const int nnl = 2 * a.nl;
  const size_t half = (size_t)a.batch * a.ef;
  const uint32_t* want;
  if (a.kind[t] == ZKP_RP_OPEN_) want = a.c + ((j ? half : 0) + t) * nnl;
  else want = a.cmul + (size_t)t * nnl;
  const uint32_t* got = a.jobs_out + (size_t)warp * nnl;
  uint32_t diff = 0;
  for (int q = lane; q < nnl; q += 32) diff |= got[q] ^ want[q];
  diff = __reduce_or_sync(0xffffffffu, diff);
  if (lane == 0 && diff) a.ok[t] = 0;
}

// accept[b] = AND_i ok[b][i], and not faulted (range_proof.rs:350-354)
__global__ void rp_accept_kernel(const uint8_t* ok, const uint8_t* fault, int batch, int ef, uint8_t* accept) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= batch) return;
  uint32_t all = 1;
  for (int i = lane; i < ef; i += 32) all &= ok[(size_t)warp * ef + i];
  all = __reduce_and_sync(0xffffffffu, all);
  if (lane == 0) accept[warp] = (uint8_t)((all & 1u) && !fault[warp]);
}

cudaError_t launch_rp_plan(const RpVerifyArgs& a, cudaStream_t st) {
  if (a.wl > kMaxW) return cudaErrorInvalidValue;
  const int total = a.batch * a.ef;
  if (total <= 0) return cudaSuccess;
  rp_plan_kernel<<<(total + 127) / 128, 128, 0, st>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_rp_bits(const RpVerifyArgs& a, cudaStream_t st) {
  const int total = a.batch * a.ef;
  if (total <= 0) return cudaSuccess;
  rp_bits_kernel<<<(total + 127) / 128, 128, 0, st>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_rp_check(const RpVerifyArgs& a, cudaStream_t st) {
  const long long maxjobs = 2ll * a.batch * a.ef;
  if (maxjobs <= 0) return cudaSuccess;
  rp_check_kernel<<<(unsigned)((maxjobs * 32 + 255) / 256), 256, 0, st>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_rp_accept(const uint8_t* ok, const uint8_t* fault, int batch, int ef, uint8_t* accept,
                             cudaStream_t st) {
  if (batch <= 0) return cudaSuccess;
  rp_accept_kernel<<<(batch * 32 + 255) / 256, 256, 0, st>>>(ok, fault, batch, ef, accept);
  return cudaGetLastError();
}

}  // namespace zkp
