// Device SHA-256 (FIPS 180-4) as a byte-stream absorber, used by the Fiat-Shamir
// transcript kernel K4 and by the NiCorrectKeyProof rho derivation.
//
// Replaces sha2::Sha256 as driven by compute_digest (reference
// src/zkproofs/utils.rs:9-22): the message is the plain concatenation of
// BigInt::to_bytes() of each item -- minimal-length big-endian magnitude, zero
// encoded as the single byte 0x00 -- with no length prefixes.
//
// One thread owns one hash.  The 16-word block buffer lives in shared memory,
// word-interleaved across the CTA (word i of thread t at w[i * stride + t]) so
// the run-time block index never forces a local-memory array and accesses are
// bank-conflict free.
#pragma once
#include <stdint.h>

namespace zkp {

static __constant__ uint32_t kSha256K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

__device__ __forceinline__ uint32_t rotr32(uint32_t x, int n) { return __funnelshift_r(x, x, n); }

struct Sha256 {
  uint32_t h[8];
  uint32_t* w;         // this thread's column of the shared block buffer
  int stride;          // distance between consecutive words of the column
  int idx;             // words filled in the current block (0..15)
  uint32_t pend;       // pending bytes (< 4), right-aligned
  int npend;
  unsigned long long total;  // message bytes absorbed

  __device__ __forceinline__ void init(uint32_t* col, int stride_) {
    h[0] = 0x6a09e667; h[1] = 0xbb67ae85; h[2] = 0x3c6ef372; h[3] = 0xa54ff53a;
    h[4] = 0x510e527f; h[5] = 0x9b05688c; h[6] = 0x1f83d9ab; h[7] = 0x5be0cd19;
    w = col;
    stride = stride_;
    idx = 0;
    pend = 0;
    npend = 0;
    total = 0;
  }

  __device__ __noinline__ void compress() {
    uint32_t m[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) m[i] = w[i * stride];
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
    for (int i = 0; i < 64; ++i) {
      uint32_t wi;
      if (i < 16) {
        wi = m[i];
      } else {
        uint32_t w15 = m[(i + 1) & 15], w2 = m[(i + 14) & 15];
        uint32_t s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3);
        uint32_t s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
        wi = m[i & 15] + s0 + m[(i + 9) & 15] + s1;
        m[i & 15] = wi;
      }
      uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
      uint32_t ch = (e & f) ^ (~e & g);
      uint32_t t1 = hh + S1 + ch + kSha256K[i] + wi;
      uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
      uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
      uint32_t t2 = S0 + mj;
      hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
  }

  __device__ __forceinline__ void emit(uint32_t word) {
    w[idx * stride] = word;
    if (++idx == 16) {
      compress();
      idx = 0;
    }
  }
  // absorb the low `nb` (1..4) bytes of v, most significant of them first
  __device__ __forceinline__ void push(uint32_t v, int nb) {
    total += (unsigned long long)nb;
    unsigned long long acc = ((unsigned long long)pend << (8 * nb)) | (nb == 4 ? v : (v & ((1u << (8 * nb)) - 1u)));
    int n = npend + nb;
    if (n >= 4) {
      emit((uint32_t)(acc >> (8 * (n - 4))));
      n -= 4;
      acc &= (1ull << (8 * n)) - 1ull;
    }
    pend = (uint32_t)acc;
    npend = n;
  }
  __device__ __forceinline__ void push_word(uint32_t v) {  // 4 bytes, big-endian
    total += 4ull;
    if (npend == 0) {
      emit(v);
    } else {
      int sh = 8 * npend;
      emit((pend << (32 - sh)) | (v >> sh));
      pend = v & ((1u << sh) - 1u);
    }
  }
  // BigInt::to_bytes() of a little-endian limb array read through `get(i)`:
  // minimal big-endian magnitude; zero -> one 0x00 byte.
  template <class Get>
  __device__ __forceinline__ void push_bigint(int limbs, Get get) {
    int top = limbs - 1;
    uint32_t v = 0;
    while (top >= 0 && (v = get(top)) == 0u) --top;
    if (top < 0) {
      push(0u, 1);
      return;
    }
    int nb = 4 - (__clz(v) >> 3);
    push(v, nb);
    for (int i = top - 1; i >= 0; --i) push_word(get(i));
  }
  // 32-byte digest as 8 big-endian words
  __device__ __forceinline__ void finish(uint32_t (&out)[8]) {
    unsigned long long bits = total * 8ull;
    push(0x80u, 1);
    while (npend != 0) push(0u, 1);
    while (idx != 14) emit(0u);
    emit((uint32_t)(bits >> 32));
    emit((uint32_t)bits);
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = h[i];
  }
};

// The same absorber for the transcript kernel K4, where the 32 lanes of a warp hash 32 transcripts of (almost) the same
// length side by side.  Items that lost a leading zero byte make the lanes drift a word or two apart; with one 16-word
// buffer per lane each lane would call the compression function in a different loop iteration and the warp would run it
// twice per block (ncu: 4 500 warp instructions per block).  Here every lane has a 32-word ring, words are emitted at ONE
// code site, and the warp compresses when EVERY converged lane holds a full block (or some lane's ring is full), so the
// 64 rounds run once per block for the whole warp.
struct Sha256Ring {
  uint32_t h[8];
  uint32_t* w;        // this thread's column of the shared ring: word i at w[(i & 31) * stride]
  int stride;
  uint32_t wr, rd;    // words written / consumed so far
  uint32_t pend;      // pending bytes (< 4), right-aligned
  int npend;
  unsigned long long total;

  __device__ __forceinline__ void init(uint32_t* col, int stride_) {
    h[0] = 0x6a09e667; h[1] = 0xbb67ae85; h[2] = 0x3c6ef372; h[3] = 0xa54ff53a;
    h[4] = 0x510e527f; h[5] = 0x9b05688c; h[6] = 0x1f83d9ab; h[7] = 0x5be0cd19;
    w = col;
    stride = stride_;
    wr = rd = 0;
    pend = 0;
    npend = 0;
    total = 0;
  }
  __device__ __noinline__ void compress() {
    uint32_t m[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) m[i] = w[((rd + i) & 31u) * stride];
    rd += 16;
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
    for (int i = 0; i < 64; ++i) {
      uint32_t wi;
      if (i < 16) {
        wi = m[i];
      } else {
        uint32_t w15 = m[(i + 1) & 15], w2 = m[(i + 14) & 15];
        uint32_t s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3);
        uint32_t s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
        wi = m[i & 15] + s0 + m[(i + 9) & 15] + s1;
        m[i & 15] = wi;
      }
      uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
      uint32_t ch = (e & f) ^ (~e & g);
      uint32_t t1 = hh + S1 + ch + kSha256K[i] + wi;
      uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
      uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
      uint32_t t2 = S0 + mj;
      hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
  }
  // absorb the low nb (0..4) bytes of v, most significant first; emits at most one word.  The only emit site of the
  // streaming phase: every lane of the warp passes here once per limb, so the compress trigger below is warp-uniform.
  __device__ __forceinline__ void push(uint32_t v, int nb) {
    total += (unsigned long long)nb;
    const unsigned long long acc = ((unsigned long long)pend << (8 * nb)) | (nb == 4 ? v : (v & ((1u << (8 * nb)) - 1u)));
    int n = npend + nb;
    const bool out = n >= 4;
    if (out) {
      w[(wr & 31u) * stride] = (uint32_t)(acc >> (8 * (n - 4)));
      ++wr;
      n -= 4;
    }
    pend = (uint32_t)(acc & ((1ull << (8 * n)) - 1ull));
    npend = n;
    const unsigned am = __activemask();
    const bool ready = wr - rd >= 16u;
    if (__all_sync(am, ready) || __any_sync(am, wr - rd >= 32u)) {
      if (ready) compress();
    }
  }
  __device__ __forceinline__ void finish(uint32_t (&out)[8]) {
    const unsigned long long bits = total * 8ull;
    push(0x80u, 1);
    while (npend != 0) push(0u, 1);
    while ((wr & 15u) != 14u) push(0u, 4);
    push((uint32_t)(bits >> 32), 4);
    push((uint32_t)bits, 4);
    while (wr - rd >= 16u) compress();
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = h[i];
  }
};

// Challenge bit i of ChallengeBits(BigInt::to_bytes(digest)) read MSB-first as
// BitVec::from_bytes does (reference range_proof.rs:221,225,267,273): the
// digest's leading zero BYTES are stripped before indexing (to_bytes of the
// BigInt), zero keeps one byte.  Returns 0/1, or 2 when the index is past the
// end of the stripped byte string (the reference panics there).
__device__ __forceinline__ uint32_t challenge_bit(const uint8_t* digest, int i) {
  int lead = 0;
  while (lead < 31 && digest[lead] == 0) ++lead;
  int byte = i >> 3;
  if (byte >= 32 - lead) return 2u;
  return (digest[lead + byte] >> (7 - (i & 7))) & 1u;
}

// The interactive RangeProof hands the verifier's ChallengeBits bytes over as they are (range_proof.rs:86-91,
// 221): BitVec::from_bytes(&e.0)[i] without any BigInt round trip, hence no stripping.
__device__ __forceinline__ uint32_t challenge_bit_raw(const uint8_t* e, int nbytes, int i) {
  const int byte = i >> 3;
  if (byte >= nbytes) return 2u;
  return (e[byte] >> (7 - (i & 7))) & 1u;
}

}  // namespace zkp
