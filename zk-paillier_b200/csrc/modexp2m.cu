// K1m: Paillier encryption c = (1 + m n) r^n mod n^2 by MONTGOMERY ARITHMETIC IN TWO-DIGIT BASE-n FORM.
//
// Replaces the same reference lines as K1 (Paillier::encrypt_with_chosen_randomness at
// range_proof.rs:165,179,280,286,330; zero_enc_proof.rs:46,73; correct_ciphertext.rs:45,73;
// multiplication_proof.rs:63,72,118,125; verlin_proof.rs:157) with half the limb products.
//
// K1 already runs the multiplier pipe at its IMAD.WIDE issue limit, so the only way to go faster is to do fewer
// multiplies.  An element x of Z_{n^2} is held as the pair (X0, X1), x = X0 + X1 n, 0 <= X0, X1 < n.  With
// W = 2^(32 S) > n, the CIOS Montgomery multiplication of X0 and Y0 modulo n produces, besides Z0 = X0 Y0 / W mod n,
// the quotient digits q it added, and  X0 Y0 = Z0' W - q n  holds between integers.  Hence modulo n^2
//     x y / W  =  Z0 + ( (X0 Y1 + X1 Y0 - q_eff) / W  mod n ) n
// and the second digit is one more Montgomery reduction modulo n whose accumulator starts at a non-negative
// representative of -q_eff (init = W + K_lo - q, K_lo = -W mod n; or W - q when Z0 = Z0' - n was taken).
// No true quotient and no Barrett: only half-width CIOS rows of mp_coop.cuh.  A squaring costs 4 S^2 limb products
// (x^2: X0 X0 and X0 (2 X1)), a multiplication 5 S^2 (X0 Y1 + X1 Y0 under one reduction), against 8 S^2 for CIOS
// modulo n^2 (S = limbs of n): 40.8 M IMAD.WIDE per 2048-bit encryption instead of 79.3 M.  Model and bounds:
// tests/models/mont2d_model.py.
//
// Layout: one encryption per group of T lanes, L limbs of each digit per lane (Mp<8,8> for a 2048-bit n), the
// two digits in separate register arrays; window table of 16 odd powers + x^2 as digit pairs in an L2-resident
// scratch (each lane re-reads only limbs it wrote); schedule, K_lo and each pass's bases / plaintexts staged by
// 1-D TMA bulk copies.  The whole exponentiation - entering Montgomery form, the table, the sliding-window
// ladder, the Paillier factor (1 + m n) = pair (1, m), leaving Montgomery form - is ONE op list run by one loop
// with a single squaring site and a single multiplication site, so the code stays small.
#include <stdlib.h>

#include <vector>

#include "kernels.h"
#include "mp_coop.cuh"
#include "tma.cuh"

namespace zkp {

constexpr uint32_t OP_NONE = 0xffu, OP_Y_CONST = 0xfeu, OP_Y_PLAIN = 0xfdu;

struct Enc2mParams {
  Enc2mKey key;
  const uint32_t* bases;
  const uint32_t* plain;
  uint32_t* out;
  uint32_t* table;
  int base_limbs, plain_limbs, out_limbs, jobs, ops_pad, slots;
  uint32_t zero;
  const unsigned* jobs_dev;
};

// MODE 0: the two halves of a squaring / multiplication are separate instances of the row loop (multiplier limbs by SHFL).
// MODE 1: both halves of a squaring run through one copy of the row loop (multiplier picked by SEL): smaller code.
// MODE 2: multiplier limbs and quotient digits go through a per-group shared-memory scratch (kSg2m words), the row loop
//         is rolled U pairs of steps deep: smallest code, no multiplier shuffles.
// MODE 4: pair rows (Mp::cios_pair: two multiplier limbs and a two-limb quotient per step).  The serial quotient chain runs
//         once per two rows: for launches with few warps per sub-partition (K2h's narrow-lane and latency layouts).
template <int T, int L, int U, int MODE = 0>
struct TwoDigit {
  using M = Mp<T, L>;
  static constexpr int S = T * L;
  static constexpr int kSg = 3 * S + 4;  // MODE 2 scratch per group: multiplier row | second multiplier row | q row (+4: bank skew)

  // init + top 2^(32 S) = W + (delta ? 0 : K_lo) - q   (non-negative, == -q_eff mod n, < W + n)
  static __device__ __forceinline__ void init_from_q(uint32_t (&init)[L], uint32_t& top, const uint32_t (&q)[L], uint32_t delta,
                                                     const uint32_t* s_klo, int lane) {
    const int g = lane & (T - 1);
    uint32_t m[L];
    M::load(m, s_klo + g * L);
#pragma unroll
    for (int j = 0; j < L; ++j) m[j] = delta ? 0u : m[j];
    const uint32_t bo = M::sub_full(init, m, q, lane);
    top = 1u - bo;
  }

  // (x0, x1) <- (x0, x1)^2 / W
  static __device__ __forceinline__ void sqr(uint32_t (&x0)[L], uint32_t (&x1)[L], const uint32_t (&n)[L], uint32_t n0inv,
                                             const uint32_t* s_klo, int lane, uint32_t zr, uint32_t* sg = nullptr, uint32_t n0hi = 0u) {
    uint32_t q[L], z0[L];
#pragma unroll
    for (int j = 0; j < L; ++j) q[j] = 0;
    if constexpr (MODE == 4) {
      const uint32_t delta = M::template mont_mul_p<false, true, 1, U>(z0, x0, x0, n, n0inv, n0hi, lane, z0, 0u, q, zr);
      uint32_t top;
      init_from_q(q, top, q, delta, s_klo, lane);
      M::mod_double(x1, n, lane);
      M::template mont_mul_p<true, false, 2, U>(x1, x0, x1, n, n0inv, n0hi, lane, q, top, q, zr);
    } else if constexpr (MODE == 0 || MODE == 3) {
      uint32_t delta;
      if constexpr (MODE == 3) {  // symmetric squaring: every pair of lane blocks multiplied once, then S reduction-only rows
        uint32_t plo[L], phi[L];
        M::sqr_product(plo, phi, x0, lane);
        delta = M::mont_redc_x(z0, plo, phi, n, n0inv, lane, q, zr);
      } else {
        delta = M::template mont_mul_x<false, true, 1, U>(z0, x0, x0, n, n0inv, lane, z0, 0u, q, zr);
      }
      uint32_t top;
      init_from_q(q, top, q, delta, s_klo, lane);
      M::mod_double(x1, n, lane);
      M::template mont_mul_x<true, false, 2, U>(x1, x0, x1, n, n0inv, lane, q, top, q, zr);
    } else {
      const int g = lane & (T - 1);
      uint32_t init[L], res[L], top = 0u;
#pragma unroll
      for (int j = 0; j < L; ++j) init[j] = 0;
#pragma unroll 1
      for (int ph = 0; ph < 2; ++ph) {
        uint32_t delta;
        if constexpr (MODE == 1) {
          delta = M::template mont_mul_sel<1>(res, x0, x0, x1, ph != 0, n, n0inv, lane, init, top, q, zr);
        } else {
          __syncwarp();
          if (ph == 0) M::store(sg + g * L, x0);
          else M::store(sg + g * L, x1);
          __syncwarp();
          delta = M::template mont_mul_s<U>(res, x0, sg, sg + 2 * S, n, n0inv, lane, init, top, zr);
        }
        if (ph == 0) {
          if (MODE == 2) {
            __syncwarp();
            M::load(q, sg + 2 * S + g * L);
          }
#pragma unroll
          for (int j = 0; j < L; ++j) z0[j] = res[j];
          init_from_q(init, top, q, delta, s_klo, lane);
          M::mod_double(x1, n, lane);
        }
      }
#pragma unroll
      for (int j = 0; j < L; ++j) x1[j] = res[j];
    }
#pragma unroll
    for (int j = 0; j < L; ++j) x0[j] = z0[j];
  }

  // (x0, x1) <- (x0, x1) (y0, y1) / W
  static __device__ __forceinline__ void mul(uint32_t (&x0)[L], uint32_t (&x1)[L], const uint32_t (&y0)[L], const uint32_t (&y1)[L],
                                             const uint32_t (&n)[L], uint32_t n0inv, const uint32_t* s_klo, int lane, uint32_t zr,
                                             uint32_t* sg = nullptr, uint32_t n0hi = 0u) {
    uint32_t q[L], z0[L];
#pragma unroll
    for (int j = 0; j < L; ++j) q[j] = 0;
    if constexpr (MODE == 4) {
      const uint32_t delta = M::template mont_mul_p<false, true, 1, U>(z0, x0, y0, n, n0inv, n0hi, lane, z0, 0u, q, zr);
      uint32_t top;
      init_from_q(q, top, q, delta, s_klo, lane);
      M::template mont_mul2_p<U>(x1, x0, y1, x1, y0, n, n0inv, n0hi, lane, q, top, zr);
    } else if constexpr (MODE != 2) {
      const uint32_t delta = M::template mont_mul_x<false, true, 1, U>(z0, x0, y0, n, n0inv, lane, z0, 0u, q, zr);
      uint32_t top;
      init_from_q(q, top, q, delta, s_klo, lane);
      M::template mont_mul2_x<U>(x1, x0, y1, x1, y0, n, n0inv, lane, q, top, zr);  // X0 Y1 + X1 Y0 under one reduction
    } else {
      const int g = lane & (T - 1);
      __syncwarp();
      M::store(sg + g * L, y0);
      M::store(sg + S + g * L, y1);
      __syncwarp();
      const uint32_t delta = M::template mont_mul_s<U>(z0, x0, sg, sg + 2 * S, n, n0inv, lane, q, 0u, zr);  // q == 0 here
      __syncwarp();
      M::load(q, sg + 2 * S + g * L);
      uint32_t top;
      init_from_q(q, top, q, delta, s_klo, lane);
      M::template mont_mul2_s<U>(x1, x0, sg + S, x1, sg, n, n0inv, lane, q, top, zr);
    }
#pragma unroll
    for (int j = 0; j < L; ++j) x0[j] = z0[j];
  }

  // x += c (c in {0,1}, uniform across the group) over the whole digit
  static __device__ __forceinline__ void add_bit(uint32_t (&x)[L], uint32_t c, int lane) {
    const int g = lane & (T - 1);
    add_cc(x[0], g == 0 ? c : 0u);
#pragma unroll
    for (int j = 1; j < L; ++j) addc_cc(x[j], 0);
    uint32_t co = addc_out();
    uint32_t top;
    uint32_t cin = M::resolve(co != 0, M::all_ones(x), lane, top);
    M::add_small(x, cin);
  }

  // (x0, x1) <- (x0, x1) + (y0, y1) mod n^2, all four digits below n
  static __device__ __forceinline__ void pair_add(uint32_t (&x0)[L], uint32_t (&x1)[L], const uint32_t (&y0)[L], const uint32_t (&y1)[L],
                                                  const uint32_t (&n)[L], int lane) {
    const uint32_t ovf = M::add_full(x0, y0, lane);
    uint32_t d[L];
    const uint32_t borrow = M::sub_full(d, x0, n, lane);
    const bool take = (ovf != 0u) || (borrow == 0u);
#pragma unroll
    for (int j = 0; j < L; ++j) x0[j] = take ? d[j] : x0[j];
    // digit 1: x1 + y1 + take < 2n.  (y1 + take) may reach n; the sum still needs one subtraction at most.
    const uint32_t ovf1 = M::add_full(x1, y1, lane);
    add_bit(x1, take ? 1u : 0u, lane);  // cannot overflow past ovf1: x1 + y1 + 1 <= 2n - 1 < 2^(32 S + 1)
    const uint32_t borrow1 = M::sub_full(d, x1, n, lane);
    const bool take1 = (ovf1 != 0u) || (borrow1 == 0u);
#pragma unroll
    for (int j = 0; j < L; ++j) x1[j] = take1 ? d[j] : x1[j];
  }

  // The value of a little-endian row of `limbs` <= 2S limbs as a digit pair.  limbs <= S: the pair (row, 0), with
  // digit 0 possibly above n (fine as the first operand of mul).  Wider rows z = Zlo + Zhi W:
  //   mul((Zlo, 0), pair(W)) + mul((Zhi, 0), pair(W^2)) = Zlo + Zhi W   (mod n^2), both digits reduced.
  // consts: K_lo | pair(W^2 mod n^2) | pair(W mod n^2).
  static __device__ __forceinline__ void entry_pair(uint32_t (&x0)[L], uint32_t (&x1)[L], const uint32_t* row, int limbs,
                                                    const uint32_t* consts, const uint32_t (&n)[L], uint32_t n0inv,
                                                    const uint32_t* s_klo, int lane, uint32_t zr, uint32_t* sg = nullptr) {
    entry_pair_u(x0, x1, row, limbs, limbs > S, consts, n, n0inv, s_klo, lane, zr, sg);
  }
  // The same with the choice of path handed in: `wide` must be uniform across the warp (the wide path shuffles), `limbs`
  // need not be - a row of at most S limbs taken through the wide path has a zero high half.
  static __device__ __forceinline__ void entry_pair_u(uint32_t (&x0)[L], uint32_t (&x1)[L], const uint32_t* row, int limbs, bool wide,
                                                      const uint32_t* consts, const uint32_t (&n)[L], uint32_t n0inv,
                                                      const uint32_t* s_klo, int lane, uint32_t zr, uint32_t* sg = nullptr, uint32_t n0hi = 0u) {
    const int g = lane & (T - 1);
    if (!wide) {
      M::load_ext(x0, row, limbs, g);
#pragma unroll
      for (int j = 0; j < L; ++j) x1[j] = 0;
      return;
    }
    uint32_t h0[L], h1[L], y0[L], y1[L];
#pragma unroll 1
    for (int half = 1; half >= 0; --half) {
      const uint32_t* cp = consts + S + (half ? 0 : 2 * S) + g * L;
      M::load(y0, cp);
      M::load(y1, cp + S);
      if (half) M::load_ext(x0, row + S, limbs - S, g);
      else M::load_ext(x0, row, limbs < S ? limbs : S, g);
#pragma unroll
      for (int j = 0; j < L; ++j) x1[j] = 0;
      mul(x0, x1, y0, y1, n, n0inv, s_klo, lane, zr, sg, n0hi);
      if (half) {
#pragma unroll
        for (int j = 0; j < L; ++j) {
          h0[j] = x0[j];
          h1[j] = x1[j];
        }
      }
    }
    pair_add(x0, x1, h0, h1, n, lane);
  }

  // out row = X0 + X1 n (< n^2): plain product rows with X0 as the initial accumulator; the limb leaving lane 0
  // at row i is limb i of the result and goes to the lane that owns it
  static __device__ __forceinline__ void assemble_store(uint32_t* o, int out_limbs, bool valid, const uint32_t (&x0)[L],
                                                        const uint32_t (&x1)[L], const uint32_t (&n)[L], int lane) {
    const int g = lane & (T - 1);
    uint32_t E[L + 2], O[L + 2], lo[L], hi[L];
#pragma unroll
    for (int j = 0; j < L; ++j) {
      E[j] = x0[j];
      O[j] = 0;
      lo[j] = 0;
    }
    E[L] = E[L + 1] = O[L] = O[L + 1] = 0;
#pragma unroll 1
    for (int owner = 0; owner < T; ++owner) {
      const bool mine = g == owner;
#pragma unroll
      for (int j = 0; j < L; j += 2) {
        const uint32_t b0 = __shfl_sync(ZKP_FULL, n[j], owner, T);
        const uint32_t b1 = __shfl_sync(ZKP_FULL, n[j + 1], owner, T);
        M::mul_step(E, O, x1, b0, g);
        const uint32_t v0 = __shfl_sync(ZKP_FULL, E[0], 0, T);
        M::mul_step(O, E, x1, b1, g);
        const uint32_t v1 = __shfl_sync(ZKP_FULL, O[0], 0, T);
        lo[j] = mine ? v0 : lo[j];
        lo[j + 1] = mine ? v1 : lo[j + 1];
      }
    }
    M::mul_finish(hi, E, O, lane);
    if (valid) {
      M::store_ext(o, lo, out_limbs, g);
      if (out_limbs > S) M::store_ext(o + S, hi, out_limbs - S, g);
    }
  }
};

// WIDE: rows wider than n (bases up to 2S limbs: ciphertexts as randomness; plaintexts up to 2S limbs: unreduced z1).
// Kept out of the common instantiation so that its extra live state does not cost the hot loop registers.
template <int T, int L, int U, int MINB, bool WIDE, int MODE = 0>
__global__ void __launch_bounds__(kCtaThreads, MINB) enc2m_kernel(const Enc2mParams p) {
  using M = Mp<T, L>;
  using TD = TwoDigit<T, L, U, MODE>;
  constexpr int S = T * L;
  constexpr int G = kCtaThreads / T;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* s_ops = reinterpret_cast<uint32_t*>(smem_raw + 16);
  uint32_t* s_klo = s_ops + p.ops_pad;
  uint32_t* s_bases = s_klo + S;
  uint32_t* s_plain = s_bases + G * p.base_limbs;
  uint32_t* s_pair = s_plain + G * p.plain_limbs;

  const int lane = threadIdx.x & 31;
  const int g = lane & (T - 1);
  const int grp = threadIdx.x / T;
  uint32_t* sg = MODE == 2 ? s_pair + G * 2 * S + grp * TD::kSg : nullptr;  // multiplier / quotient rows of this group
  const uint32_t n0inv = p.key.n0inv;
  const uint32_t zr = p.zero;  // always 0, but opaque to ptxas (see cios_step)
  uint32_t n[L];
  M::load(n, p.key.mod + g * L);

  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  uint32_t phase = 0;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, (uint32_t)(p.ops_pad + S) * 4u);
    bulk_g2s(s_ops, p.key.ops, (uint32_t)p.ops_pad * 4u, bar);
    bulk_g2s(s_klo, p.key.consts, (uint32_t)S * 4u, bar);
  }
  mbar_wait(bar, phase);
  phase ^= 1;

  uint32_t* tab = p.table + (size_t)(blockIdx.x * G + grp) * p.slots * 2 * S + g * L;
  uint32_t* pair = s_pair + grp * 2 * S;
  const int jobs = p.jobs_dev ? min((int)*p.jobs_dev, p.jobs) : p.jobs;
  const int npass = (jobs + G - 1) / G;
  for (int cj = blockIdx.x; cj < npass; cj += gridDim.x) {
    const int job0 = cj * G;
    const int nvalid = min(G, jobs - job0);
    if (threadIdx.x == 0) {
      uint32_t bb = (uint32_t)(nvalid * p.base_limbs) * 4u;
      uint32_t pb = p.plain ? (uint32_t)(nvalid * p.plain_limbs) * 4u : 0u;
      mbar_expect_tx(bar, bb + pb);
      bulk_g2s(s_bases, p.bases + (size_t)job0 * p.base_limbs, bb, bar);
      if (p.plain) bulk_g2s(s_plain, p.plain + (size_t)job0 * p.plain_limbs, pb, bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    const bool valid = grp < nvalid;
    const int src = valid ? grp : 0;

    // the last multiplier: the pair (1, m) = 1 + m n in plain form (it also takes the result out of Montgomery form)
    uint32_t x0[L], x1[L], y0[L], y1[L];
    if (WIDE && p.plain && p.plain_limbs > S) {
      // m wider than n (CiphertextProof's unreduced z1): 1 + m n = 1 + (m mod n) n, and
      // m mod n = Mlo W / W + Mhi W^2 / W  (two Montgomery products by W mod n and W^2 mod n)
      const uint32_t* mrow = s_plain + src * p.plain_limbs;
      M::load_ext(x1, mrow, S, g);
      M::load(y0, p.key.consts + 3 * S + g * L);
      M::mont_mul(x1, x1, y0, n, n0inv, lane);
      M::load_ext(y1, mrow + S, p.plain_limbs - S, g);
      M::load(y0, p.key.consts + S + g * L);
      M::mont_mul(y1, y1, y0, n, n0inv, lane);
      M::add_mod(x1, y1, n, lane);
    } else if (p.plain) {
      M::load_ext(x1, s_plain + src * p.plain_limbs, p.plain_limbs, g);
    } else {
#pragma unroll
      for (int j = 0; j < L; ++j) x1[j] = 0;
    }
    M::set_small(x0, 1u, g);
    M::store(pair + g * L, x0);
    M::store(pair + S + g * L, x1);
    __syncwarp();

    // the pair (r, 0) (r may exceed n), or the reduced pair of a base wider than n
    if (WIDE) {
      TD::entry_pair(x0, x1, s_bases + src * p.base_limbs, p.base_limbs, p.key.consts, n, n0inv, s_klo, lane, zr, sg);
    } else {
      M::load_ext(x0, s_bases + src * p.base_limbs, p.base_limbs, g);
#pragma unroll
      for (int j = 0; j < L; ++j) x1[j] = 0;
    }
#pragma unroll 1
    for (int k = 0; k < p.key.nops; ++k) {
      const uint32_t op = s_ops[k];
      const uint32_t yk = op & 0xffu, st = (op >> 8) & 0xffu, rl = (op >> 16) & 0xffu;
      const int nsq = (int)(op >> 24);
      if (yk != OP_NONE) {  // fetch the multiplier ahead of the squarings
        const uint32_t* ysrc = yk == OP_Y_CONST ? p.key.consts + S + g * L : (yk == OP_Y_PLAIN ? pair + g * L : tab + (size_t)yk * 2 * S);
        M::load(y0, ysrc);
        M::load(y1, ysrc + S);
      }
#pragma unroll 1
      for (int q = 0; q < nsq; ++q) TD::sqr(x0, x1, n, n0inv, s_klo, lane, zr, sg);
      if (yk != OP_NONE) TD::mul(x0, x1, y0, y1, n, n0inv, s_klo, lane, zr, sg);
      if (st != OP_NONE) {
        M::store(tab + (size_t)st * 2 * S, x0);
        M::store(tab + (size_t)st * 2 * S + S, x1);
      }
      if (rl != OP_NONE) {
        M::load(x0, tab + (size_t)rl * 2 * S);
        M::load(x1, tab + (size_t)rl * 2 * S + S);
      }
    }
    TD::assemble_store(p.out + (size_t)(job0 + grp) * p.out_limbs, p.out_limbs, valid, x0, x1, n, lane);
    __syncthreads();  // everyone is done with the staged inputs
    fence_proxy_async();
  }
}

// ------------------------------------------------------------------------ K2m
// out[j] = bases[j]^exps[j / exp_per] mod n^2 in the same two-digit Montgomery form, per-job exponent, fixed 5-bit
// window (the exponent is data, so every group of a warp scans the same number of windows).  Replaces
// BigInt::mod_pow(_, _, nn) / Paillier::mul at zero_enc_proof.rs:60,81; correct_ciphertext.rs:62,84;
// multiplication_proof.rs:91-94,133-138; verlin_proof.rs:89,109,147-155 (bases are ciphertexts: 2 S limbs wide).
constexpr int kMaxPhases = 32;
struct Jobs2mParams {
  Enc2mKey key;
  PowJobs jobs;
  uint32_t* table;    // phased: [total jobs][kTableVar][2 S]; else one [kTableVar][2 S] block per resident group
  uint32_t* acc;      // phased: the accumulator pair of every job between phases, [total jobs][2 S]
  unsigned* done;     // phased: phases completed per work unit, [units] (zeroed by the launcher)
  unsigned* cursor;   // next phase-unit (zeroed by the launcher on the same stream)
  const unsigned* jobs_dev;  // optional: the actual job count on the device (<= jobs.total; single-segment, unphased launches)
  int out_limbs;
  int tab_bases;      // window tables per job (the largest nbase of the launch)
  int win_per_phase;  // windows of the exponent scan per phase; >= the longest scan: one phase per unit, nothing migrates
  int nphase;         // phases of the longest unit
  unsigned pre[kMaxPhases + 1];  // pre[p] = phase-units before phase p (units are ordered longest scan first, so the
                                 // units that have a phase p are the first pre[p + 1] - pre[p])
  uint32_t zero;
};

struct Var2mParams {
  Enc2mKey key;
  const uint32_t* bases;
  const uint32_t* exps;
  uint32_t* out;
  uint32_t* table;
  int base_limbs, exp_limbs, exp_bits, exp_per, out_limbs, jobs;
  uint32_t zero;
};

__device__ __forceinline__ uint32_t exp_window2m(const uint32_t* e, int exp_limbs, int bit) {
  int limb = bit >> 5, sh = bit & 31;
  if (limb >= exp_limbs) return 0u;  // K2h pads a warp's shorter exponents with leading zero windows
  uint64_t v = e[limb];
  if (limb + 1 < exp_limbs) v |= (uint64_t)e[limb + 1] << 32;
  return (uint32_t)(v >> sh) & ((1u << kWindowVar) - 1u);
}

template <int T, int L, int MINB>
__global__ void __launch_bounds__(kCtaThreads, MINB) modexp2m_var_kernel(const Var2mParams p) {
  using M = Mp<T, L>;
  using TD = TwoDigit<T, L, 1>;
  constexpr int S = T * L;
  constexpr int G = kCtaThreads / T;
  __shared__ __align__(16) uint32_t s_klo[S];
  const int lane = threadIdx.x & 31;
  const int g = lane & (T - 1);
  const int grp = threadIdx.x / T;
  const uint32_t n0inv = p.key.n0inv;
  const uint32_t zr = p.zero;
  uint32_t n[L];
  M::load(n, p.key.mod + g * L);
  for (int i = threadIdx.x; i < S; i += kCtaThreads) s_klo[i] = p.key.consts[i];
  __syncthreads();
  uint32_t* tab = p.table + (size_t)(blockIdx.x * G + grp) * kTableVar * 2 * S + g * L;
  const int npass = (p.jobs + G - 1) / G;
  const int nwin = (p.exp_bits + kWindowVar - 1) / kWindowVar;
  for (int cj = blockIdx.x; cj < npass; cj += gridDim.x) {
    const int job = cj * G + grp;
    const bool valid = job < p.jobs;
    const int src = valid ? job : 0;
    const uint32_t* e = p.exps + (size_t)(src / p.exp_per) * p.exp_limbs;
    uint32_t x0[L], x1[L], y0[L], y1[L];
    TD::entry_pair(x0, x1, p.bases + (size_t)src * p.base_limbs, p.base_limbs, p.key.consts, n, n0inv, s_klo, lane, zr);
    M::load(y0, p.key.consts + S + g * L);  // pair(W^2): into Montgomery form
    M::load(y1, p.key.consts + 2 * S + g * L);
    TD::mul(x0, x1, y0, y1, n, n0inv, s_klo, lane, zr);
    M::store(tab + 2 * S, x0);
    M::store(tab + 2 * S + S, x1);
#pragma unroll
    for (int j = 0; j < L; ++j) {
      y0[j] = x0[j];
      y1[j] = x1[j];
    }
    M::load(x0, p.key.consts + 3 * S + g * L);  // pair(W) = 1 in Montgomery form = x^0
    M::load(x1, p.key.consts + 4 * S + g * L);
    M::store(tab, x0);
    M::store(tab + S, x1);
#pragma unroll
    for (int j = 0; j < L; ++j) {
      x0[j] = y0[j];
      x1[j] = y1[j];
    }
#pragma unroll 1
    for (int k = 2; k < kTableVar; ++k) {
      TD::mul(x0, x1, y0, y1, n, n0inv, s_klo, lane, zr);
      M::store(tab + (size_t)k * 2 * S, x0);
      M::store(tab + (size_t)k * 2 * S + S, x1);
    }
    {
      const uint32_t* t0 = tab + (size_t)exp_window2m(e, p.exp_limbs, (nwin - 1) * kWindowVar) * 2 * S;
      M::load(x0, t0);
      M::load(x1, t0 + S);
    }
    // the last pass of the loop multiplies by the plain pair (1, 0): it takes the result out of Montgomery form
#pragma unroll 1
    for (int w = nwin - 2; w >= -1; --w) {
      if (w >= 0) {
        const uint32_t* t0 = tab + (size_t)exp_window2m(e, p.exp_limbs, w * kWindowVar) * 2 * S;
        M::load(y0, t0);
        M::load(y1, t0 + S);
#pragma unroll 1
        for (int q = 0; q < kWindowVar; ++q) TD::sqr(x0, x1, n, n0inv, s_klo, lane, zr);
      } else {
        M::set_small(y0, 1u, g);
#pragma unroll
        for (int j = 0; j < L; ++j) y1[j] = 0;
      }
      TD::mul(x0, x1, y0, y1, n, n0inv, s_klo, lane, zr);
    }
    TD::assemble_store(p.out + (size_t)job * p.out_limbs, p.out_limbs, valid, x0, x1, n, lane);
  }
}

// ------------------------------------------------------------------------ K2h
// One launch for ALL the modular exponentiations of a batch of sigma-protocol proofs: a heterogeneous job list of up
// to kMaxPowSegs homogeneous segments (PowSeg: its own base rows, exponent rows / shared exponent, exponent length,
// optional Paillier plaintext and output rows), longest exponents first.  Replaces the five separate launches behind
// MulProof::verify (multiplication_proof.rs:118-138: Enc(f, z1), Enc(0, z2), e_a^e, e_c^e, e_b^f) and the four behind
// VerlinProof::verify (verlin_proof.rs:109,147-163), which at a 512-proof batch were 64 CTAs each.  Same arithmetic as
// K2m: fixed 5-bit window, two-digit Montgomery form; an Enc job is the modexp with exponent n whose final multiplier
// is the plain pair (1, m) instead of (1, 0).
//
// Scheduling.  A unit is one warp's worth of jobs (32 / T).  A batch of 512 proofs at 4096-bit n is ~5 whole modexps per
// SM sub-partition: handing out whole modexps leaves most sub-partitions idle while the unlucky ones finish their sixth
// (measured: 59-66 % of the multiplier pipe, profiles/r02_fill_curve.json).  So the exponent scan of a unit is cut into
// PHASES of win_per_phase windows, and the persistent warps take phase-units from one device-side cursor, phase-major:
// every unit's phase 0 (entry, window table, first windows), then every phase 1, ...  Between phases the state of a job
// lives in global memory - its window table (indexed by job, written once in phase 0) and its accumulator pair - so
// whichever warp is free next continues it.  done[unit] counts finished phases; the cursor hands phases out in
// dependency order, so a warp that has to wait for the previous phase of its unit waits on a warp that is already
// running (no deadlock), and with thousands of phase-units in flight it practically never waits.
template <int T, int L, int MINB, int U, int MODE>
__global__ void __launch_bounds__(kCtaThreads, MINB) modexp2m_jobs_kernel(const __grid_constant__ Jobs2mParams p) {
  using M = Mp<T, L>;
  using TD = TwoDigit<T, L, U, MODE>;  // U: lane-owner iterations of a row loop unrolled together; MODE 4: pair rows
  constexpr int S = T * L;
  constexpr int G = kCtaThreads / T;
  constexpr int GW = 32 / T;  // jobs per warp
  __shared__ __align__(16) uint32_t s_klo[S];
  const int lane = threadIdx.x & 31;
  const int g = lane & (T - 1);
  const int grp = threadIdx.x / T;
  const uint32_t n0inv = p.key.n0inv, n0hi = p.key.n0inv_hi;
  const uint32_t zr = p.zero;
  uint32_t n[L];
  M::load(n, p.key.mod + g * L);
  for (int i = threadIdx.x; i < S; i += kCtaThreads) s_klo[i] = p.key.consts[i];
  __syncthreads();
  const bool phased = p.nphase > 1;
  const int total = p.jobs_dev ? min((int)*p.jobs_dev, p.jobs.total) : p.jobs.total;
  const unsigned nwork = p.pre[p.nphase];
  for (;;) {
    unsigned work = 0;
    if (lane == 0) work = atomicAdd(p.cursor, 1u);
    work = __shfl_sync(ZKP_FULL, work, 0);
    if (work >= nwork) break;
    int ph = 0;
    while (work >= p.pre[ph + 1]) ++ph;
    const unsigned unit = work - p.pre[ph];
    if ((int)unit * GW >= total) continue;  // (a device-side job count below the host's upper bound)
    const int job = (int)unit * GW + (lane / T);
    const bool valid = job < total;
    const int src = valid ? job : total - 1;
    int si = 0;
#pragma unroll
    for (int k = 1; k < kMaxPowSegs; ++k)
      if (k < p.jobs.nseg && src >= p.jobs.seg[k].first) si = k;
    const PowSeg& sgm = p.jobs.seg[si];
    const int rel = src - sgm.first;
    int nwin = (sgm.exp_bits + kWindowVar - 1) / kWindowVar, nb = sgm.nbase;
#pragma unroll
    for (int o = T; o < 32; o <<= 1) {  // one trip count and one number of bases per warp
      nwin = max(nwin, __shfl_xor_sync(ZKP_FULL, nwin, o));
      nb = max(nb, __shfl_xor_sync(ZKP_FULL, nb, o));
    }
    // exponent rows of this job; a base the job does not have scans as zero (its windows pick table entry 0 = 1)
    const uint32_t* e[kMaxPowBases];
    int el[kMaxPowBases];
#pragma unroll
    for (int k = 0; k < kMaxPowBases; ++k) {
      const bool has = k < sgm.nbase;
      e[k] = has ? sgm.exp[k] + (size_t)rel * sgm.exp_stride[k] : nullptr;
      el[k] = has ? sgm.exp_limbs[k] : 0;
    }
    const size_t tab_base_stride = (size_t)kTableVar * 2 * S;
    // windows hi - 1 .. lo of the scan (most significant first) belong to this phase
    const int hi = nwin - ph * p.win_per_phase;
    const int lo = max(hi - p.win_per_phase, 0);
    uint32_t* tab = (phased ? p.table + (size_t)src * p.tab_bases * tab_base_stride : p.table + (size_t)(blockIdx.x * G + grp) * p.tab_bases * tab_base_stride) + g * L;
    uint32_t* accp = p.acc + (size_t)src * 2 * S + g * L;

    uint32_t x0[L], x1[L], y0[L], y1[L];
    if (ph == 0) {
#pragma unroll 1
      for (int k = 0; k < nb; ++k) {  // the window table of every base: x^0 .. x^31 in Montgomery form
        const int kk = k < sgm.nbase ? k : 0;  // (a base this job lacks: uniform work, the table is never read past entry 0)
        const int base_limbs = sgm.base_limbs[kk];
        const bool wide_base = __any_sync(ZKP_FULL, base_limbs > S);
        uint32_t* tk = tab + (size_t)k * tab_base_stride;
        TD::entry_pair_u(x0, x1, sgm.base[kk] + (size_t)rel * base_limbs, base_limbs, wide_base, p.key.consts, n, n0inv, s_klo, lane, zr, nullptr, n0hi);
        M::load(y0, p.key.consts + S + g * L);  // pair(W^2): into Montgomery form
        M::load(y1, p.key.consts + 2 * S + g * L);
        TD::mul(x0, x1, y0, y1, n, n0inv, s_klo, lane, zr, nullptr, n0hi);
        M::store(tk + 2 * S, x0);
        M::store(tk + 2 * S + S, x1);
#pragma unroll
        for (int j = 0; j < L; ++j) {
          y0[j] = x0[j];
          y1[j] = x1[j];
        }
        M::load(x0, p.key.consts + 3 * S + g * L);  // pair(W) = 1 in Montgomery form = x^0
        M::load(x1, p.key.consts + 4 * S + g * L);
        M::store(tk, x0);
        M::store(tk + S, x1);
#pragma unroll
        for (int j = 0; j < L; ++j) {
          x0[j] = y0[j];
          x1[j] = y1[j];
        }
#pragma unroll 1
        for (int t = 2; t < kTableVar; ++t) {
          TD::mul(x0, x1, y0, y1, n, n0inv, s_klo, lane, zr, nullptr, n0hi);
          M::store(tk + (size_t)t * 2 * S, x0);
          M::store(tk + (size_t)t * 2 * S + S, x1);
        }
      }
      // the top window: the accumulator starts at the product of the bases' table entries
      {
        const uint32_t* t0 = tab + (size_t)exp_window2m(e[0], el[0], (hi - 1) * kWindowVar) * 2 * S;
        M::load(x0, t0);
        M::load(x1, t0 + S);
      }
#pragma unroll 1
      for (int k = 1; k < nb; ++k) {
        const uint32_t* t0 = tab + (size_t)k * tab_base_stride + (size_t)exp_window2m(k == 1 ? e[1] : e[2], k == 1 ? el[1] : el[2], (hi - 1) * kWindowVar) * 2 * S;
        M::load(y0, t0);
        M::load(y1, t0 + S);
        TD::mul(x0, x1, y0, y1, n, n0inv, s_klo, lane, zr, nullptr, n0hi);
      }
    } else {
      if (lane == 0)
        while (atomicAdd(p.done + unit, 0u) < (unsigned)ph) __nanosleep(256);
      __syncwarp();
      __threadfence();
      M::load_cg(x0, accp);  // rewritten every phase, possibly by another SM: not through L1
      M::load_cg(x1, accp + S);
    }
#pragma unroll 1
    for (int w = hi - 1 - (ph == 0 ? 1 : 0); w >= lo; --w) {  // one squaring chain for all bases (Straus)
#pragma unroll 1
      for (int q = 0; q < kWindowVar; ++q) TD::sqr(x0, x1, n, n0inv, s_klo, lane, zr, nullptr, n0hi);
#pragma unroll 1
      for (int k = 0; k < nb; ++k) {
        const uint32_t* ek = k == 0 ? e[0] : (k == 1 ? e[1] : e[2]);
        const int elk = k == 0 ? el[0] : (k == 1 ? el[1] : el[2]);
        const uint32_t* t0 = tab + (size_t)k * tab_base_stride + (size_t)exp_window2m(ek, elk, w * kWindowVar) * 2 * S;
        M::load(y0, t0);
        M::load(y1, t0 + S);
        TD::mul(x0, x1, y0, y1, n, n0inv, s_klo, lane, zr, nullptr, n0hi);
      }
    }
    if (lo > 0) {  // hand the job over to whoever takes its next phase
      M::store(accp, x0);
      M::store(accp + S, x1);
      __threadfence();
      __syncwarp();
      if (lane == 0) atomicExch(p.done + unit, (unsigned)ph + 1u);
      continue;
    }
    // the final multiplier (1, m): the Paillier factor 1 + m n (m = 0: it only takes the result out of Montgomery form)
    const uint32_t* mrow = sgm.plain ? sgm.plain + (size_t)rel * sgm.plain_limbs : nullptr;
    const int pl = mrow ? sgm.plain_limbs : 0;
    if (__any_sync(ZKP_FULL, pl > S)) {  // m wider than n: m mod n = Mlo W / W + Mhi W^2 / W (see enc2m_kernel)
      M::load_ext(y1, mrow, pl < S ? pl : S, g);
      M::load(y0, p.key.consts + 3 * S + g * L);
      M::mont_mul(y1, y1, y0, n, n0inv, lane);
      uint32_t hi1[L];
      M::load_ext(hi1, mrow ? mrow + S : nullptr, pl - S, g);
      M::load(y0, p.key.consts + S + g * L);
      M::mont_mul(hi1, hi1, y0, n, n0inv, lane);
      M::add_mod(y1, hi1, n, lane);
    } else {
      M::load_ext(y1, mrow, pl, g);
    }
    M::set_small(y0, 1u, g);
    TD::mul(x0, x1, y0, y1, n, n0inv, s_klo, lane, zr, nullptr, n0hi);
    TD::assemble_store(sgm.out + (size_t)rel * p.out_limbs, p.out_limbs, valid, x0, x1, n, lane);
  }
}

// ------------------------------------------------------------------ host side
namespace {
using Limbs = std::vector<uint32_t>;
int cmp(const Limbs& a, const Limbs& b) {
  for (int i = (int)a.size() - 1; i >= 0; --i)
    if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
  return 0;
}
void sub_in(Limbs& a, const Limbs& b) {
  uint64_t bo = 0;
  for (size_t i = 0; i < a.size(); ++i) {
    uint64_t t = (uint64_t)a[i] - b[i] - bo;
    a[i] = (uint32_t)t;
    bo = (t >> 32) & 1u;
  }
}
// x = (2 x + cin) mod n for x < n; returns 1 iff n was subtracted
uint32_t dbl_mod(Limbs& x, uint32_t cin, const Limbs& n) {
  uint32_t c = cin;
  for (size_t i = 0; i < x.size(); ++i) {
    uint32_t v = x[i];
    x[i] = (v << 1) | c;
    c = v >> 31;
  }
  if (c || cmp(x, n) >= 0) {
    sub_in(x, n);
    return 1;
  }
  return 0;
}
}  // namespace

static bool pick_shape(int S, int& T, int& L) {
  switch (S) {
    case 32: T = 4; L = 8; return true;
    case 64: T = 8; L = 8; return true;
    case 96: T = 8; L = 12; return true;
    case 128: T = 16; L = 8; return true;
    default: return false;
  }
}

bool enc2m_supported(const uint32_t* n_host, int S) {
  int T, L;
  if (!pick_shape(S, T, L)) return false;
  // finish_x<2> needs n <= W - 4; n = 1 is meaningless
  bool top_ones = true;
  for (int i = 1; i < S; ++i) top_ones = top_ones && n_host[i] == 0xffffffffu;
  if (top_ones && n_host[0] >= 0xfffffffcu) return false;
  bool small = n_host[0] <= 1u;
  for (int i = 1; i < S; ++i) small = small && n_host[i] == 0u;
  return !small && (n_host[0] & 1u);
}

void enc2m_host_constants(const uint32_t* n_host, int S, uint32_t* consts) {
  Limbs n(n_host, n_host + S), x(S, 0u), x0(S, 0u), x1(S, 0u);
  x[0] = 1;  // W mod n by 32 S doublings
  for (int i = 0; i < 32 * S; ++i) dbl_mod(x, 0, n);
  Limbs klo = n;  // -W mod n (W mod n != 0: n is odd and > 1)
  sub_in(klo, x);
  x0[0] = 1;  // W and W^2 mod n^2 as digit pairs by pair doublings: (X0, X1) -> (2 X0 - c n, 2 X1 + c mod n)
  for (int i = 0; i < 64 * S; ++i) {
    if (i == 32 * S) {
      for (int k = 0; k < S; ++k) {
        consts[3 * S + k] = x0[k];
        consts[4 * S + k] = x1[k];
      }
    }
    uint32_t c = dbl_mod(x0, 0, n);
    dbl_mod(x1, c, n);
  }
  for (int i = 0; i < S; ++i) {
    consts[i] = klo[i];
    consts[S + i] = x0[i];
    consts[2 * S + i] = x1[i];
  }
}

// Tuning variant of K1m, fixed per process.  ZKP_B200_K1M_VARIANT picks the lane layout of a 2048-bit n
// (default 0 = Mp<8,8>, 4 CTAs per SM), ZKP_B200_K1M_WINDOW the sliding-window width (5 or 6; default kWindow2m).
// Measured on B200 (profiles/r01_k1m_variants2.jsonl): the <4,16> layouts are 1-12 % slower than <8,8> (3 resident CTAs,
// or spills at 4; larger loops against the 32 KB instruction cache), the shared-memory rows are within 0.4 %, and a
// 6-bit window is 1.6 % faster than a 5-bit one (32 odd powers: 17 KB of table per encryption in flight).
//   variant  layout  CTAs/SM  MODE (TwoDigit)
//   0        <8,8>   4        0   separate row loops, multiplier by SHFL
//   1        <4,16>  3        0
//   2        <4,16>  4        0   (128 registers: a few spills)
//   3        <4,16>  3        1   one row loop per squaring
//   4        <8,8>   4        1
//   5        <4,16>  3        2   multiplier / quotient rows in shared memory, row loop 4 steps deep
//   6        <4,16>  3        2   ... 8 steps deep
//   7        <8,8>   4        2   ... 8 steps deep
//   8        <4,16>  4        2   ... 4 steps deep
//   9        <8,8>   4        3   symmetric squaring of the first digit (each pair of lane blocks once) + reduction-only rows
// The other key sizes keep their layout and take only the MODE of the variant.
// The variants other than 0 and the environment knobs exist only in the lab build (make lab: -DZKP_B200_LAB ->
// libzkp_b200_lab.so); the product library holds the dispatched instantiations only and reads no environment.
struct Enc2mConfig {
  int variant, window;
};
static Enc2mConfig enc2m_config() {
#ifdef ZKP_B200_LAB
  static Enc2mConfig cfg = [] {
    Enc2mConfig c{0, kWindow2m};
    if (const char* v = getenv("ZKP_B200_K1M_VARIANT")) c.variant = atoi(v);
    if (const char* w = getenv("ZKP_B200_K1M_WINDOW")) c.window = atoi(w);
    if (c.variant < 0 || c.variant > 9) c.variant = 0;
    if (c.window != 5 && c.window != 6) c.window = kWindow2m;
    return c;
  }();
  return cfg;
#else
  return Enc2mConfig{0, kWindow2m};
#endif
}
int enc2m_window() { return enc2m_config().window; }
double enc2m_sqr_products() {  // variant 9 multiplies every pair of lane blocks once: (T/2 + 1)/T of the first product, T = 8 or 16 lanes
  return enc2m_config().variant == 9 ? 3.0 + 5.0 / 8.0 : 4.0;
}
static int enc2m_table() { return 1 << (enc2m_config().window - 1); }  // odd powers kept
static int enc2m_slots() { return enc2m_table() + 1; }                 // ... and x^2

// Sliding-window schedule of the exponent n (api_core.cu: recode_exponent, same window) -> K1m op list.
// op = nsq << 24 | reload_slot << 16 | store_slot << 8 | multiplier (0xff: none; 0xfe: W^2 pair; 0xfd: (1, m))
std::vector<uint32_t> enc2m_ops(const uint32_t* sched, int nsteps) {
  auto mk = [](uint32_t nsq, uint32_t rl, uint32_t st, uint32_t y) { return (nsq << 24) | (rl << 16) | (st << 8) | y; };
  const uint32_t tbl = (uint32_t)enc2m_table();
  std::vector<uint32_t> ops;
  ops.push_back(mk(0, OP_NONE, 0, OP_Y_CONST));                           // x W  -> slot 0
  ops.push_back(mk(1, 0, tbl, OP_NONE));                                  // x^2  -> last slot; back to x
  for (uint32_t e = 1; e < tbl; ++e) ops.push_back(mk(0, OP_NONE, e, tbl));  // x^(2e+1)
  ops.push_back(mk(0, sched[0] & 0xffu, OP_NONE, OP_NONE));               // accumulator = first window
  for (int k = 1; k < nsteps; ++k) {
    uint32_t idx = sched[k] & 0xffu, nsq = sched[k] >> 8;
    while (nsq > 255u) {
      ops.push_back(mk(255u, OP_NONE, OP_NONE, OP_NONE));
      nsq -= 255u;
    }
    ops.push_back(mk(nsq, OP_NONE, OP_NONE, idx));
  }
  ops.push_back(mk(0, OP_NONE, OP_NONE, OP_Y_PLAIN));
  return ops;
}

// measured on B200 with <8,8>: 5 and 6 resident CTAs (96 / 80 registers) are 2-4 % slower than 4, unrolling the owner loop gains nothing
constexpr int kCtasPerSm2m = 4;

static bool pick_shape_v(int S, int& T, int& L, int& minb, int& mode) {
  if (!pick_shape(S, T, L)) return false;
  static const int kMinb[10] = {4, 3, 4, 3, 4, 3, 3, 4, 4, 4}, kMode[10] = {0, 0, 0, 1, 1, 2, 2, 2, 2, 3};
  static const bool kT4[10] = {false, true, true, true, false, true, true, false, true, false};
  const int v = enc2m_config().variant;
  minb = kCtasPerSm2m;
  mode = kMode[v];
  if (S == 64) {
    minb = kMinb[v];
    if (kT4[v]) { T = 4; L = 16; }
  }
  return true;
}

int enc2m_resident_groups(int S, int num_sms) {
  int T, L, minb, mode;
  if (!pick_shape_v(S, T, L, minb, mode)) return 0;
  return num_sms * minb * (kCtaThreads / T);
}
size_t enc2m_table_limbs(int S, int num_sms) { return (size_t)enc2m_resident_groups(S, num_sms) * enc2m_slots() * 2 * S; }

template <int T, int L, int U, int MINB, int MODE>
static cudaError_t launch_one(const Enc2mParams& p, int num_sms, cudaStream_t st) {
  constexpr int G = kCtaThreads / T;
  constexpr int S = T * L;
  size_t smem = 16 + (size_t)(p.ops_pad + S) * 4 + (size_t)G * (p.base_limbs + p.plain_limbs + 2 * S) * 4;
  if (MODE == 2) smem += (size_t)G * TwoDigit<T, L, U, MODE>::kSg * 4;
  int grid = num_sms * MINB;
  int npass = (p.jobs + G - 1) / G;
  if (grid > npass) grid = npass;
  const bool wide = p.base_limbs > S || p.plain_limbs > S;
  auto kern = wide ? enc2m_kernel<T, L, U, MINB, true, MODE> : enc2m_kernel<T, L, U, MINB, false, MODE>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, kCtaThreads, smem, st>>>(p);
  return cudaGetLastError();
}

template <int T, int L>
static cudaError_t launch_mode(const Enc2mParams& p, int mode, int num_sms, cudaStream_t st) {
#ifdef ZKP_B200_LAB
  switch (mode) {
    case 1: return launch_one<T, L, 1, kCtasPerSm2m, 1>(p, num_sms, st);
    case 2: return launch_one<T, L, 2, kCtasPerSm2m, 2>(p, num_sms, st);
    case 3: return launch_one<T, L, 1, kCtasPerSm2m, 3>(p, num_sms, st);
    default: break;
  }
#endif
  (void)mode;
  return launch_one<T, L, 1, kCtasPerSm2m, 0>(p, num_sms, st);
}

cudaError_t launch_enc2m(const Enc2mKey& key, const uint32_t* bases, int base_limbs, const uint32_t* plain, int plain_limbs,
                         uint32_t* out, int out_limbs, int jobs, uint32_t* table, int num_sms, cudaStream_t st,
                         const unsigned* jobs_dev) {
  if (jobs <= 0) return cudaSuccess;
  if (base_limbs % 4 || base_limbs > 2 * key.S || (plain && (plain_limbs % 4 || plain_limbs > 2 * key.S)) || out_limbs % 2 ||
      out_limbs > 2 * key.S || key.nops <= 0)
    return cudaErrorInvalidValue;
  Enc2mParams p;
  p.key = key;
  p.bases = bases;
  p.plain = plain;
  p.out = out;
  p.table = table;
  p.base_limbs = base_limbs;
  p.plain_limbs = plain ? plain_limbs : 0;
  p.out_limbs = out_limbs;
  p.jobs = jobs;
  p.jobs_dev = jobs_dev;
  p.ops_pad = (key.nops + 3) & ~3;
  p.slots = enc2m_slots();
  p.zero = 0u;
  int T, L, minb, mode;
  if (!pick_shape_v(key.S, T, L, minb, mode)) return cudaErrorInvalidValue;
  switch (key.S) {
    case 32: return launch_mode<4, 8>(p, mode, num_sms, st);
    case 96: return launch_mode<8, 12>(p, mode, num_sms, st);
    case 128: return launch_mode<16, 8>(p, mode, num_sms, st);
    case 64:
      switch (enc2m_config().variant) {
#ifdef ZKP_B200_LAB
        case 1: return launch_one<4, 16, 1, 3, 0>(p, num_sms, st);
        case 2: return launch_one<4, 16, 1, 4, 0>(p, num_sms, st);
        case 3: return launch_one<4, 16, 1, 3, 1>(p, num_sms, st);
        case 4: return launch_one<8, 8, 1, 4, 1>(p, num_sms, st);
        case 5: return launch_one<4, 16, 2, 3, 2>(p, num_sms, st);
        case 6: return launch_one<4, 16, 4, 3, 2>(p, num_sms, st);
        case 7: return launch_one<8, 8, 4, 4, 2>(p, num_sms, st);
        case 8: return launch_one<4, 16, 2, 4, 2>(p, num_sms, st);
        case 9: return launch_one<8, 8, 1, 4, 3>(p, num_sms, st);
#endif
        default: return launch_one<8, 8, 1, 4, 0>(p, num_sms, st);
      }
    default: return cudaErrorInvalidValue;
  }
}

template <int T, int L>
static cudaError_t launch_var_one(const Var2mParams& p, int num_sms, cudaStream_t st) {
  constexpr int G = kCtaThreads / T;
  int grid = num_sms * kCtasPerSm2m;
  int npass = (p.jobs + G - 1) / G;
  if (grid > npass) grid = npass;
  modexp2m_var_kernel<T, L, kCtasPerSm2m><<<grid, kCtaThreads, 0, st>>>(p);
  return cudaGetLastError();
}

size_t var2m_table_limbs(int S, int num_sms) { return (size_t)enc2m_resident_groups(S, num_sms) * kTableVar * 2 * S; }

cudaError_t launch_modexp2m_var(const Enc2mKey& key, const uint32_t* bases, int base_limbs, const uint32_t* exps, int exp_limbs,
                                int exp_bits, int exp_per, uint32_t* out, int out_limbs, int jobs, uint32_t* table, int num_sms,
                                cudaStream_t st) {
  if (jobs <= 0) return cudaSuccess;
  if (base_limbs % 2 || base_limbs <= 0 || base_limbs > 2 * key.S || out_limbs % 2 || out_limbs > 2 * key.S || exp_per <= 0 ||
      exp_bits <= 0 || exp_bits > 32 * exp_limbs)
    return cudaErrorInvalidValue;
  Var2mParams p;
  p.key = key;
  p.bases = bases;
  p.exps = exps;
  p.out = out;
  p.table = table;
  p.base_limbs = base_limbs;
  p.exp_limbs = exp_limbs;
  p.exp_bits = exp_bits;
  p.exp_per = exp_per;
  p.out_limbs = out_limbs;
  p.jobs = jobs;
  p.zero = 0u;
  switch (key.S) {
    case 32: return launch_var_one<4, 8>(p, num_sms, st);
    case 64: return launch_var_one<8, 8>(p, num_sms, st);
    case 96: return launch_var_one<8, 12>(p, num_sms, st);
    case 128: return launch_var_one<16, 8>(p, num_sms, st);
    default: return cudaErrorInvalidValue;
  }
}

// ---- K2h launcher -------------------------------------------------------------------------------------------------
// Narrow-lane layouts: the same integer over twice the lanes (S = 128: one job per warp).  A batch of 512 MulProofs at
// 4096-bit n is 1 536 long jobs: 768 warps in <16,8>, 1.3 per SM sub-partition, against 1 536 warps in <32,4>.
constexpr int kCtasPerSmNarrow = 5;  // the narrow-lane layouts hold half the limbs per lane: more resident warps
constexpr int kPhasedUnitsPerSmsp = 12;   // phased scheduling while the launch has fewer units than this per sub-partition ...
constexpr int kPhaseTargetPerSmsp = 24;   // ... cut so that every sub-partition sees about this many phase-units
constexpr int kMinWinPerPhase = 8;

static int scan_windows(int exp_bits) { return (exp_bits + kWindowVar - 1) / kWindowVar; }

// limbs of scratch one launch needs: the window tables (by job when phased, else by resident group), the accumulators
// of a phased launch and its done[] counters
size_t jobs2m_scratch_limbs(int S, int num_sms, int total_jobs, int max_bases) {
  int T, L;
  if (!pick_shape(S, T, L)) return 0;
  const size_t wide = (size_t)num_sms * kCtasPerSm2m * (kCtaThreads / T);
  const size_t narrow = (size_t)num_sms * kCtasPerSmNarrow * (kCtaThreads / (2 * T));
  if (max_bases < 1) max_bases = 1;
  const size_t resident = (wide > narrow ? wide : narrow) * max_bases * kTableVar * 2 * S;
  // a phased launch has fewer than kPhasedUnitsPerSmsp long units per sub-partition, 32 / T jobs each; twice that leaves
  // room for the short jobs of the same launch (a launch that still does not fit runs unphased)
  size_t jobs = (size_t)num_sms * 4 * kPhasedUnitsPerSmsp * (32 / T) * 2;
  if ((size_t)total_jobs < jobs) jobs = (size_t)total_jobs;
  const size_t phased = jobs * ((size_t)max_bases * kTableVar + 1) * 2 * S + jobs + 64;
  return resident > phased ? resident : phased;
}

template <int T, int L, int MINB, int U = 1, int MODE = 0>
static cudaError_t launch_jobs_one(Jobs2mParams& p, size_t table_limbs, int num_sms, cudaStream_t st) {
  constexpr int GW = 32 / T;
  constexpr int S = T * L;
  const int nunits = (p.jobs.total + GW - 1) / GW;
  // windows scanned by unit u = those of its first job (segments are listed longest first)
  auto unit_windows = [&](int u) {
    const int job = u * GW;
    int si = 0;
    for (int k = 1; k < p.jobs.nseg; ++k)
      if (job >= p.jobs.seg[k].first) si = k;
    return scan_windows(p.jobs.seg[si].exp_bits);
  };
  const int max_win = unit_windows(0);
  int long_units = 0;  // units within a factor 2 of the longest: they decide how full the machine is
  for (int k = 0; k < p.jobs.nseg; ++k)
    if (2 * scan_windows(p.jobs.seg[k].exp_bits) >= max_win) long_units = (p.jobs.seg[k].first + p.jobs.seg[k].jobs + GW - 1) / GW;
  const int smsp = num_sms * 4;
  int nphase = 1;
  const size_t phased_limbs = (size_t)nunits * GW * ((size_t)p.tab_bases * kTableVar + 1) * 2 * S + (size_t)nunits + 64;
  if ((size_t)num_sms * MINB * (kCtaThreads / T) * p.tab_bases * kTableVar * 2 * S > table_limbs) return cudaErrorInvalidValue;
  // phases balance a launch of a few units per sub-partition; with at most one unit each there is nothing to balance
  if (long_units > smsp && long_units < kPhasedUnitsPerSmsp * smsp && phased_limbs <= table_limbs && !p.jobs_dev) {
    nphase = (kPhaseTargetPerSmsp * smsp + long_units - 1) / long_units;
    if (nphase > kMaxPhases) nphase = kMaxPhases;
    if (nphase > max_win / kMinWinPerPhase) nphase = max_win / kMinWinPerPhase;
    if (nphase < 1) nphase = 1;
  }
  p.win_per_phase = (max_win + nphase - 1) / nphase;
  p.nphase = (max_win + p.win_per_phase - 1) / p.win_per_phase;
  // pre[]: the units that have a phase ph are the first cnt(ph) (binary search over the non-increasing unit_windows)
  p.pre[0] = 0;
  for (int ph = 0; ph < p.nphase; ++ph) {
    int lo = 0, hi = nunits;  // first unit with unit_windows <= ph * win_per_phase
    while (lo < hi) {
      const int mid = (lo + hi) / 2;
      if (unit_windows(mid) > ph * p.win_per_phase) lo = mid + 1;
      else hi = mid;
    }
    p.pre[ph + 1] = p.pre[ph] + (unsigned)lo;
  }
  for (int ph = p.nphase; ph < kMaxPhases; ++ph) p.pre[ph + 1] = p.pre[p.nphase];
  if (p.nphase > 1) {
    p.acc = p.table + (size_t)nunits * GW * p.tab_bases * kTableVar * 2 * S;
    p.done = reinterpret_cast<unsigned*>(p.acc + (size_t)nunits * GW * 2 * S);
    cudaError_t e = cudaMemsetAsync(p.done, 0, sizeof(unsigned) * (size_t)nunits, st);
    if (e != cudaSuccess) return e;
  } else {
    p.acc = p.table;  // never touched
    p.done = p.cursor;
  }
  int grid = num_sms * MINB;
  const int need = (nunits + kCtaThreads / 32 - 1) / (kCtaThreads / 32);
  if (grid > need) grid = need;
  cudaError_t e = cudaMemsetAsync(p.cursor, 0, sizeof(unsigned), st);
  if (e != cudaSuccess) return e;
  modexp2m_jobs_kernel<T, L, MINB, U, MODE><<<grid, kCtaThreads, 0, st>>>(p);
  return cudaGetLastError();
}

// Row form by layout, measured (profiles/r02_pair_rows.json): pair rows make ONE encryption 11 % faster in the latency layout
// (13.3 -> 11.8 ms at 2048 bits) and the narrow throughput layout 10 % SLOWER (118 -> 133 ms, ZeroProof::verify x 1 536 at 4096
// bits): the quotient chain is not what a lone warp waits for - it issues ~20 dependent instructions per row at ~4 cycles
// each whatever their order - and with several warps per sub-partition the extra carry adds of the pair form cost issue slots.
constexpr bool kPairRowsLatency = true, kPairRowsNarrow = false;

cudaError_t launch_modexp2m_jobs(const Enc2mKey& key, const PowJobs& jobs, int out_limbs, uint32_t* table, size_t table_limbs,
                                 unsigned* cursor, int num_sms, cudaStream_t st, int shape, const unsigned* jobs_dev, int rows) {
  if (jobs_dev && jobs.nseg != 1) return cudaErrorInvalidValue;
  if (jobs.total <= 0) return cudaSuccess;
  if (jobs.nseg <= 0 || jobs.nseg > kMaxPowSegs || out_limbs % 2 || out_limbs > 2 * key.S) return cudaErrorInvalidValue;
  int first = 0;
  long long long_jobs = 0;
  int max_bits = 0;
  for (int k = 0; k < jobs.nseg; ++k) max_bits = jobs.seg[k].exp_bits > max_bits ? jobs.seg[k].exp_bits : max_bits;
  int tab_bases = 1;
  for (int k = 0; k < jobs.nseg; ++k) {
    const PowSeg& s = jobs.seg[k];
    if (s.first != first || s.jobs <= 0 || s.nbase < 1 || s.nbase > kMaxPowBases || s.exp_bits <= 0 ||
        (s.plain && (s.plain_limbs % 2 || s.plain_limbs <= 0 || s.plain_limbs > 2 * key.S)) || (k > 0 && s.exp_bits > jobs.seg[k - 1].exp_bits))
      return cudaErrorInvalidValue;
    for (int b = 0; b < s.nbase; ++b)
      if (!s.base[b] || !s.exp[b] || s.base_limbs[b] % 2 || s.base_limbs[b] <= 0 || s.base_limbs[b] > 2 * key.S || s.exp_limbs[b] <= 0)
        return cudaErrorInvalidValue;
    tab_bases = s.nbase > tab_bases ? s.nbase : tab_bases;
    first += s.jobs;
    if (2 * s.exp_bits >= max_bits) long_jobs += s.jobs;
  }
  if (first != jobs.total) return cudaErrorInvalidValue;
  Jobs2mParams p;
  p.key = key;
  p.jobs = jobs;
  p.table = table;
  p.cursor = cursor;
  p.jobs_dev = jobs_dev;
  p.out_limbs = out_limbs;
  p.tab_bases = tab_bases;
  p.zero = 0u;
  int T, L;
  if (!pick_shape(key.S, T, L)) return cudaErrorInvalidValue;
  if (shape == 0) {
    // narrow lanes while the long jobs in the wide layout would be fewer than ~4 warps per sub-partition: only whole
    // units run side by side, and the multiplier pipe needs 3-4 resident warps to fill
    const long long warps_wide = (long_jobs * T + 31) / 32;
    shape = warps_wide < (long long)num_sms * 4 * 4 ? 2 : 1;
    // at most one job per sub-partition: nothing shares the pipe, the time is ONE modexp's latency - spread every job over a
    // whole warp (the latency of a single proof: RangeProofNi::prove alone is 256 encryptions)
    if (long_jobs <= (long long)num_sms * 4) shape = 3;
  }
  // rows: 1 = single rows (as K1m / K2m), 2 = pair rows (two quotient digits per step); 0 = the default of the layout
  bool pair = rows == 0 ? kPairRowsLatency : rows == 2;
  if (shape == 3) {
    switch (key.S) {
      case 32: return pair ? launch_jobs_one<16, 2, kCtasPerSmNarrow, 2, 4>(p, table_limbs, num_sms, st) : launch_jobs_one<16, 2, kCtasPerSmNarrow, 2>(p, table_limbs, num_sms, st);
      case 64: return pair ? launch_jobs_one<32, 2, kCtasPerSmNarrow, 2, 4>(p, table_limbs, num_sms, st) : launch_jobs_one<32, 2, kCtasPerSmNarrow, 2>(p, table_limbs, num_sms, st);
      default: shape = 2; break;  // 3072 / 4096-bit n: the narrow layouts are already one job per (half-)warp
    }
  } else {
    pair = rows == 0 ? kPairRowsNarrow : rows == 2;
  }
#ifdef ZKP_B200_LAB
  if (shape == 2 && key.S == 128 && !pair) {  // lab: row-loop unrolling of the narrow layout (ZKP_B200_K2H_UNROLL = 1 | 2 | 4)
    static const int u = [] { const char* e = getenv("ZKP_B200_K2H_UNROLL"); return e ? atoi(e) : 2; }();
    if (u == 1) return launch_jobs_one<32, 4, kCtasPerSmNarrow, 1>(p, table_limbs, num_sms, st);
    if (u == 4) return launch_jobs_one<32, 4, kCtasPerSmNarrow, 4>(p, table_limbs, num_sms, st);
  }
#endif
  if (shape == 2) {
    switch (key.S) {
      case 32: return pair ? launch_jobs_one<8, 4, kCtasPerSmNarrow, 2, 4>(p, table_limbs, num_sms, st) : launch_jobs_one<8, 4, kCtasPerSmNarrow, 2>(p, table_limbs, num_sms, st);
      case 64: return pair ? launch_jobs_one<16, 4, kCtasPerSmNarrow, 2, 4>(p, table_limbs, num_sms, st) : launch_jobs_one<16, 4, kCtasPerSmNarrow, 2>(p, table_limbs, num_sms, st);
      case 96: return pair ? launch_jobs_one<16, 6, kCtasPerSmNarrow, 2, 4>(p, table_limbs, num_sms, st) : launch_jobs_one<16, 6, kCtasPerSmNarrow, 2>(p, table_limbs, num_sms, st);
      case 128: return pair ? launch_jobs_one<32, 4, kCtasPerSmNarrow, 2, 4>(p, table_limbs, num_sms, st) : launch_jobs_one<32, 4, kCtasPerSmNarrow, 2>(p, table_limbs, num_sms, st);
      default: return cudaErrorInvalidValue;
    }
  }
  switch (key.S) {
    case 32: return launch_jobs_one<4, 8, kCtasPerSm2m>(p, table_limbs, num_sms, st);
    case 64: return launch_jobs_one<8, 8, kCtasPerSm2m>(p, table_limbs, num_sms, st);
    case 96: return launch_jobs_one<8, 12, kCtasPerSm2m>(p, table_limbs, num_sms, st);
    case 128: return launch_jobs_one<16, 8, kCtasPerSm2m>(p, table_limbs, num_sms, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace zkp
