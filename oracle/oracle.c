/*
 * CPU ORACLE (test infrastructure and the timed CPU baseline; NOT a product path).
 *
 * A C restatement of the reference's hot path on the reference's own big-integer backend: GMP
 * (curv-kzen 0.10 is built with feature rust-gmp-kzen, reference Cargo.toml:41-42, so every
 * BigInt::mod_pow below is literally mpz_powm).  The image ships libgmp.so.10 (6.3.0) but no gmp.h,
 * so the handful of mpz entry points used are declared by hand.  SHA-256 is OpenSSL's.
 *
 * Parallelism mirrors the reference: one task per index of the security-parameter loop
 * (rayon par_iter at range_proof.rs:161-187,223,271 and correct_key_ni.rs:91), here a pthread pool
 * over the flattened (proof, index) grid.
 *
 * Integers use the same flat little-endian uint32 limb layout as include/zkp_b200.h so outputs can
 * be compared byte for byte with the CUDA path.  Citations are file:line into /root/reference/src.
 *
 * PARITY: the reference cannot be built here (no Rust toolchain, un-vendored crates) and has no golden
 * vectors; modexp/Enc/SHA outputs are pinned by uniqueness of the canonical residue / FIPS vectors, the
 * transcript byte encoding (minimal big-endian, zero -> 0x00) is the RECALLED curv-kzen rule and is
 * "parity unpinned" (see oracle/zkp_oracle.py).
 */
#include <openssl/sha.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

/* ---- hand-declared GMP 6 ABI (x86-64: mp_limb_t = unsigned long) ---- */
typedef struct {
  int _mp_alloc;
  int _mp_size;
  unsigned long* _mp_d;
} __mpz_struct;
typedef __mpz_struct mpz_t[1];
extern void __gmpz_init(__mpz_struct*);
extern void __gmpz_clear(__mpz_struct*);
extern void __gmpz_import(__mpz_struct*, size_t, int, size_t, int, size_t, const void*);
extern void* __gmpz_export(void*, size_t*, int, size_t, int, size_t, const __mpz_struct*);
extern void __gmpz_powm(__mpz_struct*, const __mpz_struct*, const __mpz_struct*, const __mpz_struct*);
extern void __gmpz_mul(__mpz_struct*, const __mpz_struct*, const __mpz_struct*);
extern void __gmpz_mod(__mpz_struct*, const __mpz_struct*, const __mpz_struct*);
extern void __gmpz_add(__mpz_struct*, const __mpz_struct*, const __mpz_struct*);
extern void __gmpz_sub(__mpz_struct*, const __mpz_struct*, const __mpz_struct*);
extern void __gmpz_add_ui(__mpz_struct*, const __mpz_struct*, unsigned long);
extern void __gmpz_set_ui(__mpz_struct*, unsigned long);
extern void __gmpz_set(__mpz_struct*, const __mpz_struct*);
extern int __gmpz_cmp(const __mpz_struct*, const __mpz_struct*);
extern int __gmpz_cmp_ui(const __mpz_struct*, unsigned long);
extern size_t __gmpz_sizeinbase(const __mpz_struct*, int);
extern void __gmpz_gcd(__mpz_struct*, const __mpz_struct*, const __mpz_struct*);
extern unsigned long __gmpz_fdiv_q_ui(__mpz_struct*, const __mpz_struct*, unsigned long);
extern void __gmpz_mul_2exp(__mpz_struct*, const __mpz_struct*, unsigned long);
extern const char* const __gmp_version;
#define mpz_init __gmpz_init
#define mpz_clear __gmpz_clear
#define mpz_powm __gmpz_powm
#define mpz_mul __gmpz_mul
#define mpz_mod __gmpz_mod
#define mpz_add __gmpz_add
#define mpz_sub __gmpz_sub
#define mpz_add_ui __gmpz_add_ui
#define mpz_set_ui __gmpz_set_ui
#define mpz_set __gmpz_set
#define mpz_cmp __gmpz_cmp
#define mpz_cmp_ui __gmpz_cmp_ui
#define mpz_sizeinbase __gmpz_sizeinbase
#define mpz_gcd __gmpz_gcd
#define mpz_fdiv_q_ui __gmpz_fdiv_q_ui
#define mpz_mul_2exp __gmpz_mul_2exp

#define RP_OPEN 0
#define RP_MASK1 1
#define RP_MASK2 2
#define CK_M2 11

const char* orc_gmp_version(void) { return __gmp_version; }
int orc_hw_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

static void imp(mpz_t z, const uint32_t* limbs, int n) { __gmpz_import(z, (size_t)n, -1, 4, 0, 0, limbs); }
/* returns 0 if the value does not fit n limbs */
static int expo(uint32_t* limbs, int n, const mpz_t z) {
  memset(limbs, 0, (size_t)n * 4);
  if (z->_mp_size < 0) return 0;
  if (__gmpz_sizeinbase(z, 2) > (size_t)n * 32 && z->_mp_size != 0) return 0;
  size_t cnt = 0;
  __gmpz_export(limbs, &cnt, -1, 4, 0, 0, z);
  return 1;
}

/* curv-kzen BigInt::to_bytes (RECALLED): minimal big-endian magnitude, zero -> one 0x00 byte. */
static size_t to_bytes(uint8_t* buf, const mpz_t z) {
  size_t size = (__gmpz_sizeinbase(z, 2) + 7) / 8;
  memset(buf, 0, size);
  __gmpz_export(buf, NULL, 1, 1, 0, 0, z);
  return size;
}
static void sha_update_mpz(SHA256_CTX* h, const mpz_t z, uint8_t* scratch) {
  size_t n = to_bytes(scratch, z);
  SHA256_Update(h, scratch, n);
}

/* ---- thread pool over a flat index range ---- */
typedef void (*task_fn)(void* arg, long idx);
typedef struct {
  task_fn fn;
  void* arg;
  long total;
  atomic_long next;
} pool_t;
static void* pool_worker(void* p) {
  pool_t* pl = (pool_t*)p;
  for (;;) {
    long i = atomic_fetch_add(&pl->next, 1);
    if (i >= pl->total) break;
    pl->fn(pl->arg, i);
  }
  return NULL;
}
static void run_parallel(task_fn fn, void* arg, long total, int threads) {
  if (threads <= 0) threads = orc_hw_threads();
  if (threads > total) threads = (int)(total > 0 ? total : 1);
  pool_t pl;
  pl.fn = fn;
  pl.arg = arg;
  pl.total = total;
  atomic_init(&pl.next, 0);
  if (threads == 1) {
    pool_worker(&pl);
    return;
  }
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
  for (int t = 0; t < threads; ++t) pthread_create(&th[t], NULL, pool_worker, &pl);
  for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
  free(th);
}

/* Paillier::encrypt_with_chosen_randomness (kzen-paillier, RECALLED):
 * rn = r^n mod nn; gm = (m*n + 1) % nn; c = gm*rn % nn. */
static void enc(mpz_t c, const mpz_t n, const mpz_t nn, const mpz_t m, const mpz_t r, mpz_t t) {
  mpz_powm(c, r, n, nn);
  mpz_mul(t, m, n);
  mpz_add_ui(t, t, 1);
  mpz_mod(t, t, nn);
  mpz_mul(c, c, t);
  mpz_mod(c, c, nn);
}

/* ---------------- raw batches ---------------- */
typedef struct {
  const uint32_t *n, *m, *r;
  int n_limbs, m_limbs, r_limbs;
  uint32_t* out;
  mpz_t zn, znn;
} enc_job;
static void enc_task(void* a, long i) {
  enc_job* j = (enc_job*)a;
  mpz_t m, r, c, t;
  mpz_init(m); mpz_init(r); mpz_init(c); mpz_init(t);
  imp(m, j->m + (size_t)i * j->m_limbs, j->m_limbs);
  imp(r, j->r + (size_t)i * j->r_limbs, j->r_limbs);
  enc(c, j->zn, j->znn, m, r, t);
  expo(j->out + (size_t)i * 2 * j->n_limbs, 2 * j->n_limbs, c);
  mpz_clear(m); mpz_clear(r); mpz_clear(c); mpz_clear(t);
}
void orc_paillier_enc(const uint32_t* n, int n_limbs, const uint32_t* m, int m_limbs, const uint32_t* r, int r_limbs,
                      int batch, uint32_t* out, int threads) {
  enc_job j = {n, m, r, n_limbs, m_limbs, r_limbs, out};
  mpz_init(j.zn); mpz_init(j.znn);
  imp(j.zn, n, n_limbs);
  mpz_mul(j.znn, j.zn, j.zn);
  run_parallel(enc_task, &j, batch, threads);
  mpz_clear(j.zn); mpz_clear(j.znn);
}

typedef struct {
  const uint32_t *bases, *exps, *mods;
  int mod_limbs, exp_limbs, per;
  uint32_t* out;
} pow_job;
static void pow_task(void* a, long i) {
  pow_job* j = (pow_job*)a;
  mpz_t b, e, m;
  mpz_init(b); mpz_init(e); mpz_init(m);
  imp(b, j->bases + (size_t)i * j->mod_limbs, j->mod_limbs);
  imp(e, j->exps + (size_t)(i / j->per) * j->exp_limbs, j->exp_limbs);
  imp(m, j->mods + (size_t)(i / j->per) * j->mod_limbs, j->mod_limbs);
  mpz_powm(b, b, e, m);  /* BigInt::mod_pow (correct_key_ni.rs:92 and the sigma-protocol sites) */
  expo(j->out + (size_t)i * j->mod_limbs, j->mod_limbs, b);
  mpz_clear(b); mpz_clear(e); mpz_clear(m);
}
void orc_modexp(const uint32_t* bases, const uint32_t* exps, int exp_limbs, const uint32_t* mods, int mod_limbs, int per,
                int batch, uint32_t* out, int threads) {
  pow_job j = {bases, exps, mods, mod_limbs, exp_limbs, per, out};
  run_parallel(pow_task, &j, batch, threads);
}

/* compute_digest (utils.rs:9-22) over items[b][count][limbs] */
void orc_sha256_transcript(const uint32_t* items, int limbs, int count, int batch, uint8_t* digest) {
  uint8_t* scratch = (uint8_t*)malloc((size_t)limbs * 4 + 8);
  mpz_t z;
  mpz_init(z);
  for (int b = 0; b < batch; ++b) {
    SHA256_CTX h;
    SHA256_Init(&h);
    for (int k = 0; k < count; ++k) {
      imp(z, items + ((size_t)b * count + k) * limbs, limbs);
      sha_update_mpz(&h, z, scratch);
    }
    SHA256_Final(digest + (size_t)b * 32, &h);
  }
  mpz_clear(z);
  free(scratch);
}

/* ---------------- RangeProofNi ---------------- */
typedef struct {
  int batch, ef, w, nl;
  const uint32_t *range, *x, *r, *w1, *r1, *r2, *cx, *c1in, *c2in, *resp_w_in, *resp_r_in;
  const uint8_t *swap, *kind_in;
  uint32_t *c1, *c2, *resp_w, *resp_r;
  uint8_t *digest, *kind, *ok, *fault;
  mpz_t zn, znn;
  uint8_t* ebits; /* [batch][ef] challenge bits, 2 = index out of range */
  const uint8_t* chal; /* interactive proof: raw ChallengeBits bytes [batch][chal_bytes], or NULL */
  int chal_bytes;
} rp_job;

/* third = range.div_floor(3), two_thirds = 2*third (range_proof.rs:133-134) */
static void thirds(mpz_t third, mpz_t two, const uint32_t* range, int w) {
  imp(third, range, w);
  mpz_fdiv_q_ui(third, third, 3);
  mpz_mul_2exp(two, third, 1);
}

/* range_proof.rs:136-187 for one (proof, i): w2 = w1 - third, coin swap, two encryptions. */
static void rp_pairs_task(void* a, long t) {
  rp_job* j = (rp_job*)a;
  long b = t / j->ef;
  mpz_t third, two, w1, w2, r, c, tmp;
  mpz_init(third); mpz_init(two); mpz_init(w1); mpz_init(w2); mpz_init(r); mpz_init(c); mpz_init(tmp);
  thirds(third, two, j->range + (size_t)b * j->w, j->w);
  imp(w1, j->w1 + (size_t)t * j->w, j->w);
  mpz_sub(w2, w1, third);
  if (j->swap[t]) { mpz_set(tmp, w1); mpz_set(w1, w2); mpz_set(w2, tmp); }
  imp(r, j->r1 + (size_t)t * j->nl, j->nl);
  enc(c, j->zn, j->znn, w1, r, tmp);
  expo(j->c1 + (size_t)t * 2 * j->nl, 2 * j->nl, c);
  imp(r, j->r2 + (size_t)t * j->nl, j->nl);
  enc(c, j->zn, j->znn, w2, r, tmp);
  expo(j->c2 + (size_t)t * 2 * j->nl, 2 * j->nl, c);
  /* park w1', w2' in resp_w so the response pass can read them */
  expo(j->resp_w + (size_t)t * 2 * j->w, j->w, w1);
  expo(j->resp_w + ((size_t)t * 2 + 1) * j->w, j->w, w2);
  mpz_clear(third); mpz_clear(two); mpz_clear(w1); mpz_clear(w2); mpz_clear(r); mpz_clear(c); mpz_clear(tmp);
}

/* e = to_bytes(compute_digest([n] ++ c1 ++ c2)) (range_proof_ni.rs:58-61); bit i of BitVec::from_bytes(e). */
static void rp_challenge(rp_job* j, const uint32_t* c1, const uint32_t* c2) {
  int nnl = 2 * j->nl;
  uint8_t* scratch = (uint8_t*)malloc((size_t)nnl * 4 + 8);
  mpz_t z;
  mpz_init(z);
  for (int b = 0; b < j->batch; ++b) {
    SHA256_CTX h;
    SHA256_Init(&h);
    sha_update_mpz(&h, j->zn, scratch);
    for (int i = 0; i < j->ef; ++i) {
      imp(z, c1 + ((size_t)b * j->ef + i) * nnl, nnl);
      sha_update_mpz(&h, z, scratch);
    }
    for (int i = 0; i < j->ef; ++i) {
      imp(z, c2 + ((size_t)b * j->ef + i) * nnl, nnl);
      sha_update_mpz(&h, z, scratch);
    }
    uint8_t d[32];
    SHA256_Final(d, &h);
    if (j->digest) memcpy(j->digest + (size_t)b * 32, d, 32);
    int lead = 0; /* BigInt::from_bytes then to_bytes strips leading zero bytes; zero -> [0] */
    while (lead < 31 && d[lead] == 0) ++lead;
    int len = 32 - lead;
    for (int i = 0; i < j->ef; ++i) {
      int byte = i / 8;
      if (j->chal) /* BitVec::from_bytes(&e.0)[i] on the verifier's bytes as they are (range_proof.rs:221,267) */
        j->ebits[(size_t)b * j->ef + i] =
            byte < j->chal_bytes ? (uint8_t)((j->chal[(size_t)b * j->chal_bytes + byte] >> (7 - i % 8)) & 1) : 2;
      else
        j->ebits[(size_t)b * j->ef + i] = byte < len ? (uint8_t)((d[lead + byte] >> (7 - i % 8)) & 1) : 2;
    }
  }
  mpz_clear(z);
  free(scratch);
}

/* range_proof.rs:210-252 for one (proof, i) */
static void rp_response_task(void* a, long t) {
  rp_job* j = (rp_job*)a;
  long b = t / j->ef;
  uint32_t* rw = j->resp_w + (size_t)t * 2 * j->w;
  uint32_t* rr = j->resp_r + (size_t)t * 2 * j->nl;
  uint8_t e = j->ebits[t];
  if (e == 2) { j->fault[b] = 1; e = 0; }
  if (!e) {
    j->kind[t] = RP_OPEN;
    memcpy(rr, j->r1 + (size_t)t * j->nl, (size_t)j->nl * 4);
    memcpy(rr + j->nl, j->r2 + (size_t)t * j->nl, (size_t)j->nl * 4);
    return; /* resp_w already holds (w1', w2') */
  }
  mpz_t third, two, x, w1, w2, s, r, ri;
  mpz_init(third); mpz_init(two); mpz_init(x); mpz_init(w1); mpz_init(w2); mpz_init(s); mpz_init(r); mpz_init(ri);
  thirds(third, two, j->range + (size_t)b * j->w, j->w);
  imp(x, j->x + (size_t)b * j->w, j->w);
  imp(w1, rw, j->w);
  imp(w2, rw + j->w, j->w);
  imp(r, j->r + (size_t)b * j->nl, j->nl);
  mpz_add(s, x, w1);
  int first = mpz_cmp(s, third) > 0 && mpz_cmp(s, two) < 0;
  if (first) {
    j->kind[t] = RP_MASK1;
    imp(ri, j->r1 + (size_t)t * j->nl, j->nl);
  } else {
    j->kind[t] = RP_MASK2;
    mpz_add(s, x, w2);
    imp(ri, j->r2 + (size_t)t * j->nl, j->nl);
  }
  mpz_mul(ri, r, ri);
  mpz_mod(ri, ri, j->zn);
  if (!expo(rw, j->w, s)) j->fault[b] = 1;
  memset(rw + j->w, 0, (size_t)j->w * 4);
  expo(rr, j->nl, ri);
  memset(rr + j->nl, 0, (size_t)j->nl * 4);
  mpz_clear(third); mpz_clear(two); mpz_clear(x); mpz_clear(w1); mpz_clear(w2); mpz_clear(s); mpz_clear(r); mpz_clear(ri);
}

/* RangeProofNi::prove (range_proof_ni.rs:47-82) for a batch under one key; layouts as zkp_rangeproof_ni_prove. */
void orc_rangeproof_prove(const uint32_t* n, int n_limbs, int batch, int ef, int w_limbs, const uint32_t* range,
                          const uint32_t* x, const uint32_t* r, const uint32_t* w1, const uint8_t* swap, const uint32_t* r1,
                          const uint32_t* r2, const uint8_t* challenge, int chal_bytes, uint32_t* c1, uint32_t* c2,
                          uint8_t* digest, uint8_t* kind, uint32_t* resp_w, uint32_t* resp_r, uint8_t* fault, int threads);
void orc_rangeproof_ni_prove(const uint32_t* n, int n_limbs, int batch, int ef, int w_limbs, const uint32_t* range,
                             const uint32_t* x, const uint32_t* r, const uint32_t* w1, const uint8_t* swap,
                             const uint32_t* r1, const uint32_t* r2, uint32_t* c1, uint32_t* c2, uint8_t* digest,
                             uint8_t* kind, uint32_t* resp_w, uint32_t* resp_r, uint8_t* fault, int threads) {
  orc_rangeproof_prove(n, n_limbs, batch, ef, w_limbs, range, x, r, w1, swap, r1, r2, NULL, 0, c1, c2, digest, kind, resp_w,
                       resp_r, fault, threads);
}
/* challenge != NULL: the interactive RangeProof (generate_encrypted_pairs + generate_proof with the verifier's e). */
void orc_rangeproof_prove(const uint32_t* n, int n_limbs, int batch, int ef, int w_limbs, const uint32_t* range,
                          const uint32_t* x, const uint32_t* r, const uint32_t* w1, const uint8_t* swap, const uint32_t* r1,
                          const uint32_t* r2, const uint8_t* challenge, int chal_bytes, uint32_t* c1, uint32_t* c2,
                          uint8_t* digest, uint8_t* kind, uint32_t* resp_w, uint32_t* resp_r, uint8_t* fault, int threads) {
  rp_job j;
  memset(&j, 0, sizeof j);
  j.chal = challenge; j.chal_bytes = chal_bytes;
  j.batch = batch; j.ef = ef; j.w = w_limbs; j.nl = n_limbs;
  j.range = range; j.x = x; j.r = r; j.w1 = w1; j.swap = swap; j.r1 = r1; j.r2 = r2;
  j.c1 = c1; j.c2 = c2; j.digest = digest; j.kind = kind; j.resp_w = resp_w; j.resp_r = resp_r;
  uint8_t* own_fault = NULL;
  if (!fault) fault = own_fault = (uint8_t*)malloc((size_t)batch);
  memset(fault, 0, (size_t)batch);
  j.fault = fault;
  j.ebits = (uint8_t*)malloc((size_t)batch * ef);
  mpz_init(j.zn); mpz_init(j.znn);
  imp(j.zn, n, n_limbs);
  mpz_mul(j.znn, j.zn, j.zn);
  run_parallel(rp_pairs_task, &j, (long)batch * ef, threads);
  rp_challenge(&j, c1, c2);
  run_parallel(rp_response_task, &j, (long)batch * ef, threads);
  mpz_clear(j.zn); mpz_clear(j.znn);
  free(j.ebits);
  free(own_fault);
}

/* range_proof.rs:270-348 for one (proof, i) */
static void rp_verify_task(void* a, long t) {
  rp_job* j = (rp_job*)a;
  long b = t / j->ef;
  int nnl = 2 * j->nl;
  const uint32_t* rw = j->resp_w_in + (size_t)t * 2 * j->w;
  const uint32_t* rr = j->resp_r_in + (size_t)t * 2 * j->nl;
  uint8_t e = j->ebits[t];
  uint8_t k = j->kind_in[t];
  if (e == 2 || k > RP_MASK2) { j->fault[b] = 1; j->ok[t] = 0; return; }
  if ((e == 0) != (k == RP_OPEN)) { j->ok[t] = 0; return; } /* `_ => false` :345 */
  mpz_t third, two, m, r, c, tmp, want;
  mpz_init(third); mpz_init(two); mpz_init(m); mpz_init(r); mpz_init(c); mpz_init(tmp); mpz_init(want);
  thirds(third, two, j->range + (size_t)b * j->w, j->w);
  int res = 1;
  if (k == RP_OPEN) {
    mpz_t w2;
    mpz_init(w2);
    imp(m, rw, j->w);
    imp(r, rr, j->nl);
    enc(c, j->zn, j->znn, m, r, tmp);
    imp(want, j->c1in + (size_t)t * nnl, nnl);
    if (mpz_cmp(c, want) != 0) res = 0;
    imp(w2, rw + j->w, j->w);
    imp(r, rr + j->nl, j->nl);
    enc(c, j->zn, j->znn, w2, r, tmp);
    imp(want, j->c2in + (size_t)t * nnl, nnl);
    if (mpz_cmp(c, want) != 0) res = 0;
    int f1 = mpz_cmp(w2, third) < 0 && mpz_cmp(m, third) > 0 && mpz_cmp(m, two) < 0;
    int f2 = mpz_cmp(m, third) < 0 && mpz_cmp(w2, third) > 0 && mpz_cmp(w2, two) < 0;
    if (!(f1 || f2)) res = 0;
    mpz_clear(w2);
  } else {
    imp(want, (k == RP_MASK1 ? j->c1in : j->c2in) + (size_t)t * nnl, nnl);
    imp(tmp, j->cx + (size_t)b * nnl, nnl);
    mpz_mul(want, want, tmp);
    mpz_mod(want, want, j->znn);
    imp(m, rw, j->w);
    imp(r, rr, j->nl);
    enc(c, j->zn, j->znn, m, r, tmp);
    if (mpz_cmp(c, want) != 0) res = 0;
    if (mpz_cmp(m, third) < 0 || mpz_cmp(m, two) > 0) res = 0;
  }
  j->ok[t] = (uint8_t)res;
  mpz_clear(third); mpz_clear(two); mpz_clear(m); mpz_clear(r); mpz_clear(c); mpz_clear(tmp); mpz_clear(want);
}

/* RangeProofNi::verify (range_proof_ni.rs:84-107); layouts as zkp_rangeproof_ni_verify.
 * Returns the number of Paillier encryptions performed. */
long long orc_rangeproof_verify(const uint32_t* n, int n_limbs, int batch, int ef, int w_limbs, const uint32_t* range,
                                const uint32_t* cipher_x, const uint32_t* c1, const uint32_t* c2, const uint8_t* kind,
                                const uint32_t* resp_w, const uint32_t* resp_r, const uint8_t* challenge, int chal_bytes,
                                uint8_t* accept, uint8_t* fault, uint8_t* digest, int threads);
long long orc_rangeproof_ni_verify(const uint32_t* n, int n_limbs, int batch, int ef, int w_limbs, const uint32_t* range,
                                   const uint32_t* cipher_x, const uint32_t* c1, const uint32_t* c2, const uint8_t* kind,
                                   const uint32_t* resp_w, const uint32_t* resp_r, uint8_t* accept, uint8_t* fault,
                                   uint8_t* digest, int threads) {
  return orc_rangeproof_verify(n, n_limbs, batch, ef, w_limbs, range, cipher_x, c1, c2, kind, resp_w, resp_r, NULL, 0, accept,
                               fault, digest, threads);
}
/* challenge != NULL: RangeProof::verifier_output with the verifier's own e (interactive proof). */
long long orc_rangeproof_verify(const uint32_t* n, int n_limbs, int batch, int ef, int w_limbs, const uint32_t* range,
                                const uint32_t* cipher_x, const uint32_t* c1, const uint32_t* c2, const uint8_t* kind,
                                const uint32_t* resp_w, const uint32_t* resp_r, const uint8_t* challenge, int chal_bytes,
                                uint8_t* accept, uint8_t* fault, uint8_t* digest, int threads) {
  rp_job j;
  memset(&j, 0, sizeof j);
  j.chal = challenge; j.chal_bytes = chal_bytes;
  j.batch = batch; j.ef = ef; j.w = w_limbs; j.nl = n_limbs;
  j.range = range; j.cx = cipher_x; j.c1in = c1; j.c2in = c2; j.kind_in = kind; j.resp_w_in = resp_w; j.resp_r_in = resp_r;
  j.digest = digest;
  uint8_t* own_fault = NULL;
  if (!fault) fault = own_fault = (uint8_t*)malloc((size_t)batch);
  memset(fault, 0, (size_t)batch);
  j.fault = fault;
  j.ebits = (uint8_t*)malloc((size_t)batch * ef);
  j.ok = (uint8_t*)malloc((size_t)batch * ef);
  mpz_init(j.zn); mpz_init(j.znn);
  imp(j.zn, n, n_limbs);
  mpz_mul(j.znn, j.zn, j.zn);
  rp_challenge(&j, c1, c2);
  run_parallel(rp_verify_task, &j, (long)batch * ef, threads);
  long long encs = 0;
  for (int b = 0; b < batch; ++b) {
    uint8_t all = 1;
    for (int i = 0; i < ef; ++i) {
      size_t t = (size_t)b * ef + i;
      all &= j.ok[t];
      if (j.ebits[t] != 2 && kind[t] <= RP_MASK2 && ((j.ebits[t] == 0) == (kind[t] == RP_OPEN))) encs += kind[t] == RP_OPEN ? 2 : 1;
    }
    accept[b] = (uint8_t)(all && !fault[b]);
  }
  mpz_clear(j.zn); mpz_clear(j.znn);
  free(j.ebits); free(j.ok); free(own_fault);
  return encs;
}

/* ---------------- NiCorrectKeyProof::verify (correct_key_ni.rs:73-117) ---------------- */
typedef struct {
  int batch, nl;
  const uint32_t *n, *sigma;
  uint32_t *rho_out, *derived;
  uint8_t* rho_ok;
} ck_job;

static void ck_pow_task(void* a, long t) {
  ck_job* j = (ck_job*)a;
  long b = t / CK_M2;
  mpz_t s, n;
  mpz_init(s); mpz_init(n);
  imp(s, j->sigma + (size_t)t * j->nl, j->nl);
  imp(n, j->n + (size_t)b * j->nl, j->nl);
  mpz_powm(s, s, n, n); /* :92 */
  expo(j->derived + (size_t)t * j->nl, j->nl, s);
  mpz_clear(s); mpz_clear(n);
}

static void digest_to_mpz(mpz_t z, SHA256_CTX* h) {
  uint8_t d[32];
  SHA256_Final(d, h);
  __gmpz_import(z, 32, 1, 1, 0, 0, d);
}

static const unsigned short* small_primes(int* count) {
  static unsigned short pr[1024];
  static int n = 0;
  if (!n) {
    for (int v = 2; v < 6370; ++v) {
      int is = 1;
      for (int d = 2; d * d <= v; ++d)
        if (v % d == 0) { is = 0; break; }
      if (is) pr[n++] = (unsigned short)v;
    }
  }
  *count = n;
  return pr;
}

void orc_correct_key_ni_verify(int batch, int n_limbs, const uint32_t* n, const uint32_t* sigma, const uint8_t* salt,
                               int salt_len, uint8_t* accept, uint32_t* rho_out, int threads) {
  ck_job j;
  j.batch = batch; j.nl = n_limbs; j.n = n; j.sigma = sigma; j.rho_out = rho_out;
  j.derived = (uint32_t*)malloc((size_t)batch * CK_M2 * n_limbs * 4);
  run_parallel(ck_pow_task, &j, (long)batch * CK_M2, threads);
  /* P = product of primes < 6370 (correct_key_ni.rs:26), rebuilt rather than parsed */
  mpz_t P, zn, salt_bn, seed, h, acc, idx, g, want;
  mpz_init(P); mpz_init(zn); mpz_init(salt_bn); mpz_init(seed); mpz_init(h); mpz_init(acc); mpz_init(idx); mpz_init(g); mpz_init(want);
  int np = 0;
  const unsigned short* pr = small_primes(&np);
  mpz_set_ui(P, 1);
  for (int k = 0; k < np; ++k) { mpz_set_ui(idx, pr[k]); mpz_mul(P, P, idx); }
  uint8_t* scratch = (uint8_t*)malloc((size_t)n_limbs * 4 + 64 + (size_t)salt_len);
  /* salt_bn = H(from_bytes(salt)) (:75) */
  {
    SHA256_CTX c;
    SHA256_Init(&c);
    __gmpz_import(salt_bn, (size_t)salt_len, 1, 1, 0, 0, salt);
    sha_update_mpz(&c, salt_bn, scratch);
    digest_to_mpz(salt_bn, &c);
  }
  for (int b = 0; b < batch; ++b) {
    imp(zn, n + (size_t)b * n_limbs, n_limbs);
    size_t key_length = mpz_cmp_ui(zn, 0) == 0 ? 0 : mpz_sizeinbase(zn, 2);
    int msklen = (int)(key_length / 256 + 1);
    int ok = 1;
    for (int i = 0; i < CK_M2; ++i) {
      SHA256_CTX c;
      SHA256_Init(&c);
      sha_update_mpz(&c, zn, scratch);
      sha_update_mpz(&c, salt_bn, scratch);
      mpz_set_ui(idx, (unsigned long)i);
      sha_update_mpz(&c, idx, scratch);
      digest_to_mpz(seed, &c); /* :78-83 */
      mpz_set_ui(acc, 0);
      for (int k = 0; k < msklen; ++k) { /* mask_generation :105-117 */
        SHA256_Init(&c);
        sha_update_mpz(&c, seed, scratch);
        mpz_set_ui(idx, (unsigned long)k);
        sha_update_mpz(&c, idx, scratch);
        digest_to_mpz(h, &c);
        mpz_mul_2exp(h, h, (unsigned long)k * 256);
        mpz_add(acc, acc, h);
      }
      mpz_mod(acc, acc, zn); /* :84 */
      if (rho_out) expo(rho_out + ((size_t)b * CK_M2 + i) * n_limbs, n_limbs, acc);
      imp(want, j.derived + ((size_t)b * CK_M2 + i) * n_limbs, n_limbs);
      if (mpz_cmp(acc, want) != 0) ok = 0;
    }
    mpz_gcd(g, P, zn); /* :87-88 */
    if (mpz_cmp_ui(g, 1) != 0) ok = 0;
    accept[b] = (uint8_t)ok;
  }
  free(scratch);
  free(j.derived);
  mpz_clear(P); mpz_clear(zn); mpz_clear(salt_bn); mpz_clear(seed); mpz_clear(h); mpz_clear(acc); mpz_clear(idx); mpz_clear(g); mpz_clear(want);
}

/* ======================================================================================================
 * Sigma-protocol verifiers on GMP (a second restatement next to oracle/zkp_oracle.py; CPU tests compare the two):
 *   MulProof::verify            multiplication_proof.rs:108-145
 *   VerlinProof::verify         verlin_proof.rs:101-134, gen_phi :138-165
 *   CompositeDLogProof::verify  wi_dlog_proof.rs:66-91
 *   CorrectMessageProof::verify correct_message.rs:126-161
 * verdict[b]: 1 = Ok(()), 0 = Err(IncorrectProof), 2 = the reference panics (unwrap() on a missing inverse, assert!).
 * ====================================================================================================== */
extern int __gmpz_invert(__mpz_struct*, const __mpz_struct*, const __mpz_struct*);
#define mpz_invert __gmpz_invert

/* e = compute_digest(items...) as a BigInt (utils.rs:9-22) */
static void digest_of(mpz_t e, const __mpz_struct* const* items, int count, uint8_t* scratch) {
  SHA256_CTX h;
  SHA256_Init(&h);
  for (int i = 0; i < count; ++i) sha_update_mpz(&h, items[i], scratch);
  uint8_t d[32];
  SHA256_Final(d, &h);
  __gmpz_import(e, 32, 1, 1, 0, 0, d);
}

typedef struct {
  int nl, zl, M, el, ml;
  const uint32_t *n, *a0, *a1, *a2, *a3, *a4, *a5, *a6, *a7;
  uint8_t* verdict;
} sig_job;

static void mul_verify_task(void* arg, long b) {
  sig_job* j = (sig_job*)arg;
  const int nl = j->nl, nnl = 2 * nl;
  mpz_t n, nn, e_a, e_b, e_c, f, z1, z2, e_d, e_db, e, t, u, v, lhs;
  __mpz_struct* all[] = {n, nn, e_a, e_b, e_c, f, z1, z2, e_d, e_db, e, t, u, v, lhs};
  for (size_t k = 0; k < sizeof(all) / sizeof(all[0]); ++k) mpz_init(all[k]);
  uint8_t* scratch = (uint8_t*)malloc((size_t)nnl * 4 + 64);
  imp(n, j->n, nl);
  mpz_mul(nn, n, n);
  imp(e_a, j->a0 + (size_t)b * nnl, nnl); imp(e_b, j->a1 + (size_t)b * nnl, nnl); imp(e_c, j->a2 + (size_t)b * nnl, nnl);
  imp(f, j->a3 + (size_t)b * nl, nl); imp(z1, j->a4 + (size_t)b * nnl, nnl); imp(z2, j->a5 + (size_t)b * nnl, nnl);
  imp(e_d, j->a6 + (size_t)b * nnl, nnl); imp(e_db, j->a7 + (size_t)b * nnl, nnl);
  const __mpz_struct* items[] = {n, e_a, e_b, e_c, e_d, e_db};
  digest_of(e, items, 6, scratch);                                       /* :109-116 */
  uint8_t verdict = 1;
  enc(u, n, nn, f, z1, t);                                               /* enc_f_z1  :118-124 */
  mpz_powm(lhs, e_a, e, nn); mpz_mul(lhs, lhs, e_d); mpz_mod(lhs, lhs, nn);  /* :133-134 */
  if (mpz_cmp(lhs, u) != 0) verdict = 0;
  mpz_powm(v, e_c, e, nn); mpz_mul(v, v, e_db); mpz_mod(v, v, nn);       /* :135-136 */
  if (!mpz_invert(v, v, nn)) verdict = 2;                                /* :137 unwrap() */
  else {
    mpz_set_ui(t, 0);
    mpz_powm(u, z2, n, nn);                                              /* enc_0_z2 = z2^n  :125-131 */
    mpz_powm(lhs, e_b, f, nn); mpz_mul(lhs, lhs, v); mpz_mod(lhs, lhs, nn);  /* :138-139 */
    if (mpz_cmp(lhs, u) != 0) verdict = 0;
  }
  j->verdict[b] = verdict;
  free(scratch);
  for (size_t k = 0; k < sizeof(all) / sizeof(all[0]); ++k) mpz_clear(all[k]);
}
void orc_mul_verify(const uint32_t* n, int nl, int batch, const uint32_t* e_a, const uint32_t* e_b, const uint32_t* e_c, const uint32_t* f,
                    const uint32_t* z1, const uint32_t* z2, const uint32_t* e_d, const uint32_t* e_db, uint8_t* verdict, int threads) {
  sig_job j = {nl, 0, 0, 0, 0, n, e_a, e_b, e_c, f, z1, z2, e_d, e_db, verdict};
  run_parallel(mul_verify_task, &j, batch, threads);
}

static void verlin_verify_task(void* arg, long b) {
  sig_job* j = (sig_job*)arg;
  const int nl = j->nl, nnl = 2 * nl, zl = j->zl;
  mpz_t n, nn, c, cp, phi_x, phi_a, z, zp, zdp, r_z, e, t, u, rhs, phi;
  __mpz_struct* all[] = {n, nn, c, cp, phi_x, phi_a, z, zp, zdp, r_z, e, t, u, rhs, phi};
  for (size_t k = 0; k < sizeof(all) / sizeof(all[0]); ++k) mpz_init(all[k]);
  uint8_t* scratch = (uint8_t*)malloc((size_t)nnl * 4 + 64);
  imp(n, j->n, nl);
  mpz_mul(nn, n, n);
  imp(c, j->a0 + (size_t)b * nnl, nnl); imp(cp, j->a1 + (size_t)b * nnl, nnl); imp(phi_x, j->a2 + (size_t)b * nnl, nnl);
  imp(phi_a, j->a3 + (size_t)b * nnl, nnl); imp(z, j->a4 + (size_t)b * zl, zl); imp(zp, j->a5 + (size_t)b * zl, zl);
  imp(zdp, j->a6 + (size_t)b * zl, zl); imp(r_z, j->a7 + (size_t)b * nnl, nnl);
  const __mpz_struct* items[] = {n, c, cp, phi_x, phi_a};
  digest_of(e, items, 5, scratch);                                       /* :102-108 */
  mpz_powm(rhs, phi_x, e, nn); mpz_mul(rhs, rhs, phi_a); mpz_mod(rhs, rhs, nn);   /* :109-118 */
  mpz_powm(phi, c, z, nn);                                               /* gen_phi :147-163 */
  mpz_powm(u, cp, zp, nn); mpz_mul(phi, phi, u); mpz_mod(phi, phi, nn);
  enc(u, n, nn, zdp, r_z, t); mpz_mul(phi, phi, u); mpz_mod(phi, phi, nn);
  j->verdict[b] = mpz_cmp(phi, rhs) == 0 ? 1 : 0;
  free(scratch);
  for (size_t k = 0; k < sizeof(all) / sizeof(all[0]); ++k) mpz_clear(all[k]);
}
void orc_verlin_verify(const uint32_t* n, int nl, int zl, int batch, const uint32_t* c, const uint32_t* cp, const uint32_t* phi_x,
                       const uint32_t* phi_a, const uint32_t* z, const uint32_t* zp, const uint32_t* zdp, const uint32_t* r_z,
                       uint8_t* verdict, int threads) {
  sig_job j = {nl, zl, 0, 0, 0, n, c, cp, phi_x, phi_a, z, zp, zdp, r_z, verdict};
  run_parallel(verlin_verify_task, &j, batch, threads);
}

static void dlog_verify_task(void* arg, long b) {
  sig_job* j = (sig_job*)arg;
  const int nl = j->nl, yl = j->zl;
  mpz_t N, g, ni, x, y, e, t, u;
  __mpz_struct* all[] = {N, g, ni, x, y, e, t, u};
  for (size_t k = 0; k < sizeof(all) / sizeof(all[0]); ++k) mpz_init(all[k]);
  uint8_t* scratch = (uint8_t*)malloc((size_t)nl * 4 + 64);
  imp(N, j->a0 + (size_t)b * nl, nl); imp(g, j->a1 + (size_t)b * nl, nl); imp(ni, j->a2 + (size_t)b * nl, nl);
  imp(x, j->a3 + (size_t)b * nl, nl); imp(y, j->a4 + (size_t)b * yl, yl);
  uint8_t verdict = 1;
  mpz_set_ui(t, 1); mpz_mul_2exp(t, t, 128);
  if (mpz_cmp(N, t) <= 0) verdict = 2;                                   /* :68 assert!(N > 2^K) */
  mpz_gcd(t, g, N); if (mpz_cmp_ui(t, 1) != 0) verdict = 2;              /* :71 */
  mpz_gcd(t, ni, N); if (mpz_cmp_ui(t, 1) != 0) verdict = 2;             /* :72 */
  if (verdict == 1) {
    const __mpz_struct* items[] = {x, g, N, ni};
    digest_of(e, items, 4, scratch);                                     /* :74-79 */
    mpz_powm(t, ni, e, N); mpz_powm(u, g, y, N); mpz_mul(u, u, t); mpz_mod(u, u, N);   /* :80-82 */
    verdict = mpz_cmp(x, u) == 0 ? 1 : 0;                                /* :85 */
  }
  j->verdict[b] = verdict;
  free(scratch);
  for (size_t k = 0; k < sizeof(all) / sizeof(all[0]); ++k) mpz_clear(all[k]);
}
void orc_dlog_verify(int nl, int yl, int batch, const uint32_t* N, const uint32_t* g, const uint32_t* ni, const uint32_t* x, const uint32_t* y,
                     uint8_t* verdict, int threads) {
  sig_job j = {nl, yl, 0, 0, 0, NULL, N, g, ni, x, y, NULL, NULL, NULL, verdict};
  run_parallel(dlog_verify_task, &j, batch, threads);
}

static void cmsg_verify_task(void* arg, long b) {
  sig_job* j = (sig_job*)arg;
  const int nl = j->nl, nnl = 2 * nl, M = j->M, el = j->el, ml = j->ml;
  mpz_t n, nn, c, chal, esum, two_b, m, e, z, a, gm, u, t, lhs;
  __mpz_struct* all[] = {n, nn, c, chal, esum, two_b, m, e, z, a, gm, u, t, lhs};
  for (size_t k = 0; k < sizeof(all) / sizeof(all[0]); ++k) mpz_init(all[k]);
  uint8_t* scratch = (uint8_t*)malloc((size_t)nnl * 4 + 64);
  imp(n, j->n, nl);
  mpz_mul(nn, n, n);
  imp(c, j->a0 + (size_t)b * nnl, nnl);
  mpz_set_ui(two_b, 1); mpz_mul_2exp(two_b, two_b, 256);
  SHA256_CTX h;                                                           /* chal = H(a_vec) mod 2^256  :128-129 */
  SHA256_Init(&h);
  for (int i = 0; i < M; ++i) {
    imp(a, j->a4 + ((size_t)b * M + i) * nnl, nnl);
    sha_update_mpz(&h, a, scratch);
  }
  uint8_t d[32];
  SHA256_Final(d, &h);
  __gmpz_import(chal, 32, 1, 1, 0, 0, d);
  mpz_mod(chal, chal, two_b);
  mpz_set_ui(esum, 0);
  for (int i = 0; i < M; ++i) {                                           /* :130-131 */
    imp(e, j->a2 + ((size_t)b * M + i) * el, el);
    mpz_add(esum, esum, e);
  }
  mpz_mod(esum, esum, two_b);
  uint8_t verdict = 1;
  if (mpz_cmp(chal, esum) != 0) verdict = 2;                              /* :133 assert_eq! */
  for (int i = 0; i < M && verdict != 2; ++i) {  /* u_vec is built for every slot before any comparison (:134-142) */
    imp(m, j->a1 + ((size_t)b * M + i) * ml, ml);
    imp(e, j->a2 + ((size_t)b * M + i) * el, el);
    imp(z, j->a3 + ((size_t)b * M + i) * nl, nl);
    imp(a, j->a4 + ((size_t)b * M + i) * nnl, nnl);
    mpz_mul(gm, m, n); mpz_add_ui(gm, gm, 1); mpz_mod(gm, gm, nn);        /* :136-139 */
    if (!mpz_invert(gm, gm, nn)) { verdict = 2; break; }                  /* unwrap() */
    mpz_mul(u, c, gm); mpz_mod(u, u, nn);                                 /* u_i  :140 */
    mpz_powm(t, z, n, nn);                                                /* z_i^n  :145 */
    mpz_powm(lhs, u, e, nn); mpz_mul(lhs, lhs, a); mpz_mod(lhs, lhs, nn); /* :146-147 */
    if (mpz_cmp(lhs, t) != 0) verdict = 0;                                /* :148-151 */
  }
  j->verdict[b] = verdict;
  free(scratch);
  for (size_t k = 0; k < sizeof(all) / sizeof(all[0]); ++k) mpz_clear(all[k]);
}
void orc_correct_message_verify(const uint32_t* n, int nl, int batch, int M, int ml, int el, const uint32_t* ciphertext, const uint32_t* valid,
                                const uint32_t* e_vec, const uint32_t* z_vec, const uint32_t* a_vec, uint8_t* verdict, int threads) {
  sig_job j = {nl, 0, M, el, ml, n, ciphertext, valid, e_vec, z_vec, a_vec, NULL, NULL, NULL, verdict};
  run_parallel(cmsg_verify_task, &j, batch, threads);
}

/* ======================================================================================================
 * Provers and the ZeroProof verifier on GMP: the CPU baselines of bench.py's secondary lines (configs[0] is
 * ZeroProof prove + verify on the CPU; the dlog / correct_message lines measure prove + verify).
 * Randomness is an input, as everywhere.  Compared with the Python restatement in tests/test_oracle_more.py.
 * ====================================================================================================== */
/* ZeroProof::prove (zero_enc_proof.rs:44-64): a = r'^n mod nn, e = H(n, c, a), z = r' r^e mod nn */
static void zero_prove_task(void* arg, long b) {
  sig_job* j = (sig_job*)arg;
  const int nl = j->nl, nnl = 2 * nl;
  mpz_t n, nn, r, c, rp, a, e, z, t;
  __mpz_struct* all[] = {n, nn, r, c, rp, a, e, z, t};
  for (size_t k = 0; k < sizeof(all) / sizeof(all[0]); ++k) mpz_init(all[k]);
  uint8_t* scratch = (uint8_t*)malloc((size_t)nnl * 4 + 64);
  imp(n, j->n, nl);
  mpz_mul(nn, n, n);
  imp(r, j->a0 + (size_t)b * nl, nl); imp(c, j->a1 + (size_t)b * nnl, nnl); imp(rp, j->a2 + (size_t)b * nl, nl);
  mpz_powm(a, rp, n, nn);                                                 /* :46-52 */
  const __mpz_struct* items[] = {n, c, a};
  digest_of(e, items, 3, scratch);                                        /* :54-58 */
  mpz_powm(z, r, e, nn); mpz_mul(z, z, rp); mpz_mod(z, z, nn);            /* :60-61 */
  expo((uint32_t*)j->a3 + (size_t)b * nnl, nnl, z);
  expo((uint32_t*)j->a4 + (size_t)b * nnl, nnl, a);
  free(scratch);
  for (size_t k = 0; k < sizeof(all) / sizeof(all[0]); ++k) mpz_clear(all[k]);
}
void orc_zero_prove(const uint32_t* n, int nl, int batch, const uint32_t* r, const uint32_t* c, const uint32_t* r_prime, uint32_t* z, uint32_t* a,
                    int threads) {
  sig_job j = {nl, 0, 0, 0, 0, n, r, c, r_prime, z, a, NULL, NULL, NULL, NULL};
  run_parallel(zero_prove_task, &j, batch, threads);
}
/* ZeroProof::verify (zero_enc_proof.rs:66-94): z^n == c^e a mod nn */
static void zero_verify_task(void* arg, long b) {
  sig_job* j = (sig_job*)arg;
  const int nl = j->nl, nnl = 2 * nl;
  mpz_t n, nn, c, z, a, e, t, u;
  __mpz_struct* all[] = {n, nn, c, z, a, e, t, u};
  for (size_t k = 0; k < sizeof(all) / sizeof(all[0]); ++k) mpz_init(all[k]);
  uint8_t* scratch = (uint8_t*)malloc((size_t)nnl * 4 + 64);
  imp(n, j->n, nl);
  mpz_mul(nn, n, n);
  imp(c, j->a0 + (size_t)b * nnl, nnl); imp(z, j->a1 + (size_t)b * nnl, nnl); imp(a, j->a2 + (size_t)b * nnl, nnl);
  const __mpz_struct* items[] = {n, c, a};
  digest_of(e, items, 3, scratch);                                        /* :67-71 */
  mpz_powm(t, z, n, nn);                                                  /* Enc(0, z) :73-79 */
  mpz_powm(u, c, e, nn); mpz_mul(u, u, a); mpz_mod(u, u, nn);             /* :81-88 */
  j->verdict[b] = mpz_cmp(t, u) == 0 ? 1 : 0;
  free(scratch);
  for (size_t k = 0; k < sizeof(all) / sizeof(all[0]); ++k) mpz_clear(all[k]);
}
void orc_zero_verify(const uint32_t* n, int nl, int batch, const uint32_t* c, const uint32_t* z, const uint32_t* a, uint8_t* verdict, int threads) {
  sig_job j = {nl, 0, 0, 0, 0, n, c, z, a, NULL, NULL, NULL, NULL, NULL, verdict};
  run_parallel(zero_verify_task, &j, batch, threads);
}

/* CompositeDLogProof::prove (wi_dlog_proof.rs:46-65): x = g^r mod N, e = H(x, g, N, ni), y = r + e * secret (unreduced) */
static void dlog_prove_task(void* arg, long b) {
  sig_job* j = (sig_job*)arg;
  const int nl = j->nl, yl = j->zl, sl = j->el, rl = j->ml;
  mpz_t N, g, ni, s, r, x, e, y;
  __mpz_struct* all[] = {N, g, ni, s, r, x, e, y};
  for (size_t k = 0; k < sizeof(all) / sizeof(all[0]); ++k) mpz_init(all[k]);
  uint8_t* scratch = (uint8_t*)malloc((size_t)nl * 4 + 64);
  imp(N, j->a0 + (size_t)b * nl, nl); imp(g, j->a1 + (size_t)b * nl, nl); imp(ni, j->a2 + (size_t)b * nl, nl);
  imp(s, j->a3 + (size_t)b * sl, sl); imp(r, j->a4 + (size_t)b * rl, rl);
  mpz_powm(x, g, r, N);                                                   /* :54 */
  const __mpz_struct* items[] = {x, g, N, ni};
  digest_of(e, items, 4, scratch);                                        /* :55-60 */
  mpz_mul(y, e, s); mpz_add(y, y, r);                                     /* :61 */
  expo((uint32_t*)j->a5 + (size_t)b * nl, nl, x);
  expo((uint32_t*)j->a6 + (size_t)b * yl, yl, y);
  free(scratch);
  for (size_t k = 0; k < sizeof(all) / sizeof(all[0]); ++k) mpz_clear(all[k]);
}
void orc_dlog_prove(int nl, int sl, int rl, int yl, int batch, const uint32_t* N, const uint32_t* g, const uint32_t* ni, const uint32_t* secret,
                    const uint32_t* r, uint32_t* x, uint32_t* y, int threads) {
  sig_job j = {nl, yl, 0, sl, rl, NULL, N, g, ni, secret, r, x, y, NULL, NULL};
  run_parallel(dlog_prove_task, &j, batch, threads);
}

/* CorrectMessageProof::prove (correct_message.rs:35-128); the message must be one of the valid ones (the caller's workload
 * guarantees it: the reference indexes out of bounds otherwise).  e_rand / z_rand: [batch][M-1] rows. */
typedef struct {
  int nl, M, ml, el;
  const uint32_t *n, *valid, *msg, *r, *e_rand, *z_rand, *w;
  uint32_t *ciphertext, *e_vec, *z_vec, *a_vec;
} cmsg_prove_job;
static void cmsg_prove_task(void* arg, long b) {
  cmsg_prove_job* j = (cmsg_prove_job*)arg;
  const int nl = j->nl, nnl = 2 * nl, M = j->M, ml = j->ml, el = j->el;
  mpz_t n, nn, msg, r, w, c, m, gm, u, e, z, a, t, chal, esum, two_b, ei, zi;
  __mpz_struct* all[] = {n, nn, msg, r, w, c, m, gm, u, e, z, a, t, chal, esum, two_b, ei, zi};
  for (size_t k = 0; k < sizeof(all) / sizeof(all[0]); ++k) mpz_init(all[k]);
  uint8_t* scratch = (uint8_t*)malloc((size_t)nnl * 4 + 64);
  imp(n, j->n, nl);
  mpz_mul(nn, n, n);
  imp(msg, j->msg + (size_t)b * ml, ml); imp(r, j->r + (size_t)b * nl, nl); imp(w, j->w + (size_t)b * nl, nl);
  enc(c, n, nn, msg, r, t);                                               /* :41-49 */
  mpz_set_ui(two_b, 1); mpz_mul_2exp(two_b, two_b, 256);
  mpz_set_ui(esum, 0);
  SHA256_CTX h;
  SHA256_Init(&h);
  int k = 0;
  uint8_t* hit = (uint8_t*)calloc((size_t)M, 1);                          /* slots whose valid message equals the message (:70,:98,:110) */
  for (int i = 0; i < M; ++i) {                                           /* a_vec :68-83 */
    imp(m, j->valid + ((size_t)b * M + i) * ml, ml);
    if (mpz_cmp(m, msg) == 0) {
      hit[i] = 1;
      mpz_powm(a, w, n, nn);
    } else {
      mpz_mul(gm, m, n); mpz_add_ui(gm, gm, 1); mpz_mod(gm, gm, nn);      /* u_i :51-57 */
      mpz_invert(gm, gm, nn);
      mpz_mul(u, c, gm); mpz_mod(u, u, nn);
      imp(e, j->e_rand + ((size_t)b * (M - 1) + k) * el, el);
      imp(z, j->z_rand + ((size_t)b * (M - 1) + k) * nl, nl);
      ++k;
      mpz_add(esum, esum, e);
      mpz_powm(t, z, n, nn);
      mpz_powm(u, u, e, nn);
      mpz_invert(u, u, nn);
      mpz_mul(a, t, u); mpz_mod(a, a, nn);
    }
    sha_update_mpz(&h, a, scratch);
    expo(j->a_vec + ((size_t)b * M + i) * nnl, nnl, a);
  }
  uint8_t d[32];
  SHA256_Final(d, &h);
  __gmpz_import(chal, 32, 1, 1, 0, 0, d);
  mpz_mod(chal, chal, two_b);                                             /* :85-88 */
  mpz_mod(esum, esum, two_b);
  mpz_sub(ei, chal, esum); mpz_mod(ei, ei, two_b);                        /* :93 */
  mpz_powm(zi, r, ei, n); mpz_mul(zi, zi, w); mpz_mod(zi, zi, n);         /* :94-95 */
  k = 0;
  for (int i = 0; i < M; ++i) {                                           /* :97-121 */
    if (hit[i]) {
      expo(j->e_vec + ((size_t)b * M + i) * 8, 8, ei);
      expo(j->z_vec + ((size_t)b * M + i) * nl, nl, zi);
    } else {
      memcpy(j->e_vec + ((size_t)b * M + i) * 8, j->e_rand + ((size_t)b * (M - 1) + k) * el, 32);
      memcpy(j->z_vec + ((size_t)b * M + i) * nl, j->z_rand + ((size_t)b * (M - 1) + k) * nl, (size_t)nl * 4);
      ++k;
    }
  }
  expo(j->ciphertext + (size_t)b * nnl, nnl, c);
  free(hit);
  free(scratch);
  for (size_t q = 0; q < sizeof(all) / sizeof(all[0]); ++q) mpz_clear(all[q]);
}
void orc_correct_message_prove(const uint32_t* n, int nl, int batch, int M, int ml, const uint32_t* valid, const uint32_t* msg, const uint32_t* r,
                               const uint32_t* e_rand, const uint32_t* z_rand, const uint32_t* w, uint32_t* ciphertext, uint32_t* e_vec,
                               uint32_t* z_vec, uint32_t* a_vec, int threads) {
  cmsg_prove_job j = {nl, M, ml, 8, n, valid, msg, r, e_rand, z_rand, w, ciphertext, e_vec, z_vec, a_vec};
  run_parallel(cmsg_prove_task, &j, batch, threads);
}
