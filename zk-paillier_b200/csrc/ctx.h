// Internal state behind the opaque zkp_ctx of include/zkp_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/zkp_b200.h"
#include "kernels.h"

namespace zkp {

// Growable device allocation (never shrinks; freed with the context).
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + (bytes >> 3) + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class U>
  U* as() const { return reinterpret_cast<U*>(p); }
};

// One shared (modulus, exponent) pair resident on the device.
struct KeySlot {
  int S = 0;          // kernel width (limbs)
  int limbs = 0;      // caller's width of the modulus
  DevBuf mod, r2, nR, sched, n0;
  std::vector<uint32_t> h_mod;  // host copy, S limbs
  int nsteps = 0;
  uint32_t n0inv = 0;
  bool ready = false;
  SharedKey view() const {
    SharedKey k;
    k.mod = mod.as<uint32_t>();
    k.r2 = r2.as<uint32_t>();
    k.nR = nR.as<uint32_t>();
    k.sched = sched.as<uint32_t>();
    k.nsteps = nsteps;
    k.n0inv = n0inv;
    k.S = S;
    return k;
  }
  void release() {
    mod.release(); r2.release(); nR.release(); sched.release(); n0.release();
    ready = false;
  }
};

constexpr int kAuxStreams = 4;  // concurrent modexp launches of one call (fork_stream / join_streams)

// KID_CALL: the device span of a whole sigma-protocol call - first kernel to last kernel on the main stream, copies excluded
enum KernelId { KID_MODEXP_SHARED = 0, KID_MODEXP_VAR = 1, KID_MODMUL = 2, KID_SHA = 3, KID_OTHER = 4, KID_CALL = 5, KID_COUNT = 6 };

struct ProfEntry {
  int kid;
  cudaEvent_t a, b;
  double units;
};

struct RpState {  // RangeProofNi staging (api_rangeproof.cu)
  int batch = 0, ef = 0, wl = 0;  // prove shape
  bool prove_staged = false, prove_done = false, verify_staged = false, verify_done = false;
  // prove: inputs, then outputs
  DevBuf range, x, r, w1in, w, swap, rr;    // w = [w1' | w2'] plaintexts, rr = [r1 | r2] bases
  DevBuf c, digest, kind, resp_w, resp_r;   // c = [c1 | c2]
  DevBuf rmul, fault;                       // r*r1 | r*r2 mod n
  DevBuf chal, v_chal;                      // interactive proof: the verifier's raw challenge bytes
  int chal_bytes = 0, v_chal_bytes = 0;
  bool pairs_done = false;
  // verify: owned copies of host inputs
  DevBuf v_range, v_cx, v_c, v_kind, v_resp_w, v_resp_r;
  // verify: work and outputs
  DevBuf v_digest, v_jobs_base, v_jobs_plain, v_tag, v_jobs_out, v_count, v_cmul, v_sel, v_ok, v_accept, v_fault;
  // verify: views of the inputs (own copies, or the prove buffers when chained on the device)
  int vbatch = 0, vef = 0, vwl = 0;
  const uint32_t* pv_range = nullptr;
  const uint32_t* pv_c = nullptr;
  const uint32_t* pv_resp_w = nullptr;
  const uint32_t* pv_resp_r = nullptr;
  const uint8_t* pv_kind = nullptr;
  long long enc_count = 0;
  std::vector<DevBuf*> all() {
    return {&range, &x, &r, &w1in, &w, &swap, &rr, &c, &digest, &kind, &resp_w, &resp_r, &rmul, &fault,
            &v_range, &v_cx, &v_c, &v_kind, &v_resp_w, &v_resp_r, &v_digest, &v_jobs_base, &v_jobs_plain,
            &v_tag, &v_jobs_out, &v_count, &v_cmul, &v_sel, &v_ok, &v_accept, &v_fault, &chal, &v_chal};
  }
};

struct CkState {  // NiCorrectKeyProof staging (api_correctkey.cu)
  int batch = 0, nl = 0, S = 0, salt_len = 0;
  bool staged = false, done = false;
  DevBuf n, sigma, salt, r2, n0inv, mask, rho, derived, accept, primes;
  int nprimes = 0;
  std::vector<DevBuf*> all() { return {&n, &sigma, &salt, &r2, &n0inv, &mask, &rho, &derived, &accept, &primes}; }
};

}  // namespace zkp

struct zkp_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int num_sms = 0;
  std::string err;
  zkp::KeySlot nn;   // modulus n^2 (or the generic shared modulus), exponent n
  zkp::KeySlot n;    // modulus n
  bool paillier = false;
  int n_limbs = 0;   // caller's width of n after zkp_set_key
  // K1m (two-digit Montgomery, modexp2m.cu): the default Paillier encryption kernel when the key qualifies.
  // ZKP_B200_ENC=k1 forces K1 (Montgomery modulo n^2).
  bool enc2m_key = false, enc2m_enabled = true;
  bool enc_small_k2h = true;  // launches of a few hundred encryptions go to K2h's latency layout (zkp_tune ZKP_TUNE_ENC_KERNEL)
  zkp::DevBuf enc2m_consts, enc2m_ops;
  int enc2m_nops = 0;
  long long enc2m_launches = 0, k1_launches = 0;
  double enc2m_mads = 0, k1_mads = 0;  // IMAD.WIDE one encryption executes (zkp_enc_executed_mads)
  zkp::DevBuf table;                 // window-table scratch: one region per stream that may run a modexp kernel
  size_t table_region_limbs = 0, table_off = 0;  // region size; offset of the current stream's region (limbs)
  cudaStream_t main_stream = nullptr, aux[zkp::kAuxStreams] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t aux_ev[zkp::kAuxStreams] = {nullptr, nullptr, nullptr, nullptr}, fork_ev = nullptr;
  unsigned aux_used = 0;
  zkp::DevBuf cursor;                // K2h work-unit cursors, one 256-byte slot per stream (ensure_table)
  int jobs_rows = 0;                 // K2h row form of the narrow / latency layouts: 0 = default, 1 = single rows, 2 = pair rows
  int jobs_shape = 0;                // K2h lane layout: 0 = by job count, 1 = wide lanes, 2 = narrow lanes (zkp_set_jobs_shape)
  zkp::DevBuf in0, in1, in2, in3, out0;  // generic staging for the one-shot calls
  bool profiling = false;
  std::vector<zkp::ProfEntry> prof;
  std::vector<cudaEvent_t> ev_pool;
  zkp::RpState rp;
  zkp::CkState ck;
};

namespace zkp {

int fail(zkp_ctx* c, int code, const char* what);
int fail_cuda(zkp_ctx* c, cudaError_t e, const char* what);

#define ZKP_CU(ctx, call)                                        \
  do {                                                           \
    cudaError_t e__ = (call);                                    \
    if (e__ != cudaSuccess) return zkp::fail_cuda(ctx, e__, #call); \
  } while (0)

// Scoped device-time accounting of one kernel launch (no-op unless profiling).
struct ProfScope {
  zkp_ctx* c;
  int idx = -1;
  ProfScope(zkp_ctx* ctx, int kid, double units);
  ~ProfScope();
};

// table scratch large enough for K1/K2 at width S
// pow_jobs > 0: also room for a K2h launch of that many jobs of up to pow_bases bases each (launch_pow_jobs)
cudaError_t ensure_table(zkp_ctx* c, int S, int entries, int pow_jobs = 0, int pow_bases = 1);
// Paillier::encrypt_with_chosen_randomness for `jobs` rows under the current key: picks K1v2 (two-digit base-n)
// when the key and row widths qualify, else K1 (Montgomery mod n^2).  plain == nullptr encrypts 0.
// Returns 1 if K1v2 ran, 0 if K1 ran, negative cudaError as -(int)err - 1000 on failure (see enc_failed()).
cudaError_t launch_enc(zkp_ctx* c, const uint32_t* bases, int base_limbs, const uint32_t* plain, int plain_limbs, uint32_t* out,
                       int jobs, const unsigned* jobs_dev = nullptr);

// BigInt::mod_pow(base, exp, nn) / Paillier::mul for `jobs` rows with per-row exponents (row j uses exps[j / exp_per]) under
// the current key: K2m (two-digit Montgomery form) when the key qualifies, else K2 on the modulus n^2.
// The table scratch must have been sized with ensure_table(c, c->nn.S, kTableVar).
cudaError_t launch_pow_nn(zkp_ctx* c, const uint32_t* base, int base_limbs, const uint32_t* exp, int exp_limbs, int exp_bits,
                          int exp_per, uint32_t* out, int jobs);

// K2h: every modexp / Enc of a batch of sigma-protocol proofs in one launch (modexp2m.cu: modexp2m_jobs_kernel).  Needs a key
// that K1m takes (c->enc2m_key); callers fall back to launch_enc / launch_pow_nn otherwise.
bool jobs_supported(const zkp_ctx* c);
cudaError_t launch_pow_jobs(zkp_ctx* c, const PowJobs& jobs, const unsigned* jobs_dev = nullptr);

// Concurrent modexp launches inside one call (api_core.cu)
cudaError_t fork_stream(zkp_ctx* c, int k);
cudaError_t main_stream(zkp_ctx* c);
cudaError_t join_streams(zkp_ctx* c);

// Montgomery/key helpers (api_core.cu)
int setup_slot(zkp_ctx* c, KeySlot& slot, const uint32_t* mod, int limbs, const uint32_t* exp, int exp_limbs);

}  // namespace zkp
