// C-ABI layer, part 1: context, key setup, the raw modexp / modmul / Enc calls and
// the measurement hooks declared in include/zkp_b200.h.
//
// Host code here only marshals buffers and derives per-key constants; every
// big-integer operation on caller data runs in the CUDA kernels of modexp.cu.
// There is deliberately no CPU fallback: zkp_ctx_create fails without a device.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>

#include "ctx.h"

using namespace zkp;

namespace zkp {

int fail(zkp_ctx* c, int code, const char* what) {
  if (c) c->err = what;
  return code;
}
int fail_cuda(zkp_ctx* c, cudaError_t e, const char* what) {
  if (c) {
    c->err = std::string(cudaGetErrorName(e)) + ": " + cudaGetErrorString(e) + " in " + what;
  }
  cudaGetLastError();
  return ZKP_E_CUDA;
}

ProfScope::ProfScope(zkp_ctx* ctx, int kid, double units) : c(ctx) {
  if (!c->profiling) return;
  ProfEntry e;
  e.kid = kid;
  e.units = units;
  auto get = [&]() {
    cudaEvent_t ev;
    if (!c->ev_pool.empty()) {
      ev = c->ev_pool.back();
      c->ev_pool.pop_back();
    } else {
      cudaEventCreate(&ev);
    }
    return ev;
  };
  e.a = get();
  e.b = get();
  cudaEventRecord(e.a, c->stream);
  c->prof.push_back(e);
  idx = (int)c->prof.size() - 1;
}
ProfScope::~ProfScope() {
  if (idx >= 0) cudaEventRecord(c->prof[idx].b, c->stream);
}

cudaError_t ensure_table(zkp_ctx* c, int S, int entries, int pow_jobs, int pow_bases) {
  size_t bytes = (size_t)resident_groups(S, c->num_sms) * entries * S * sizeof(uint32_t);
  if (c->enc2m_key && c->enc2m_enabled) {  // a call may run K1m (Enc) and K2m (mod_pow) back to back: size for both
    const size_t b1 = enc2m_table_limbs(c->n.S, c->num_sms) * sizeof(uint32_t);
    const size_t b2 = var2m_table_limbs(c->n.S, c->num_sms) * sizeof(uint32_t);
    const size_t b3 = pow_jobs > 0 ? jobs2m_scratch_limbs(c->n.S, c->num_sms, pow_jobs, pow_bases) * sizeof(uint32_t) : 0;
    bytes = std::max(std::max(bytes, b3), std::max(b1, b2));
  }
  bytes = (bytes + 255) & ~size_t(255);
  // one region per stream that may run a modexp kernel at the same time (fork_stream below)
  if (bytes / sizeof(uint32_t) < c->table_region_limbs) bytes = c->table_region_limbs * sizeof(uint32_t);  // regions never shrink
  c->table_region_limbs = bytes / sizeof(uint32_t);
  cudaError_t e = c->cursor.ensure(256 * (1 + kAuxStreams));
  if (e != cudaSuccess) return e;
  return c->table.ensure(bytes * (1 + kAuxStreams));
}

// ---- concurrent modexp launches of one call (the independent modexps of a sigma-protocol proof) -----------------
// fork_stream(c, k): later launches go to auxiliary stream k (which first waits for everything queued on the main stream
// so far) and use window-table region k + 1; join_streams(c): back on the main stream, which waits for every auxiliary
// stream used since the last join.
cudaError_t fork_stream(zkp_ctx* c, int k) {
  if (k < 0 || k >= kAuxStreams) return cudaErrorInvalidValue;
  if (!c->main_stream) c->main_stream = c->stream;
  if (!c->aux[k]) {
    cudaError_t e = cudaStreamCreateWithFlags(&c->aux[k], cudaStreamNonBlocking);
    if (e != cudaSuccess) return e;
    e = cudaEventCreateWithFlags(&c->aux_ev[k], cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
  }
  if (!c->fork_ev) {
    cudaError_t e = cudaEventCreateWithFlags(&c->fork_ev, cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
  }
  cudaError_t e = cudaEventRecord(c->fork_ev, c->main_stream);
  if (e != cudaSuccess) return e;
  e = cudaStreamWaitEvent(c->aux[k], c->fork_ev, 0);
  if (e != cudaSuccess) return e;
  c->stream = c->aux[k];
  c->table_off = (size_t)(k + 1) * c->table_region_limbs;
  c->aux_used |= 1u << k;
  return cudaSuccess;
}
cudaError_t main_stream(zkp_ctx* c) {  // back to the main stream without waiting for the auxiliary ones
  if (c->main_stream) c->stream = c->main_stream;
  c->table_off = 0;
  return cudaSuccess;
}
cudaError_t join_streams(zkp_ctx* c) {
  main_stream(c);
  for (int k = 0; k < kAuxStreams; ++k) {
    if (!(c->aux_used & (1u << k))) continue;
    cudaError_t e = cudaEventRecord(c->aux_ev[k], c->aux[k]);
    if (e != cudaSuccess) return e;
    e = cudaStreamWaitEvent(c->stream, c->aux_ev[k], 0);
    if (e != cudaSuccess) return e;
  }
  c->aux_used = 0;
  return cudaSuccess;
}

static Enc2mKey enc2m_view(const zkp_ctx* c) {
  Enc2mKey k;
  k.mod = c->n.mod.as<uint32_t>();
  k.consts = c->enc2m_consts.as<uint32_t>();
  k.ops = c->enc2m_ops.as<uint32_t>();
  k.nops = c->enc2m_nops;
  k.n0inv = c->n.n0inv;
  {  // -n^{-1} mod 2^64 by Newton iteration from the low 64 bits of n (odd)
    const uint64_t n64 = (uint64_t)c->n.h_mod[0] | ((uint64_t)(c->n.S > 1 ? c->n.h_mod[1] : 0u) << 32);
    uint64_t x = n64;  // correct to 3 bits
    for (int i = 0; i < 6; ++i) x *= 2 - n64 * x;
    k.n0inv_hi = (uint32_t)((0 - x) >> 32);
  }
  k.S = c->n.S;
  return k;
}

cudaError_t launch_pow_nn(zkp_ctx* c, const uint32_t* base, int base_limbs, const uint32_t* exp, int exp_limbs, int exp_bits,
                          int exp_per, uint32_t* out, int jobs) {
  if (c->enc2m_key && c->enc2m_enabled && base_limbs <= 2 * c->n.S) {
    ++c->enc2m_launches;
    return launch_modexp2m_var(enc2m_view(c), base, base_limbs, exp, exp_limbs, exp_bits, exp_per, out, c->nn.limbs, jobs,
                               (c->table.as<uint32_t>() + c->table_off), c->num_sms, c->stream);
  }
  ++c->k1_launches;
  return launch_modexp_var(base, c->nn.mod.as<uint32_t>(), c->nn.limbs, c->nn.r2.as<uint32_t>(), c->nn.n0.as<uint32_t>(), exp,
                           exp_limbs, exp_bits, exp_per, 0x7fffffff, out, jobs, c->nn.S, (c->table.as<uint32_t>() + c->table_off), c->num_sms,
                           c->stream, base_limbs);
}

bool jobs_supported(const zkp_ctx* c) { return c->enc2m_key && c->enc2m_enabled; }

cudaError_t launch_pow_jobs(zkp_ctx* c, const PowJobs& jobs, const unsigned* jobs_dev) {
  if (!jobs_supported(c)) return cudaErrorNotSupported;
  ++c->enc2m_launches;
  const size_t region = c->table_region_limbs ? c->table_off / c->table_region_limbs : 0;  // 0 = main stream, k + 1 = auxiliary stream k
  return launch_modexp2m_jobs(enc2m_view(c), jobs, c->nn.limbs, c->table.as<uint32_t>() + c->table_off, c->table_region_limbs,
                              reinterpret_cast<unsigned*>(c->cursor.as<uint8_t>() + 256 * region), c->num_sms, c->stream, c->jobs_shape, jobs_dev, c->jobs_rows);
}

// A launch of so few encryptions that K1m's layout would leave sub-partitions with less than two warps (one proof: 256
// encryptions are 64 warps of Mp<8,8> on 592 sub-partitions) is latency-bound: K2h spreads every job over more lanes.
static bool enc_is_small(const zkp_ctx* c, int jobs) {
  const int T = c->n.S <= 32 ? 4 : (c->n.S <= 96 ? 8 : 16);  // lanes per job of K1m (modexp2m.cu: pick_shape)
  return (long long)jobs * T / 32 < 2ll * 4 * c->num_sms;
}

cudaError_t launch_enc(zkp_ctx* c, const uint32_t* bases, int base_limbs, const uint32_t* plain, int plain_limbs, uint32_t* out,
                       int jobs, const unsigned* jobs_dev) {
  if (c->enc2m_key && c->enc2m_enabled && base_limbs <= 2 * c->n.S && (!plain || plain_limbs <= 2 * c->n.S) && c->enc_small_k2h && enc_is_small(c, jobs) &&
      c->cursor.p && jobs2m_scratch_limbs(c->n.S, c->num_sms, jobs) <= c->table_region_limbs) {
    PowJobs pj;
    PowSeg& g = pj.seg[0];
    int n_bits = 32 * c->n.S;
    while (n_bits > 1 && !((c->n.h_mod[(n_bits - 1) >> 5] >> ((n_bits - 1) & 31)) & 1u)) --n_bits;
    for (int k = 0; k < kMaxPowBases; ++k) {
      g.base[k] = bases;
      g.base_limbs[k] = base_limbs;
      g.exp[k] = c->n.mod.as<uint32_t>();
      g.exp_limbs[k] = c->n.S;
      g.exp_stride[k] = 0;
    }
    g.nbase = 1;
    g.exp_bits = n_bits;
    g.plain = plain;
    g.plain_limbs = plain ? plain_limbs : 0;
    g.out = out;
    g.jobs = jobs;
    g.first = 0;
    pj.nseg = 1;
    pj.total = jobs;
    return launch_pow_jobs(c, pj, jobs_dev);
  }
  if (c->enc2m_key && c->enc2m_enabled && base_limbs <= 2 * c->n.S && (!plain || plain_limbs <= 2 * c->n.S)) {
    const Enc2mKey k = enc2m_view(c);
    ++c->enc2m_launches;
    return launch_enc2m(k, bases, base_limbs, plain, plain_limbs, out, c->nn.limbs, jobs, (c->table.as<uint32_t>() + c->table_off), c->num_sms,
                        c->stream, jobs_dev);
  }
  ++c->k1_launches;
  return launch_modexp_shared(c->nn.view(), bases, base_limbs, plain, plain_limbs, out, c->nn.limbs, jobs, (c->table.as<uint32_t>() + c->table_off),
                              c->num_sms, c->stream, jobs_dev);
}

// Sliding-window (width kWindowShared) recoding of a public exponent, most
// significant bit first.  Entry k: low byte = table index of the odd power
// x^(2*idx+1) to multiply by (0xff = none), upper 24 bits = squarings to do first.
// Entry 0 only loads the accumulator.
std::vector<uint32_t> recode_exponent(const uint32_t* e, int limbs, int window = kWindowShared) {
  std::vector<uint32_t> out;
  int top = limbs * 32 - 1;
  auto bit = [&](int i) { return (e[i >> 5] >> (i & 31)) & 1u; };
  while (top >= 0 && !bit(top)) --top;
  if (top < 0) return out;  // exponent zero
  int i = top;
  uint32_t pending_sq = 0;
  bool first = true;
  while (i >= 0) {
    if (!bit(i)) {
      ++pending_sq;
      --i;
      continue;
    }
    int l = std::max(i - window + 1, 0);
    while (!bit(l)) ++l;  // window [i .. l] ends in a one
    uint32_t v = 0;
    for (int k = i; k >= l; --k) v = (v << 1) | bit(k);
    uint32_t width = (uint32_t)(i - l + 1);
    if (first) {
      out.push_back(v >> 1);
      first = false;
    } else {
      out.push_back(((pending_sq + width) << 8) | (v >> 1));
    }
    pending_sq = 0;
    i = l - 1;
  }
  if (pending_sq) out.push_back((pending_sq << 8) | 0xffu);
  return out;
}

int setup_slot(zkp_ctx* c, KeySlot& slot, const uint32_t* mod, int limbs, const uint32_t* exp, int exp_limbs) {
  slot.ready = false;
  if (!mod || limbs <= 0 || (limbs % 4)) return fail(c, ZKP_E_ARG, "modulus width must be a positive multiple of 4 limbs");
  if (!(mod[0] & 1u)) return fail(c, ZKP_E_ARG, "modulus must be odd");
  int S = pick_width(limbs);
  if (S < 0) return fail(c, ZKP_E_ARG, "modulus wider than 8192 bits");
  slot.S = S;
  slot.limbs = limbs;
  slot.h_mod.assign(S, 0u);
  memcpy(slot.h_mod.data(), mod, (size_t)limbs * 4);
  ZKP_CU(c, slot.mod.ensure((size_t)S * 4));
  ZKP_CU(c, slot.r2.ensure((size_t)S * 4));
  ZKP_CU(c, slot.nR.ensure((size_t)S * 4));
  ZKP_CU(c, slot.n0.ensure(16));
  ZKP_CU(c, cudaMemcpyAsync(slot.mod.p, slot.h_mod.data(), (size_t)S * 4, cudaMemcpyHostToDevice, c->stream));
  ZKP_CU(c, cudaMemsetAsync(slot.nR.p, 0, (size_t)S * 4, c->stream));
  ZKP_CU(c, launch_mont_setup(slot.mod.as<uint32_t>(), S, S, 1, slot.r2.as<uint32_t>(), slot.n0.as<uint32_t>(), c->stream));
  ZKP_CU(c, cudaMemcpyAsync(&slot.n0inv, slot.n0.p, 4, cudaMemcpyDeviceToHost, c->stream));
  slot.nsteps = 0;
  if (exp) {
    std::vector<uint32_t> sched = recode_exponent(exp, exp_limbs);
    if (sched.empty()) return fail(c, ZKP_E_ARG, "shared exponent must be non-zero");
    if ((int)sched.size() > kMaxSchedSteps) return fail(c, ZKP_E_ARG, "shared exponent too long");
    slot.nsteps = (int)sched.size();
    sched.resize((sched.size() + 3) & ~size_t(3), 0xffu);
    ZKP_CU(c, slot.sched.ensure(sched.size() * 4));
    ZKP_CU(c, cudaMemcpyAsync(slot.sched.p, sched.data(), sched.size() * 4, cudaMemcpyHostToDevice, c->stream));
  }
  ZKP_CU(c, cudaStreamSynchronize(c->stream));
  slot.ready = true;
  return ZKP_OK;
}

}  // namespace zkp

extern "C" {

int zkp_version(void) { return 100; }

int zkp_ctx_create(int device, void* stream, zkp_ctx** out) {
  if (!out) return ZKP_E_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
    cudaGetLastError();
    return ZKP_E_CUDA;  // no CUDA device: there is no CPU path behind this ABI
  }
  zkp_ctx* c = new (std::nothrow) zkp_ctx();
  if (!c) return ZKP_E_NOMEM;
  c->device = device;
  if (cudaSetDevice(device) != cudaSuccess) {
    delete c;
    return ZKP_E_CUDA;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    delete c;
    return ZKP_E_CUDA;
  }
  c->num_sms = prop.multiProcessorCount;
  if (stream) {
    c->stream = (cudaStream_t)stream;
  } else {
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
      delete c;
      return ZKP_E_CUDA;
    }
    c->own_stream = true;
  }
  *out = c;
  return ZKP_OK;
}

void zkp_ctx_destroy(zkp_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (auto& e : c->prof) {
    cudaEventDestroy(e.a);
    cudaEventDestroy(e.b);
  }
  for (auto e : c->ev_pool) cudaEventDestroy(e);
  c->nn.release();
  c->n.release();
  std::vector<DevBuf*> bufs = {&c->enc2m_consts, &c->enc2m_ops, &c->table, &c->cursor, &c->in0, &c->in1, &c->in2, &c->in3, &c->out0};
  for (DevBuf* b : c->rp.all()) bufs.push_back(b);
  for (DevBuf* b : c->ck.all()) bufs.push_back(b);
  for (DevBuf* b : bufs) b->release();
  if (c->main_stream) c->stream = c->main_stream;
  for (int k = 0; k < zkp::kAuxStreams; ++k) {
    if (c->aux[k]) cudaStreamDestroy(c->aux[k]);
    if (c->aux_ev[k]) cudaEventDestroy(c->aux_ev[k]);
  }
  if (c->fork_ev) cudaEventDestroy(c->fork_ev);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
}

const char* zkp_last_error(const zkp_ctx* c) { return c ? c->err.c_str() : "null context"; }
int zkp_sm_count(const zkp_ctx* c) { return c ? c->num_sms : 0; }

int zkp_sync(zkp_ctx* c) {
  if (!c) return ZKP_E_ARG;
  ZKP_CU(c, cudaStreamSynchronize(c->stream));
  return ZKP_OK;
}

int zkp_profile_enable(zkp_ctx* c, int on) {
  if (!c) return ZKP_E_ARG;
  c->profiling = on != 0;
  return ZKP_OK;
}
int zkp_profile_reset(zkp_ctx* c) {
  if (!c) return ZKP_E_ARG;
  cudaStreamSynchronize(c->stream);
  for (auto& e : c->prof) {
    c->ev_pool.push_back(e.a);
    c->ev_pool.push_back(e.b);
  }
  c->prof.clear();
  return ZKP_OK;
}
int zkp_profile_get(zkp_ctx* c, int kernel, double* ms_total, long long* launches, double* units) {
  if (!c || kernel < 0 || kernel >= KID_COUNT) return ZKP_E_ARG;
  ZKP_CU(c, cudaStreamSynchronize(c->stream));
  double ms = 0, u = 0;
  long long n = 0;
  for (auto& e : c->prof) {
    if (e.kid != kernel) continue;
    float t = 0;
    ZKP_CU(c, cudaEventElapsedTime(&t, e.a, e.b));
    ms += t;
    u += e.units;
    ++n;
  }
  if (ms_total) *ms_total = ms;
  if (launches) *launches = n;
  if (units) *units = u;
  return ZKP_OK;
}

int zkp_set_key(zkp_ctx* c, const uint32_t* n, int n_limbs) {
  if (!c) return ZKP_E_ARG;
  c->paillier = false;
  if (!n || n_limbs <= 0 || (n_limbs % 4) || 2 * n_limbs > 256) return fail(c, ZKP_E_ARG, "n_limbs must be a multiple of 4, at most 128");
  ZKP_CU(c, cudaSetDevice(c->device));
  // nn = n * n (host schoolbook; once per key)
  std::vector<uint32_t> nn((size_t)2 * n_limbs, 0u);
  for (int i = 0; i < n_limbs; ++i) {
    uint64_t carry = 0;
    for (int j = 0; j < n_limbs; ++j) {
      uint64_t t = (uint64_t)n[i] * n[j] + nn[i + j] + carry;
      nn[i + j] = (uint32_t)t;
      carry = t >> 32;
    }
    nn[i + n_limbs] = (uint32_t)carry;
  }
  int rc = setup_slot(c, c->nn, nn.data(), 2 * n_limbs, n, n_limbs);
  if (rc) return rc;
  rc = setup_slot(c, c->n, n, n_limbs, nullptr, 0);
  if (rc) return rc;
  // nR = n * R mod nn (Montgomery form of n, used by the Enc epilogue)
  SharedKey k = c->nn.view();
  ZKP_CU(c, launch_modmul_shared(k, 1, c->n.mod.as<uint32_t>(), c->n.S, nullptr, 0, 1, c->nn.nR.as<uint32_t>(), k.S, 1,
                                 c->stream));
  ZKP_CU(c, cudaStreamSynchronize(c->stream));
  c->n_limbs = n_limbs;
  c->paillier = true;
  {
    // K1m: two-digit Montgomery form (modexp2m.cu), the default encryption kernel
    c->enc2m_key = false;
    if (enc2m_supported(c->n.h_mod.data(), c->n.S)) {
      const int S = c->n.S;
      std::vector<uint32_t> consts((size_t)5 * S);
      enc2m_host_constants(c->n.h_mod.data(), S, consts.data());
      std::vector<uint32_t> sched = recode_exponent(n, n_limbs, enc2m_window());
      std::vector<uint32_t> ops = enc2m_ops(sched.data(), (int)sched.size());
      c->enc2m_nops = (int)ops.size();
      ops.resize((ops.size() + 3) & ~size_t(3), 0xffffffu);
      ZKP_CU(c, c->enc2m_consts.ensure(consts.size() * 4));
      ZKP_CU(c, c->enc2m_ops.ensure(ops.size() * 4));
      ZKP_CU(c, cudaMemcpyAsync(c->enc2m_consts.p, consts.data(), consts.size() * 4, cudaMemcpyHostToDevice, c->stream));
      ZKP_CU(c, cudaMemcpyAsync(c->enc2m_ops.p, ops.data(), ops.size() * 4, cudaMemcpyHostToDevice, c->stream));
      ZKP_CU(c, cudaStreamSynchronize(c->stream));
      c->enc2m_key = true;
      double sq = 0, mu = 0;
      for (int k = 0; k < c->enc2m_nops; ++k) {
        sq += ops[k] >> 24;
        mu += (ops[k] & 0xffu) != 0xffu;
      }
      c->enc2m_mads = ((double)S * S) * (enc2m_sqr_products() * sq + 5.0 * mu + 1.0);
    }
    {
      std::vector<uint32_t> sched = recode_exponent(n, n_limbs);
      double mm = 2 + (kTableShared - 1) + 2;  // into Montgomery form, x^2, the odd powers, m n and the final product
      for (size_t k = 1; k < sched.size(); ++k) mm += (sched[k] >> 8) + ((sched[k] & 0xffu) != 0xffu);
      c->k1_mads = mm * 2.0 * c->nn.S * c->nn.S;
    }
  }
  c->rp.prove_staged = c->rp.prove_done = c->rp.verify_staged = c->rp.verify_done = false;
  return ZKP_OK;
}

int zkp_set_modulus(zkp_ctx* c, const uint32_t* mod, int mod_limbs, const uint32_t* exp, int exp_limbs) {
  if (!c) return ZKP_E_ARG;
  if (!exp || exp_limbs <= 0) return fail(c, ZKP_E_ARG, "exponent required");
  ZKP_CU(c, cudaSetDevice(c->device));
  c->paillier = false;
  c->n.ready = false;
  return setup_slot(c, c->nn, mod, mod_limbs, exp, exp_limbs);
}

int zkp_nn_limbs(const zkp_ctx* c) { return (c && c->nn.ready) ? c->nn.limbs : 0; }

int zkp_modexp_shared(zkp_ctx* c, const uint32_t* bases, int base_limbs, int batch, uint32_t* out) {
  if (!c) return ZKP_E_ARG;
  if (!c->nn.ready) return fail(c, ZKP_E_STATE, "no key / modulus set");
  if (batch < 0 || !bases || !out) return fail(c, ZKP_E_ARG, "null buffer or negative batch");
  if (base_limbs <= 0 || base_limbs % 4 || base_limbs > c->nn.S) return fail(c, ZKP_E_ARG, "bad base_limbs");
  if (batch == 0) return ZKP_OK;
  ZKP_CU(c, cudaSetDevice(c->device));
  const int ol = c->nn.limbs;
  ZKP_CU(c, c->in0.ensure((size_t)batch * base_limbs * 4));
  ZKP_CU(c, c->out0.ensure((size_t)batch * ol * 4));
  ZKP_CU(c, ensure_table(c, c->nn.S, kTableShared));
  ZKP_CU(c, cudaMemcpyAsync(c->in0.p, bases, (size_t)batch * base_limbs * 4, cudaMemcpyHostToDevice, c->stream));
  {
    ProfScope ps(c, KID_MODEXP_SHARED, batch);
    ZKP_CU(c, launch_modexp_shared(c->nn.view(), c->in0.as<uint32_t>(), base_limbs, nullptr, 0, c->out0.as<uint32_t>(), ol,
                                   batch, (c->table.as<uint32_t>() + c->table_off), c->num_sms, c->stream));
  }
  ZKP_CU(c, cudaMemcpyAsync(out, c->out0.p, (size_t)batch * ol * 4, cudaMemcpyDeviceToHost, c->stream));
  ZKP_CU(c, cudaStreamSynchronize(c->stream));
  return ZKP_OK;
}

int zkp_paillier_enc(zkp_ctx* c, const uint32_t* m, int m_limbs, const uint32_t* r, int r_limbs, int batch,
                     uint32_t* out) {
  if (!c) return ZKP_E_ARG;
  if (!c->paillier) return fail(c, ZKP_E_STATE, "zkp_set_key not called");
  if (batch < 0 || !m || !r || !out) return fail(c, ZKP_E_ARG, "null buffer or negative batch");
  if (m_limbs <= 0 || m_limbs % 4 || m_limbs > c->nn.S || r_limbs <= 0 || r_limbs % 4 || r_limbs > c->nn.S)
    return fail(c, ZKP_E_ARG, "bad m_limbs / r_limbs");
  if (batch == 0) return ZKP_OK;
  ZKP_CU(c, cudaSetDevice(c->device));
  const int ol = c->nn.limbs;
  ZKP_CU(c, c->in0.ensure((size_t)batch * r_limbs * 4));
  ZKP_CU(c, c->in1.ensure((size_t)batch * m_limbs * 4));
  ZKP_CU(c, c->out0.ensure((size_t)batch * ol * 4));
  ZKP_CU(c, ensure_table(c, c->nn.S, kTableShared));
  ZKP_CU(c, cudaMemcpyAsync(c->in0.p, r, (size_t)batch * r_limbs * 4, cudaMemcpyHostToDevice, c->stream));
  ZKP_CU(c, cudaMemcpyAsync(c->in1.p, m, (size_t)batch * m_limbs * 4, cudaMemcpyHostToDevice, c->stream));
  {
    ProfScope ps(c, KID_MODEXP_SHARED, batch);
    ZKP_CU(c, launch_enc(c, c->in0.as<uint32_t>(), r_limbs, c->in1.as<uint32_t>(), m_limbs, c->out0.as<uint32_t>(), batch));
  }
  ZKP_CU(c, cudaMemcpyAsync(out, c->out0.p, (size_t)batch * ol * 4, cudaMemcpyDeviceToHost, c->stream));
  ZKP_CU(c, cudaStreamSynchronize(c->stream));
  return ZKP_OK;
}

int zkp_modexp_var(zkp_ctx* c, const uint32_t* bases, const uint32_t* exps, int exp_limbs, int exp_bits, int exp_per,
                   const uint32_t* mods, int mod_limbs, int mod_per, int batch, uint32_t* out) {
  if (!c) return ZKP_E_ARG;
  if (batch < 0 || !bases || !exps || !mods || !out) return fail(c, ZKP_E_ARG, "null buffer or negative batch");
  if (mod_limbs <= 0 || mod_limbs % 4 || exp_limbs <= 0 || exp_bits <= 0 || exp_bits > 32 * exp_limbs || exp_per <= 0 ||
      mod_per <= 0)
    return fail(c, ZKP_E_ARG, "bad widths");
  int S = pick_width(mod_limbs);
  if (S < 0) return fail(c, ZKP_E_ARG, "modulus wider than 8192 bits");
  if (batch == 0) return ZKP_OK;
  const int nmod = (batch + mod_per - 1) / mod_per;
  const int nexp = (batch + exp_per - 1) / exp_per;
  for (int i = 0; i < nmod; ++i)
    if (!(mods[(size_t)i * mod_limbs] & 1u)) return fail(c, ZKP_E_ARG, "every modulus must be odd");
  ZKP_CU(c, cudaSetDevice(c->device));
  ZKP_CU(c, c->in0.ensure((size_t)batch * mod_limbs * 4));
  ZKP_CU(c, c->in1.ensure((size_t)nexp * exp_limbs * 4));
  ZKP_CU(c, c->in2.ensure((size_t)nmod * mod_limbs * 4));
  ZKP_CU(c, c->in3.ensure((size_t)nmod * (S + 1) * 4));
  ZKP_CU(c, c->out0.ensure((size_t)batch * mod_limbs * 4));
  ZKP_CU(c, ensure_table(c, S, kTableVar));
  uint32_t* d_r2 = c->in3.as<uint32_t>();
  uint32_t* d_n0 = d_r2 + (size_t)nmod * S;
  ZKP_CU(c, cudaMemcpyAsync(c->in0.p, bases, (size_t)batch * mod_limbs * 4, cudaMemcpyHostToDevice, c->stream));
  ZKP_CU(c, cudaMemcpyAsync(c->in1.p, exps, (size_t)nexp * exp_limbs * 4, cudaMemcpyHostToDevice, c->stream));
  ZKP_CU(c, cudaMemcpyAsync(c->in2.p, mods, (size_t)nmod * mod_limbs * 4, cudaMemcpyHostToDevice, c->stream));
  {
    ProfScope ps(c, KID_OTHER, nmod);
    ZKP_CU(c, launch_mont_setup(c->in2.as<uint32_t>(), mod_limbs, S, nmod, d_r2, d_n0, c->stream));
  }
  {
    ProfScope ps(c, KID_MODEXP_VAR, batch);
    ZKP_CU(c, launch_modexp_var(c->in0.as<uint32_t>(), c->in2.as<uint32_t>(), mod_limbs, d_r2, d_n0, c->in1.as<uint32_t>(),
                                exp_limbs, exp_bits, exp_per, mod_per, c->out0.as<uint32_t>(), batch, S,
                                (c->table.as<uint32_t>() + c->table_off), c->num_sms, c->stream));
  }
  ZKP_CU(c, cudaMemcpyAsync(out, c->out0.p, (size_t)batch * mod_limbs * 4, cudaMemcpyDeviceToHost, c->stream));
  ZKP_CU(c, cudaStreamSynchronize(c->stream));
  return ZKP_OK;
}

int zkp_modmul(zkp_ctx* c, int which_nn, const uint32_t* a, const uint32_t* b, int b_per, int batch, uint32_t* out) {
  if (!c) return ZKP_E_ARG;
  KeySlot& slot = which_nn ? c->nn : c->n;
  if (!slot.ready) return fail(c, ZKP_E_STATE, "no key set");
  if (batch < 0 || !a || !b || !out || b_per <= 0) return fail(c, ZKP_E_ARG, "null buffer / bad batch");
  if (batch == 0) return ZKP_OK;
  ZKP_CU(c, cudaSetDevice(c->device));
  const int w = slot.limbs;
  const int nb = (batch + b_per - 1) / b_per;
  ZKP_CU(c, c->in0.ensure((size_t)batch * w * 4));
  ZKP_CU(c, c->in1.ensure((size_t)nb * w * 4));
  ZKP_CU(c, c->out0.ensure((size_t)batch * w * 4));
  ZKP_CU(c, cudaMemcpyAsync(c->in0.p, a, (size_t)batch * w * 4, cudaMemcpyHostToDevice, c->stream));
  ZKP_CU(c, cudaMemcpyAsync(c->in1.p, b, (size_t)nb * w * 4, cudaMemcpyHostToDevice, c->stream));
  {
    ProfScope ps(c, KID_MODMUL, batch);
    ZKP_CU(c, launch_modmul_shared(slot.view(), 0, c->in0.as<uint32_t>(), w, c->in1.as<uint32_t>(), w, b_per,
                                   c->out0.as<uint32_t>(), w, batch, c->stream));
  }
  ZKP_CU(c, cudaMemcpyAsync(out, c->out0.p, (size_t)batch * w * 4, cudaMemcpyDeviceToHost, c->stream));
  ZKP_CU(c, cudaStreamSynchronize(c->stream));
  return ZKP_OK;
}

int zkp_tune(zkp_ctx* c, int knob, int value) {
  if (!c) return ZKP_E_ARG;
  switch (knob) {
    case ZKP_TUNE_ENC_KERNEL:
      if (value < 0 || value > 2) return fail(c, ZKP_E_ARG, "ZKP_TUNE_ENC_KERNEL: 0 (by launch size: K1m, K2h for a few hundred jobs), 1 (K1) or 2 (K1m)");
      c->enc2m_enabled = value != 1;
      c->enc_small_k2h = value == 0;
      return ZKP_OK;
    case ZKP_TUNE_JOBS_SHAPE:
      if (value < 0 || value > 2) return fail(c, ZKP_E_ARG, "ZKP_TUNE_JOBS_SHAPE: 0 (by job count), 1 (wide lanes) or 2 (narrow lanes)");
      c->jobs_shape = value;
      return ZKP_OK;
    case ZKP_TUNE_JOBS_ROWS:
      if (value < 0 || value > 2) return fail(c, ZKP_E_ARG, "ZKP_TUNE_JOBS_ROWS: 0 (default), 1 (single rows) or 2 (pair rows)");
      c->jobs_rows = value;
      return ZKP_OK;
    default:
      return fail(c, ZKP_E_ARG, "unknown tuning knob");
  }
}

int zkp_enc_kernel_launches(const zkp_ctx* c, long long* k1m, long long* k1) {
  if (!c) return ZKP_E_ARG;
  if (k1m) *k1m = c->enc2m_launches;
  if (k1) *k1 = c->k1_launches;
  return ZKP_OK;
}

int zkp_enc_executed_mads(const zkp_ctx* c, double* k1m, double* k1) {
  if (!c || !c->paillier) return ZKP_E_ARG;
  if (k1m) *k1m = c->enc2m_key ? c->enc2m_mads : 0.0;
  if (k1) *k1 = c->k1_mads;
  return ZKP_OK;
}

int zkp_imad_peak(zkp_ctx* c, int variant, double* mads_per_s) {
  if (!c || !mads_per_s) return ZKP_E_ARG;
  ZKP_CU(c, cudaSetDevice(c->device));
  ZKP_CU(c, c->out0.ensure(256));
  const int blocks = c->num_sms * 8;
  double ops = 0;
  cudaEvent_t a, b;
  ZKP_CU(c, cudaEventCreate(&a));
  ZKP_CU(c, cudaEventCreate(&b));
  // warm-up, then a launch long enough (~100+ ms) to settle at sustained clocks
  ZKP_CU(c, launch_imad_peak(variant, blocks, 2000, c->out0.as<uint32_t>(), &ops, c->stream));
  ZKP_CU(c, cudaStreamSynchronize(c->stream));
  ZKP_CU(c, cudaEventRecord(a, c->stream));
  ZKP_CU(c, launch_imad_peak(variant, blocks, 400000, c->out0.as<uint32_t>(), &ops, c->stream));
  ZKP_CU(c, cudaEventRecord(b, c->stream));
  ZKP_CU(c, cudaStreamSynchronize(c->stream));
  float ms = 0;
  ZKP_CU(c, cudaEventElapsedTime(&ms, a, b));
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  *mads_per_s = ops / (ms * 1e-3);
  return ZKP_OK;
}

}  // extern "C"
