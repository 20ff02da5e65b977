#!/usr/bin/env python
"""bench.py -- RangeProofNi proofs+verifies/sec at 2048-bit n on 1..8 B200 (BASELINE.json metric), with every other
BASELINE.json config measured in the same run.

Headline (the driver's line): one "step" = RangeProofNi::prove followed by RangeProofNi::verify over one batch of
synthetic statements under the reference's fixed 2048-bit test key (range_proof_ni.rs:141-145), error factor 128
(range_proof_ni.rs:23), 256-bit ranges (range_proof_ni.rs:133) -- the shape of the reference's Criterion bench
(benches/all.rs:55-71), batched.  Workload = BASELINE.json configs[1]: batch 1024 per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA, sm_100a)
  python bench.py --impl reference ...                            the CPU path (GMP mpz_powm, all host cores)
  python bench.py --config correct_key|sigma|dlog|correct_message one of the other proofs as the printed line

Numbers printed (one JSON line, rank 0):
  value       device-resident throughput: inputs already in HBM, prove_run + verify_run chained on the device,
              CUDA events on the launch stream, max over ranks
  e2e         the same metric through the public C ABI with HOST (pinned) buffers: H2D of the statement and the
              randomness, prove, D2H of the whole proof, H2D of the proof again (the verifier is another party),
              verify, D2H of the verdicts -- all inside the timed region
  roofline    the dominant kernel (K1m, the two-digit Montgomery encryption kernel): EXECUTED multiply-adds per second
              of its device time against the IMAD.WIDE.U32 issue peak measured in this run (frac); the algorithmic count
              of SURVEY.md section 8d is reported beside it (algorithmic_ratio exceeds 1: K1m executes half of it)
  cpu_baseline  oracle/oracle.c (the reference's loops on the reference's own backend, GMP) on a bounded sample
  secondary   configs[2] (NiCorrectKeyProof verify, batch 4096, 3072-bit n, 4096 distinct moduli), configs[4]'s per-GPU
              share (MulProof + VerlinProof verify, 512 + 512, 4096-bit n), configs[0] (ZeroProof prove + verify at
              1024 bits on the CPU), the latency of ONE RangeProofNi, and at N = 8 one step of configs[3]'s share
              (batch 8192 per GPU) -- each with its own value, e2e, roofline, GMP cpu_baseline and clock record
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_BITS = 2048
EF = 128
METRIC = "RangeProofNi proofs+verifies/sec at 2048-bit n"
UNIT = "proofs+verifies/s"
DTYPE = "u32 limbs (32x32+64 IMAD)"


# ---- algorithmic work ---------------------------------------------------
def mm(s):
    return 2 * s * s + s


def modexp_imads(mod_bits, exp_bits):
    return (exp_bits + -(-exp_bits // 5) + 32) * mm(mod_bits // 32)


def k2m_executed(n_bits, exp_bits, nbase=1):
    """IMAD.WIDE one K2m / K2h job executes: fixed 5-bit window in two-digit form (4 S^2 per squaring, 5 S^2 per multiplication);
    per base 30 table products, the wide entry (2) and the step into Montgomery form (1); one squaring chain for all bases of
    a simultaneous-exponentiation job; the final multiplier and the assembly X0 + X1 n."""
    s2 = (n_bits // 32) ** 2
    nwin = -(-exp_bits // 5)
    return s2 * (4 * 5 * (nwin - 1) + 5 * (nbase * (nwin + 30 + 3) + 1) + 1)


ENC_IMADS = modexp_imads(2 * N_BITS, N_BITS)  # 81.9 M at 2048-bit n
# algorithmic HBM bytes of one Enc inside RangeProofNi: base (|n|) + plaintext row in, ciphertext (2|n|) out
ENC_BYTES = N_BITS // 8 + 48 + 2 * N_BITS // 8


def oracle_path():
    for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)


def fixture_key(bits, which=0):
    k = json.load(open(os.path.join(ROOT, "tests", "golden", "keys.json")))[str(bits)][which]
    return int(k["p"]) * int(k["q"])


def test_key():
    """2048-bit n: the reference's fixed test primes (range_proof_ni.rs:141-145).  --n-bits 3072 / 4096 (the other key sizes
    BASELINE.json's target names): the first committed fixture key of that size (tests/golden/keys.json)."""
    oracle_path()
    import zkp_oracle as po  # constants only (the reference's fixed test primes)

    if N_BITS == 2048:
        return po.TEST_P * po.TEST_Q
    return fixture_key(N_BITS)


def set_key_size(bits):
    global N_BITS, METRIC, ENC_IMADS, ENC_BYTES
    N_BITS = bits
    METRIC = f"RangeProofNi proofs+verifies/sec at {bits}-bit n"
    ENC_IMADS = modexp_imads(2 * bits, bits)
    ENC_BYTES = bits // 8 + 48 + 2 * bits // 8


def workload_name(batch):
    key = "reference test key" if N_BITS == 2048 else "committed fixture key"
    return f"RangeProofNi prove+verify, batch={batch} per GPU, {N_BITS}-bit n ({key}), error_factor=128, 256-bit range"


def headline_config(batch):
    """The same dict in both arms (the driver compares them): a static description of the workload only."""
    return {"workload": workload_name(batch), "batch_per_gpu": batch, "n_bits": N_BITS, "error_factor": EF,
            "l2": "working set per step (approx 0.5 GB of bases, ciphertexts and responses) exceeds the 126 MB L2; no explicit flush",
            "sharding": "independent proofs, contiguous shard per rank; NCCL broadcast of n before; after the last step of the e2e region an "
                        "all_gather of the proof bytes (device to device) and of the verdicts; no collective on the modexp path"}


def cpu_sample(n_int, work, cx, sel, threads):
    """prove+verify of the proofs `sel` on the CPU oracle; returns seconds."""
    import c_oracle
    from zk_paillier_b200.native import to_limbs

    nl = work["n_limbs"]
    nlimbs = to_limbs(n_int, nl)
    t0 = time.perf_counter()
    pr = c_oracle.rangeproof_ni_prove(nlimbs, EF, work["range"][sel], work["x"][sel], work["r"][sel], work["w1"][sel],
                                      work["swap"][sel], work["r1"][sel], work["r2"][sel], threads)
    acc, fault, dig, encs = c_oracle.rangeproof_ni_verify(nlimbs, EF, work["range"][sel], cx[sel], pr["c1"], pr["c2"], pr["kind"],
                                                          pr["resp_w"], pr["resp_r"], threads)
    dt = time.perf_counter() - t0
    return dt, pr, acc


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except Exception:
                continue
            for nm, v in zip(names, r[3:7]):
                if v.strip().lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), power_w_max=float(max(pw)), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


def pinned(shape, dtype):
    import torch

    t = torch.empty(tuple(shape), dtype={np.uint32: torch.int32, np.uint8: torch.uint8}[dtype], pin_memory=True)
    return t.numpy().view(dtype)


def pin(a):
    out = pinned(a.shape, a.dtype.type)
    out[...] = a
    return out


class Env:
    """One process per GPU: rank / world from torchrun, the launch stream, one engine context on it."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        import zk_paillier_b200 as zk

        self.torch, self.dist, self.zk = torch, dist, zk
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the engine has no CPU path (use --impl reference for the CPU baseline)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            # the GPU box exports NCCL_DEBUG=VERSION, so NCCL prints its version banner on stdout before rank 0's JSON line; it is
            # left alone (it is the launcher's setting, and the evidence that NCCL initialised)
            dist.init_process_group("nccl", device_id=self.dev)
        self.stream = torch.cuda.Stream(self.dev)  # the library launches on this stream, so torch events on it time the kernels
        torch.cuda.set_stream(self.stream)
        self.ctx = zk.native.Context(self.local, stream=self.stream.cuda_stream)
        self._imad_peak = None
        # solo: measure this rank alone, no collectives (the secondary entries: a failure on one rank must neither hang the others
        # nor cost the headline line; their per-rank rates are combined afterwards by ONE reduction, combine_ranks)
        self.solo = False

    @property
    def nranks(self):
        return 1 if self.solo else self.world

    def barrier(self):
        self.torch.cuda.synchronize(self.dev)
        if self.world > 1 and not self.solo:
            self.dist.barrier()
            self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        if self.world > 1 and not self.solo:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def combine_ranks(self, rates):
        """Per-rank rates of identical shards (weak scaling) -> whole-job rates: world x the slowest rank's (NaN if any rank failed)."""
        t = self.torch.tensor([r if r is not None else float("nan") for r in rates], dtype=self.torch.float64, device=self.dev)
        t = self.torch.where(self.torch.isnan(t), self.torch.full_like(t, -1.0), t)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return [None if v < 0 else v * self.world for v in t.tolist()]

    def imad_peak(self):
        if self._imad_peak is None:
            # variant 1: the carry-chained IMAD.WIDE.U32.X rows the kernels are made of (the plain mad.wide of variant 0 is split by
            # ptxas into IMAD.WIDE + IADD3 pairs and measures the ALU pipe: profiles/r02_imad_probes.json)
            self._imad_peak = self.ctx.imad_peak(1)
        return self._imad_peak

    def sampler(self):
        return ClockSampler(self.local) if self.rank == 0 else None

    def events(self):
        return self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0), "MEASURED_PEAKS.json"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def two_contexts(E, makers, rounds):
    """Double-buffered pipelining through the plain ABI: two engine contexts (each with its own stream and staging buffers) on
    two host threads, so that while one context's kernels run the other uploads its next batch / downloads its last verdicts.
    makers: two callables ctx -> (callable doing one batch).  Returns (seconds, batches done)."""
    ctxs = [E.zk.native.Context(E.local) for _ in range(2)]
    workers = [mk(c) for mk, c in zip(makers, ctxs)]
    for w in workers:
        w()  # warm: allocations, key setup
    E.barrier()
    err = []

    def run(w):
        try:
            for _ in range(rounds):
                w()
        except BaseException as ex:  # pragma: no cover
            err.append(ex)

    t0 = time.perf_counter()
    th = [threading.Thread(target=run, args=(w,)) for w in workers]
    for t in th:
        t.start()
    for t in th:
        t.join()
    E.torch.cuda.synchronize(E.dev)
    dt = time.perf_counter() - t0
    for c in ctxs:
        c.close()
    if err:
        raise err[0]
    return dt, 2 * rounds


# =================================================================================================== headline
def measure_rangeproof(E, batch, steps, warmup, e2e_steps, cpu_seconds, want_cpu=True, want_gather=True):
    from zk_paillier_b200 import sharding, workload
    from zk_paillier_b200.native import KID_MODEXP_SHARED, KID_MODMUL, KID_OTHER, KID_SHA, _p32, _p8, to_limbs

    ctx, torch, rank, world, dev, stream = E.ctx, E.torch, E.rank, E.world, E.dev, E.stream
    nl = N_BITS // 32
    # public key: rank 0 owns it, one NCCL broadcast (the only collective before the hot path)
    n_limbs_arr = sharding.broadcast_key(to_limbs(test_key(), nl) if rank == 0 else np.zeros(1, np.uint32), dev)
    n_int = int.from_bytes(n_limbs_arr.tobytes(), "little")
    ctx.set_key(n_limbs_arr)

    # this rank's shard of independent statements (weak scaling: `batch` proofs per GPU)
    work = workload.rangeproof_batch(n_int, batch, ef=EF, seed=workload.DEFAULT_SEED + rank, reject_every=100)
    wl = work["w_limbs"]
    cx = ctx.paillier_enc(work["x_n"], work["r"])  # statement ciphertexts c = Enc(x, r): not on the measured path

    # ---------------- device-resident steps (value) ----------------
    ctx.rp_prove_stage(EF, work["range"], work["x"], work["r"], work["w1"], work["swap"], work["r1"], work["r2"])
    ctx.rp_prove_run()
    ctx.rp_verify_stage_from_prove(cx)

    def step_device():
        ctx.rp_prove_run()
        ctx.rp_verify_run()

    for _ in range(warmup):
        step_device()
    ctx.profile_enable(True)
    ctx.profile_reset()
    E.barrier()
    sampler = E.sampler()
    e0, e1 = E.events()
    e0.record(stream)
    for _ in range(steps):
        step_device()
    e1.record(stream)
    E.barrier()
    clocks = sampler.stop() if sampler else None
    ms = e0.elapsed_time(e1)
    prof = {k: ctx.profile_get(k) for k in (KID_MODEXP_SHARED, KID_MODMUL, KID_SHA, KID_OTHER)}
    ctx.profile_enable(False)
    ctx.profile_reset()
    accept, fault, digest = ctx.rp_verify_fetch()
    expect = np.array([0 if b % 100 == 99 else 1 for b in range(batch)], np.uint8)
    if not np.array_equal(accept, expect) or fault.any():
        raise SystemExit("bench.py: wrong verdicts from the device path")
    enc_verify = ctx.rp_verify_enc_count()

    # ---------------- end-to-end steps through the C ABI with host buffers (e2e) ----------------
    host_in = {k: pin(work[k]) for k in ("range", "x", "r", "w1", "swap", "r1", "r2")}
    cx_h = pin(cx)
    lib, h = ctx._lib, ctx._h
    nnl = 2 * nl
    out = {"c1": pinned((batch, EF, nnl), np.uint32), "c2": pinned((batch, EF, nnl), np.uint32), "digest": pinned((batch, 32), np.uint8),
           "kind": pinned((batch, EF), np.uint8), "resp_w": pinned((batch, EF, 2, wl), np.uint32), "resp_r": pinned((batch, EF, 2, nl), np.uint32)}
    acc_h, fault_h, dig_h = pinned((batch,), np.uint8), pinned((batch,), np.uint8), pinned((batch, 32), np.uint8)

    def step_e2e():
        ctx._ck(lib.zkp_rangeproof_ni_prove(h, batch, EF, wl, _p32(host_in["range"]), _p32(host_in["x"]), _p32(host_in["r"]),
                                            _p32(host_in["w1"]), _p8(host_in["swap"]), _p32(host_in["r1"]), _p32(host_in["r2"]),
                                            _p32(out["c1"]), _p32(out["c2"]), _p8(out["digest"]), _p8(out["kind"]), _p32(out["resp_w"]),
                                            _p32(out["resp_r"])))
        ctx._ck(lib.zkp_rangeproof_ni_verify(h, batch, EF, wl, _p32(host_in["range"]), _p32(cx_h), _p32(out["c1"]), _p32(out["c2"]),
                                             _p8(out["kind"]), _p32(out["resp_w"]), _p32(out["resp_r"]), _p8(acc_h), _p8(fault_h), _p8(dig_h)))

    h2d = sum(host_in[k].nbytes for k in host_in) + cx_h.nbytes + host_in["range"].nbytes + sum(out[k].nbytes for k in ("c1", "c2", "kind", "resp_w", "resp_r"))
    d2h = sum(v.nbytes for v in out.values()) + acc_h.nbytes + fault_h.nbytes + dig_h.nbytes
    e2e_ms = float("nan")
    if e2e_steps > 0:
        step_e2e()
        E.barrier()
        g0, g1 = E.events()
        g0.record(stream)
        for _ in range(e2e_steps):
            step_e2e()
        if world > 1 and want_gather:  # final gather over NVLink: the proofs (device to device, 0.2 MB each), then verdicts + challenge hashes (33 B per proof)
            allproofs = sharding.gather_proof_bytes([out[k] for k in ("c1", "c2", "kind", "resp_w", "resp_r")], dev)
            torch.cuda.synchronize(dev)
            assert allproofs.shape[:2] == (world, batch)
            del allproofs
            allrec = sharding.gather_records(np.concatenate([acc_h[:, None], dig_h], axis=1), dev, counts=[batch] * world)
            assert allrec.shape == (world * batch, 33)
        g1.record(stream)
        E.barrier()
        e2e_ms = g0.elapsed_time(g1)  # the ABI calls are host-synchronous, so the event pair brackets copies + kernels + host gaps
        if not np.array_equal(acc_h, expect):
            raise SystemExit("bench.py: wrong verdicts from the e2e path")

    ms, e2e_ms = E.max_over_ranks([ms, e2e_ms])
    res = {"ms": ms, "e2e_ms": e2e_ms, "value": world * batch * steps / (ms * 1e-3),
           "e2e_value": world * batch * e2e_steps / (e2e_ms * 1e-3) if e2e_steps > 0 else None,
           "h2d": int(h2d), "d2h": int(d2h), "clocks": clocks, "enc_verify": int(enc_verify)}
    if rank != 0:
        return res

    # ---------------- roofline of the dominant kernel ----------------
    k1_ms, k1_launches, k1_units = prof[KID_MODEXP_SHARED]
    imad_peak = E.imad_peak()
    achieved = k1_units * ENC_IMADS / (k1_ms * 1e-3)
    used = ctx.enc_kernel_launches()
    which = "k1m" if used["k1m"] and not used["k1"] else ("k1" if used["k1"] and not used["k1m"] else "mixed")
    exec_mads = ctx.enc_executed_mads().get(which, 0.0)
    executed = k1_units * exec_mads / (k1_ms * 1e-3)
    shape = {1024: "<4,8>", 2048: "<8,8>", 3072: "<8,12>", 4096: "<16,8>"}.get(N_BITS, "")
    kernel_name = {"k1m": f"enc2m_kernel{shape} (K1m, two-digit Montgomery form)", "k1": "modexp_shared_kernel (K1)"}.get(which, which)
    hbm, hbm_src = hbm_peak()
    traffic = None
    try:  # DRAM bytes of K1 from the committed ncu capture, scaled to this run's Enc per launch
        tr = json.load(open(os.path.join(ROOT, "profiles", "k1m_traffic.json" if which == "k1m" else "k1_traffic.json")))
        if N_BITS == 2048:  # the capture is of the 2048-bit kernel
            traffic = tr["dram_bytes_per_enc"] * k1_units / max(k1_launches, 1)
    except Exception:
        pass
    hbm_ach = k1_units * ENC_BYTES / (k1_ms * 1e-3) / 1e9
    res["roofline"] = {
        "bound": "imad", "kernel": kernel_name, "achieved": executed / 1e12, "peak": imad_peak / 1e12, "unit": "T IMAD.WIDE.U32/s",
        "frac": executed / imad_peak, "traffic": traffic,
        "executed_imads_per_enc": exec_mads, "alg_imads_per_enc": ENC_IMADS, "algorithmic": achieved / 1e12,
        "algorithmic_ratio": achieved / imad_peak,
        "frac_note": "achieved / frac = multiply-adds the kernel EXECUTES per second of its device time against the IMAD.WIDE.U32 issue peak measured in "
                     "this run; algorithmic / algorithmic_ratio = the SURVEY.md 8d count (fixed-window schoolbook CIOS modulo n^2) per second over the same "
                     "peak - above 1 because K1m executes about half of it (two-digit base-n Montgomery form, 6-bit sliding window)",
        "traffic_note": "DRAM bytes per launch = ncu dram_bytes per Enc (profiles/k1m_traffic.json or k1_traffic.json) x Enc per launch",
        "peak_source": "IMAD.WIDE.U32.X issue-rate microbenchmark (register-only carry-chained rows) measured in this run; MEASURED_PEAKS.json has no "
                       "integer entry; 4.6 cycles per warp instruction per sub-partition on B200 (profiles/r02_imad_probes.json)",
        "k1_launches": int(k1_launches), "k1_ms_avg": k1_ms / max(k1_launches, 1), "k1_share_of_step": k1_ms / ms}
    res["roofline_hbm"] = {"bound": "hbm", "achieved": hbm_ach, "peak": hbm, "unit": "GB/s", "frac": hbm_ach / hbm, "traffic": traffic,
                           "alg_bytes_per_enc": ENC_BYTES, "peak_source": hbm_src}
    res["kernel_ms"] = {"modexp_shared": k1_ms, "modmul": prof[KID_MODMUL][0], "sha256_transcript": prof[KID_SHA][0], "other": prof[KID_OTHER][0]}
    res["launches"] = int(sum(p[1] for p in prof.values()))

    # ---------------- CPU baseline on a bounded sample (rank 0, N=1 only) ----------------
    res["cpu"] = None
    if world == 1 and want_cpu and e2e_steps > 0:
        oracle_path()
        import c_oracle

        cores = c_oracle.hw_threads()
        t1, pr1, acc1 = cpu_sample(n_int, work, cx, np.arange(1), cores)
        m = int(max(1, min(64, batch, cpu_seconds / max(t1, 1e-3))))
        dt, pr, acc = cpu_sample(n_int, work, cx, np.arange(m), cores)
        same = all(np.array_equal(pr[k], out[k][:m]) for k in ("c1", "c2", "digest", "kind", "resp_w", "resp_r")) and np.array_equal(acc, acc_h[:m])
        if not same:
            raise SystemExit("bench.py: CUDA outputs differ from the CPU oracle on the baseline sample")
        res["cpu"] = {"value": m / dt, "unit": UNIT, "cores": cores, "kind": "port",
                      "sample": f"{m} of the {batch} proofs (prove+verify, {m * 256} + {int((pr['kind'] == 0).sum()) + m * EF} Enc), GMP {c_oracle.gmp_version()} mpz_powm, outputs byte-identical to the GPU's"}
    return res


def measure_latency(E, reps=5):
    """ONE RangeProofNi through the one-shot ABI (the reference's call shape, range_proof_ni.rs:47-107): prove, then verify."""
    from zk_paillier_b200 import workload
    from zk_paillier_b200.native import to_limbs

    ctx = E.ctx
    n_int = test_key()
    ctx.set_key(to_limbs(n_int, N_BITS // 32))
    work = workload.rangeproof_batch(n_int, 1, ef=EF, seed=workload.DEFAULT_SEED + 77)
    cx = ctx.paillier_enc(work["x_n"], work["r"])
    args = (EF, work["range"], work["x"], work["r"], work["w1"], work["swap"], work["r1"], work["r2"])
    tp, tv = [], []
    for i in range(reps + 2):
        t0 = time.perf_counter()
        pr = ctx.rangeproof_ni_prove(*args)
        t1 = time.perf_counter()
        acc, fault, dig = ctx.rangeproof_ni_verify(EF, work["range"], cx, pr["c1"], pr["c2"], pr["kind"], pr["resp_w"], pr["resp_r"])
        t2 = time.perf_counter()
        assert acc.tolist() == [1]
        if i >= 2:
            tp.append(t1 - t0); tv.append(t2 - t1)
    out = {"metric": "latency of one RangeProofNi at 2048-bit n (batch = 1 through the one-shot ABI, host buffers)", "unit": "ms", "higher_is_better": False,
           "prove_ms": float(np.median(tp)) * 1e3, "verify_ms": float(np.median(tv)) * 1e3, "value": float(np.median(tp) + np.median(tv)) * 1e3,
           "config": {"workload": "one RangeProofNi prove, then verify, 2048-bit n (reference test key), error_factor=128; median of 5 after 2 warm-up calls"}}
    if E.rank == 0 and E.world == 1:
        oracle_path()
        import c_oracle

        cores = c_oracle.hw_threads()
        ts = [cpu_sample(n_int, work, cx, np.arange(1), cores)[0] for _ in range(3)]
        out["cpu_baseline"] = {"value": float(np.median(ts)) * 1e3, "unit": "ms", "cores": cores, "kind": "port",
                               "sample": "the same single proof, prove + verify, GMP mpz_powm over the security parameter on all host threads (median of 3)"}
    return out


# =================================================================================================== configs[2]
def measure_correct_key(E, batch=4096, bits=3072, steps=5, warmup=3, e2e_rounds=3, want_cpu=True):
    """NiCorrectKeyProof::verify (correct_key_ni.rs:73-100), a DISTINCT modulus per proof (device keygen), salt 'Zen Go X'."""
    from zk_paillier_b200 import workload
    from zk_paillier_b200.native import KID_MODEXP_VAR

    ctx, stream = E.ctx, E.stream
    salt = b"Zen Go X"
    t0 = time.perf_counter()
    work = workload.correct_key_distinct(bits, batch, salt, seed=workload.DEFAULT_SEED + 1000 * E.rank, bad_every=64, device=E.local)
    gen_s = time.perf_counter() - t0
    assert len({r.tobytes() for r in work["n"]}) == batch, "moduli are not distinct"
    expect = [0 if b % 64 == 63 else 1 for b in range(batch)]
    ctx.ck_verify_stage(work["n"], work["sigma"], salt)
    for _ in range(warmup):
        ctx.ck_verify_run()
    ctx.profile_enable(True); ctx.profile_reset()
    E.barrier()
    sampler = E.sampler()
    e0, e1 = E.events()
    e0.record(stream)
    for _ in range(steps):
        ctx.ck_verify_run()
    e1.record(stream)
    E.barrier()
    ms = e0.elapsed_time(e1)
    k2_ms, k2_n, k2_units = ctx.profile_get(KID_MODEXP_VAR)
    ctx.profile_enable(False); ctx.profile_reset()
    assert ctx.ck_verify_fetch().tolist() == expect
    # e2e: the one-shot call from pinned host buffers; two contexts pipeline upload / kernels / download
    n_h, s_h = pin(work["n"]), pin(work["sigma"])

    def maker(c):
        def one():
            assert c.correct_key_ni_verify(n_h, s_h, salt).tolist() == expect
        return one

    e2e_s, e2e_batches = two_contexts(E, [maker, maker], e2e_rounds)
    clocks = sampler.stop() if sampler else None
    ms, e2e_s = E.max_over_ranks([ms, e2e_s])
    per = modexp_imads(bits, bits)
    out = {"metric": f"NiCorrectKeyProof verifies/sec at {bits}-bit n", "unit": "verifies/s", "value": E.nranks * batch * steps / (ms * 1e-3),
           "ms_per_step": ms / steps, "steps": steps, "warmup": warmup, "n_gpus": E.nranks, "higher_is_better": True, "scaling": "weak", "dtype": DTYPE,
           "config": {"workload": f"NiCorrectKeyProof verify, batch={batch} per GPU, {bits}-bit n, {batch} distinct moduli (device keygen), salt 'Zen Go X', 1/64 bad proofs",
                      "batch_per_gpu": batch, "n_bits": bits, "distinct_moduli": batch},
           "e2e": {"value": E.nranks * batch * e2e_batches / e2e_s, "unit": "verifies/s", "h2d_bytes_per_step": int(work["n"].nbytes + work["sigma"].nbytes),
                   "d2h_bytes_per_step": batch, "steps": e2e_batches,
                   "how": "zkp_correct_key_ni_verify from pinned host buffers on two contexts / two host threads (double-buffered: one context's copies under "
                          "the other's kernels); host wall clock around synchronised calls"},
           "roofline": {"bound": "imad", "kernel": "modexp_var_kernel<8,12> (K2)", "achieved": k2_units * per / (k2_ms * 1e-3) / 1e12, "peak": E.imad_peak() / 1e12,
                        "unit": "T IMAD.WIDE.U32/s", "frac": k2_units * per / (k2_ms * 1e-3) / E.imad_peak(), "traffic": None, "imads_per_modexp": per,
                        "frac_note": "K2 runs the fixed-window schoolbook CIOS that the SURVEY.md 8d formula counts: executed = algorithmic",
                        "k2_share_of_step": k2_ms / ms},
           "clocks": clocks, "gpu_launches": int(k2_n), "workload_generation_s": gen_s, "cpu_baseline": None}
    if E.rank == 0 and E.world == 1 and want_cpu:
        out["cpu_baseline"] = cpu_correct_key(work["n"], work["sigma"], salt, expect)
    return out


def cpu_correct_key(n, sigma, salt, expect, seconds=8.0):
    oracle_path()
    import c_oracle

    cores = c_oracle.hw_threads()
    m0 = min(len(n), 2 * cores)
    t0 = time.perf_counter()
    c_oracle.correct_key_ni_verify(n[:m0], sigma[:m0], salt, cores)
    t1 = time.perf_counter() - t0
    m = int(max(m0, min(len(n), seconds / max(t1 / m0, 1e-6))))
    t0 = time.perf_counter()
    acc_c, _ = c_oracle.correct_key_ni_verify(n[:m], sigma[:m], salt, cores)
    dt = time.perf_counter() - t0
    if expect is not None and acc_c.tolist() != list(expect[:m]):
        raise SystemExit("bench.py: the CPU oracle and the GPU disagree on NiCorrectKeyProof verdicts")
    return {"value": m / dt, "unit": "verifies/s", "cores": cores, "kind": "port",
            "sample": f"{m} of the proofs, GMP {c_oracle.gmp_version()} mpz_powm on {cores} threads" + (", same verdicts as the GPU" if expect is not None else "")}


# =================================================================================================== configs[4]
def sigma_workload(ctx, bits, B, seed=5):
    """MulProof and VerlinProof statements + honest proofs under one key, built on the device (not on the measured path)."""
    from zk_paillier_b200 import workload
    from zk_paillier_b200.native import ints_to_limbs, limbs_to_ints, to_limbs

    n = fixture_key(bits)
    nl, nnl = bits // 32, bits // 16
    ctx.set_key(to_limbs(n, nl))
    g = np.random.Generator(np.random.PCG64(seed))
    rows = lambda: workload._rand_limbs_below_pow2(g, (B,), nl, bits - 1)
    a, b = rows(), rows()
    c = ints_to_limbs([x * y % n for x, y in zip(limbs_to_ints(a), limbs_to_ints(b))], nl)
    r_a, r_b, r_c, d, r_d = (rows() | 1 for _ in range(5))
    e_a, e_b, e_c = ctx.paillier_enc(a, r_a), ctx.paillier_enc(b, r_b), ctx.paillier_enc(c, r_c)
    f, z1, z2, e_d, e_db, fault = ctx.mul_prove(a, b, r_a, r_b, r_c, e_a, e_b, e_c, d, r_d)
    assert not fault.any()
    x, xp, xdp, r_x = rows(), rows(), rows(), rows() | 1
    cc, cp = ctx.paillier_enc(rows(), rows() | 1), ctx.paillier_enc(rows(), rows() | 1)
    pad = lambda v: np.concatenate([v, np.zeros((B, nnl - nl), np.uint32)], axis=1)
    nn_rows = to_limbs(n * n, nnl)[None, :]
    phi_x = ctx.modmul(ctx.modmul(ctx.modexp_var(cc, pad(x), nn_rows, exp_per=1, mod_per=B, exp_bits=bits),
                                  ctx.modexp_var(cp, pad(xp), nn_rows, exp_per=1, mod_per=B, exp_bits=bits)), ctx.paillier_enc(xdp, r_x))
    phi_a, z, zp, zdp, r_z = ctx.verlin_prove(x, xp, xdp, r_x, cc, cp, phi_x, rows(), rows(), rows(), rows() | 1)
    return {"n": n, "mul": tuple(pin(v) for v in (e_a, e_b, e_c, f, z1, z2, e_d, e_db)), "verlin": tuple(pin(v) for v in (cc, cp, phi_x, phi_a, z, zp, zdp, r_z))}


def measure_sigma(E, B=512, bits=4096, steps=5, warmup=3, want_cpu=True, jobs_shape=0):
    """MulProof::verify (multiplication_proof.rs:108-145) x B + VerlinProof::verify (verlin_proof.rs:101-134) x B under one key."""
    from zk_paillier_b200.native import KID_CALL, KID_MODEXP_VAR, TUNE_JOBS_SHAPE, to_limbs

    ctx = E.ctx
    ctx.tune(TUNE_JOBS_SHAPE, jobs_shape)
    w = sigma_workload(ctx, bits, B, seed=5 + E.rank)
    nl = bits // 32

    def step(c):
        acc1, flt = c.mul_verify(*w["mul"])
        acc2 = c.verlin_verify(*w["verlin"])
        if not (acc1.all() and acc2.all()) or flt.any():
            raise SystemExit("bench.py: the device rejects honest MulProof / VerlinProof proofs")

    # ---- value: device spans of the calls (first kernel to last kernel, copies excluded), one context, calls back to back
    for _ in range(warmup):
        step(ctx)
    ctx.profile_enable(True); ctx.profile_reset()
    E.barrier()
    sampler = E.sampler()
    t0 = time.perf_counter()
    for _ in range(steps):
        step(ctx)
    E.torch.cuda.synchronize(E.dev)
    seq_s = time.perf_counter() - t0
    span_ms, span_n, _ = ctx.profile_get(KID_CALL)
    k2h_ms, k2h_n, k2h_jobs = ctx.profile_get(KID_MODEXP_VAR)
    ctx.profile_enable(False); ctx.profile_reset()
    ctx.tune(TUNE_JOBS_SHAPE, 0)

    # ---- e2e: the mixed batch as a caller would run it - the MulProof verifies on one context, the VerlinProof verifies on another,
    # two host threads: the modexps of both calls are resident together (3 072 long jobs instead of 1 536 at a time)
    def maker(kind):
        def make(c):
            c.set_key(to_limbs(w["n"], nl))
            c.tune(TUNE_JOBS_SHAPE, jobs_shape)
            if kind == "mul":
                def one():
                    acc1, flt = c.mul_verify(*w["mul"])
                    assert acc1.all() and not flt.any()
            else:
                def one():
                    assert c.verlin_verify(*w["verlin"]).all()
            return one
        return make

    conc_s, _ = two_contexts(E, [maker("mul"), maker("verlin")], steps)
    clocks = sampler.stop() if sampler else None
    span_ms, seq_s, conc_s = E.max_over_ranks([span_ms, seq_s, conc_s])
    mul_alg = 3 * modexp_imads(2 * bits, bits) + 2 * modexp_imads(2 * bits, 256)
    ver_alg = modexp_imads(2 * bits, 256) + 2 * modexp_imads(2 * bits, bits + 256) + modexp_imads(2 * bits, bits)
    zbits = 32 * (nl + 12)
    mul_exe = 3 * k2m_executed(bits, bits) + 2 * k2m_executed(bits, 256)
    ver_exe = k2m_executed(bits, 256) + k2m_executed(bits, zbits, nbase=3)  # gen_phi is ONE three-base job (Straus)
    best_s = min(seq_s, conc_s)
    bytes_in = sum(v.nbytes for v in w["mul"]) + sum(v.nbytes for v in w["verlin"])
    peak = E.imad_peak()
    out = {"metric": f"MulProof+VerlinProof verifies/sec at {bits}-bit n", "unit": "verifies/s", "value": E.nranks * 2 * B * steps / (span_ms * 1e-3),
           "ms_per_step": span_ms / steps, "steps": steps, "warmup": warmup, "n_gpus": E.nranks, "higher_is_better": True, "scaling": "weak", "dtype": DTYPE,
           "value_how": "device spans (CUDA events, first kernel to last kernel of each zkp_mul_verify / zkp_verlin_verify call; copies excluded), calls back to back on one context",
           "config": {"workload": f"MulProof verify x{B} + VerlinProof verify x{B} per GPU, {bits}-bit n ({2 * bits}-bit modulus), one key (the per-GPU share of BASELINE configs[4])",
                      "batch_per_gpu": 2 * B, "n_bits": bits},
           "e2e": {"value": E.nranks * 2 * B * steps / best_s, "unit": "verifies/s", "h2d_bytes_per_step": int(bytes_in), "d2h_bytes_per_step": 3 * B, "steps": steps,
                   "sequential": E.nranks * 2 * B * steps / seq_s, "two_contexts": E.nranks * 2 * B * steps / conc_s,
                   "how": "one-shot calls from pinned host buffers, host wall clock around synchronised calls; `sequential` = both calls on one context, "
                          "`two_contexts` = the MulProof batch and the VerlinProof batch on two contexts / host threads at the same time; value = the faster"},
           "roofline": {"bound": "imad", "kernel": "modexp2m_jobs_kernel (K2h: every modexp of a call in one or two phased launches; gen_phi as one three-base simultaneous exponentiation)",
                        "achieved": B * (mul_exe + ver_exe) * steps / (span_ms * 1e-3) / 1e12, "peak": peak / 1e12, "unit": "T IMAD.WIDE.U32/s",
                        "frac": B * (mul_exe + ver_exe) * steps / (span_ms * 1e-3) / peak, "traffic": None,
                        "algorithmic_ratio": B * (mul_alg + ver_alg) * steps / (span_ms * 1e-3) / peak,
                        "frac_two_contexts": B * (mul_exe + ver_exe) * steps / conc_s / peak,
                        "frac_note": "executed multiply-adds of the two-digit fixed-window kernels over the device span of the calls (frac_two_contexts: over the wall time "
                                     "of the two-context run, copies included); algorithmic_ratio counts the SURVEY.md 8d formula instead",
                        "k2h_launches": int(k2h_n), "k2h_ms_summed": k2h_ms},
           "clocks": clocks, "gpu_launches": int(k2h_n), "cpu_baseline": None}
    if E.rank == 0 and E.world == 1 and want_cpu:
        oracle_path()
        import c_oracle

        cores = c_oracle.hw_threads()
        m = min(B, 4 * cores)
        nlimbs = to_limbs(w["n"], nl)
        t0 = time.perf_counter()
        v1 = c_oracle.mul_verify(nlimbs, *[v[:m] for v in w["mul"]], cores)
        v2 = c_oracle.verlin_verify(nlimbs, *[v[:m] for v in w["verlin"]], cores)
        cpu_s = time.perf_counter() - t0
        if not (v1 == 1).all() or not (v2 == 1).all():
            raise SystemExit("bench.py: the CPU oracle rejects proofs the GPU accepted")
        out["cpu_baseline"] = {"value": 2 * m / cpu_s, "unit": "verifies/s", "cores": cores, "kind": "port",
                               "sample": f"{m} MulProof + {m} VerlinProof verifies of the batch, GMP {c_oracle.gmp_version()} mpz_powm, same verdicts as the GPU"}
    return out


# =================================================================================================== configs[0]
def zero_workload(bits, B, seed=3):
    from zk_paillier_b200 import workload
    from zk_paillier_b200.native import to_limbs

    oracle_path()
    import c_oracle

    n = fixture_key(bits)
    nl = bits // 32
    rng = np.random.Generator(np.random.PCG64(seed))
    r = workload._rand_limbs_below_pow2(rng, (B,), nl, bits - 1) | 1
    rp = workload._rand_limbs_below_pow2(rng, (B,), nl, bits - 1) | 1
    c = c_oracle.paillier_enc(to_limbs(n, nl), np.zeros((B, 4), np.uint32), r, 0)
    return n, nl, r, rp, c


def measure_zero_cpu(seconds=4.0):
    """BASELINE configs[0]: ZeroProof prove + verify (zero_enc_proof.rs:44-94), 1024-bit n, on the CPU (the GMP restatement)."""
    oracle_path()
    import c_oracle
    from zk_paillier_b200.native import to_limbs

    cores = c_oracle.hw_threads()
    n, nl, r, rp, c = zero_workload(1024, 64 * cores)
    nlimbs = to_limbs(n, nl)

    def run(m, threads):
        t0 = time.perf_counter()
        z, a = c_oracle.zero_prove(nlimbs, r[:m], c[:m], rp[:m], threads)
        acc = c_oracle.zero_verify(nlimbs, c[:m], z, a, threads)
        dt = time.perf_counter() - t0
        assert (np.asarray(acc) == 1).all()
        return dt

    run(1, 1)
    single = float(np.median([run(1, 1) for _ in range(9)]))
    t = run(2 * cores, cores)
    m = int(max(2 * cores, min(len(r), seconds / max(t / (2 * cores), 1e-6))))
    dt = run(m, cores)
    return {"metric": "ZeroProof proofs+verifies/sec at 1024-bit n (CPU)", "unit": "proofs+verifies/s", "value": m / dt, "higher_is_better": True,
            "single_proof_ms": single * 1e3, "cores": cores, "kind": "port",
            "config": {"workload": "ZeroProof prove + verify, 1024-bit n (committed fixture key): one proof on one core, and a batch on all host threads; "
                                   f"GMP {c_oracle.gmp_version()} mpz_powm (BASELINE configs[0])"},
            "sample": f"{m} proofs on {cores} threads; one proof alone (1 thread, median of 9): {single * 1e3:.2f} ms"}


def measure_zero_gpu(E, B=8192):
    """The same proof on the GPU, for the ratio: batch prove + verify through the one-shot ABI, and one proof alone."""
    from zk_paillier_b200.native import to_limbs

    ctx = E.ctx
    n, nl, r, rp, c = zero_workload(1024, B)
    ctx.set_key(to_limbs(n, nl))
    r_h, rp_h, c_h = pin(r), pin(rp), pin(c)

    def run(m):
        t0 = time.perf_counter()
        z, a = ctx.zero_prove(r_h[:m], c_h[:m], rp_h[:m])
        acc = ctx.zero_verify(c_h[:m], z, a)
        dt = time.perf_counter() - t0
        assert acc.all()
        return dt

    run(B); run(1)
    single = float(np.median([run(1) for _ in range(9)]))
    dt = min(run(B) for _ in range(3))
    return {"value": B / dt, "unit": "proofs+verifies/s", "batch": B, "single_proof_ms": single * 1e3,
            "how": "zkp_zero_prove + zkp_zero_verify from pinned host buffers (host wall clock, best of 3); single proof: batch = 1, median of 9"}


# =================================================================================================== lines
def run_b200(args):
    E = Env()
    if args.config == "rangeproof":
        line = headline_line(E, args)
    elif args.config == "correct_key":
        line = measure_correct_key(E, batch=args.batch if args.batch != 1024 else 4096, steps=max(args.steps, 1), warmup=max(args.warmup, 3), want_cpu=not args.no_cpu)
        line.update(vs_baseline=None, data="synthetic (device keygen, seeded)")  # (collective timing: barrier + max over ranks inside)
    elif args.config == "sigma":
        line = measure_sigma(E, B=args.batch if args.batch != 1024 else 512, steps=max(args.steps, 1), warmup=max(args.warmup, 3), want_cpu=not args.no_cpu,
                             jobs_shape=args.jobs_shape)
        line.update(vs_baseline=None, data="synthetic (seeded PCG64; committed fixture key)")
    else:
        line = run_other(E, args)
    if E.rank == 0:
        print(json.dumps(line))
    E.close()


def headline_line(E, args):
    batch = args.batch
    r = measure_rangeproof(E, batch, args.steps, args.warmup, args.e2e_steps, args.cpu_seconds, want_cpu=not args.no_cpu)
    secondary = None
    if not args.no_secondary and N_BITS == 2048:
        secondary = run_secondary(E, args, batch)
    if E.rank != 0:
        return None
    return {
        "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": E.world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms"] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE,
        "data": "synthetic (seeded PCG64; " + ("reference test key" if N_BITS == 2048 else "committed fixture key") + "; 1% reject-path statements)",
        "config": headline_config(batch),
        "e2e": {"value": r["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"], "steps": args.e2e_steps},
        "gpu_launches": r["launches"],
        "roofline": r["roofline"], "roofline_hbm": r["roofline_hbm"], "kernel_ms": r["kernel_ms"],
        "clocks": r["clocks"], "cpu_baseline": r["cpu"],
        "detail": {"enc_per_step_per_gpu": int(2 * batch * EF + r["enc_verify"]), "e2e_steps": args.e2e_steps,
                   "e2e_note": f"e2e is timed over {args.e2e_steps} steps (value over {args.steps}); a step is seconds long, so the host buffers' copies are <1 % of it"},
        "secondary": secondary,
    }


def run_secondary(E, args, batch):
    """The other BASELINE configs.  Every entry is measured by each rank ALONE (E.solo: no collectives inside, a failure is recorded
    in the entry instead of taking the line - or the other ranks - down); one reduction at the end turns the per-rank rates of the
    identical shards into whole-job rates (world x the slowest rank)."""
    import traceback

    sec = {}
    E.solo = True

    def guarded(name, fn):
        try:
            sec[name] = fn()
        except BaseException as ex:  # SystemExit from a failed check included
            sec[name] = {"error": f"{type(ex).__name__}: {ex}", "trace": traceback.format_exc(limit=3)}

    guarded("correct_key_3072", lambda: measure_correct_key(E, want_cpu=not args.no_cpu))
    guarded("mul_verlin_4096", lambda: measure_sigma(E, want_cpu=not args.no_cpu))
    if E.world == 1:
        guarded("latency_one_proof", lambda: measure_latency(E))
        if not args.no_cpu:
            def zero():
                z = measure_zero_cpu()
                z["b200"] = measure_zero_gpu(E)
                return z
            guarded("zero_1024_cpu", zero)
    E.solo = False
    if E.world > 1:
        keys = [("correct_key_3072", "value"), ("correct_key_3072", "e2e"), ("mul_verlin_4096", "value"), ("mul_verlin_4096", "e2e")]

        def local(name, what):
            ent = sec.get(name, {})
            if "error" in ent:
                return None
            return ent["value"] if what == "value" else ent["e2e"]["value"]

        agg = E.combine_ranks([local(n, w) for n, w in keys])
        for (name, what), v in zip(keys, agg):
            ent = sec.get(name, {})
            if "error" in ent:
                continue
            if v is None:
                ent["error"] = "another rank failed this entry"
            elif what == "value":
                ent["value"], ent["n_gpus"] = v, E.world
            else:
                ent["e2e"]["value"] = v
            ent["aggregation"] = f"{E.world} ranks, identical shards: whole-job rate = {E.world} x the slowest rank's rate"
    if E.world == 8 and batch != 8192:
        # configs[3] as stated: 65 536 proofs over 8 GPUs = 8 192 per GPU; one timed step (22 s) after one untimed, kernels already warm
        big = measure_rangeproof(E, 8192, 1, 1, 0, 0, want_cpu=False, want_gather=False)
        sec["rangeproof_65536_over_8"] = {"metric": METRIC, "unit": UNIT, "value": big["value"], "ms_per_step": big["ms"], "steps": 1, "warmup": 1,
                                          "n_gpus": E.world, "clocks": big["clocks"],
                                          "config": {"workload": workload_name(8192), "batch_per_gpu": 8192, "total_proofs": 8192 * E.world}}
    return sec


def run_reference(args):
    """The reference's CPU path (its loops restated on its own backend, GMP) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    oracle_path()
    import c_oracle
    from zk_paillier_b200 import workload
    from zk_paillier_b200.native import to_limbs

    n_int = test_key()
    cores = c_oracle.hw_threads()
    batch = args.batch
    work = workload.rangeproof_batch(n_int, min(batch, 64), ef=EF, seed=workload.DEFAULT_SEED, reject_every=100)
    cx = c_oracle.paillier_enc(to_limbs(n_int, work["n_limbs"]), work["x_n"], work["r"], cores)
    t1, _, _ = cpu_sample(n_int, work, cx, np.arange(1), cores)
    budget = args.ref_seconds / max(1, args.steps + args.warmup)
    m = int(max(1, min(len(work["range"]), budget / max(t1, 1e-3))))
    sel = np.arange(m)
    for _ in range(args.warmup):
        cpu_sample(n_int, work, cx, sel, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_sample(n_int, work, cx, sel, cores)
    dt = time.perf_counter() - t0
    value = m * args.steps / dt
    sample = (f"each step = {m} of the {batch} proofs of the workload (prove+verify; the per-proof CPU cost does not depend on the batch), "
              f"GMP {c_oracle.gmp_version()} mpz_powm on {cores} threads")
    secondary = None
    if not args.no_secondary and N_BITS == 2048:
        import zkp_oracle as po
        from util import keys

        salt = b"Zen Go X"
        ks = keys(3072)
        wk = workload.correct_key_batch(ks, 4 * len(ks), salt, lambda p, q, s: po.NiCorrectKeyProof.proof(p, q, s).sigma_vec, 96)
        secondary = {"correct_key_3072": cpu_correct_key(wk["n"], wk["sigma"], salt, None, seconds=6.0), "zero_1024_cpu": measure_zero_cpu()}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "GMP mpz (64-bit limbs)",
        "data": "synthetic (seeded PCG64; " + ("reference test key" if N_BITS == 2048 else "committed fixture key") + "; 1% reject-path statements)",
        "config": headline_config(batch),
        "reference_note": "the Rust reference cannot be built offline (no cargo; curv-kzen / kzen-paillier un-vendored): this arm is its loops restated in C on its own "
                          "bigint backend (GMP), parallel over the security parameter like rayon; " + sample,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "secondary": secondary,
    }))


def kernel_time_ms(ctx):
    from zk_paillier_b200.native import KID_MODEXP_SHARED, KID_MODEXP_VAR, KID_MODMUL, KID_OTHER, KID_SHA

    tot, n = 0.0, 0
    for k in (KID_MODEXP_SHARED, KID_MODEXP_VAR, KID_MODMUL, KID_SHA, KID_OTHER):
        ms, cnt, _ = ctx.profile_get(k)
        tot += ms
        n += cnt
    return tot, n


def run_other(E, args):
    """The remaining public proofs, single GPU, through the one-shot host-buffer ABI:
      --config dlog            : CompositeDLogProof prove + verify (wi_dlog_proof.rs:46-91), one 2048-bit modulus N per statement
      --config correct_message : CorrectMessageProof prove + verify (correct_message.rs:35-162), 4 valid messages, 2048-bit n
    value = proofs over the summed device time of the kernels (CUDA events per launch; copies and host gaps excluded);
    e2e = the same calls from pinned host buffers on two contexts / two host threads, wall clock."""
    from zk_paillier_b200 import workload
    from zk_paillier_b200.native import ints_to_limbs, to_limbs

    oracle_path()
    import c_oracle
    from util import keys

    ctx = E.ctx
    imad_peak = E.imad_peak()
    cores = c_oracle.hw_threads()
    steps = max(args.steps, 1)
    line = {"n_gpus": 1, "steps": steps, "warmup": max(args.warmup, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": DTYPE, "data": "synthetic"}
    if args.config == "dlog":
        bits, B = 2048, args.batch if args.batch != 1024 else 4096
        nl = bits // 32
        rng = __import__("random").Random(11)
        ks = keys(bits)
        st = []
        for i in range(B):
            pp, qq = ks[i % len(ks)]
            N = pp * qq
            g = rng.randrange(2, N - 1)
            sec = rng.getrandbits(256)
            st.append((N, g, pow(pow(g, -1, N), sec, N), sec, rng.getrandbits(512)))
        N_, g_, ni_, s_, r_ = (pin(ints_to_limbs([t[k] for t in st], w)) for k, w in enumerate((nl, nl, nl, 8, 16)))

        def make(c):
            def step():
                x, y, flt = c.dlog_prove(N_, g_, ni_, s_, r_, 20)
                acc, flt2 = c.dlog_verify(N_, g_, ni_, x, y)
                assert acc.all() and not flt.any() and not flt2.any()
                return x, y
            return step

        alg = B * (modexp_imads(bits, 512) + modexp_imads(bits, 256) + modexp_imads(bits, 513))
        name = "CompositeDLogProof proofs+verifies/sec at 2048-bit N"
        wl_name = f"CompositeDLogProof prove + verify x{B}, 2048-bit N, {len(ks)} distinct moduli cycled, 256-bit secrets"
        kern = "modexp_var_kernel<8,8> (K2, short exponents)"

        def cpu(m):
            t0 = time.perf_counter()
            x, y = c_oracle.dlog_prove(N_[:m], g_[:m], ni_[:m], s_[:m], r_[:m], 20, cores)
            v = c_oracle.dlog_verify(N_[:m], g_[:m], ni_[:m], x, y, cores)
            dt = time.perf_counter() - t0
            assert (np.asarray(v) == 1).all()
            return dt, (x, y)

        def same(gpu, cpu_out, m):
            return np.array_equal(gpu[0][:m], cpu_out[0]) and np.array_equal(gpu[1][:m], cpu_out[1])
    else:
        bits, B, M = 2048, args.batch if args.batch != 1024 else 1024, 4
        nl = bits // 32
        n = test_key()
        g = np.random.Generator(np.random.PCG64(7))
        rows = lambda *shape: workload._rand_limbs_below_pow2(g, shape, nl, bits - 1)
        valid = pin(ints_to_limbs([[3, 4, 5, 6]] * B, 4))
        msgs = pin(ints_to_limbs([3 + (i % M) for i in range(B)], 4))
        r, w, z_rand = pin(rows(B) | 1), pin(rows(B) | 1), pin(rows(B, M - 1) | 1)
        e_rand = pin(np.frombuffer(g.bytes(B * (M - 1) * 32), dtype=np.uint32).reshape(B, M - 1, 8).copy())
        nlimbs = to_limbs(n, nl)

        def make(c):
            c.set_key(nlimbs)

            def step():
                out = c.correct_message_prove(valid, msgs, r, e_rand, z_rand, w)
                acc, flt = c.correct_message_verify(out["ciphertext"], valid, out["e_vec"], out["z_vec"], out["a_vec"])
                assert acc.all() and not flt.any() and not out["fault"].any()
                return out
            return step

        slot = modexp_imads(2 * bits, bits) + modexp_imads(2 * bits, 256)
        alg = B * (2 * M * slot + modexp_imads(2 * bits, bits) + modexp_imads(bits, 256))
        name = "CorrectMessageProof proofs+verifies/sec at 2048-bit n"
        wl_name = f"CorrectMessageProof prove + verify x{B}, {M} valid messages, 2048-bit n (reference test key)"
        kern = "enc2m_kernel<8,8> (K1m) + modexp2m_var_kernel<8,8> (K2m)"

        def cpu(m):
            t0 = time.perf_counter()
            o = c_oracle.correct_message_prove(nlimbs, valid[:m], msgs[:m], r[:m], e_rand[:m], z_rand[:m], w[:m], cores)
            v = c_oracle.correct_message_verify(nlimbs, o["ciphertext"], valid[:m], o["e_vec"], o["z_vec"], o["a_vec"], cores)
            dt = time.perf_counter() - t0
            assert (np.asarray(v) == 1).all()
            return dt, o

        def same(gpu, cpu_out, m):
            return all(np.array_equal(gpu[k][:m], cpu_out[k]) for k in ("ciphertext", "e_vec", "z_vec", "a_vec"))

    step = make(ctx)
    for _ in range(max(args.warmup, 3)):
        step()
    ctx.profile_enable(True); ctx.profile_reset()
    sampler = E.sampler()
    for _ in range(steps):
        gpu_out = step()
    dev_ms, launches = kernel_time_ms(ctx)
    ctx.profile_enable(False); ctx.profile_reset()
    e2e_s, batches = two_contexts(E, [make, make], max(steps, 2))
    clocks = sampler.stop() if sampler else None
    m = min(B, 8 * cores)
    t, _ = cpu(m)
    m = int(max(m, min(B, 6.0 / max(t / m, 1e-6))))
    cpu_s, cpu_out = cpu(m)
    if not same(gpu_out, cpu_out, m):
        raise SystemExit("bench.py: CUDA outputs differ from the GMP oracle on the baseline sample")
    line.update(metric=name, unit="proofs+verifies/s", value=B * steps / (dev_ms * 1e-3), ms_per_step=dev_ms / steps,
                value_how="summed device time of the kernels (CUDA events around every launch); host<->device copies and host gaps excluded",
                config={"workload": wl_name, "batch_per_gpu": B, "n_bits": bits},
                e2e={"value": B * batches / e2e_s, "unit": "proofs+verifies/s", "steps": batches,
                     "how": "prove + verify calls from pinned host buffers on two contexts / two host threads (double-buffered), host wall clock around synchronised calls"},
                roofline={"bound": "imad", "kernel": kern, "achieved": alg * steps / (dev_ms * 1e-3) / 1e12, "peak": imad_peak / 1e12,
                          "unit": "T IMAD.WIDE.U32/s", "frac": alg * steps / (dev_ms * 1e-3) / imad_peak, "traffic": None,
                          "frac_note": "ALGORITHMIC multiply-adds (SURVEY.md 8d formula) per second of summed kernel time; the two-digit kernels execute about half of them"},
                clocks=clocks, gpu_launches=int(launches),
                cpu_baseline={"value": m / cpu_s, "unit": "proofs+verifies/s", "cores": cores, "kind": "port",
                              "sample": f"{m} of the {B} proofs, prove + verify, GMP {c_oracle.gmp_version()} mpz_powm on {cores} threads; proof fields identical to the GPU's"})
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="rangeproof", choices=["rangeproof", "correct_key", "sigma", "dlog", "correct_message"])
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="proofs per GPU per step (configs[1])")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--ref-seconds", type=float, default=120.0, help="target total time of the --impl reference run")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="headline only (skip the other BASELINE configs)")
    ap.add_argument("--jobs-shape", type=int, default=0, choices=[0, 1, 2], help="K2h lane layout (zkp_tune ZKP_TUNE_JOBS_SHAPE): 0 = by job count")
    ap.add_argument("--n-bits", type=int, default=2048, choices=[1024, 2048, 3072, 4096],
                    help="key size of the RangeProofNi workload (the headline is 2048; the others are secondary lines)")
    args = ap.parse_args()
    if args.n_bits != N_BITS:
        set_key_size(args.n_bits)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
