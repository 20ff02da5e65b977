#!/usr/bin/env python
"""bench.py -- RangeProofNi proofs+verifies/sec at 2048-bit n on 1..8 B200 (BASELINE.json metric).

One "step" = RangeProofNi::prove followed by RangeProofNi::verify over one batch of synthetic statements
under the reference's fixed 2048-bit test key (range_proof_ni.rs:141-145), error factor 128
(range_proof_ni.rs:23), 256-bit ranges (range_proof_ni.rs:133) -- the shape of the reference's Criterion
bench (benches/all.rs:55-71), batched.  Default workload = BASELINE.json configs[1]: batch 1024 per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA, sm_100a)
  python bench.py --impl reference ...                            the CPU path (GMP mpz_powm, all host cores)

Numbers printed (one JSON line, rank 0):
  value       device-resident throughput: inputs already in HBM, prove_run + verify_run chained on the device,
              CUDA events on the launch stream, max over ranks
  e2e         the same metric through the public C ABI with HOST (pinned) buffers: H2D of the statement and the
              randomness, prove, D2H of the whole proof, H2D of the proof again (the verifier is another party),
              verify, D2H of the verdicts -- all inside the timed region
  roofline    the dominant kernel (K1m, the two-digit Montgomery encryption kernel): algorithmic multiply-adds
               per second of its device time against the IMAD.WIDE.U32 issue peak measured in
              this run, the multiply-adds it actually executes (executed_frac), and its HBM view
  cpu_baseline  oracle/oracle.c (the reference's loops on the reference's own backend, GMP) on a bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_BITS = 2048
EF = 128
METRIC = "RangeProofNi proofs+verifies/sec at 2048-bit n"
UNIT = "proofs+verifies/s"


# ---- algorithmic work ---------------------------------------------------
def mm(s):
    return 2 * s * s + s


def modexp_imads(mod_bits, exp_bits):
    return (exp_bits + -(-exp_bits // 5) + 32) * mm(mod_bits // 32)


ENC_IMADS = modexp_imads(2 * N_BITS, N_BITS)  # 81.9 M at 2048-bit n
# algorithmic HBM bytes of one Enc inside RangeProofNi: base (|n|) + plaintext row in, ciphertext (2|n|) out
ENC_BYTES = N_BITS // 8 + 48 + 2 * N_BITS // 8


def test_key():
    """2048-bit n: the reference's fixed test primes (range_proof_ni.rs:141-145).  --n-bits 3072 / 4096 (secondary lines, the
    other key sizes BASELINE.json's target names): the first committed fixture key of that size (tests/golden/keys.json)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import zkp_oracle as po  # constants only (the reference's fixed test primes)

    if N_BITS == 2048:
        return po.TEST_P * po.TEST_Q
    k = json.load(open(os.path.join(ROOT, "tests", "golden", "keys.json")))[str(N_BITS)][0]
    return int(k["p"]) * int(k["q"])


def set_key_size(bits):
    global N_BITS, METRIC, ENC_IMADS, ENC_BYTES
    N_BITS = bits
    METRIC = f"RangeProofNi proofs+verifies/sec at {bits}-bit n"
    ENC_IMADS = modexp_imads(2 * bits, bits)
    ENC_BYTES = bits // 8 + 48 + 2 * bits // 8


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


def run_b200(args):
    import torch
    import torch.distributed as dist
    import zk_paillier_b200 as zk
    from zk_paillier_b200 import workload
    from zk_paillier_b200.native import KID_MODEXP_SHARED, KID_MODMUL, KID_OTHER, KID_SHA, to_limbs

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # the GPU box exports NCCL_DEBUG=VERSION, so NCCL prints its version banner on stdout before rank 0's JSON line; it is
        # left alone (it is the launcher's setting, and the evidence that NCCL initialised)
        dist.init_process_group("nccl", device_id=dev)
    batch = args.batch
    nl = N_BITS // 32

    # public key: rank 0 owns it, one NCCL broadcast (the only collective before the hot path)
    from zk_paillier_b200 import sharding

    n_limbs_arr = sharding.broadcast_key(to_limbs(test_key(), nl) if rank == 0 else np.zeros(1, np.uint32), dev)
    n_int = int.from_bytes(n_limbs_arr.tobytes(), "little")

    stream = torch.cuda.Stream(dev)  # the library launches on this stream, so torch events on it time the kernels
    torch.cuda.set_stream(stream)
    ctx = zk.native.Context(local, stream=stream.cuda_stream)
    ctx.set_key(n_limbs_arr)

    # this rank's shard of independent statements (weak scaling: `batch` proofs per GPU)
    work = workload.rangeproof_batch(n_int, batch, ef=EF, seed=workload.DEFAULT_SEED + rank, reject_every=100)
    wl = work["w_limbs"]
    cx = ctx.paillier_enc(work["x_n"], work["r"])  # statement ciphertexts c = Enc(x, r): not on the measured path

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---------------- device-resident steps (value) ----------------
    ctx.rp_prove_stage(EF, work["range"], work["x"], work["r"], work["w1"], work["swap"], work["r1"], work["r2"])
    ctx.rp_prove_run()
    ctx.rp_verify_stage_from_prove(cx)

    def step_device():
        ctx.rp_prove_run()
        ctx.rp_verify_run()

    for _ in range(args.warmup):
        step_device()
    ctx.profile_enable(True)
    ctx.profile_reset()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
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
    host_in = {k: pinned(work[k].shape, work[k].dtype.type) for k in ("range", "x", "r", "w1", "swap", "r1", "r2")}
    for k in host_in:
        host_in[k][...] = work[k]
    cx_h = pinned(cx.shape, np.uint32)
    cx_h[...] = cx
    lib, h = ctx._lib, ctx._h
    from zk_paillier_b200.native import _p32, _p8

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
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record(stream)
    for _ in range(args.e2e_steps):
        step_e2e()
    if world > 1:  # final gather over NVLink: the proofs (device to device, 0.2 MB each), then verdicts + challenge hashes (33 B per proof)
        allproofs = sharding.gather_proof_bytes([out[k] for k in ("c1", "c2", "kind", "resp_w", "resp_r")], dev)
        torch.cuda.synchronize(dev)
        assert allproofs.shape[:2] == (world, batch)
        gathered_bytes = int(allproofs.numel())
        del allproofs
        allrec = sharding.gather_records(np.concatenate([acc_h[:, None], dig_h], axis=1), dev, counts=[batch] * world)
        assert allrec.shape == (world * batch, 33)
    g1.record(stream)
    barrier()
    e2e_ms = g0.elapsed_time(g1)  # the ABI calls are host-synchronous, so the event pair brackets copies + kernels + host gaps
    if not np.array_equal(acc_h, expect):
        raise SystemExit("bench.py: wrong verdicts from the e2e path")

    # max over ranks
    times = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, e2e_ms = times.tolist()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = world * batch * args.steps / (ms * 1e-3)
    e2e_value = world * batch * args.e2e_steps / (e2e_ms * 1e-3)

    # ---------------- roofline of the dominant kernel ----------------
    k1_ms, k1_launches, k1_units = prof[KID_MODEXP_SHARED]
    imad_peak = ctx.imad_peak(0)
    achieved = k1_units * ENC_IMADS / (k1_ms * 1e-3)
    used = ctx.enc_kernel_launches()
    which = "k1m" if used["k1m"] and not used["k1"] else ("k1" if used["k1"] and not used["k1m"] else "mixed")
    exec_mads = ctx.enc_executed_mads().get(which, 0.0)
    executed = k1_units * exec_mads / (k1_ms * 1e-3)
    shape = {1024: "<4,8>", 2048: "<8,8>", 3072: "<8,12>", 4096: "<16,8>"}.get(N_BITS, "")
    kernel_name = {"k1m": f"enc2m_kernel{shape} (K1m, two-digit Montgomery form)", "k1": "modexp_shared_kernel (K1)"}.get(which, which)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    traffic = None
    try:  # DRAM bytes of K1 from the committed ncu capture, scaled to this run's Enc per launch
        tr = json.load(open(os.path.join(ROOT, "profiles", "k1m_traffic.json" if which == "k1m" else "k1_traffic.json")))
        if N_BITS == 2048:  # the capture is of the 2048-bit kernel
            traffic = tr["dram_bytes_per_enc"] * k1_units / max(k1_launches, 1)
    except Exception:
        pass
    hbm_ach = k1_units * ENC_BYTES / (k1_ms * 1e-3) / 1e9
    kernel_ms = {"modexp_shared": k1_ms, "modmul": prof[KID_MODMUL][0], "sha256_transcript": prof[KID_SHA][0], "other": prof[KID_OTHER][0]}
    launches = int(sum(p[1] for p in prof.values()))

    # ---------------- CPU baseline on a bounded sample (rank 0, N=1 only) ----------------
    cpu = None
    if world == 1 and not args.no_cpu:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import c_oracle

        cores = c_oracle.hw_threads()
        t1, pr1, acc1 = cpu_sample(n_int, work, cx, np.arange(1), cores)
        m = int(max(1, min(64, args.cpu_seconds / max(t1, 1e-3))))
        dt, pr, acc = cpu_sample(n_int, work, cx, np.arange(m), cores)
        same = all(np.array_equal(pr[k], out[k][:m]) for k in ("c1", "c2", "digest", "kind", "resp_w", "resp_r")) and np.array_equal(acc, acc_h[:m])
        if not same:
            raise SystemExit("bench.py: CUDA outputs differ from the CPU oracle on the baseline sample")
        cpu = {"value": m / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{m} of the {batch} proofs (prove+verify, {m * 256} + {int((pr['kind'] == 0).sum()) + m * EF} Enc), GMP {c_oracle.gmp_version()} mpz_powm, outputs byte-identical to the GPU's"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (32x32+64 IMAD)",
        "data": "synthetic (seeded PCG64; " + ("reference test key" if N_BITS == 2048 else "committed fixture key") + "; 1% reject-path statements)",
        "config": {"workload": f"RangeProofNi prove+verify, batch={batch} per GPU, {N_BITS}-bit n ({'reference test key' if N_BITS == 2048 else 'committed fixture key'}), error_factor=128, 256-bit range",
                   "batch_per_gpu": batch, "n_bits": N_BITS, "error_factor": EF, "enc_per_step_per_gpu": int(2 * batch * EF + enc_verify),
                   "l2": "working set per step (approx 0.5 GB of bases, ciphertexts and responses) exceeds the 126 MB L2; no explicit flush",
                   "sharding": "independent proofs, contiguous shard per rank; NCCL broadcast of n before; after the last step of the e2e region an all_gather of the proof bytes (device to device) and of the verdicts; no collective on the modexp path"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": args.e2e_steps},
        "gpu_launches": launches,
        "roofline": {"bound": "imad", "kernel": kernel_name, "achieved": achieved / 1e12, "peak": imad_peak / 1e12, "unit": "T IMAD.WIDE.U32/s",
                     "frac": achieved / imad_peak, "traffic": traffic,
                     "executed": executed / 1e12, "executed_frac": executed / imad_peak, "executed_imads_per_enc": exec_mads,
                     "frac_note": "achieved = ALGORITHMIC multiply-adds (SURVEY.md 8d: fixed-window schoolbook CIOS modulo n^2) per second of kernel time; "
                                  "it exceeds the pipe peak because K1m executes about half of them (two-digit base-n Montgomery form, sliding window): "
                                  "executed_frac is the share of the measured IMAD.WIDE issue peak the kernel actually runs at",
                     "traffic_note": "DRAM bytes per launch = ncu dram_bytes per Enc (profiles/k1m_traffic.json or k1_traffic.json) x Enc per launch",
                     "peak_source": "IMAD.WIDE.U32 issue-rate microbenchmark measured in this run (MEASURED_PEAKS.json has no integer entry)",
                     "alg_imads_per_enc": ENC_IMADS, "k1_launches": int(k1_launches), "k1_ms_avg": k1_ms / max(k1_launches, 1),
                     "k1_share_of_step": k1_ms / ms},
        "roofline_hbm": {"bound": "hbm", "achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak, "traffic": traffic,
                         "alg_bytes_per_enc": ENC_BYTES, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
        "kernel_ms": kernel_ms,
        "clocks": clocks,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    """The reference's CPU path (its loops restated on its own backend, GMP) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_oracle
    from zk_paillier_b200 import workload

    n_int = test_key()
    cores = c_oracle.hw_threads()
    batch = args.batch
    work = workload.rangeproof_batch(n_int, min(batch, 64), ef=EF, seed=workload.DEFAULT_SEED, reject_every=100)
    from zk_paillier_b200.native import to_limbs

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
    sample = f"{m} of the {batch} proofs per step (prove+verify), GMP {c_oracle.gmp_version()} mpz_powm on {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "GMP mpz (64-bit limbs)",
        "data": "synthetic (seeded PCG64; " + ("reference test key)" if N_BITS == 2048 else "committed fixture key)"),
        "config": {"workload": f"RangeProofNi prove+verify, batch={batch} per GPU, {N_BITS}-bit n ({'reference test key' if N_BITS == 2048 else 'committed fixture key'}), error_factor=128, 256-bit range",
                   "note": "the Rust reference cannot be built offline (no cargo; curv-kzen / kzen-paillier un-vendored): this arm is its loops restated in C on its own bigint backend (GMP), parallel over the security parameter like rayon"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_other(args):
    """Secondary configs (not the driver's headline): BASELINE.json configs[2] and configs[4] on one GPU.
      --config correct_key : NiCorrectKeyProof verify, batch 4096, 3072-bit n (16 committed keys cycled; distinct-modulus kernel path)
      --config sigma       : MulProof + VerlinProof verify, 4096-bit n, batch 512 + 512 (the per-GPU share of 8192 over 8 GPUs)"""
    import torch
    import zk_paillier_b200 as zk
    from zk_paillier_b200 import workload
    from zk_paillier_b200.native import KID_MODEXP_SHARED, KID_MODEXP_VAR, to_limbs, ints_to_limbs, limbs_to_ints

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import c_oracle
    import zkp_oracle as po
    from util import keys

    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    ctx = zk.native.Context(0, stream=stream.cuda_stream)
    imad_peak = ctx.imad_peak(0)
    ctx.tune(zk.native.TUNE_JOBS_SHAPE, args.jobs_shape)
    line = {"n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 limbs (32x32+64 IMAD)", "data": "synthetic"}
    if args.config == "correct_key":
        bits, batch, salt = 3072, args.batch if args.batch != 1024 else 4096, b"Zen Go X"
        nl = bits // 32
        work = workload.correct_key_batch(keys(bits), batch, salt, lambda p, q, s: po.NiCorrectKeyProof.proof(p, q, s).sigma_vec, nl, bad_every=64)
        ctx.ck_verify_stage(work["n"], work["sigma"], salt)
        for _ in range(args.warmup):
            ctx.ck_verify_run()
        ctx.profile_enable(True); ctx.profile_reset()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            ctx.ck_verify_run()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        k2_ms, k2_n, k2_units = ctx.profile_get(KID_MODEXP_VAR)
        ctx.profile_enable(False)
        acc = ctx.ck_verify_fetch()
        assert acc.tolist() == [0 if b % 64 == 63 else 1 for b in range(batch)]
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            acc = ctx.correct_key_ni_verify(work["n"], work["sigma"], salt)
        e2e_s = time.perf_counter() - t0
        cores = c_oracle.hw_threads()
        m = min(batch, 16 * cores)
        t0 = time.perf_counter()
        acc_c, _ = c_oracle.correct_key_ni_verify(work["n"][:m], work["sigma"][:m], salt, cores)
        cpu_s = time.perf_counter() - t0
        assert np.array_equal(acc_c, acc[:m])
        per = modexp_imads(bits, bits)
        line.update(metric="NiCorrectKeyProof verifies/sec at 3072-bit n", unit="verifies/s", value=batch * args.steps / (ms * 1e-3),
                    ms_per_step=ms / args.steps,
                    config={"workload": f"NiCorrectKeyProof verify, batch={batch}, 3072-bit n, 16 distinct committed keys cycled, salt 'Zen Go X', 1/64 bad proofs"},
                    e2e={"value": batch * args.e2e_steps / e2e_s, "unit": "verifies/s", "h2d_bytes_per_step": int(work["n"].nbytes + work["sigma"].nbytes),
                         "d2h_bytes_per_step": batch},
                    roofline={"bound": "imad", "kernel": "modexp_var_kernel<8,12> (K2)", "achieved": k2_units * per / (k2_ms * 1e-3) / 1e12, "peak": imad_peak / 1e12,
                              "unit": "T IMAD.WIDE.U32/s", "frac": k2_units * per / (k2_ms * 1e-3) / imad_peak, "traffic": None, "alg_imads_per_modexp": per,
                              "k2_share_of_step": k2_ms / ms},
                    cpu_baseline={"value": m / cpu_s, "unit": "verifies/s", "cores": cores, "kind": "port", "sample": f"{m} of the {batch} proofs, GMP mpz_powm"})
    elif args.config == "dlog":
        # CompositeDLogProof prove + verify (wi_dlog_proof.rs:46-91), one modulus N per statement (the fixture keys, cycled)
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
        N_, g_, ni_, s_, r_ = (ints_to_limbs([t[k] for t in st], w) for k, w in enumerate((nl, nl, nl, 8, 16)))

        def step():
            x, y, flt = ctx.dlog_prove(N_, g_, ni_, s_, r_, 20)
            acc, flt2 = ctx.dlog_verify(N_, g_, ni_, x, y)
            assert acc.all() and not flt.any() and not flt2.any()
            return x, y

        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            x, y = step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        alg = B * (modexp_imads(bits, 512) + modexp_imads(bits, 256) + modexp_imads(bits, 513)) * args.steps
        m = 8
        t0 = time.perf_counter()
        for i in range(m):
            pr = po.CompositeDLogProof.prove(*st[i])
            assert pr.x == int.from_bytes(x[i].tobytes(), "little") and pr.y == int.from_bytes(y[i].tobytes(), "little")
            pr.verify(*st[i][:3])
        cpu_s = time.perf_counter() - t0
        line.update(metric="CompositeDLogProof proofs+verifies/sec at 2048-bit N", unit="proofs+verifies/s", value=B * args.steps / dt, ms_per_step=dt / args.steps * 1e3,
                    config={"workload": f"CompositeDLogProof prove + verify x{B}, 2048-bit N, {len(ks)} distinct moduli cycled, 256-bit secrets; through the host-buffer ABI"},
                    e2e={"value": B * args.steps / dt, "unit": "proofs+verifies/s", "h2d_bytes_per_step": int(2 * (N_.nbytes + g_.nbytes + ni_.nbytes) + s_.nbytes + r_.nbytes + x.nbytes + y.nbytes),
                         "d2h_bytes_per_step": int(x.nbytes + y.nbytes + 3 * B)},
                    roofline={"bound": "imad", "kernel": "modexp_var_kernel<8,8> (K2, short exponents)", "achieved": alg / dt / 1e12, "peak": imad_peak / 1e12,
                              "unit": "T IMAD.WIDE.U32/s", "frac": alg / dt / imad_peak, "traffic": None,
                              "frac_note": "algorithmic multiply-adds (SURVEY.md 8d formula) per second of the whole step, host copies included"},
                    cpu_baseline={"value": m / cpu_s, "unit": "proofs+verifies/s", "cores": 1, "kind": "port",
                                  "sample": f"{m} of the {B} statements on the Python-int oracle (CPython pow); x and y identical to the GPU's"})
    elif args.config == "correct_message":
        # CorrectMessageProof prove + verify (correct_message.rs:35-162), M = 4 valid messages as in the reference's test
        bits, B, M = 2048, args.batch if args.batch != 1024 else 1024, 4
        nl, nnl = bits // 32, bits // 16
        n = test_key()
        ctx.set_key(to_limbs(n, nl))
        g = np.random.Generator(np.random.PCG64(7))
        rows = lambda *shape: workload._rand_limbs_below_pow2(g, shape, nl, bits - 1)
        valid = ints_to_limbs([[3, 4, 5, 6]] * B, 4)
        msgs = ints_to_limbs([3 + (i % M) for i in range(B)], 4)
        r, w, z_rand = rows(B) | 1, rows(B) | 1, rows(B, M - 1) | 1
        e_rand = np.frombuffer(g.bytes(B * (M - 1) * 32), dtype=np.uint32).reshape(B, M - 1, 8).copy()

        def step():
            out = ctx.correct_message_prove(valid, msgs, r, e_rand, z_rand, w)
            acc, flt = ctx.correct_message_verify(out["ciphertext"], valid, out["e_vec"], out["z_vec"], out["a_vec"])
            assert acc.all() and not flt.any() and not out["fault"].any()
            return out

        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out = step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        slot = modexp_imads(2 * bits, bits) + modexp_imads(2 * bits, 256)
        alg = B * (2 * M * slot + modexp_imads(2 * bits, bits) + modexp_imads(bits, 256)) * args.steps
        m = 2
        L2I = lambda a: [int.from_bytes(np.ascontiguousarray(v).tobytes(), "little") for v in a]
        t0 = time.perf_counter()
        for i in range(m):
            pr = po.CorrectMessageProof.prove(n, [3, 4, 5, 6], 3 + (i % M), L2I(r[i:i + 1])[0], L2I(e_rand[i]), L2I(z_rand[i]), L2I(w[i:i + 1])[0])
            assert pr.a_vec == L2I(out["a_vec"][i]) and pr.e_vec == L2I(out["e_vec"][i]) and pr.z_vec == L2I(out["z_vec"][i])
            pr.verify()
        cpu_s = time.perf_counter() - t0
        line.update(metric="CorrectMessageProof proofs+verifies/sec at 2048-bit n", unit="proofs+verifies/s", value=B * args.steps / dt, ms_per_step=dt / args.steps * 1e3,
                    config={"workload": f"CorrectMessageProof prove + verify x{B}, {M} valid messages, 2048-bit n (reference test key); through the host-buffer ABI"},
                    e2e={"value": B * args.steps / dt, "unit": "proofs+verifies/s",
                         "h2d_bytes_per_step": int(2 * valid.nbytes + msgs.nbytes + r.nbytes + w.nbytes + z_rand.nbytes + e_rand.nbytes + sum(out[k].nbytes for k in ("ciphertext", "e_vec", "z_vec", "a_vec"))),
                         "d2h_bytes_per_step": int(sum(out[k].nbytes for k in ("ciphertext", "e_vec", "z_vec", "a_vec")) + 3 * B)},
                    roofline={"bound": "imad", "kernel": "enc2m_kernel<8,8> (K1m) + modexp2m_var_kernel<8,8> (K2m)", "achieved": alg / dt / 1e12, "peak": imad_peak / 1e12,
                              "unit": "T IMAD.WIDE.U32/s", "frac": alg / dt / imad_peak, "traffic": None,
                              "frac_note": "algorithmic multiply-adds (SURVEY.md 8d formula) per second of the whole step, host copies included; the two-digit kernels execute about half of them"},
                    cpu_baseline={"value": m / cpu_s, "unit": "proofs+verifies/s", "cores": 1, "kind": "port",
                                  "sample": f"{m} of the {B} proofs on the Python-int oracle (CPython pow); proof fields identical to the GPU's"})
    else:
        bits, B = 4096, args.batch if args.batch != 1024 else 512
        p, q = keys(bits)[0]
        n = p * q
        nl, nnl = bits // 32, bits // 16
        ctx.set_key(to_limbs(n, nl))
        g = np.random.Generator(np.random.PCG64(5))
        rows = lambda: workload._rand_limbs_below_pow2(g, (B,), nl, bits - 1)
        a, b = rows(), rows()
        c = ints_to_limbs([x * y % n for x, y in zip(limbs_to_ints(a), limbs_to_ints(b))], nl)
        r_a, r_b, r_c, d, r_d = (rows() | 1 for _ in range(5))
        e_a, e_b, e_c = ctx.paillier_enc(a, r_a), ctx.paillier_enc(b, r_b), ctx.paillier_enc(c, r_c)
        f, z1, z2, e_d, e_db, fault = ctx.mul_prove(a, b, r_a, r_b, r_c, e_a, e_b, e_c, d, r_d)
        x, xp, xdp, r_x = rows(), rows(), rows(), rows() | 1
        cc, cp = ctx.paillier_enc(rows(), rows() | 1), ctx.paillier_enc(rows(), rows() | 1)
        pad = lambda v: np.concatenate([v, np.zeros((B, nnl - nl), np.uint32)], axis=1)
        nn_rows = to_limbs(n * n, nnl)[None, :]
        phi_x = ctx.modmul(ctx.modmul(ctx.modexp_var(cc, pad(x), nn_rows, exp_per=1, mod_per=B, exp_bits=bits),
                                      ctx.modexp_var(cp, pad(xp), nn_rows, exp_per=1, mod_per=B, exp_bits=bits)), ctx.paillier_enc(xdp, r_x))
        phi_a, z, zp, zdp, r_z = ctx.verlin_prove(x, xp, xdp, r_x, cc, cp, phi_x, rows(), rows(), rows(), rows() | 1)

        def step():
            acc1, flt = ctx.mul_verify(e_a, e_b, e_c, f, z1, z2, e_d, e_db)
            acc2 = ctx.verlin_verify(cc, cp, phi_x, phi_a, z, zp, zdp, r_z)
            assert acc1.all() and acc2.all() and not flt.any()

        for _ in range(args.warmup):
            step()
        ctx.profile_enable(True); ctx.profile_reset()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        k1_ms, _, k1_units = ctx.profile_get(KID_MODEXP_SHARED)
        k2_ms, _, k2_units = ctx.profile_get(KID_MODEXP_VAR)
        mul_imads = 3 * modexp_imads(2 * bits, bits) + 2 * modexp_imads(2 * bits, 256)
        ver_imads = modexp_imads(2 * bits, 256) + 2 * modexp_imads(2 * bits, bits + 256) + modexp_imads(2 * bits, bits)
        alg = B * (mul_imads + ver_imads) * args.steps
        bytes_in = sum(v.nbytes for v in (e_a, e_b, e_c, f, z1, z2, e_d, e_db, cc, cp, phi_x, phi_a, z, zp, zdp, r_z))
        line.update(metric="MulProof+VerlinProof verifies/sec at 4096-bit n", unit="verifies/s", value=2 * B * args.steps / dt, ms_per_step=dt / args.steps * 1e3,
                    config={"workload": f"MulProof verify x{B} + VerlinProof verify x{B}, 4096-bit n (8192-bit modulus), one key; through the host-buffer ABI"},
                    e2e={"value": 2 * B * args.steps / dt, "unit": "verifies/s", "h2d_bytes_per_step": int(bytes_in), "d2h_bytes_per_step": 3 * B},
                    roofline={"bound": "imad", "kernel": "enc2m_kernel<16,8> (K1m) + modexp2m_var_kernel<16,8> (K2m), side by side on forked streams",
                              "achieved": alg / dt / 1e12, "peak": imad_peak / 1e12, "unit": "T IMAD.WIDE.U32/s", "frac": alg / dt / imad_peak, "traffic": None,
                              "frac_note": "ALGORITHMIC multiply-adds (SURVEY.md 8d) per second of the whole step (the modexp kernels of one proof overlap, so "
                                           "their summed durations exceed the wall time); the two-digit kernels execute about half of them",
                              "modexp_kernel_ms_summed": k1_ms + k2_ms},
                    cpu_baseline=None)
        if not args.no_cpu:
            # CPU baseline for this secondary line: the GMP restatement of the two verifiers (oracle/oracle.c), all host threads
            cores = c_oracle.hw_threads()
            m = min(B, 4 * cores)
            nlimbs = to_limbs(n, nl)
            t0 = time.perf_counter()
            v1 = c_oracle.mul_verify(nlimbs, e_a[:m], e_b[:m], e_c[:m], f[:m], z1[:m], z2[:m], e_d[:m], e_db[:m], cores)
            v2 = c_oracle.verlin_verify(nlimbs, cc[:m], cp[:m], phi_x[:m], phi_a[:m], z[:m], zp[:m], zdp[:m], r_z[:m], cores)
            cpu_s = time.perf_counter() - t0
            if not (v1 == 1).all() or not (v2 == 1).all():
                raise SystemExit("bench.py: the CPU oracle rejects proofs the GPU accepted")
            line["cpu_baseline"] = {"value": 2 * m / cpu_s, "unit": "verifies/s", "cores": cores, "kind": "port",
                                    "sample": f"{m} MulProof + {m} VerlinProof verifies of the batch, GMP {c_oracle.gmp_version()} mpz_powm, same verdicts as the GPU"}
    print(json.dumps(line))


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
    ap.add_argument("--jobs-shape", type=int, default=0, choices=[0, 1, 2], help="K2h lane layout (zkp_tune ZKP_TUNE_JOBS_SHAPE): 0 = by job count")
    ap.add_argument("--n-bits", type=int, default=2048, choices=[1024, 2048, 3072, 4096],
                    help="key size of the RangeProofNi workload (the headline is 2048; the others are secondary lines)")
    args = ap.parse_args()
    if args.n_bits != N_BITS:
        set_key_size(args.n_bits)
    if args.impl == "reference":
        run_reference(args)
    elif args.config != "rangeproof":
        run_other(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
