"""Where the time of ONE RangeProofNi goes (batch = 1 through the one-shot ABI): per-kernel device time from the profile scopes
next to the host wall time of the call."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from zk_paillier_b200 import workload
from zk_paillier_b200.native import KID_MODEXP_SHARED, KID_MODMUL, KID_OTHER, KID_SHA, to_limbs
import zk_paillier_b200 as zk

ctx = zk.native.Context(0)
n_int = bench.test_key()
ctx.set_key(to_limbs(n_int, 64))
EF = 128
for batch, rows in ((1, 1), (1, 2), (4, 2), (16, 2)):
    ctx.tune(zk.native.TUNE_JOBS_ROWS, rows)
    work = workload.rangeproof_batch(n_int, batch, ef=EF, seed=5)
    cx = ctx.paillier_enc(work["x_n"], work["r"])
    args = (EF, work["range"], work["x"], work["r"], work["w1"], work["swap"], work["r1"], work["r2"])
    for _ in range(2):
        pr = ctx.rangeproof_ni_prove(*args)
        ctx.rangeproof_ni_verify(EF, work["range"], cx, pr["c1"], pr["c2"], pr["kind"], pr["resp_w"], pr["resp_r"])
    out = {"batch": batch, "rows": "single" if rows == 1 else "pair"}
    for name, fn in (("prove", lambda: ctx.rangeproof_ni_prove(*args)),
                     ("verify", lambda: ctx.rangeproof_ni_verify(EF, work["range"], cx, pr["c1"], pr["c2"], pr["kind"], pr["resp_w"], pr["resp_r"]))):
        ctx.profile_enable(True); ctx.profile_reset()
        t0 = time.perf_counter(); fn(); wall = (time.perf_counter() - t0) * 1e3
        k = {nm: round(ctx.profile_get(kid)[0], 3) for nm, kid in (("enc", KID_MODEXP_SHARED), ("modmul", KID_MODMUL), ("sha", KID_SHA), ("other", KID_OTHER))}
        ctx.profile_enable(False); ctx.profile_reset()
        out[name] = {"wall_ms": round(wall, 3), "kernels_ms": k, "host_and_copies_ms": round(wall - sum(k.values()), 3)}
    print(json.dumps(out), flush=True)
