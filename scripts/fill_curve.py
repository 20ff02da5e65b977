"""How full does a launch of B long modexps make the GPU?  ZeroProof::verify at <bits>-bit n for a range of batch sizes,
K2h in its wide-lane and narrow-lane layouts: device time of the one K2h launch per call (Enc(0, z) with the |n|-bit
exponent + c^e with the 256-bit challenge per proof).  usage: python scripts/fill_curve.py [bits]"""
import json, os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import zk_paillier_b200 as zk
from zk_paillier_b200.native import to_limbs, KID_MODEXP_VAR, TUNE_JOBS_SHAPE, TUNE_JOBS_ROWS
from util import keys

bits = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
p, q = keys(bits)[0]
n = p * q
nl, nnl = bits // 32, bits // 16
ctx = zk.native.Context(0)
ctx.set_key(to_limbs(n, nl))
g = np.random.default_rng(3)
res = []
for B in [int(x) for x in os.environ.get("FILL_B", "148,296,512,768,1024,1536,2048,3072").split(",")]:
    c = np.frombuffer(g.bytes(B * nnl * 4), dtype=np.uint32).reshape(B, nnl).copy(); c[:, -1] &= 0x0fffffff
    z = c.copy(); a = c.copy()
    for shape, rows in ((1, 0), (2, 1), (2, 2)):
        ctx.tune(TUNE_JOBS_SHAPE, shape)
        ctx.tune(TUNE_JOBS_ROWS, rows)
        ctx.zero_verify(c, z, a)
        ctx.profile_enable(True); ctx.profile_reset()
        ctx.zero_verify(c, z, a)
        ms, launches, units = ctx.profile_get(KID_MODEXP_VAR)
        ctx.profile_enable(False)
        r = {"bits": bits, "batch": B, "shape": "wide" if shape == 1 else ("narrow, single rows" if rows == 1 else "narrow, pair rows"), "k2h_ms": round(ms, 2), "launches": launches, "proofs_per_s_kernel": round(B / (ms * 1e-3), 1)}
        print(json.dumps(r), flush=True)
        res.append(r)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open(f"gpurun_out/fill_curve_{bits}.json", "w"), indent=1)
