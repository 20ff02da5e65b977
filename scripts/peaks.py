"""Integer-pipe microbenchmarks (roofline denominators and design probes)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zk_paillier_b200 as zk
ctx = zk.native.Context(0)
names = {0: "IMAD.WIDE.U32 independent", 1: "IMAD.WIDE.U32.X carry-chained rows", 2: "IMAD 32-bit",
         3: "block_mul 32x32 full, 16 warps/SM", 4: "block_mul 32x32 full, 8 warps/SM", 5: "block_mul 32x32 lower-tri, 16 warps/SM",
         6: "block_mul 32x32 full, 12 warps/SM"}
res = {}
for v in sorted(names):
    r = ctx.imad_peak(v)
    res[names[v]] = r
    print(f"{names[v]:45s} {r/1e12:8.3f} T mads/s")
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/peaks.json", "w"), indent=1)
