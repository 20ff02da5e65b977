"""First-light measurement on the GPU box: IMAD peak variants and raw K1 throughput."""
import json, random, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import zk_paillier_b200 as zk
from zk_paillier_b200.native import to_limbs, ints_to_limbs, KID_MODEXP_SHARED, KID_MODEXP_VAR

ctx = zk.native.Context(0)
res = {"sm_count": ctx.sm_count}
for v in (0, 1, 2):
    res[f"imad_peak_v{v}"] = ctx.imad_peak(v)
rng = random.Random(1)
for n_bits in (2048, 3072, 4096):
    nl = n_bits // 32
    n = rng.getrandbits(n_bits) | 1 | (1 << (n_bits - 1))
    ctx.set_key(to_limbs(n, nl))
    batch = ctx.sm_count * 64 * 2
    r = np.frombuffer(np.random.default_rng(0).bytes(batch * nl * 4), dtype=np.uint32).reshape(batch, nl).copy()
    r[:, -1] &= 0x7fffffff
    m = np.zeros((batch, 8), np.uint32); m[:, 0] = 5
    ctx.paillier_enc(m[:64], r[:64])
    ctx.profile_enable(True); ctx.profile_reset()
    t = time.time(); ctx.paillier_enc(m, r); wall = time.time() - t
    ms, launches, units = ctx.profile_get(KID_MODEXP_SHARED)
    s = 2 * nl
    E = n_bits
    imads = (E + -(-E // 5) + 32) * (2 * s * s + s)
    res[f"enc_{n_bits}"] = {"batch": batch, "kernel_ms": ms, "wall_s": wall, "enc_per_s": batch / (ms * 1e-3),
                            "alg_imad_per_s": batch * imads / (ms * 1e-3)}
    ctx.profile_enable(False)
print(json.dumps(res, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/quick_core.json", "w"), indent=1)
