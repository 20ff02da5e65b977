"""Raw throughput of the encryption kernels (device time of one launch, CUDA events) per tuning variant.
usage: python scripts/k1m_variants.py <n_bits> <tag>   (ZKP_B200_ENC / ZKP_B200_K1M_VARIANT select the kernel)"""
import json, os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import zk_paillier_b200 as zk
from zk_paillier_b200.native import to_limbs, from_limbs, KID_MODEXP_SHARED

n_bits = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
tag = sys.argv[2] if len(sys.argv) > 2 else "default"
ctx = zk.native.Context(0)
rng = random.Random(1)
nl = n_bits // 32
n = rng.getrandbits(n_bits) | 1 | (1 << (n_bits - 1))
ctx.set_key(to_limbs(n, nl))
per_wave_lcm = ctx.sm_count * 32 * 60          # whole waves for every variant: 3 / 4 / 5 / 6 resident CTAs, 16 or 32 groups per CTA
batch = per_wave_lcm if n_bits <= 2048 else per_wave_lcm // 8
batch = int(os.environ.get("K1M_BATCH", batch))
r = np.frombuffer(np.random.default_rng(0).bytes(batch * nl * 4), dtype=np.uint32).reshape(batch, nl).copy()
r[:, -1] &= 0x7fffffff
m = np.zeros((batch, 8), np.uint32); m[:, 0] = 5; m[:, 7] = 0x1234
out = ctx.paillier_enc(m[:64], r[:64])
CHECK = os.environ.get("K1M_NO_CHECK") is None   # the _nosub lab build computes wrong residues on purpose
for j in (0, 3, 63) if CHECK else ():
    assert from_limbs(out[j]) == ((from_limbs(m[j]) * n + 1) * pow(from_limbs(r[j]), n, n * n)) % (n * n), f"variant {tag}: wrong ciphertext"
ctx.profile_enable(True); ctx.profile_reset()
big = ctx.paillier_enc(m, r)
for j in (0, batch // 2 + 17, batch - 1) if CHECK else ():
    assert from_limbs(big[j]) == ((from_limbs(m[j]) * n + 1) * pow(from_limbs(r[j]), n, n * n)) % (n * n), f"variant {tag}: wrong ciphertext in the big batch"
ms, launches, units = ctx.profile_get(KID_MODEXP_SHARED)
res = {"tag": tag, "variant": os.environ.get("ZKP_B200_K1M_VARIANT", "0"), "window": os.environ.get("ZKP_B200_K1M_WINDOW", "5"), "n_bits": n_bits, "batch": batch, "kernel_ms": round(ms, 2), "enc_per_s": round(batch / (ms * 1e-3)),
       "kernels": ctx.enc_kernel_launches()}
print(json.dumps(res))
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/k1m_variants.jsonl", "a").write(json.dumps(res) + "\n")
