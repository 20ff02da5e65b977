"""Generate tests/golden/vectors.json by running the Python oracle (oracle/zkp_oracle.py) on fixed seeds.
The reference has no golden vectors and cannot be run here (no Rust toolchain), so these vectors pin the
ORACLE against regressions and the C oracle / CUDA path against the oracle; they do not pin the oracle
against the reference (see the PARITY note in oracle/zkp_oracle.py)."""
import hashlib, json, os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import zkp_oracle as po

n = po.TEST_P * po.TEST_Q
rng = random.Random(0x5A4B50)
g = {"n": str(n), "enc": [], "digest": []}
for m, r in [(0, 1), (1, 1), (n - 1, n - 1), (rng.getrandbits(256), rng.randrange(n)), (rng.getrandbits(2300), rng.randrange(n * n))]:
    g["enc"].append({"m": str(m), "r": str(r), "c": str(po.paillier_encrypt(n, m, r))})
for items in [[0], [0, 0, 1], [n, 255, 256], [rng.getrandbits(4096) >> 9, rng.getrandbits(4090), 1 << 4088]]:
    g["digest"].append({"items": [str(x) for x in items], "sha256": "%064x" % po.compute_digest(items)})
ef = 128
q = rng.getrandbits(256) | (1 << 255)
third = q // 3
x, r = rng.randrange(third), rng.randrange(n)
c = po.paillier_encrypt(n, x, r)
w1 = [rng.randrange(third, 2 * third) for _ in range(ef)]
swap = [rng.getrandbits(1) for _ in range(ef)]
r1 = [rng.randrange(n) for _ in range(ef)]
r2 = [rng.randrange(n) for _ in range(ef)]
proof = po.RangeProofNi.prove(n, q, c, x, r, w1, swap, r1, r2)
proof.verify(n, c)
g["range_proof_ni"] = {"range": str(q), "x": str(x), "r": str(r), "ciphertext": str(c), "w1": [str(v) for v in w1], "swap": swap,
                       "r1": [str(v) for v in r1], "r2": [str(v) for v in r2],
                       "challenge_digest": po.range_digest32(n, proof.encrypted_pairs).hex(),
                       "kinds": [0 if t[0] == "Open" else t[1] for t in proof.proof],
                       "proof_json_sha256": hashlib.sha256(proof.to_json().encode()).hexdigest()}
salt = b"Zen Go X"
ck = po.NiCorrectKeyProof.proof(po.TEST_P, po.TEST_Q, salt)
ck.verify(n, salt)
g["correct_key_ni"] = {"p": str(po.TEST_P), "q": str(po.TEST_Q), "salt_hex": salt.hex(), "sigma_vec": [str(s) for s in ck.sigma_vec],
                       "rho_vec": [str(s) for s in po.correct_key_rho(n, salt)]}
# ---- the remaining public proofs (row f3): a second file so that vectors.json stays byte-identical
rng = random.Random(0xF3)
R = 1 << (po.DLOG_K + po.DLOG_K_PRIME + po.DLOG_SAMPLE_S)
leg = lambda a, pr: 1 if pow(a, (pr - 1) // 2, pr) == 1 else -1
while True:
    h1 = rng.randrange(1, n - 1)
    if leg(h1, po.TEST_P) * leg(h1, po.TEST_Q) == -1:
        break
secret = rng.randrange(1 << po.DLOG_SAMPLE_S)
h2 = pow(pow(h1, -1, n), secret, n)
rr = rng.randrange(R)
dl = po.CompositeDLogProof.prove(n, h1, h2, secret, rr)
dl.verify(n, h1, h2)
valid = [3, 4, 5, 6]
cm_in = {"message": 5, "r": rng.randrange(1, n), "e_rand": [rng.getrandbits(256) for _ in valid[1:]],
         "z_rand": [rng.randrange(1, n) for _ in valid[1:]], "w": rng.randrange(1, n)}
cm = po.CorrectMessageProof.prove(n, valid, **cm_in)
cm.verify()
g3 = {"n": str(n),
      "dlog": {"g": str(h1), "ni": str(h2), "secret": str(secret), "r": str(rr), "x": str(dl.x), "y": str(dl.y), "json": dl.to_json()},
      "correct_message": {"valid": valid, "message": cm_in["message"], "r": str(cm_in["r"]), "e_rand": [str(v) for v in cm_in["e_rand"]],
                          "z_rand": [str(v) for v in cm_in["z_rand"]], "w": str(cm_in["w"]), "ciphertext": str(cm.ciphertext),
                          "e_vec": [str(v) for v in cm.e_vec], "z_vec": [str(v) for v in cm.z_vec],
                          "a_vec_sha256": hashlib.sha256(",".join(str(v) for v in cm.a_vec).encode()).hexdigest()}}
json.dump(g3, open(os.path.join(ROOT, "tests", "golden", "vectors_more.json"), "w"), indent=0)
json.dump(g, open(os.path.join(ROOT, "tests", "golden", "vectors.json"), "w"), indent=0)
print("wrote vectors.json, vectors_more.json")
