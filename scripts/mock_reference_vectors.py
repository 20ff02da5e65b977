"""Writes a file in the schema of tests/golden/reference_vectors.json from the PYTHON ORACLE (not from the reference).

Purpose: exercise tests/test_reference_vectors.py end to end while the real file is absent (no cargo here).  The output is
marked `"generator": "oracle-mock"` and must never be committed as tests/golden/reference_vectors.json - a pin of the
oracle against itself pins nothing.  The real file comes from rust/gen_vectors (the unmodified reference crate).

usage: python scripts/mock_reference_vectors.py out.json"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import zkp_oracle as po


def sigma_json(**f):
    return json.dumps({k: po.serde_bigint_native(v) for k, v in f.items()}, separators=(",", ":"))


def verdict(fn):
    try:
        fn()
        return "ok"
    except po.IncorrectProof:
        return "incorrect"


def build(seed=1):
    rng = random.Random(seed)
    p, q = po.TEST_P, po.TEST_Q
    n = p * q
    nn = n * n
    rnd = lambda: rng.randrange(1, n)
    probes = [0, 1, 255, 256, 2**32, 2**32 - 1, 2**64, 2**255, 2**256 - 1, 2**2047 + 12345, n, nn]
    lists = [[0], [n, 0, 1], [2**64, 255, nn], [int.from_bytes(bytes([75, 90, 101, 110]), "big")], [int.from_bytes(bytes([0, 0, 7, 9]), "big")], list(range(40))]
    doc = {
        "generator": "oracle-mock",
        "crates": {},
        "key": {"p": str(p), "q": str(q), "n": str(n)},
        "to_bytes": [{"dec": str(x), "hex": po.bigint_to_bytes(x).hex()} for x in probes],
        "compute_digest": [{"items": [str(v) for v in l], "digest": str(po.compute_digest(l))} for l in lists],
        "bigint_serde": [{"dec": str(x), "json": json.dumps(po.serde_bigint_native(x))} for x in probes],
        "encryption_key": {"n": str(n), "json": json.dumps(po.serde_encryption_key(n), separators=(",", ":"))},
        "paillier_enc": [{"m": str(m), "r": str(r), "c": str(po.paillier_encrypt(n, m, r))} for m, r in [(0, 1), (1, 2), (n - 1, n - 1), (rnd(), rnd()), (rng.getrandbits(256), rnd())]],
    }
    doc["ni_correct_key"] = []
    for salt in (po.SALT_STRING, bytes([0, 0, 1])):
        pr = po.NiCorrectKeyProof.proof(p, q, salt)
        doc["ni_correct_key"].append({"p": str(p), "q": str(q), "salt_hex": salt.hex(), "json": pr.to_json(), "verdict": verdict(lambda: pr.verify(n, salt)),
                                      "verdict_wrong_salt": verdict(lambda: pr.verify(n, bytes([1, 2, 3])))})
    doc["range_proof_ni"] = []
    ef = 16  # the reference's SECURITY_PARAMETER is 128; the mock only exercises the consuming code and keeps the CPU suite short
    for honest in (True, False):
        rg = rng.getrandbits(256) | (1 << 255)
        r = rnd()
        x = rng.randrange(rg // 3) if honest else rng.randrange(100 * rg, 10000 * rg)
        c = po.paillier_encrypt(n, x, r)
        third = rg // 3
        pr = po.RangeProofNi.prove(n, rg, c, x, r, [rng.randrange(third, 2 * third) for _ in range(ef)], [rng.getrandbits(1) for _ in range(ef)],
                                   [rnd() for _ in range(ef)], [rnd() for _ in range(ef)])
        doc["range_proof_ni"].append({"range": str(rg), "x": str(x), "r": str(r), "ciphertext": str(c), "json": pr.to_json(), "verdict": verdict(lambda: pr.verify(n, c))})
    doc["zero"] = []
    for m in (0, 1):
        r = rnd()
        c = po.paillier_encrypt(n, m, r)
        pr = po.ZeroProof.prove(r, n, c, rnd())
        doc["zero"].append({"c": str(c), "json": sigma_json(z=pr.z, a=pr.a), "verdict": verdict(lambda: pr.verify(n, c))})
    doc["ciphertext"] = []
    for bad in (0, 1):
        x, r = rnd(), rnd()
        c = po.paillier_encrypt(n, x, r)
        pr = po.CiphertextProof.prove(x, r + bad, n, c, rnd(), rnd())
        doc["ciphertext"].append({"c": str(c), "json": sigma_json(z1=pr.z1, z2=pr.z2, c_prime=pr.c_prime), "verdict": verdict(lambda: pr.verify(n, c))})
    doc["mul"] = []
    for bad in (0, 1):
        a, b = rnd(), rnd()
        c = (a * b + bad) % n
        r_a, r_b, r_c = rnd(), rnd(), rnd()
        e_a, e_b, e_c = po.paillier_encrypt(n, a, r_a), po.paillier_encrypt(n, b, r_b), po.paillier_encrypt(n, c, r_c)
        pr = po.MulProof.prove(a, b, c, r_a, r_b, r_c, n, e_a, e_b, e_c, rnd(), rnd())
        doc["mul"].append({"e_a": str(e_a), "e_b": str(e_b), "e_c": str(e_c), "json": sigma_json(f=pr.f, z1=pr.z1, z2=pr.z2, e_d=pr.e_d, e_db=pr.e_db),
                           "verdict": verdict(lambda: pr.verify(n, e_a, e_b, e_c))})
    doc["verlin"] = []
    for bad in (0, 1):
        x, xp, xdp, r_x = rnd(), rnd(), rnd(), rnd()
        c, cp = po.paillier_encrypt(n, rnd(), rnd()), po.paillier_encrypt(n, rnd(), rnd())
        phi_x = po.gen_phi(n, c, cp, x, xp, xdp, r_x)
        pr = po.VerlinProof.prove(x + bad, xp, xdp, r_x, n, c, cp, phi_x, rnd(), rnd(), rnd(), rnd())
        doc["verlin"].append({"c": str(c), "c_prime": str(cp), "phi_x": str(phi_x),
                              "json": sigma_json(phi_a=pr.phi_a, z=pr.z, z_prime=pr.z_prime, z_double_prime=pr.z_double_prime, r_z=pr.r_z),
                              "verdict": verdict(lambda: pr.verify(n, c, cp, phi_x))})
    doc["dlog"] = []
    for bad in (0, 1):
        g = rnd()
        secret = rng.getrandbits(256)
        ni = pow(pow(g, -1, n), secret, n)
        pr = po.CompositeDLogProof.prove(n, g, ni, secret + bad, rng.getrandbits(512))
        doc["dlog"].append({"N": str(n), "g": str(g), "ni": str(ni), "json": pr.to_json(), "statement_json": sigma_json(N=n, g=g, ni=ni),
                            "verdict": verdict(lambda: pr.verify(n, g, ni))})
    return doc


if __name__ == "__main__":
    out = sys.argv[1]
    assert os.path.basename(out) != "reference_vectors.json" or "golden" not in os.path.abspath(out), "never write the mock over the golden file"
    json.dump(build(), open(out, "w"), indent=1)
    print("wrote", out)
