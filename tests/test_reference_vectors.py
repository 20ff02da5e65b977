"""The pin of the oracle - and of the CUDA path - against the REAL reference crate.

tests/golden/reference_vectors.json is produced by rust/gen_vectors, a small cargo project that links the unmodified
zk-paillier 0.4.4 / curv-kzen 0.10 / kzen-paillier 0.4.3 and dumps known-answer vectors for every rule SURVEY.md 8c lists
as RECALLED (BigInt::to_bytes, the transcript hash, curv's native BigInt serde, the EncryptionKey serde) plus one honest and
one dishonest proof of every kind with the reference's own verdict.  The build image has no Rust toolchain, so the file
cannot be produced here: while it is absent these tests are reported as xfail ("parity unpinned") and a mock in the same
schema, generated from the oracle (scripts/mock_reference_vectors.py), keeps the consuming code itself exercised.
Once a maintainer commits the real file, the same checks pin the oracle (CPU tests) and the engine (-m gpu tests)."""
import importlib.util
import json
import os

import pytest

from util import ROOT, po

GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_vectors.json")
UNPINNED = "parity unpinned: tests/golden/reference_vectors.json is absent (generate it with `cargo run --release` in rust/gen_vectors)"


def _mock(tmp_path_factory):
    spec = importlib.util.spec_from_file_location("mock_reference_vectors", os.path.join(ROOT, "scripts", "mock_reference_vectors.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


@pytest.fixture(scope="module", params=["reference", "mock"])
def vectors(request, tmp_path_factory):
    if request.param == "reference":
        if not os.path.exists(GOLDEN):
            pytest.xfail(UNPINNED)
        doc = json.load(open(GOLDEN))
        assert doc["generator"] != "oracle-mock", "the golden file must come from the reference crate, not from the oracle"
        return doc
    return _mock(tmp_path_factory)


def _verdict(fn):
    try:
        fn()
        return "ok"
    except po.IncorrectProof:
        return "incorrect"


def _native(d):
    return {k: po.serde_bigint_native_parse(v) for k, v in json.loads(d).items()}


def _compact(s):
    return json.dumps(json.loads(s), separators=(",", ":"))


# ------------------------------------------------------------------------------------------ oracle (CPU)
def test_oracle_to_bytes_and_digest(vectors):
    for v in vectors["to_bytes"]:
        assert po.bigint_to_bytes(int(v["dec"])).hex() == v["hex"], v["dec"][:40]
    for v in vectors["compute_digest"]:
        assert po.compute_digest([int(x) for x in v["items"]]) == int(v["digest"])


def test_oracle_serde_rules(vectors):
    for v in vectors["bigint_serde"]:
        assert json.dumps(po.serde_bigint_native(int(v["dec"]))) == v["json"], v["dec"][:40]
    ek = vectors["encryption_key"]
    assert json.dumps(po.serde_encryption_key(int(ek["n"])), separators=(",", ":")) == _compact(ek["json"])


def test_oracle_paillier_and_correct_key(vectors):
    n = int(vectors["key"]["n"])
    for v in vectors["paillier_enc"]:
        assert po.paillier_encrypt(n, int(v["m"]), int(v["r"])) == int(v["c"])
    for v in vectors["ni_correct_key"]:
        salt = bytes.fromhex(v["salt_hex"])
        pr = po.NiCorrectKeyProof.proof(int(v["p"]), int(v["q"]), salt)      # deterministic: bit for bit
        assert pr.to_json() == _compact(v["json"])
        assert _verdict(lambda: po.NiCorrectKeyProof.from_json(v["json"]).verify(n, salt)) == v["verdict"]
        assert _verdict(lambda: po.NiCorrectKeyProof.from_json(v["json"]).verify(n, bytes([1, 2, 3]))) == v["verdict_wrong_salt"]


def test_oracle_verdicts_and_reserialization(vectors):
    n = int(vectors["key"]["n"])
    for v in vectors["range_proof_ni"]:
        pr = po.RangeProofNi.from_json(v["json"])
        assert pr.to_json() == _compact(v["json"])                             # wire format, byte for byte
        assert _verdict(lambda: pr.verify(n, int(v["ciphertext"]))) == v["verdict"]
    for v in vectors["zero"]:
        f = _native(v["json"])
        assert _verdict(lambda: po.ZeroProof(f["z"], f["a"]).verify(n, int(v["c"]))) == v["verdict"]
        assert json.dumps({k: po.serde_bigint_native(x) for k, x in f.items()}, separators=(",", ":")) == _compact(v["json"])
    for v in vectors["ciphertext"]:
        f = _native(v["json"])
        assert _verdict(lambda: po.CiphertextProof(f["z1"], f["z2"], f["c_prime"]).verify(n, int(v["c"]))) == v["verdict"]
    for v in vectors["mul"]:
        f = _native(v["json"])
        assert _verdict(lambda: po.MulProof(f["f"], f["z1"], f["z2"], f["e_d"], f["e_db"]).verify(n, int(v["e_a"]), int(v["e_b"]), int(v["e_c"]))) == v["verdict"]
    for v in vectors["verlin"]:
        f = _native(v["json"])
        pr = po.VerlinProof(f["phi_a"], f["z"], f["z_prime"], f["z_double_prime"], f["r_z"])
        assert _verdict(lambda: pr.verify(n, int(v["c"]), int(v["c_prime"]), int(v["phi_x"]))) == v["verdict"]
    for v in vectors["dlog"]:
        pr = po.CompositeDLogProof.from_json(v["json"])
        assert pr.to_json() == _compact(v["json"])
        assert _verdict(lambda: pr.verify(int(v["N"]), int(v["g"]), int(v["ni"]))) == v["verdict"]
        assert _native(v["statement_json"]) == {"N": int(v["N"]), "g": int(v["g"]), "ni": int(v["ni"])}


# ------------------------------------------------------------------------------------------ engine (GPU)
@pytest.mark.gpu
def test_engine_against_the_vectors(vectors):
    """The CUDA path through the C++ mirror: same digests, ciphertexts, NiCorrectKeyProof bytes, verdicts and JSON."""
    import numpy as np

    import zk_paillier_b200 as zk
    from hostlib import call
    from zk_paillier_b200.native import ints_to_limbs, limbs_to_ints, to_limbs

    n = int(vectors["key"]["n"])
    with zk.native.Context(0) as ctx:
        for v in vectors["compute_digest"]:
            items = [int(x) for x in v["items"]]
            limbs = max(4, -(-max(x.bit_length() for x in items) // 128) * 4)
            dig = ctx.sha256_transcript(ints_to_limbs([items], limbs))
            assert int.from_bytes(bytes(dig[0]), "big") == int(v["digest"])
        ctx.set_key(to_limbs(n, 64))
        m = ints_to_limbs([int(v["m"]) % n for v in vectors["paillier_enc"]], 64)
        r = ints_to_limbs([int(v["r"]) for v in vectors["paillier_enc"]], 64)
        assert limbs_to_ints(ctx.paillier_enc(m, r)) == [int(v["c"]) for v in vectors["paillier_enc"]]
    for v in vectors["ni_correct_key"]:
        pr = call("correct_key_ni.proof", p=v["p"], q=v["q"], salt_hex=v["salt_hex"])
        assert pr["ok"] and pr["proof"] == _compact(v["json"])
        res = call("correct_key_ni.verify", proofs=[_compact(v["json"])], n=[str(n)], salt_hex=v["salt_hex"])
        assert res["results"] == [v["verdict"]]
    rp = vectors["range_proof_ni"]
    res = call("rangeproof_ni.verify", n=str(n), proofs=[_compact(v["json"]) for v in rp], ciphertexts=[v["ciphertext"] for v in rp])
    assert res["results"] == [v["verdict"] for v in rp]
    for name, keys in (("zero", ["c"]), ("ciphertext", ["c"]), ("mul", ["e_a", "e_b", "e_c"]), ("verlin", ["c", "c_prime", "phi_x"])):
        items = [dict({k: v[k] for k in keys}, proof=_compact(v["json"])) for v in vectors[name]]
        res = call(f"{name}.verify", n=str(n), items=items)
        assert res["ok"], res
        assert res["results"] == [v["verdict"] for v in vectors[name]], name
    res = call("dlog.verify", items=[{"N": v["N"], "g": v["g"], "ni": v["ni"], "proof": _compact(v["json"])} for v in vectors["dlog"]])
    assert res["ok"], res
    assert res["results"] == [v["verdict"] for v in vectors["dlog"]]
