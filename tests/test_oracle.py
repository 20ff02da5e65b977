"""CPU suite (-m "not gpu"): pins the oracles.

* the Python-int restatement against the fixtures the reference ships (test primes, primorial P, SALT_STRING),
  FIPS 180-4 SHA-256 vectors, the accept/reject classes of the reference's own tests,
  and the committed golden vectors (tests/golden/*.json, made by scripts/gen_golden.py);
* the C/GMP restatement (oracle/oracle.c, the timed CPU baseline) against the Python one, byte for byte;
* GMP mpz_powm against CPython pow() and OpenSSL BN_mod_exp (three independent bigints).
"""
import ctypes
import hashlib
import json
import os
import random
import re

import numpy as np
import pytest

from util import GOLDEN, ROOT, c_oracle, keys, limbs_for, po
from zk_paillier_b200.native import from_limbs, ints_to_limbs, limbs_to_ints, to_limbs
from zk_paillier_b200 import workload

N2048 = po.TEST_P * po.TEST_Q


def test_reference_fixtures():
    assert po.TEST_P.bit_length() == 1024 and po.TEST_Q.bit_length() == 1024 and N2048.bit_length() == 2048
    assert po.primorial().bit_length() == 9095
    assert po.SALT_STRING == b"KZen"
    assert len(po.SMALL_PRIMES) == 830 and po.SMALL_PRIMES[-1] == 6367
    assert (po.SECURITY_PARAMETER, po.M2, po.DIGEST_SIZE) == (128, 11, 256)
    ref = "/root/reference/src/zkproofs"
    if os.path.isdir(ref):  # only in the build container; the constants are also pinned above
        src = open(os.path.join(ref, "correct_key_ni.rs")).read()
        assert int(re.search(r'const P: &str = "(\d+)"', src).group(1)) == po.primorial()
        src = open(os.path.join(ref, "range_proof_ni.rs")).read()
        p, q = re.findall(r'from_str_radix\("(\d+)", 10\)', src)[:2]
        assert (int(p), int(q)) == (po.TEST_P, po.TEST_Q)


def test_to_bytes_rule_and_digest_kats():
    assert po.bigint_to_bytes(0) == b"\x00"
    assert po.bigint_to_bytes(255) == b"\xff" and po.bigint_to_bytes(256) == b"\x01\x00"
    # FIPS 180-4 "abc" through the BigInt route: from_bytes("abc") -> to_bytes -> SHA-256
    abc = int.from_bytes(b"abc", "big")
    assert po.compute_digest([abc]) == int("ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad", 16)
    # two-block message "abcdbcde...nopq" split over several items (plain concatenation, no length prefixes)
    msg = b"abcdbcdecdefdefgefghfghighijhijkijkljklmklmnlmnomnopnopq"
    parts = [int.from_bytes(msg[:10], "big"), int.from_bytes(msg[10:33], "big"), int.from_bytes(msg[33:], "big")]
    assert po.compute_digest(parts) == int("248d6a61d20638b8e5c026930c3e6039a33ce45964ff2167f6ecedd419db06c1", 16)
    # a leading zero byte inside an item is dropped: (0x00ab) hashes as (0xab)
    assert po.compute_digest([0xAB]) == int.from_bytes(hashlib.sha256(b"\xab").digest(), "big")


def test_challenge_bits_strip_leading_zero_bytes():
    e = po.bigint_to_bytes(int.from_bytes(b"\x00\x00\x80" + b"\x01" * 29, "big"))
    assert len(e) == 30 and po.challenge_bit(e, 0) == 1 and po.challenge_bit(e, 1) == 0
    with pytest.raises(po.ReferencePanic):
        po.challenge_bit(e, 240)


def _rp_case(seed, x_big=False, ef=16, n=N2048):
    rng = random.Random(seed)
    q = rng.getrandbits(256) | (1 << 255)
    third = q // 3
    x = (q * rng.randrange(100, 10000)) if x_big else rng.randrange(third)
    r = rng.randrange(n)
    c = po.paillier_encrypt(n, x, r)
    w1 = [rng.randrange(third, 2 * third) for _ in range(ef)]
    swap = [rng.getrandbits(1) for _ in range(ef)]
    r1 = [rng.randrange(n) for _ in range(ef)]
    r2 = [rng.randrange(n) for _ in range(ef)]
    return q, x, r, c, w1, swap, r1, r2


def test_range_proof_ni_accepts_and_rejects():
    # range_proof_ni.rs:163-179 (x < q/3 => Ok) and :181-199 (x in [100q, 10000q) => Err)
    q, x, r, c, w1, swap, r1, r2 = _rp_case(1)
    proof = po.RangeProofNi.prove(N2048, q, c, x, r, w1, swap, r1, r2)
    proof.verify(N2048, c)
    q, x, r, c, w1, swap, r1, r2 = _rp_case(2, x_big=True, ef=128)
    proof = po.RangeProofNi.prove(N2048, q, c, x, r, w1, swap, r1, r2)
    with pytest.raises(po.IncorrectProof):
        proof.verify(N2048, c)
    with pytest.raises(po.ReferencePanic):  # assert_eq!(ciphertext, self.ciphertext)
        proof.verify(N2048, c + 1)


def test_range_proof_serde_round_trip():
    q, x, r, c, w1, swap, r1, r2 = _rp_case(3, ef=8)
    proof = po.RangeProofNi.prove(N2048, q, c, x, r, w1, swap, r1, r2)
    s = proof.to_json()
    d = json.loads(s)
    assert list(d.keys()) == ["ek", "range", "ciphertext", "encrypted_pairs", "proof", "error_factor"]
    assert all(isinstance(v, str) and v.isdigit() for v in d["encrypted_pairs"]["c1"])
    kinds = {list(o.keys())[0] for o in d["proof"]}
    assert kinds <= {"Open", "Mask"}
    back = po.RangeProofNi.from_json(s)
    back.verify(N2048, c)
    assert back.to_json() == s


@pytest.mark.parametrize("salt", [po.SALT_STRING, b"Zen Go X", b"", b"\x00\x00Zen"])
def test_correct_key_round_trip(salt):
    # correct_key_ni.rs:126-138
    proof = po.NiCorrectKeyProof.proof(po.TEST_P, po.TEST_Q, salt)
    proof.verify(N2048, salt)
    assert po.NiCorrectKeyProof.from_json(proof.to_json()).sigma_vec == proof.sigma_vec
    bad = po.NiCorrectKeyProof([s for s in proof.sigma_vec])
    bad.sigma_vec[5] += 1
    with pytest.raises(po.IncorrectProof):
        bad.verify(N2048, salt)
    with pytest.raises(po.IncorrectProof):  # a modulus with a small prime factor fails the gcd test
        po.NiCorrectKeyProof.proof(3, po.TEST_Q, salt).verify(3 * po.TEST_Q, salt)


def test_sigma_protocols_accept_and_reject():
    p, q = keys(1024)[0]
    n = p * q
    nn = n * n
    rng = random.Random(7)
    rnd = lambda: rng.randrange(1, n)
    # ZeroProof (zero_enc_proof.rs:112-155)
    r = rnd()
    c = po.paillier_encrypt(n, 0, r)
    po.ZeroProof.prove(r, n, c, rnd()).verify(n, c)
    with pytest.raises(po.IncorrectProof):
        c1 = po.paillier_encrypt(n, 1, r)
        po.ZeroProof.prove(r, n, c1, rnd()).verify(n, c1)
    # CiphertextProof (correct_ciphertext.rs:116-163)
    x, r = rnd(), rnd()
    c = po.paillier_encrypt(n, x, r)
    po.CiphertextProof.prove(x, r, n, c, rnd(), rnd()).verify(n, c)
    with pytest.raises(po.IncorrectProof):
        po.CiphertextProof.prove(x, r + 1, n, c, rnd(), rnd()).verify(n, c)
    # MulProof (multiplication_proof.rs:172-272)
    a, b = rnd(), rnd()
    cc = a * b % n
    r_a, r_b, r_c = rnd(), rnd(), rnd()
    e_a, e_b, e_c = (po.paillier_encrypt(n, v, rr) for v, rr in ((a, r_a), (b, r_b), (cc, r_c)))
    po.MulProof.prove(a, b, cc, r_a, r_b, r_c, n, e_a, e_b, e_c, rnd(), rnd()).verify(n, e_a, e_b, e_c)
    with pytest.raises(po.IncorrectProof):
        e_bad = po.paillier_encrypt(n, cc + 1, r_c)
        po.MulProof.prove(a, b, cc + 1, r_a, r_b, r_c, n, e_a, e_b, e_bad, rnd(), rnd()).verify(n, e_a, e_b, e_bad)
    # VerlinProof (verlin_proof.rs:182-307)
    x, xp, xdp, r_x = rnd(), rnd(), rnd(), rnd()
    c, cp = po.paillier_encrypt(n, rnd(), rnd()), po.paillier_encrypt(n, rnd(), rnd())
    phi_x = po.gen_phi(n, c, cp, x, xp, xdp, r_x)
    po.VerlinProof.prove(x, xp, xdp, r_x, n, c, cp, phi_x, rnd(), rnd(), rnd(), rnd()).verify(n, c, cp, phi_x)
    with pytest.raises(po.IncorrectProof):
        po.VerlinProof.prove(x + 1, xp, xdp, r_x, n, c, cp, phi_x, rnd(), rnd(), rnd(), rnd()).verify(n, c, cp, phi_x)
    with pytest.raises(po.IncorrectProof):
        po.VerlinProof.prove(x, xp, xdp, r_x + 1, n, c, cp, phi_x, rnd(), rnd(), rnd(), rnd()).verify(n, c, cp, phi_x)
    # paillier_open recovers (m, r) (correct_opening.rs:49-56)
    m, r = rnd(), rnd()
    assert po.paillier_open(p, q, po.paillier_encrypt(n, m, r)) == (m, r)


def test_three_bigints_agree_on_modexp():
    """GMP mpz_powm (the reference's backend) == CPython pow == OpenSSL BN_mod_exp."""
    assert c_oracle.gmp_version().startswith("6.")
    crypto = ctypes.CDLL("libcrypto.so.3")
    crypto.BN_new.restype = ctypes.c_void_p
    crypto.BN_CTX_new.restype = ctypes.c_void_p
    crypto.BN_bin2bn.restype = ctypes.c_void_p
    crypto.BN_bin2bn.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p]
    crypto.BN_mod_exp.argtypes = [ctypes.c_void_p] * 5
    crypto.BN_bn2bin.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
    crypto.BN_num_bits.argtypes = [ctypes.c_void_p]
    bnctx = crypto.BN_CTX_new()

    def ossl_pow(b, e, m):
        conv = lambda v: crypto.BN_bin2bn(v.to_bytes((v.bit_length() + 7) // 8 or 1, "big"), (v.bit_length() + 7) // 8 or 1, None)
        r = crypto.BN_new()
        assert crypto.BN_mod_exp(r, conv(b), conv(e), conv(m), bnctx) == 1
        nb = (crypto.BN_num_bits(r) + 7) // 8
        buf = ctypes.create_string_buffer(nb or 1)
        crypto.BN_bn2bin(r, buf)
        return int.from_bytes(buf.raw[:nb], "big")

    rng = random.Random(11)
    for bits in (1024, 2048, 3072, 4096):
        l = bits // 32
        mods = [rng.getrandbits(bits) | 1 | (1 << (bits - 1)) for _ in range(3)]
        exps = [rng.getrandbits(bits) for _ in range(3)]
        bases = [rng.getrandbits(bits) for _ in range(3)]
        got = limbs_to_ints(c_oracle.modexp(ints_to_limbs(bases, l), ints_to_limbs(exps, l), ints_to_limbs(mods, l)))
        for b, e, m, g in zip(bases, exps, mods, got):
            assert g == pow(b, e, m) == ossl_pow(b, e, m)


def test_c_oracle_enc_and_sha_match_python():
    p, q = keys(1024)[1]
    n = p * q
    rng = random.Random(5)
    m = [rng.getrandbits(300) for _ in range(12)] + [0, n - 1]
    r = [rng.randrange(n) for _ in range(14)]
    out = limbs_to_ints(c_oracle.paillier_enc(to_limbs(n, 32), ints_to_limbs(m, 32), ints_to_limbs(r, 32)))
    assert out == [po.paillier_encrypt(n, a, b) for a, b in zip(m, r)]
    items = [[0, 1, 255, 256, rng.getrandbits(500), rng.getrandbits(24), 1 << 511, (1 << 512) - 1] for _ in range(5)]
    dig = c_oracle.sha256_transcript(ints_to_limbs(items, 16))
    for row, d in zip(items, dig):
        assert bytes(d) == hashlib.sha256(po.transcript_bytes(row)).digest()


def _py_prove(n, wl, nl, work, b, ef):
    """RangeProofNi::prove through the Python oracle for proof b of a workload."""
    w1 = limbs_to_ints(work["w1"][b])
    r1 = limbs_to_ints(work["r1"][b])
    r2 = limbs_to_ints(work["r2"][b])
    r = from_limbs(work["r"][b])
    x = work["x_int"][b]
    q = work["range_int"][b]
    c = po.paillier_encrypt(n, x, r)
    return po.RangeProofNi.prove(n, q, c, x, r, w1, [int(v) for v in work["swap"][b]], r1, r2), c


def test_c_oracle_rangeproof_matches_python():
    p, q = keys(1024)[0]
    n = p * q
    nl, ef, batch = 32, 16, 6
    work = workload.rangeproof_batch(n, batch, ef=ef, seed=42, reject_every=3)
    nlimbs = to_limbs(n, nl)
    got = c_oracle.rangeproof_ni_prove(nlimbs, ef, work["range"], work["x"], work["r"], work["w1"], work["swap"], work["r1"], work["r2"])
    cx = []
    for b in range(batch):
        proof, c = _py_prove(n, work["w_limbs"], nl, work, b, ef)
        cx.append(c)
        assert limbs_to_ints(got["c1"][b]) == proof.encrypted_pairs["c1"]
        assert limbs_to_ints(got["c2"][b]) == proof.encrypted_pairs["c2"]
        assert bytes(got["digest"][b]) == po.range_digest32(n, proof.encrypted_pairs)
        for i, resp in enumerate(proof.proof):
            k = int(got["kind"][b, i])
            if resp[0] == "Open":
                assert k == 0
                assert (from_limbs(got["resp_w"][b, i, 0]), from_limbs(got["resp_r"][b, i, 0]),
                        from_limbs(got["resp_w"][b, i, 1]), from_limbs(got["resp_r"][b, i, 1])) == resp[1:]
            else:
                assert k == resp[1]
                assert (from_limbs(got["resp_w"][b, i, 0]), from_limbs(got["resp_r"][b, i, 0])) == resp[2:]
        ok = True
        try:
            proof.verify(n, c)
        except po.IncorrectProof:
            ok = False
        assert ok == (b % 3 != 2)
    acc, fault, dig, encs = c_oracle.rangeproof_ni_verify(nlimbs, ef, work["range"], ints_to_limbs(cx, 2 * nl), got["c1"], got["c2"],
                                                          got["kind"], got["resp_w"], got["resp_r"])
    assert acc.tolist() == [1, 1, 0, 1, 1, 0] and not fault.any()
    assert (dig == got["digest"]).all()
    assert encs == int((got["kind"] == 0).sum()) + batch * ef


def test_c_oracle_correct_key_matches_python():
    ks = keys(1024)[:3]
    salt = b"Zen Go X"
    work = workload.correct_key_batch(ks, 5, salt, lambda p, q, s: po.NiCorrectKeyProof.proof(p, q, s).sigma_vec, 32, bad_every=4)
    acc, rho = c_oracle.correct_key_ni_verify(work["n"], work["sigma"], salt)
    assert acc.tolist() == [1, 1, 1, 0, 1]
    for b in range(5):
        assert limbs_to_ints(rho[b]) == po.correct_key_rho(work["n_int"][b], salt)
    # a modulus divisible by a small prime is rejected by the gcd test even with "matching" sigma
    n_bad = 6367 * ks[0][1]
    sig = [1] * 11
    acc, _ = c_oracle.correct_key_ni_verify(ints_to_limbs([n_bad], 32), ints_to_limbs([sig], 32), salt)
    assert acc.tolist() == [0]


def test_golden_vectors():
    """Committed vectors (scripts/gen_golden.py ran the Python oracle): both oracles must still reproduce them."""
    g = json.load(open(os.path.join(GOLDEN, "vectors.json")))
    n = int(g["n"])
    for v in g["enc"]:
        assert po.paillier_encrypt(n, int(v["m"]), int(v["r"])) == int(v["c"])
    for v in g["digest"]:
        assert "%064x" % po.compute_digest([int(x) for x in v["items"]]) == v["sha256"]
    rp = g["range_proof_ni"]
    w1, r1, r2 = ([int(x) for x in rp[k]] for k in ("w1", "r1", "r2"))
    proof = po.RangeProofNi.prove(n, int(rp["range"]), int(rp["ciphertext"]), int(rp["x"]), int(rp["r"]), w1, rp["swap"], r1, r2)
    assert hashlib.sha256(proof.to_json().encode()).hexdigest() == rp["proof_json_sha256"]
    assert po.range_digest32(n, proof.encrypted_pairs).hex() == rp["challenge_digest"]
    ck = g["correct_key_ni"]
    assert [str(s) for s in po.NiCorrectKeyProof.proof(int(ck["p"]), int(ck["q"]), bytes.fromhex(ck["salt_hex"])).sigma_vec] == ck["sigma_vec"]
    assert [str(s) for s in po.correct_key_rho(int(ck["p"]) * int(ck["q"]), bytes.fromhex(ck["salt_hex"]))] == ck["rho_vec"]


def test_interactive_range_proof_oracles_agree():
    """The interactive RangeProof (range_proof.rs:431-525, error factor 40): the verifier's ChallengeBits bytes are used
    as they are.  C oracle vs the Python restatement of generate_proof / verifier_output."""
    p, q = keys(1024)[0]
    n = p * q
    nl, ef, batch = 32, 40, 4
    work = workload.rangeproof_batch(n, batch, ef=ef, seed=40, reject_every=4)
    chal = np.frombuffer(random.Random(1).randbytes(batch * 5), np.uint8).reshape(batch, 5).copy()
    chal[1] = 0            # all Open, first byte zero: would be stripped if it went through a BigInt
    nlimbs = to_limbs(n, nl)
    got = c_oracle.rangeproof_ni_prove(nlimbs, ef, work["range"], work["x"], work["r"], work["w1"], work["swap"], work["r1"], work["r2"],
                                       challenge=chal)
    cx = []
    for b in range(batch):
        w1 = limbs_to_ints(work["w1"][b]); r1 = limbs_to_ints(work["r1"][b]); r2 = limbs_to_ints(work["r2"][b])
        r, x, rg = from_limbs(work["r"][b]), work["x_int"][b], work["range_int"][b]
        pairs, data = po.generate_encrypted_pairs(n, rg, w1, [int(v) for v in work["swap"][b]], r1, r2)
        e = bytes(chal[b])
        proof = po.generate_proof(n, x, r, e, rg, data, ef)
        c = po.paillier_encrypt(n, x, r)
        cx.append(c)
        assert limbs_to_ints(got["c1"][b]) == pairs["c1"]
        assert [int(k) for k in got["kind"][b]] == [0 if t[0] == "Open" else t[1] for t in proof]
        bits = po.verifier_output_bits(n, e, pairs, proof, rg, c, ef)
        assert all(bits) == (b != 3)
    assert (got["kind"][1] == 0).all()
    acc, fault, _, _ = c_oracle.rangeproof_ni_verify(nlimbs, ef, work["range"], ints_to_limbs(cx, 2 * nl), got["c1"], got["c2"], got["kind"],
                                                     got["resp_w"], got["resp_r"], challenge=chal)
    assert acc.tolist() == [1, 1, 1, 0] and not fault.any()
    # a different challenge than the one answered -> variant mismatch -> reject
    chal2 = chal.copy()
    chal2[0, 0] ^= 0x80
    acc, _, _, _ = c_oracle.rangeproof_ni_verify(nlimbs, ef, work["range"], ints_to_limbs(cx, 2 * nl), got["c1"], got["c2"], got["kind"],
                                                 got["resp_w"], got["resp_r"], challenge=chal2)
    assert acc.tolist() == [0, 1, 1, 0]
    # challenge shorter than error_factor bits: index out of range panics in the reference
    acc, fault, _, _ = c_oracle.rangeproof_ni_verify(nlimbs, ef, work["range"], ints_to_limbs(cx, 2 * nl), got["c1"], got["c2"], got["kind"],
                                                     got["resp_w"], got["resp_r"], challenge=chal[:, :4])
    assert fault.all() and not acc.any()


def test_correct_key_interactive_oracle():
    # correct_key.rs:199-232: honest key accepts; a tampered challenge is refused by the prover
    p, q = keys(1024)[0]
    n = p * q
    rng = random.Random(40)
    s = [rng.randrange(1, n) for _ in range(40)]
    r = [rng.randrange(1, n) for _ in range(40)]
    ch, va = po.CorrectKey.challenge(n, s, r)
    proof = po.CorrectKey.prove(p, q, ch)
    po.CorrectKey.verify(proof, va)
    bad = dict(ch, e=ch["e"] + 1)
    with pytest.raises(po.CorrectKeyProveError):
        po.CorrectKey.prove(p, q, bad)
    with pytest.raises(po.IncorrectProof):
        po.CorrectKey.verify({"s_digest": proof["s_digest"] + 1}, va)
