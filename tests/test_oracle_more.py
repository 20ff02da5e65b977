"""CPU suite (-m "not gpu") for the oracle restatements of the remaining public proofs:
CorrectOpening, CompositeDLogProof, CorrectMessageProof -- the accept / reject / panic classes of the reference's own
tests (correct_opening.rs:47-56, wi_dlog_proof.rs:112-196, correct_message.rs:170-197)."""
import random

import pytest

from util import keys, po


def jacobi_minus_one_base(rng, p, q):
    """h1 with Jacobi symbol -1, as wi_dlog_proof.rs:121-126 picks it (legendre_symbol :93-105)."""
    leg = lambda a, pr: 1 if pow(a, (pr - 1) // 2, pr) == 1 else -1
    n = p * q
    while True:
        h1 = rng.randrange(1, n - 1)
        if leg(h1, p) * leg(h1, q) == -1:
            return h1


def dlog_statement(rng, p, q, kind="good"):
    n = p * q
    h1 = jacobi_minus_one_base(rng, p, q)
    secret = rng.randrange(1 << po.DLOG_SAMPLE_S)
    if kind == "good":
        h2 = pow(pow(h1, -1, n), secret, n)        # wi_dlog_proof.rs:128-129
    elif kind == "plus":
        h2 = pow(h1, secret, n)                    # test_bad_dlog_proof: +secret instead of -secret
    else:
        h2 = rng.randrange(1, n - 1)               # test_bad_dlog_proof_2: random ni
    return n, h1, h2, secret


def test_verify_opening():
    p, q = keys(2048)[0]
    n = p * q
    r = 0x1234567 * 3 + 1
    c = po.paillier_encrypt(n, 10, r)
    m2, r2 = po.paillier_open(p, q, c)             # correct_opening.rs:50-55
    assert (m2, r2 % n) == (10, r % n) and po.verify_opening(n, m2, r2, c)
    assert not po.verify_opening(n, 11, r2, c) and not po.verify_opening(n, 10, r2 + 1, c)


@pytest.mark.parametrize("bits", [1024, 2048])
def test_dlog_proof_classes(bits):
    rng = random.Random(bits)
    p, q = keys(bits)[0]
    R = 1 << (po.DLOG_K + po.DLOG_K_PRIME + po.DLOG_SAMPLE_S)
    N, g, ni, s = dlog_statement(rng, p, q, "good")
    proof = po.CompositeDLogProof.prove(N, g, ni, s, rng.randrange(R))
    proof.verify(N, g, ni)
    assert po.CompositeDLogProof.from_json(proof.to_json()).__dict__ == proof.__dict__
    for kind in ("plus", "random"):
        N, g, ni, s = dlog_statement(rng, p, q, kind)
        with pytest.raises(po.IncorrectProof):
            po.CompositeDLogProof.prove(N, g, ni, s, rng.randrange(R)).verify(N, g, ni)
    with pytest.raises(po.ReferencePanic):       # g shares a factor with N: assert_eq!(gcd, 1)
        po.CompositeDLogProof(1, 1).verify(N, p, ni)
    with pytest.raises(po.ReferencePanic):       # N <= 2^K
        po.CompositeDLogProof(1, 1).verify((1 << 128) - 159, 3, 5)


@pytest.mark.parametrize("bits", [1024, 2048])
def test_correct_message_proof_classes(bits):
    rng = random.Random(bits + 1)
    p, q = keys(bits)[0]
    n = p * q
    valid = [3, 4, 5, 6]                           # correct_message.rs:172-179
    rand = lambda: dict(r=rng.randrange(1, n), e_rand=[rng.getrandbits(256) for _ in valid[1:]],
                        z_rand=[rng.randrange(1, n) for _ in valid[1:]], w=rng.randrange(1, n))
    for msg in valid:
        proof = po.CorrectMessageProof.prove(n, valid, msg, **rand())
        proof.verify()
        assert sum(proof.e_vec) % (1 << 256) == po.compute_digest(proof.a_vec)
    with pytest.raises(po.ReferencePanic):       # test_bad_message_zk_proof: 7 is not a valid message
        po.CorrectMessageProof.prove(n, valid, 7, **rand())
    proof = po.CorrectMessageProof.prove(n, valid, 5, **rand())
    proof.z_vec[1] = (proof.z_vec[1] + 1) % n
    with pytest.raises(po.IncorrectProof):
        proof.verify()
    proof = po.CorrectMessageProof.prove(n, valid, 5, **rand())
    proof.e_vec[0] ^= 1                            # breaks chal == sum e: assert_eq! panics
    with pytest.raises(po.ReferencePanic):
        proof.verify()
    # a ciphertext of a message outside the list cannot be proven even by a cheating prover who claims slot 0
    forged = po.CorrectMessageProof.prove(n, valid, 3, **rand())
    forged.ciphertext = po.paillier_encrypt(n, 7, rng.randrange(1, n))
    with pytest.raises(po.IncorrectProof):
        forged.verify()


def test_golden_vectors_more():
    """Committed vectors of the row-f3 proofs (scripts/gen_golden.py ran the Python oracle): pins it against regressions."""
    import hashlib
    import json
    import os

    from util import GOLDEN

    g = json.load(open(os.path.join(GOLDEN, "vectors_more.json")))
    n = int(g["n"])
    d = g["dlog"]
    pr = po.CompositeDLogProof.prove(n, int(d["g"]), int(d["ni"]), int(d["secret"]), int(d["r"]))
    assert (pr.x, pr.y, pr.to_json()) == (int(d["x"]), int(d["y"]), d["json"])
    pr.verify(n, int(d["g"]), int(d["ni"]))
    c = g["correct_message"]
    cm = po.CorrectMessageProof.prove(n, c["valid"], c["message"], int(c["r"]), [int(v) for v in c["e_rand"]], [int(v) for v in c["z_rand"]], int(c["w"]))
    assert cm.ciphertext == int(c["ciphertext"]) and cm.e_vec == [int(v) for v in c["e_vec"]] and cm.z_vec == [int(v) for v in c["z_vec"]]
    assert hashlib.sha256(",".join(str(v) for v in cm.a_vec).encode()).hexdigest() == c["a_vec_sha256"]
    cm.verify()


def _verdict(fn):
    try:
        fn()
        return 1
    except po.IncorrectProof:
        return 0
    except po.ReferencePanic:
        return 2


def test_gmp_verifiers_agree_with_the_python_oracle():
    """oracle/oracle.c restates MulProof / VerlinProof / CompositeDLogProof / CorrectMessageProof verify on GMP; the two
    restatements must give the same Ok / Err / panic verdict on honest, tampered and panicking inputs."""
    import numpy as np

    from util import c_oracle
    from zk_paillier_b200.native import ints_to_limbs, to_limbs

    rng = random.Random(77)
    p, q = keys(1024)[1]
    n = p * q
    nn, nl, nnl, zl = n * n, 32, 64, 44
    L = ints_to_limbs
    rnd = lambda: rng.randrange(1, n)
    # MulProof: honest, c != a b, e_db = 0 (no inverse: unwrap panics)
    proofs, st = [], []
    for kind in ("ok", "bad", "panic"):
        a, b = rnd(), rnd()
        c = a * b % n if kind != "bad" else (a * b + 1) % n
        r_a, r_b, r_c = rnd(), rnd(), rnd()
        e_a, e_b, e_c = (po.paillier_encrypt(n, v, r) for v, r in ((a, r_a), (b, r_b), (c, r_c)))
        pr = po.MulProof.prove(a, b, c, r_a, r_b, r_c, n, e_a, e_b, e_c, rnd(), rnd())
        if kind == "panic":
            pr.e_db = 0
        proofs.append(pr)
        st.append((e_a, e_b, e_c))
    want = [_verdict(lambda pr=pr, s=s: pr.verify(n, *s)) for pr, s in zip(proofs, st)]
    got = c_oracle.mul_verify(to_limbs(n, nl), L([s[0] for s in st], nnl), L([s[1] for s in st], nnl), L([s[2] for s in st], nnl),
                              L([pr.f for pr in proofs], nl), L([pr.z1 for pr in proofs], nnl), L([pr.z2 for pr in proofs], nnl),
                              L([pr.e_d for pr in proofs], nnl), L([pr.e_db for pr in proofs], nnl), 2)
    assert got.tolist() == want == [1, 0, 2]
    # VerlinProof: honest, wrong x
    proofs, st = [], []
    for kind in ("ok", "bad"):
        x, xp, xdp, r_x = rnd(), rnd(), rnd(), rnd()
        c, cp = po.paillier_encrypt(n, rnd(), rnd()), po.paillier_encrypt(n, rnd(), rnd())
        phi_x = po.gen_phi(n, c, cp, x, xp, xdp, r_x)
        pr = po.VerlinProof.prove(x + (kind == "bad"), xp, xdp, r_x, n, c, cp, phi_x, rnd(), rnd(), rnd(), rnd())
        proofs.append(pr)
        st.append((c, cp, phi_x))
    want = [_verdict(lambda pr=pr, s=s: pr.verify(n, *s)) for pr, s in zip(proofs, st)]
    got = c_oracle.verlin_verify(to_limbs(n, nl), L([s[0] for s in st], nnl), L([s[1] for s in st], nnl), L([s[2] for s in st], nnl),
                                 L([pr.phi_a for pr in proofs], nnl), L([pr.z for pr in proofs], zl), L([pr.z_prime for pr in proofs], zl),
                                 L([pr.z_double_prime for pr in proofs], zl), L([pr.r_z for pr in proofs], nnl), 2)
    assert got.tolist() == want == [1, 0]
    # CompositeDLogProof: good, +secret, g | N shares a factor, N <= 2^128
    R = 1 << 512
    rows = []
    for kind in ("good", "plus", "good", "good"):
        N, g, ni, s = dlog_statement(rng, p, q, kind)
        pr = po.CompositeDLogProof.prove(N, g, ni, s, rng.randrange(R))
        rows.append([N, g, ni, pr.x, pr.y])
    rows[2][1] = p * 7
    rows[3][0] = (1 << 128) - 159
    want = [_verdict(lambda r=r: po.CompositeDLogProof(r[3], r[4]).verify(r[0], r[1], r[2])) for r in rows]
    got = c_oracle.dlog_verify(*(L([r[k] for r in rows], w) for k, w in enumerate((nl, nl, nl, nl, 20))), 2)
    assert got.tolist() == want == [1, 0, 2, 2]
    # CorrectMessageProof: honest, tampered z (Err), tampered e (assert_eq! panics)
    valid = [3, 4, 5, 6]
    prs = [po.CorrectMessageProof.prove(n, valid, 3 + k, rnd(), [rng.getrandbits(256) for _ in range(3)], [rnd() for _ in range(3)], rnd()) for k in range(3)]
    prs[1].z_vec[0] = (prs[1].z_vec[0] + 1) % n
    prs[2].e_vec[3] ^= 2
    want = [_verdict(pr.verify) for pr in prs]
    got = c_oracle.correct_message_verify(to_limbs(n, nl), L([pr.ciphertext for pr in prs], nnl), L([valid] * 3, 4), L([pr.e_vec for pr in prs], 8),
                                          L([pr.z_vec for pr in prs], nl), L([pr.a_vec for pr in prs], nnl), 2)
    assert got.tolist() == want == [1, 0, 2]
    assert isinstance(got, np.ndarray)


def test_c_oracle_provers_match_the_python_oracle():
    """The GMP provers behind bench.py's CPU baselines (ZeroProof prove / verify, CompositeDLogProof::prove,
    CorrectMessageProof::prove) against the Python-int restatement, field for field on identical randomness."""
    import numpy as np

    import c_oracle
    from zk_paillier_b200.native import ints_to_limbs, limbs_to_ints, to_limbs

    p, q = keys(1024)[1]
    n = p * q
    nl = 32
    rng = random.Random(5)
    B = 5
    # ZeroProof; statement 3 encrypts 1
    r = [rng.randrange(1, n) for _ in range(B)]
    c = [po.paillier_encrypt(n, 1 if i == 3 else 0, ri) for i, ri in enumerate(r)]
    rp = [rng.randrange(1, n) for _ in range(B)]
    z, a = c_oracle.zero_prove(to_limbs(n, nl), ints_to_limbs(r, nl), ints_to_limbs(c, 2 * nl), ints_to_limbs(rp, nl))
    want = [po.ZeroProof.prove(ri, n, ci, rpi) for ri, ci, rpi in zip(r, c, rp)]
    assert limbs_to_ints(z) == [w.z for w in want] and limbs_to_ints(a) == [w.a for w in want]
    assert c_oracle.zero_verify(to_limbs(n, nl), ints_to_limbs(c, 2 * nl), z, a).tolist() == [1, 1, 1, 0, 1]
    # CompositeDLogProof::prove
    st = [dlog_statement(rng, p, q) for _ in range(B)]
    rr = [rng.getrandbits(512) for _ in range(B)]
    x, y = c_oracle.dlog_prove(ints_to_limbs([s[0] for s in st], nl), ints_to_limbs([s[1] for s in st], nl), ints_to_limbs([s[2] for s in st], nl),
                               ints_to_limbs([s[3] for s in st], 8), ints_to_limbs(rr, 16), 20)
    want = [po.CompositeDLogProof.prove(s[0], s[1], s[2], s[3], ri) for s, ri in zip(st, rr)]
    assert limbs_to_ints(x) == [w.x for w in want] and limbs_to_ints(y) == [w.y for w in want]
    # CorrectMessageProof::prove, the message in every position of the ring
    M = 4
    valid = [3, 4, 5, 6]
    msgs = [valid[i % M] for i in range(B)]
    r = [rng.randrange(1, n) for _ in range(B)]
    w = [rng.randrange(1, n) for _ in range(B)]
    e_rand = [[rng.getrandbits(256) for _ in range(M - 1)] for _ in range(B)]
    z_rand = [[rng.randrange(1, n) for _ in range(M - 1)] for _ in range(B)]
    out = c_oracle.correct_message_prove(to_limbs(n, nl), ints_to_limbs([valid] * B, 4), ints_to_limbs(msgs, 4), ints_to_limbs(r, nl),
                                         ints_to_limbs(e_rand, 8), ints_to_limbs(z_rand, nl), ints_to_limbs(w, nl))
    for i in range(B):
        pr = po.CorrectMessageProof.prove(n, valid, msgs[i], r[i], e_rand[i], z_rand[i], w[i])
        assert limbs_to_ints(out["a_vec"][i]) == pr.a_vec and limbs_to_ints(out["e_vec"][i]) == pr.e_vec
        assert limbs_to_ints(out["z_vec"][i]) == pr.z_vec and limbs_to_ints(out["ciphertext"][i:i + 1])[0] == pr.ciphertext
    v = c_oracle.correct_message_verify(to_limbs(n, nl), out["ciphertext"], ints_to_limbs([valid] * B, 4), out["e_vec"], out["z_vec"], out["a_vec"])
    assert (np.asarray(v) == 1).all()
