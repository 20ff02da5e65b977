"""GPU parity (-m gpu) of the remaining public proofs through the C ABI against the Python-int oracle (SURVEY.md section 8,
row f3): CorrectOpening::verify_opening, CompositeDLogProof prove / verify, CorrectMessageProof prove / verify --
bit-exact proof fields on identical randomness, and accept / reject / panic (fault) as the reference's own tests
classify them (correct_opening.rs:47-56, wi_dlog_proof.rs:112-196, correct_message.rs:170-197)."""
import random

import numpy as np
import pytest

from test_oracle_more import dlog_statement
from util import keys, limbs_for, po
from zk_paillier_b200.native import ints_to_limbs, limbs_to_ints, to_limbs

pytestmark = pytest.mark.gpu


def verdict(fn):
    """1 accept, 0 Err(IncorrectProof), -1 panic."""
    try:
        fn()
        return 1
    except po.IncorrectProof:
        return 0
    except po.ReferencePanic:
        return -1


@pytest.fixture(scope="module", params=[1024, 2048])
def keyed(request, ctx):
    bits = request.param
    p, q = keys(bits)[0]
    n = p * q
    nl = limbs_for(bits)
    ctx.set_key(to_limbs(n, nl))
    return ctx, p, q, n, nl, random.Random(bits + 7)


def test_verify_opening(keyed):
    ctx, p, q, n, nl, rng = keyed
    B = 6
    m = [rng.randrange(n) for _ in range(B)]
    r = [rng.randrange(1, n) for _ in range(B)]
    c = [po.paillier_encrypt(n, mi, ri) for mi, ri in zip(m, r)]
    m_claim, r_claim = list(m), list(r)
    m_claim[1] = (m[1] + 1) % n
    r_claim[4] = (r[4] + 1) % n
    ok = ctx.verify_opening(ints_to_limbs(m_claim, nl), ints_to_limbs(r_claim, nl), ints_to_limbs(c, 2 * nl))
    assert ok.tolist() == [int(po.verify_opening(n, a, b, cc)) for a, b, cc in zip(m_claim, r_claim, c)] == [1, 0, 1, 1, 0, 1]
    mo, ro = po.paillier_open(p, q, c[0])              # correct_opening.rs:50-55: open, then verify the opening
    assert ctx.verify_opening(ints_to_limbs([mo], nl), ints_to_limbs([ro], nl), ints_to_limbs(c[:1], 2 * nl)).tolist() == [1]


@pytest.mark.parametrize("bits", [1024, 2048, 3072])
def test_composite_dlog_proof(ctx, bits):
    rng = random.Random(bits)
    ks = keys(bits)
    nl = limbs_for(bits)
    R = 1 << (po.DLOG_K + po.DLOG_K_PRIME + po.DLOG_SAMPLE_S)
    kinds = ["good", "plus", "good", "random", "good", "good"]
    st = [dlog_statement(rng, *ks[i % len(ks)], kind) for i, kind in enumerate(kinds)]   # a distinct N per statement where fixtures allow
    N, g, ni, s = (list(v) for v in zip(*st))
    r = [rng.randrange(R) for _ in st]
    L = lambda v, w: ints_to_limbs(v, w)
    x, y, fault = ctx.dlog_prove(L(N, nl), L(g, nl), L(ni, nl), L(s, 8), L(r, 16), 20)
    want = [po.CompositeDLogProof.prove(N[i], g[i], ni[i], s[i], r[i]) for i in range(len(st))]
    assert limbs_to_ints(x) == [w.x for w in want] and limbs_to_ints(y) == [w.y for w in want] and not fault.any()
    acc, fault = ctx.dlog_verify(L(N, nl), L(g, nl), L(ni, nl), x, y)
    assert acc.tolist() == [verdict(lambda i=i: want[i].verify(N[i], g[i], ni[i])) for i in range(len(st))] == [1, 0, 1, 0, 1, 1]
    assert not fault.any()
    # the reference's asserts: g not coprime to N, ni not coprime to N, N <= 2^128
    p0 = ks[0][0]
    g2, ni2, N2 = list(g), list(ni), list(N)
    g2[0] = p0 * 3
    ni2[2] = ks[2 % len(ks)][1]
    N2[4] = (1 << 128) - 159
    acc, fault = ctx.dlog_verify(L(N2, nl), L(g2, nl), L(ni2, nl), x, y)
    expect = [verdict(lambda i=i: want[i].verify(N2[i], g2[i], ni2[i])) for i in range(len(st))]
    assert expect == [-1, 0, -1, 0, -1, 1]
    assert fault.tolist() == [int(e == -1) for e in expect] and acc.tolist() == [int(e == 1) for e in expect]
    # y too wide for the caller's rows -> fault, not a silent truncation
    _, _, fault = ctx.dlog_prove(L(N, nl), L(g, nl), L(ni, nl), L([(1 << 256) - 1] * len(st), 8), L([R - 1] * len(st), 16), 16)
    assert fault.all()


def test_correct_message_proof(keyed):
    ctx, p, q, n, nl, rng = keyed
    nnl = 2 * nl
    valid = [3, 4, 5, 6]                                 # correct_message.rs:172-179
    M = len(valid)
    msgs = [4, 3, 6, 7, 5, 4]                            # statement 3: test_bad_message_zk_proof (7 is not valid)
    B = len(msgs)
    r = [rng.randrange(1, n) for _ in range(B)]
    w = [rng.randrange(1, n) for _ in range(B)]
    e_rand = [[rng.getrandbits(256) for _ in range(M - 1)] for _ in range(B)]
    z_rand = [[rng.randrange(1, n) for _ in range(M - 1)] for _ in range(B)]
    L = lambda v, wd: ints_to_limbs(v, wd)
    out = ctx.correct_message_prove(L([valid] * B, 4), L(msgs, 4), L(r, nl), L(e_rand, 8), L(z_rand, nl), L(w, nl))
    want = []
    for b in range(B):
        try:
            want.append(po.CorrectMessageProof.prove(n, valid, msgs[b], r[b], e_rand[b], z_rand[b], w[b]))
        except po.ReferencePanic:
            want.append(None)
    assert out["fault"].tolist() == [int(wb is None) for wb in want] == [0, 0, 0, 1, 0, 0]
    good = [b for b in range(B) if want[b] is not None]
    for b in good:
        assert limbs_to_ints(out["ciphertext"][b:b + 1]) == [want[b].ciphertext]
        assert limbs_to_ints(out["e_vec"][b]) == want[b].e_vec
        assert limbs_to_ints(out["z_vec"][b]) == want[b].z_vec
        assert limbs_to_ints(out["a_vec"][b]) == want[b].a_vec
    sel = np.array(good)
    validL = L([valid] * len(good), 4)
    acc, fault = ctx.correct_message_verify(out["ciphertext"][sel], validL, out["e_vec"][sel], out["z_vec"][sel], out["a_vec"][sel])
    assert acc.tolist() == [1] * len(good) and not fault.any()
    # tamper: a wrong response (Err), a wrong challenge share (assert_eq! panic), a ciphertext of an invalid message (Err)
    e2, z2, c2 = out["e_vec"][sel].copy(), out["z_vec"][sel].copy(), out["ciphertext"][sel].copy()
    z2[0, 1, 0] ^= 1
    e2[1, 2, 0] ^= 1
    c2[2] = to_limbs(po.paillier_encrypt(n, 7, 12345), nnl)
    acc, fault = ctx.correct_message_verify(c2, validL, e2, z2, out["a_vec"][sel])
    expect = []
    for k, b in enumerate(good):
        pr = po.CorrectMessageProof(limbs_to_ints(e2[k]), limbs_to_ints(z2[k]), want[b].a_vec, limbs_to_ints(c2[k:k + 1])[0], valid, n)
        expect.append(verdict(pr.verify))
    assert expect == [0, -1, 0, 1, 1]
    assert acc.tolist() == [int(e == 1) for e in expect] and fault.tolist() == [int(e == -1) for e in expect]


def test_correct_message_single_and_wide(keyed):
    """M = 1 (no simulated branch) and messages as wide as n."""
    ctx, p, q, n, nl, rng = keyed
    big = [n - 2, n // 3, 1]
    B = 2
    r = [rng.randrange(1, n) for _ in range(B)]
    w = [rng.randrange(1, n) for _ in range(B)]
    L = lambda v, wd: ints_to_limbs(v, wd)
    out = ctx.correct_message_prove(L([[5]] * B, 4), L([5] * B, 4), L(r, nl), np.zeros((B, 0, 8), np.uint32), np.zeros((B, 0, nl), np.uint32), L(w, nl))
    for b in range(B):
        wp = po.CorrectMessageProof.prove(n, [5], 5, r[b], [], [], w[b])
        assert limbs_to_ints(out["e_vec"][b]) == wp.e_vec and limbs_to_ints(out["z_vec"][b]) == wp.z_vec and limbs_to_ints(out["a_vec"][b]) == wp.a_vec
    acc, fault = ctx.correct_message_verify(out["ciphertext"], L([[5]] * B, 4), out["e_vec"], out["z_vec"], out["a_vec"])
    assert acc.tolist() == [1, 1] and not fault.any()
    e_rand = [[rng.getrandbits(256) for _ in range(2)] for _ in range(B)]
    z_rand = [[rng.randrange(1, n) for _ in range(2)] for _ in range(B)]
    out = ctx.correct_message_prove(L([big] * B, nl), L([big[0], big[2]], nl), L(r, nl), L(e_rand, 8), L(z_rand, nl), L(w, nl))
    for b, msg in enumerate((big[0], big[2])):
        wp = po.CorrectMessageProof.prove(n, big, msg, r[b], e_rand[b], z_rand[b], w[b])
        assert limbs_to_ints(out["a_vec"][b]) == wp.a_vec and limbs_to_ints(out["e_vec"][b]) == wp.e_vec
    acc, fault = ctx.correct_message_verify(out["ciphertext"], L([big] * B, nl), out["e_vec"], out["z_vec"], out["a_vec"])
    assert acc.tolist() == [1, 1] and not fault.any()
