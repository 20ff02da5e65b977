"""GPU parity (-m gpu) of the sigma-protocol proofs through the C ABI against the Python-int oracle:
ZeroProof, CiphertextProof, MulProof, VerlinProof prove (bit-exact proof fields on identical randomness)
and verify (accept / reject as the reference's own tests classify them: zero_enc_proof.rs:112-155,
correct_ciphertext.rs:116-163, multiplication_proof.rs:172-272, verlin_proof.rs:182-307)."""
import random

import numpy as np
import pytest

from util import keys, limbs_for, po
from zk_paillier_b200.native import from_limbs, ints_to_limbs, limbs_to_ints, to_limbs

pytestmark = pytest.mark.gpu


def verdict(fn):
    try:
        fn()
        return 1
    except po.IncorrectProof:
        return 0


# The fixture below runs every test once per (layout, row form) of the same key; the statements come from a generator
# seeded by the key size, so the Python-int oracle's side (seconds per proof at 4096 bits) is computed once per key size.
_ORACLE = {}


def once(key, fn):
    if key not in _ORACLE:
        _ORACLE[key] = fn()
    return _ORACLE[key]


# (key size, K2h lane layout, row form)
#   layout: 0 = by job count (the one-job-per-warp latency layout at these batch sizes), 1 = wide lanes, 2 = narrow lanes,
#           "k1" = the two-digit kernels switched off: every modexp by the single-purpose K1 / K2 launches, products by K3
#   rows:   0 = the layout's default (pair rows in the narrow / latency layouts), 1 = single rows, 2 = pair rows
@pytest.fixture(scope="module", params=[(1024, 0, 0), (1024, 1, 0), (1024, "k1", 0), (1024, 0, 1), (2048, 1, 0), (2048, 2, 0), (2048, 2, 1), (2048, 0, 0),
                                        (3072, 1, 0), (3072, 2, 0), (3072, 2, 1), (4096, 1, 0), (4096, 2, 0), (4096, 2, 1)],
                ids=lambda p: f"{p[0]}-shape{p[1]}-rows{p[2]}")
def keyed(request, ctx):
    import zk_paillier_b200 as zk

    bits, shape, rows = request.param
    p, q = keys(bits)[0]
    n = p * q
    nl = limbs_for(bits)
    ctx.tune(zk.native.TUNE_ENC_KERNEL, 1 if shape == "k1" else 0)
    ctx.set_key(to_limbs(n, nl))
    ctx.tune(zk.native.TUNE_JOBS_SHAPE, 0 if shape == "k1" else shape)
    ctx.tune(zk.native.TUNE_JOBS_ROWS, rows)
    yield ctx, n, nl, random.Random(bits)
    ctx.tune(zk.native.TUNE_JOBS_SHAPE, 0)
    ctx.tune(zk.native.TUNE_JOBS_ROWS, 0)
    ctx.tune(zk.native.TUNE_ENC_KERNEL, 0)


def test_zero_proof(keyed):
    ctx, n, nl, rng = keyed
    nnl, B = 2 * nl, 7
    r = [rng.randrange(1, n) for _ in range(B)]
    msg = [0, 0, 1, 0, 0, 5, 0]                      # statements 2 and 5 do not encrypt zero (test_one_proof)
    c = once(("zero.c", n, tuple(r)), lambda: [po.paillier_encrypt(n, m, ri) for m, ri in zip(msg, r)])
    rp = [rng.randrange(1, n) for _ in range(B)]
    z, a = ctx.zero_prove(ints_to_limbs(r, nl), ints_to_limbs(c, nnl), ints_to_limbs(rp, nl))
    want = once(("zero.want", n, tuple(r), tuple(rp)), lambda: [po.ZeroProof.prove(ri, n, ci, rpi) for ri, ci, rpi in zip(r, c, rp)])
    assert limbs_to_ints(z) == [w.z for w in want] and limbs_to_ints(a) == [w.a for w in want]
    acc = ctx.zero_verify(ints_to_limbs(c, nnl), z, a)
    cpu = once(("zero.verdict", n, tuple(r), tuple(rp)), lambda: [verdict(lambda w=w, ci=ci: w.verify(n, ci)) for w, ci in zip(want, c)])
    assert acc.tolist() == cpu == [1, 1, 0, 1, 1, 0, 1]
    z2 = z.copy()
    z2[0, 3] ^= 1
    assert ctx.zero_verify(ints_to_limbs(c, nnl), z2, a).tolist() == [0, 1, 0, 1, 1, 0, 1]


def test_ciphertext_proof(keyed):
    ctx, n, nl, rng = keyed
    nnl, B = 2 * nl, 6
    x = [rng.randrange(n) for _ in range(B)]
    r = [rng.randrange(1, n) for _ in range(B)]
    c = once(("ct.c", n, tuple(x), tuple(r)), lambda: [po.paillier_encrypt(n, xi, ri) for xi, ri in zip(x, r)])
    r_used = list(r)
    r_used[4] = (r[4] + 1) % n                          # test_bad_ciphertext_proof: witness r + 1
    xp = [rng.randrange(n) for _ in range(B)]
    rp = [rng.randrange(1, n) for _ in range(B)]
    z1, z2, cp = ctx.ciphertext_prove(ints_to_limbs(x, nl), ints_to_limbs(r_used, nl), ints_to_limbs(c, nnl), ints_to_limbs(xp, nl),
                                      ints_to_limbs(rp, nl))
    want = once(("ct.want", n, tuple(x), tuple(r), tuple(xp), tuple(rp)),
                lambda: [po.CiphertextProof.prove(x[i], r_used[i], n, c[i], xp[i], rp[i]) for i in range(B)])
    assert limbs_to_ints(z1) == [w.z1 for w in want]
    assert limbs_to_ints(z2) == [w.z2 for w in want] and limbs_to_ints(cp) == [w.c_prime for w in want]
    assert any(w.z1 >= n for w in want)                 # the unreduced response really exceeds n
    acc = ctx.ciphertext_verify(ints_to_limbs(c, nnl), z1, z2, cp)
    cpu = once(("ct.verdict", n, tuple(x), tuple(r), tuple(xp), tuple(rp)), lambda: [verdict(lambda w=w, ci=ci: w.verify(n, ci)) for w, ci in zip(want, c)])
    assert acc.tolist() == cpu == [1, 1, 1, 1, 0, 1]


def test_mul_proof(keyed):
    ctx, n, nl, rng = keyed
    nnl, B = 2 * nl, 5
    a = [rng.randrange(n) for _ in range(B)]
    b = [rng.randrange(n) for _ in range(B)]
    cc = [ai * bi % n for ai, bi in zip(a, b)]
    cc[3] = (cc[3] + 1) % n                             # test_bad_mul_proof: c != a*b
    r_a, r_b, r_c = ([rng.randrange(1, n) for _ in range(B)] for _ in range(3))
    e_a, e_b, e_c = once(("mul.enc", n, tuple(a), tuple(r_a), tuple(r_b), tuple(r_c)),
                         lambda: tuple([po.paillier_encrypt(n, v, rr) for v, rr in zip(vals, rs)] for vals, rs in ((a, r_a), (b, r_b), (cc, r_c))))
    d = [rng.randrange(n) for _ in range(B)]
    r_d = [rng.randrange(1, n) for _ in range(B)]
    key = (n, tuple(a), tuple(r_a), tuple(d), tuple(r_d))
    L = lambda v, w: ints_to_limbs(v, w)
    f, z1, z2, e_d, e_db, fault = ctx.mul_prove(L(a, nl), L(b, nl), L(r_a, nl), L(r_b, nl), L(r_c, nl), L(e_a, nnl), L(e_b, nnl),
                                                L(e_c, nnl), L(d, nl), L(r_d, nl))
    assert not fault.any()
    want = once(("mul.want",) + key,
                lambda: [po.MulProof.prove(a[i], b[i], cc[i], r_a[i], r_b[i], r_c[i], n, e_a[i], e_b[i], e_c[i], d[i], r_d[i]) for i in range(B)])
    for name, got in (("f", f), ("z1", z1), ("z2", z2), ("e_d", e_d), ("e_db", e_db)):
        assert limbs_to_ints(got) == [getattr(w, name) for w in want], name
    acc, fault = ctx.mul_verify(L(e_a, nnl), L(e_b, nnl), L(e_c, nnl), f, z1, z2, e_d, e_db)
    cpu = once(("mul.verdict",) + key, lambda: [verdict(lambda i=i: want[i].verify(n, e_a[i], e_b[i], e_c[i])) for i in range(B)])
    assert acc.tolist() == cpu == [1, 1, 1, 0, 1]
    assert not fault.any()
    # a non-invertible value: e_db = 0 makes e_db * e_c^e = 0, where the reference's unwrap() panics
    e_db0 = e_db.copy()
    e_db0[1] = 0
    acc, fault = ctx.mul_verify(L(e_a, nnl), L(e_b, nnl), L(e_c, nnl), f, z1, z2, e_d, e_db0)
    assert fault.tolist() == [0, 1, 0, 0, 0] and acc.tolist() == [1, 0, 1, 0, 1]
    def panics():
        with pytest.raises(po.ReferencePanic):
            po.MulProof(want[1].f, want[1].z1, want[1].z2, want[1].e_d, 0).verify(n, e_a[1], e_b[1], e_c[1])
        return True

    assert once(("mul.panic",) + key, panics)


def _verlin_statement(n, mc, rc, mcp, rcp, x, xp, xdp, r_x):
    c = [po.paillier_encrypt(n, m, r) for m, r in zip(mc, rc)]
    cp = [po.paillier_encrypt(n, m, r) for m, r in zip(mcp, rcp)]
    return c, cp, [po.gen_phi(n, c[i], cp[i], x[i], xp[i], xdp[i], r_x[i]) for i in range(len(c))]


def test_verlin_proof(keyed):
    ctx, n, nl, rng = keyed
    nnl, B = 2 * nl, 5
    rnd = lambda: [rng.randrange(1, n) for _ in range(B)]
    x, xp, xdp, r_x = rnd(), rnd(), rnd(), rnd()
    mc, rc = zip(*[(rng.randrange(n), rng.randrange(1, n)) for _ in range(B)])
    mcp, rcp = zip(*[(rng.randrange(n), rng.randrange(1, n)) for _ in range(B)])
    key = (n, tuple(x), mc, rc, mcp, rcp)
    c, cp, phi_x = once(("verlin.statement",) + key, lambda: _verlin_statement(n, mc, rc, mcp, rcp, x, xp, xdp, r_x))
    xw, rw = list(x), list(r_x)
    xw[1] = x[1] + 1                                     # test_bad_verlin_proof: wrong x
    rw[3] = r_x[3] + 1                                   # test_bad_verlin_proof_2: wrong r_x
    a, ap, adp, r_a = rnd(), rnd(), rnd(), rnd()
    L = lambda v, w: ints_to_limbs(v, w)
    phi_a, z, zp, zdp, r_z = ctx.verlin_prove(L(xw, nl), L(xp, nl), L(xdp, nl), L(rw, nl), L(c, nnl), L(cp, nnl), L(phi_x, nnl), L(a, nl),
                                              L(ap, nl), L(adp, nl), L(r_a, nl))
    want = once(("verlin.want",) + key + (tuple(a), tuple(r_a)),
                lambda: [po.VerlinProof.prove(xw[i], xp[i], xdp[i], rw[i], n, c[i], cp[i], phi_x[i], a[i], ap[i], adp[i], r_a[i]) for i in range(B)])
    for name, got in (("phi_a", phi_a), ("z", z), ("z_prime", zp), ("z_double_prime", zdp), ("r_z", r_z)):
        assert limbs_to_ints(got) == [getattr(w, name) for w in want], name
    acc = ctx.verlin_verify(L(c, nnl), L(cp, nnl), L(phi_x, nnl), phi_a, z, zp, zdp, r_z)
    cpu = once(("verlin.verdict",) + key + (tuple(a), tuple(r_a)), lambda: [verdict(lambda i=i: want[i].verify(n, c[i], cp[i], phi_x[i])) for i in range(B)])
    assert acc.tolist() == cpu == [1, 0, 1, 0, 1]
