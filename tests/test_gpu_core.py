"""GPU parity of the raw kernels K1 (modexp_shared / Paillier Enc), K2 (modexp_var) and K3 (modmul)
through the C ABI, against Python big-int arithmetic (the canonical residue is unique, so any
correct bigint is an exact oracle here)."""
import random

import numpy as np
import pytest

import zk_paillier_b200 as zk
from util import c_oracle
from zk_paillier_b200.native import from_limbs, ints_to_limbs, to_limbs

pytestmark = pytest.mark.gpu

SPOT = 6  # rows of every batch also checked against CPython pow(); the rest against GMP (oracle/oracle.c), which the CPU suite
          # pins against CPython pow and OpenSSL BN_mod_exp (tests/test_oracle.py) -- 8192-bit pow() in Python is ~0.1 s a row


def want_enc(n, nl, m, m_limbs, r, r_limbs):
    """Expected ciphertext rows: GMP for all, CPython for the first SPOT."""
    nn = n * n
    out = c_oracle.paillier_enc(to_limbs(n, nl), ints_to_limbs(m, m_limbs), ints_to_limbs(r, r_limbs))
    for j in range(min(SPOT, len(m))):
        assert from_limbs(out[j]) == ((m[j] * n + 1) % nn) * pow(r[j], n, nn) % nn
    return out


def rand_odd(rng, bits, full=True):
    v = rng.getrandbits(bits) | 1
    if full:
        v |= 1 << (bits - 1)
    return v


def rows(a):
    return [from_limbs(a[i]) for i in range(a.shape[0])]


@pytest.mark.parametrize("mod_bits", [1024, 2048, 3072, 4096, 6144, 8192, 1536, 2176])
def test_modexp_shared_matches_pow(ctx, mod_bits):
    rng = random.Random(mod_bits)
    limbs = mod_bits // 32
    M = rand_odd(rng, mod_bits)
    exp_bits = mod_bits // 2
    E = rng.getrandbits(exp_bits) | (1 << (exp_bits - 1))
    ctx.set_modulus(to_limbs(M, limbs), to_limbs(E, (exp_bits // 32 + 3) // 4 * 4))
    batch = 70
    bases = [rng.getrandbits(mod_bits) for _ in range(batch)]   # includes values >= M
    bases[0], bases[1], bases[2], bases[3] = 0, 1, M - 1, M
    out = ctx.modexp_shared(ints_to_limbs(bases, limbs))
    want = c_oracle.modexp(ints_to_limbs(bases, limbs), ints_to_limbs([E], limbs), ints_to_limbs([M], limbs), per=batch)
    assert np.array_equal(out, want)
    for b, g in list(zip(bases, rows(out)))[:SPOT]:
        assert g == pow(b, E, M)


@pytest.mark.parametrize("exp", [1, 2, 3, 31, 32, 33, (1 << 64) - 1, 1 << 200, (1 << 200) + 1, 0x8000000000000001])
def test_modexp_shared_small_exponents(ctx, exp):
    rng = random.Random(exp % 1000)
    M = rand_odd(rng, 2048)
    ctx.set_modulus(to_limbs(M, 64), to_limbs(exp, 8))
    bases = [rng.getrandbits(2048) % M for _ in range(20)]
    got = rows(ctx.modexp_shared(ints_to_limbs(bases, 64)))
    assert got == [pow(b, exp, M) for b in bases]


@pytest.mark.parametrize("n_bits", [1024, 2048, 3072, 4096])
def test_paillier_enc(ctx, n_bits):
    rng = random.Random(n_bits + 1)
    nl = n_bits // 32
    n = rand_odd(rng, n_bits)
    nn = n * n
    ctx.set_key(to_limbs(n, nl))
    assert ctx.nn_limbs == 2 * nl
    batch = 133
    m = [rng.getrandbits(256) for _ in range(batch)]
    r = [rng.randrange(n) for _ in range(batch)]
    m[0], r[0] = 0, 1
    m[1], r[1] = n - 1, n - 1
    m[2] = 0
    out = ctx.paillier_enc(ints_to_limbs(m, nl), ints_to_limbs(r, nl))
    assert np.array_equal(out, want_enc(n, nl, m, nl, r, nl))
    # narrow plaintext rows, full-width randomness rows (ZeroProof: Enc(0, z) with z < n^2)
    z = [rng.randrange(nn) for _ in range(9)]
    mw = [rng.getrandbits(n_bits + 256) for _ in range(9)]   # unreduced plaintext (CiphertextProof z1)
    out = ctx.paillier_enc(ints_to_limbs(mw, nl + 8), ints_to_limbs(z, 2 * nl))
    assert np.array_equal(out, want_enc(n, nl, mw, nl + 8, z, 2 * nl))


@pytest.mark.parametrize("mod_bits,per", [(1024, 1), (2048, 11), (3072, 11), (4096, 3), (8192, 1), (2560, 2)])
def test_modexp_var(ctx, mod_bits, per):
    rng = random.Random(mod_bits * 7 + per)
    limbs = mod_bits // 32
    count = 13
    batch = count * per - (per // 2)
    mods = [rand_odd(rng, mod_bits, full=(i % 3 != 0)) for i in range(count)]
    exps = [rng.getrandbits(mod_bits) for _ in range(count)]
    exps[0] = 0
    exps[1] = 1
    bases = [rng.getrandbits(mod_bits) for _ in range(batch)]
    out = ctx.modexp_var(ints_to_limbs(bases, limbs), ints_to_limbs(exps, limbs), ints_to_limbs(mods, limbs), per=per)
    assert np.array_equal(out, c_oracle.modexp(ints_to_limbs(bases, limbs), ints_to_limbs(exps, limbs), ints_to_limbs(mods, limbs), per=per))
    for j, g in list(enumerate(rows(out)))[:SPOT] + [(batch - 1, from_limbs(out[batch - 1]))]:
        assert g == pow(bases[j], exps[j // per], mods[j // per]), j


def test_modexp_var_short_exponent(ctx):
    rng = random.Random(5)
    limbs = 128
    mods = [rand_odd(rng, 4096) for _ in range(10)]
    exps = [rng.getrandbits(256) for _ in range(10)]
    bases = [rng.getrandbits(4096) for _ in range(10)]
    out = ctx.modexp_var(ints_to_limbs(bases, limbs), ints_to_limbs(exps, 8), ints_to_limbs(mods, limbs), per=1, exp_bits=256)
    assert rows(out) == [pow(b, e, m) for b, e, m in zip(bases, exps, mods)]


@pytest.mark.parametrize("n_bits", [1024, 2048, 4096])
def test_modmul(ctx, n_bits):
    rng = random.Random(n_bits + 3)
    nl = n_bits // 32
    n = rand_odd(rng, n_bits)
    nn = n * n
    ctx.set_key(to_limbs(n, nl))
    a = [rng.getrandbits(2 * n_bits) for _ in range(50)]
    b = [rng.getrandbits(2 * n_bits) for _ in range(10)]
    out = ctx.modmul(ints_to_limbs(a, 2 * nl), ints_to_limbs(b, 2 * nl), which_nn=True, b_per=5)
    assert rows(out) == [a[j] * b[j // 5] % nn for j in range(50)]
    a = [rng.getrandbits(n_bits) for _ in range(33)]
    b = [rng.getrandbits(n_bits) for _ in range(33)]
    out = ctx.modmul(ints_to_limbs(a, nl), ints_to_limbs(b, nl), which_nn=False, b_per=1)
    assert rows(out) == [x * y % n for x, y in zip(a, b)]


def test_large_batch_all_groups(ctx):
    """More jobs than one persistent wave so every CTA loops, with a ragged tail."""
    rng = random.Random(99)
    n = rand_odd(rng, 1024)
    nn = n * n
    ctx.set_key(to_limbs(n, 32))
    batch = ctx.sm_count * 4 * 16 * 2 + 37
    r = [rng.randrange(n) for _ in range(batch)]
    m = [rng.getrandbits(200) for _ in range(batch)]
    out = rows(ctx.paillier_enc(ints_to_limbs(m, 32), ints_to_limbs(r, 32)))
    idx = list(range(0, batch, 97)) + [batch - 1, batch - 2]
    for j in idx:
        assert out[j] == ((m[j] * n + 1) % nn) * pow(r[j], n, nn) % nn


@pytest.mark.parametrize("n_bits,nl", [(1024, 32), (2048, 64), (2047, 64), (3072, 96), (4096, 128), (1536, 48)])
def test_two_digit_montgomery_kernel(n_bits, nl):
    """K1m (modexp2m.cu, the default encryption kernel) against K1 (zkp_tune ZKP_TUNE_ENC_KERNEL = 1) and Python pow: moduli that do
    not fill their top limb, widths that run zero-extended, bases >= n, base 0 / 1, plaintext 0 / n-1 / >= n / none,
    narrow plaintext rows, more than one persistent wave of a small batch's worth of CTAs and a ragged tail."""
    rng = random.Random(n_bits)
    n = rand_odd(rng, n_bits)
    nn = n * n
    batch = 150
    cap = 1 << (32 * nl)
    r = [rng.getrandbits(32 * nl) for _ in range(batch)]  # about half of them >= n
    r[0], r[1], r[2], r[3], r[4] = 1, n - 1, 0, min(n + 5, cap - 1), cap - 1
    m = [rng.getrandbits(32 * nl) if j % 7 == 0 else rng.getrandbits(300) for j in range(batch)]
    m[0], m[1], m[5] = 0, n - 1, min(n + 1, cap - 1)
    want = want_enc(n, nl, m, nl, r, nl)
    outs = {}
    for mode in ("k1m", "k1", "auto"):   # auto: a 150-job launch takes K2h's one-job-per-warp layout
        with zk.native.Context(0) as c:
            c.tune(zk.native.TUNE_ENC_KERNEL, {"k1m": 2, "k1": 1, "auto": 0}[mode])
            c.set_key(to_limbs(n, nl))
            outs[mode] = c.paillier_enc(ints_to_limbs(m, nl), ints_to_limbs(r, nl))
            used = c.enc_kernel_launches()
            assert (used["k1m"] > 0) == (mode != "k1") and (used["k1"] > 0) == (mode == "k1")
            if mode == "k1m":
                narrow = [v & ((1 << 256) - 1) for v in m[:9]]
                o = c.paillier_enc(ints_to_limbs(narrow, 8), ints_to_limbs(r[:9], nl))
                assert np.array_equal(o, want_enc(n, nl, narrow, 8, r[:9], nl))
    assert np.array_equal(outs["k1m"], want)
    assert np.array_equal(outs["k1m"], outs["k1"])
    assert np.array_equal(outs["auto"], want)
