"""GPU suite at BASELINE.json's full sizes, through size-independent properties (the oracle cannot run these
sizes in seconds): prove->verify round trips, the Paillier homomorphism, challenge-hash agreement between
prover and verifier, reject-path statements, and oracle spot checks on sampled rows.
  configs[1]  RangeProofNi batch=1024, 2048-bit n
  configs[3]  RangeProofNi batch=65536 over 8 GPUs: one GPU's contiguous shard (8192 proofs), sharded as bench.py does
  configs[2]  NiCorrectKeyProof verify batch=4096, 3072-bit n
  configs[4]  MulProof + VerlinProof verify at 4096-bit n (per-GPU share of the 8192 mixed batch: 1024 = 512 + 512)
"""
import random

import numpy as np
import pytest

from util import c_oracle, keys, limbs_for, po
from zk_paillier_b200 import workload
from zk_paillier_b200.native import from_limbs, ints_to_limbs, limbs_to_ints, to_limbs

pytestmark = pytest.mark.gpu


def test_config1_rangeproof_batch1024_2048(ctx):
    n = po.TEST_P * po.TEST_Q
    nl, ef, batch = 64, 128, 1024
    ctx.set_key(to_limbs(n, nl))
    work = workload.rangeproof_batch(n, batch, ef=ef, seed=workload.DEFAULT_SEED, reject_every=100)
    cx = ctx.paillier_enc(work["x_n"], work["r"])
    out = ctx.rangeproof_ni_prove(ef, work["range"], work["x"], work["r"], work["w1"], work["swap"], work["r1"], work["r2"])
    acc, fault, dig = ctx.rangeproof_ni_verify(ef, work["range"], cx, out["c1"], out["c2"], out["kind"], out["resp_w"], out["resp_r"])
    assert acc.tolist() == [0 if b % 100 == 99 else 1 for b in range(batch)] and not fault.any()
    assert np.array_equal(dig, out["digest"])                       # prover and verifier hash the same transcript
    opens = int((out["kind"] == 0).sum())
    assert ctx.rp_verify_enc_count() == batch * ef + opens
    assert abs(opens / (batch * ef) - 0.5) < 0.01                   # challenge bits look like a hash output
    # homomorphism on the whole batch: c1[b][0] * c2[b][0] = Enc(w1 + w2, r1 * r2)  (all on the device)
    open0 = out["kind"][:, 0] == 0
    w_sum = ints_to_limbs([from_limbs(out["resp_w"][b, 0, 0]) + from_limbs(out["resp_w"][b, 0, 1]) if open0[b] else 0 for b in range(batch)], nl)
    r_prod = ctx.modmul(np.ascontiguousarray(out["resp_r"][:, 0, 0]), np.ascontiguousarray(out["resp_r"][:, 0, 1]), which_nn=False)
    lhs = ctx.modmul(np.ascontiguousarray(out["c1"][:, 0]), np.ascontiguousarray(out["c2"][:, 0]), which_nn=True)
    rhs = ctx.paillier_enc(w_sum, r_prod)
    assert np.array_equal(lhs[open0], rhs[open0]) and open0.sum() > 400
    # oracle spot check on a few whole proofs (byte-exact)
    sel = np.array([0, 99, 511, 1023])
    cpu = c_oracle.rangeproof_ni_prove(to_limbs(n, nl), ef, work["range"][sel], work["x"][sel], work["r"][sel], work["w1"][sel],
                                       work["swap"][sel], work["r1"][sel], work["r2"][sel])
    for k in ("c1", "c2", "digest", "kind", "resp_w", "resp_r"):
        assert np.array_equal(out[k][sel], cpu[k]), k
    # one flipped limb anywhere in a proof is caught
    bad = {k: v.copy() for k, v in out.items()}
    bad["c2"][7, 77, 100] ^= 0x10
    bad["resp_r"][300, 5, 0, 63] ^= 1
    acc2, _, _ = ctx.rangeproof_ni_verify(ef, work["range"], cx, bad["c1"], bad["c2"], bad["kind"], bad["resp_w"], bad["resp_r"])
    want = acc.copy()
    want[7] = 0
    want[300] = 0
    assert np.array_equal(acc2, want)


def test_config3_one_shard_of_65536(ctx):
    """The per-GPU share of configs[3]: rank 5 of 8 owns proofs [40960, 49152) of the 65 536 (sharding.shard_bounds); its
    statements come from the rank's own seeded stream, as in bench.py.  3.7 M encryptions: properties, plus whole proofs
    against the GMP oracle at the shard's ends."""
    from zk_paillier_b200 import sharding

    n = po.TEST_P * po.TEST_Q
    nl, ef, world, rank = 64, 128, 8, 5
    lo, hi = sharding.shard_bounds(65536, world, rank)
    batch = hi - lo
    assert (lo, batch) == (40960, 8192)
    ctx.set_key(to_limbs(n, nl))
    work = workload.rangeproof_batch(n, batch, ef=ef, seed=workload.DEFAULT_SEED + rank, reject_every=100)
    cx = ctx.paillier_enc(work["x_n"], work["r"])
    out = ctx.rangeproof_ni_prove(ef, work["range"], work["x"], work["r"], work["w1"], work["swap"], work["r1"], work["r2"])
    acc, fault, dig = ctx.rangeproof_ni_verify(ef, work["range"], cx, out["c1"], out["c2"], out["kind"], out["resp_w"], out["resp_r"])
    assert acc.tolist() == [0 if b % 100 == 99 else 1 for b in range(batch)] and not fault.any()
    assert np.array_equal(dig, out["digest"])
    opens = int((out["kind"] == 0).sum())
    assert ctx.rp_verify_enc_count() == batch * ef + opens and abs(opens / (batch * ef) - 0.5) < 0.005
    assert len({bytes(d) for d in dig}) == batch                       # 8192 distinct challenges
    sel = np.array([0, batch // 2, batch - 1])
    cpu = c_oracle.rangeproof_ni_prove(to_limbs(n, nl), ef, work["range"][sel], work["x"][sel], work["r"][sel], work["w1"][sel],
                                       work["swap"][sel], work["r1"][sel], work["r2"][sel])
    for k in ("c1", "c2", "digest", "kind", "resp_w", "resp_r"):
        assert np.array_equal(out[k][sel], cpu[k]), k


def test_config2_correct_key_batch4096_3072(ctx):
    """BASELINE configs[2] as stated: 4096 proofs, 3072-bit n, a DISTINCT modulus per proof.  The keys come from the device
    keygen (Paillier::keypairs_batch: Miller-Rabin waves on K2) and the honest proofs from NiCorrectKeyProof::proof_batch
    (C++ mirror); verdicts and rho against the GMP oracle on a sample, primality of sampled p, q against sympy."""
    import sympy

    nl, batch, salt = 96, 4096, b"Zen Go X"
    work = workload.correct_key_distinct(3072, batch, salt, seed=20261017, bad_every=64, with_primes=True)
    assert len({r.tobytes() for r in work["n"]}) == batch                    # 4096 distinct moduli
    acc, rho = ctx.correct_key_ni_verify(work["n"], work["sigma"], salt, want_rho=True)
    assert acc.tolist() == [0 if b % 64 == 63 else 1 for b in range(batch)]
    sel = np.array([0, 1, 62, 63, 64, 1000, 2047, 4031, 4095])
    acc_c, rho_c = c_oracle.correct_key_ni_verify(work["n"][sel], work["sigma"][sel], salt)
    assert np.array_equal(acc[sel], acc_c) and np.array_equal(rho[sel], rho_c)
    for b in (0, 1777, 4095):
        pp, qq = from_limbs(work["pq"][b, 0]), from_limbs(work["pq"][b, 1])
        assert pp != qq and pp.bit_length() == qq.bit_length() == 1536 and pp % 4 == 3 and qq % 4 == 3
        assert sympy.isprime(pp) and sympy.isprime(qq) and pp * qq == from_limbs(work["n"][b])
        assert limbs_to_ints(rho[b]) == po.correct_key_rho(pp * qq, salt)
        # the device-built proof is the oracle's proof (NiCorrectKeyProof::proof is deterministic)
        want = po.NiCorrectKeyProof.proof(pp, qq, salt).sigma_vec
        if b % 64 != 63:
            assert limbs_to_ints(work["sigma"][b]) == want
    # the prover-side derivation agrees with the verifier's
    assert np.array_equal(ctx.correct_key_ni_rho(work["n"][:32], salt), rho[:32])
    # the same seed and count give the same keys (the byte stream behind every sample is SHA-256(seed || counter))
    small = workload.correct_key_distinct(1024, 6, salt, seed=7)
    assert np.array_equal(small["n"], workload.correct_key_distinct(1024, 6, salt, seed=7)["n"])
    assert not np.array_equal(small["n"], workload.correct_key_distinct(1024, 6, salt, seed=8)["n"])


def test_config5_mul_and_verlin_4096(ctx):
    p, q = keys(4096)[0]
    n = p * q
    nl, nnl, B = 128, 256, 512
    ctx.set_key(to_limbs(n, nl))
    g = np.random.Generator(np.random.PCG64(5))

    def rows(count, limbs=nl, bits=4095):
        return workload._rand_limbs_below_pow2(g, (count,), limbs, bits)

    # MulProof statements built on the device (Paillier encryptions are not on the verify path)
    a, b = rows(B), rows(B)
    a_i, b_i = limbs_to_ints(a), limbs_to_ints(b)
    c_i = [x * y % n for x, y in zip(a_i, b_i)]
    c_i[5] = (c_i[5] + 1) % n                                     # multiplication_proof.rs:223
    c = ints_to_limbs(c_i, nl)
    r_a, r_b, r_c, d, r_d = (rows(B) | 1 for _ in range(5))
    e_a, e_b, e_c = ctx.paillier_enc(a, r_a), ctx.paillier_enc(b, r_b), ctx.paillier_enc(c, r_c)
    f, z1, z2, e_d, e_db, fault = ctx.mul_prove(a, b, r_a, r_b, r_c, e_a, e_b, e_c, d, r_d)
    assert not fault.any()
    acc, fault = ctx.mul_verify(e_a, e_b, e_c, f, z1, z2, e_d, e_db)
    assert acc.tolist() == [0 if i == 5 else 1 for i in range(B)] and not fault.any()
    z1b = z1.copy()
    z1b[100, 17] ^= 2
    acc, _ = ctx.mul_verify(e_a, e_b, e_c, f, z1b, z2, e_d, e_db)
    assert acc.tolist() == [0 if i in (5, 100) else 1 for i in range(B)]
    # one proof against the Python oracle, bit for bit
    w = po.MulProof.prove(a_i[0], b_i[0], c_i[0], from_limbs(r_a[0]), from_limbs(r_b[0]), from_limbs(r_c[0]), n, from_limbs(e_a[0]),
                          from_limbs(e_b[0]), from_limbs(e_c[0]), from_limbs(d[0]), from_limbs(r_d[0]))
    assert (from_limbs(f[0]), from_limbs(z1[0]), from_limbs(z2[0]), from_limbs(e_d[0]), from_limbs(e_db[0])) == (w.f, w.z1, w.z2, w.e_d, w.e_db)

    # VerlinProof: phi_x = gen_phi(c, c', x, x', x'', r_x) assembled from device calls
    x, xp, xdp, r_x = rows(B), rows(B), rows(B), rows(B) | 1
    cc, cp = ctx.paillier_enc(rows(B), rows(B) | 1), ctx.paillier_enc(rows(B), rows(B) | 1)
    nn_rows = np.tile(to_limbs(n * n, nnl), (1, 1))
    pad = lambda v: np.concatenate([v, np.zeros((B, nnl - nl), np.uint32)], axis=1)
    t1 = ctx.modexp_var(cc, pad(x), nn_rows, exp_per=1, mod_per=B, exp_bits=4096)
    t2 = ctx.modexp_var(cp, pad(xp), nn_rows, exp_per=1, mod_per=B, exp_bits=4096)
    phi_x = ctx.modmul(ctx.modmul(t1, t2), ctx.paillier_enc(xdp, r_x))
    xw = x.copy()
    xw[9, 0] ^= 1                                                   # verlin_proof.rs:219: wrong witness x
    a0, a1, a2, r_a2 = rows(B), rows(B), rows(B), rows(B) | 1
    phi_a, z, zp, zdp, r_z = ctx.verlin_prove(xw, xp, xdp, r_x, cc, cp, phi_x, a0, a1, a2, r_a2)
    acc = ctx.verlin_verify(cc, cp, phi_x, phi_a, z, zp, zdp, r_z)
    assert acc.tolist() == [0 if i == 9 else 1 for i in range(B)]
    assert from_limbs(phi_x[3]) == po.gen_phi(n, from_limbs(cc[3]), from_limbs(cp[3]), from_limbs(x[3]), from_limbs(xp[3]),
                                              from_limbs(xdp[3]), from_limbs(r_x[3]))
