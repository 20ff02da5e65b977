"""ctypes binding of oracle/liboracle.so (the GMP + OpenSSL C restatement in oracle/oracle.c).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Array layouts are those of include/zkp_b200.h so results compare byte for byte.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_u32p = C.POINTER(C.c_uint32)
_u8p = C.POINTER(C.c_uint8)
_lib = None
M2 = 11


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = C.CDLL(LIB_PATH)
        lib.orc_gmp_version.restype = C.c_char_p
        lib.orc_hw_threads.restype = C.c_int
        lib.orc_rangeproof_ni_verify.restype = C.c_longlong
        lib.orc_rangeproof_verify.restype = C.c_longlong
        _lib = lib
    return _lib


def _p32(a):
    return None if a is None else a.ctypes.data_as(_u32p)


def _p8(a):
    return None if a is None else a.ctypes.data_as(_u8p)


def _c32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _c8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def gmp_version():
    return load().orc_gmp_version().decode()


def hw_threads():
    return load().orc_hw_threads()


def paillier_enc(n, m, r, threads=0):
    n, m, r = _c32(n), _c32(m), _c32(r)
    batch, nl = m.shape[0], n.shape[-1]
    out = np.empty((batch, 2 * nl), np.uint32)
    load().orc_paillier_enc(_p32(n), nl, _p32(m), m.shape[1], _p32(r), r.shape[1], batch, _p32(out), threads)
    return out


def modexp(bases, exps, mods, per=1, threads=0):
    bases, exps, mods = _c32(bases), _c32(exps), _c32(mods)
    batch, ml = bases.shape
    out = np.empty((batch, ml), np.uint32)
    load().orc_modexp(_p32(bases), _p32(exps), exps.shape[1], _p32(mods), ml, per, batch, _p32(out), threads)
    return out


def sha256_transcript(items):
    items = _c32(items)
    batch, count, limbs = items.shape
    out = np.empty((batch, 32), np.uint8)
    load().orc_sha256_transcript(_p32(items), limbs, count, batch, _p8(out))
    return out


def rangeproof_ni_prove(n, ef, range_, x, r, w1, swap, r1, r2, threads=0, challenge=None):
    """challenge: uint8 [batch, nbytes] -> the interactive RangeProof with the verifier's raw ChallengeBits."""
    n, range_, x, r, w1, r1, r2 = map(_c32, (n, range_, x, r, w1, r1, r2))
    swap = _c8(swap)
    batch, wl = range_.shape
    nl = n.shape[-1]
    out = {
        "c1": np.empty((batch, ef, 2 * nl), np.uint32),
        "c2": np.empty((batch, ef, 2 * nl), np.uint32),
        "digest": np.empty((batch, 32), np.uint8),
        "kind": np.empty((batch, ef), np.uint8),
        "resp_w": np.zeros((batch, ef, 2, wl), np.uint32),
        "resp_r": np.zeros((batch, ef, 2, nl), np.uint32),
        "fault": np.zeros(batch, np.uint8),
    }
    ch = None if challenge is None else _c8(challenge)
    load().orc_rangeproof_prove(_p32(n), nl, batch, ef, wl, _p32(range_), _p32(x), _p32(r), _p32(w1), _p8(swap), _p32(r1),
                                _p32(r2), _p8(ch), 0 if ch is None else ch.shape[1], _p32(out["c1"]), _p32(out["c2"]),
                                _p8(out["digest"]), _p8(out["kind"]), _p32(out["resp_w"]), _p32(out["resp_r"]), _p8(out["fault"]), threads)
    return out


def rangeproof_ni_verify(n, ef, range_, cipher_x, c1, c2, kind, resp_w, resp_r, threads=0, challenge=None):
    n, range_, cipher_x, c1, c2, resp_w, resp_r = map(_c32, (n, range_, cipher_x, c1, c2, resp_w, resp_r))
    kind = _c8(kind)
    batch, wl = range_.shape
    nl = n.shape[-1]
    accept, fault, digest = np.empty(batch, np.uint8), np.empty(batch, np.uint8), np.empty((batch, 32), np.uint8)
    ch = None if challenge is None else _c8(challenge)
    encs = load().orc_rangeproof_verify(_p32(n), nl, batch, ef, wl, _p32(range_), _p32(cipher_x), _p32(c1), _p32(c2),
                                        _p8(kind), _p32(resp_w), _p32(resp_r), _p8(ch), 0 if ch is None else ch.shape[1],
                                        _p8(accept), _p8(fault), _p8(digest), threads)
    return accept, fault, digest, encs


def correct_key_ni_verify(n, sigma, salt: bytes, threads=0):
    n, sigma = _c32(n), _c32(sigma)
    batch, nl = n.shape
    accept = np.empty(batch, np.uint8)
    rho = np.empty((batch, M2, nl), np.uint32)
    s = np.frombuffer(bytes(salt), dtype=np.uint8).copy() if len(salt) else np.zeros(1, np.uint8)
    load().orc_correct_key_ni_verify(batch, nl, _p32(n), _p32(sigma), _p8(s), len(salt), _p8(accept), _p32(rho), threads)
    return accept, rho


# ---- sigma-protocol verifiers on GMP (verdict: 1 Ok, 0 Err(IncorrectProof), 2 where the reference panics)
def mul_verify(n, e_a, e_b, e_c, f, z1, z2, e_d, e_db, threads=0):
    n = _c32(n)
    arrs = [_c32(a) for a in (e_a, e_b, e_c, f, z1, z2, e_d, e_db)]
    batch = arrs[0].shape[0]
    out = np.empty(batch, np.uint8)
    load().orc_mul_verify(_p32(n), n.shape[-1], batch, *[_p32(a) for a in arrs], _p8(out), threads)
    return out


def verlin_verify(n, c, c_prime, phi_x, phi_a, z, z_prime, z_dp, r_z, threads=0):
    n = _c32(n)
    arrs = [_c32(a) for a in (c, c_prime, phi_x, phi_a, z, z_prime, z_dp, r_z)]
    batch = arrs[0].shape[0]
    out = np.empty(batch, np.uint8)
    load().orc_verlin_verify(_p32(n), n.shape[-1], arrs[4].shape[1], batch, *[_p32(a) for a in arrs], _p8(out), threads)
    return out


def dlog_verify(N, g, ni, x, y, threads=0):
    arrs = [_c32(a) for a in (N, g, ni, x, y)]
    batch, nl = arrs[0].shape
    out = np.empty(batch, np.uint8)
    load().orc_dlog_verify(nl, arrs[4].shape[1], batch, *[_p32(a) for a in arrs], _p8(out), threads)
    return out


def correct_message_verify(n, ciphertext, valid, e_vec, z_vec, a_vec, threads=0):
    n = _c32(n)
    ciphertext, valid, e_vec, z_vec, a_vec = (_c32(a) for a in (ciphertext, valid, e_vec, z_vec, a_vec))
    batch, M, ml = valid.shape
    out = np.empty(batch, np.uint8)
    load().orc_correct_message_verify(_p32(n), n.shape[-1], batch, M, ml, e_vec.shape[2], _p32(ciphertext), _p32(valid), _p32(e_vec),
                                      _p32(z_vec), _p32(a_vec), _p8(out), threads)
    return out



# ---- provers on GMP (bench.py baselines; compared with the Python oracle in tests/test_oracle_more.py)
def zero_prove(n, r, c, r_prime, threads=0):
    n, r, c, r_prime = _c32(n), _c32(r), _c32(c), _c32(r_prime)
    batch, nl = r.shape
    z, a = np.empty((batch, 2 * nl), np.uint32), np.empty((batch, 2 * nl), np.uint32)
    load().orc_zero_prove(_p32(n), nl, batch, _p32(r), _p32(c), _p32(r_prime), _p32(z), _p32(a), threads)
    return z, a


def zero_verify(n, c, z, a, threads=0):
    n, c, z, a = _c32(n), _c32(c), _c32(z), _c32(a)
    batch = c.shape[0]
    out = np.empty(batch, np.uint8)
    load().orc_zero_verify(_p32(n), n.shape[-1], batch, _p32(c), _p32(z), _p32(a), _p8(out), threads)
    return out


def dlog_prove(N, g, ni, secret, r, y_limbs, threads=0):
    N, g, ni, secret, r = (_c32(a) for a in (N, g, ni, secret, r))
    batch, nl = N.shape
    x, y = np.empty((batch, nl), np.uint32), np.empty((batch, y_limbs), np.uint32)
    load().orc_dlog_prove(nl, secret.shape[1], r.shape[1], y_limbs, batch, _p32(N), _p32(g), _p32(ni), _p32(secret), _p32(r), _p32(x), _p32(y), threads)
    return x, y


def correct_message_prove(n, valid, msg, r, e_rand, z_rand, w, threads=0):
    n, valid, msg, r, e_rand, z_rand, w = (_c32(a) for a in (n, valid, msg, r, e_rand, z_rand, w))
    batch, M, ml = valid.shape
    nl = n.shape[-1]
    assert e_rand.shape == (batch, M - 1, 8) and z_rand.shape == (batch, M - 1, nl)
    out = {"ciphertext": np.empty((batch, 2 * nl), np.uint32), "e_vec": np.empty((batch, M, 8), np.uint32),
           "z_vec": np.empty((batch, M, nl), np.uint32), "a_vec": np.empty((batch, M, 2 * nl), np.uint32)}
    load().orc_correct_message_prove(_p32(n), nl, batch, M, ml, _p32(valid), _p32(msg), _p32(r), _p32(e_rand), _p32(z_rand), _p32(w),
                                     _p32(out["ciphertext"]), _p32(out["e_vec"]), _p32(out["z_vec"]), _p32(out["a_vec"]), threads)
    return out
