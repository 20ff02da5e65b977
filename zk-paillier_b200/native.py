"""ctypes binding of the C ABI in include/zkp_b200.h.

Integers cross as little-endian uint32 limb arrays (numpy, C-contiguous).  There
is no CPU implementation behind any of these calls: if libzkp_b200.so is missing
or no CUDA device is present, they raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libzkp_b200.so")
# measurement scripts may point at the lab build (make lab); the product path is always the in-tree library
if os.environ.get("ZKP_B200_LIB"):
    LIB_PATH = os.path.abspath(os.environ["ZKP_B200_LIB"])
TUNE_ENC_KERNEL, TUNE_JOBS_SHAPE, TUNE_JOBS_ROWS = 0, 1, 2

ZKP_OK = 0
RP_OPEN, RP_MASK1, RP_MASK2 = 0, 1, 2
CK_M2 = 11
KID_MODEXP_SHARED, KID_MODEXP_VAR, KID_MODMUL, KID_SHA, KID_OTHER, KID_CALL = range(6)

_u32p = C.POINTER(C.c_uint32)
_u8p = C.POINTER(C.c_uint8)

# name -> (restype, argtypes); every symbol include/zkp_b200.h declares
SIGNATURES = {
    "zkp_ctx_create": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "zkp_ctx_destroy": (None, [C.c_void_p]),
    "zkp_last_error": (C.c_char_p, [C.c_void_p]),
    "zkp_version": (C.c_int, []),
    "zkp_sm_count": (C.c_int, [C.c_void_p]),
    "zkp_sync": (C.c_int, [C.c_void_p]),
    "zkp_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "zkp_profile_reset": (C.c_int, [C.c_void_p]),
    "zkp_profile_get": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.POINTER(C.c_double)]),
    "zkp_set_key": (C.c_int, [C.c_void_p, _u32p, C.c_int]),
    "zkp_set_modulus": (C.c_int, [C.c_void_p, _u32p, C.c_int, _u32p, C.c_int]),
    "zkp_nn_limbs": (C.c_int, [C.c_void_p]),
    "zkp_modexp_shared": (C.c_int, [C.c_void_p, _u32p, C.c_int, C.c_int, _u32p]),
    "zkp_paillier_enc": (C.c_int, [C.c_void_p, _u32p, C.c_int, _u32p, C.c_int, C.c_int, _u32p]),
    "zkp_modexp_var": (C.c_int, [C.c_void_p, _u32p, _u32p, C.c_int, C.c_int, C.c_int, _u32p, C.c_int, C.c_int, C.c_int, _u32p]),
    "zkp_modmul": (C.c_int, [C.c_void_p, C.c_int, _u32p, _u32p, C.c_int, C.c_int, _u32p]),
    "zkp_sha256_transcript": (C.c_int, [C.c_void_p, _u32p, C.c_int, C.c_int, C.c_int, _u8p]),
    "zkp_rangeproof_ni_prove": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, _u32p, _u32p, _u32p, _u32p, _u8p, _u32p, _u32p,
                                          _u32p, _u32p, _u8p, _u8p, _u32p, _u32p]),
    "zkp_rp_prove_stage": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, _u32p, _u32p, _u32p, _u32p, _u8p, _u32p, _u32p]),
    "zkp_rp_prove_run": (C.c_int, [C.c_void_p]),
    "zkp_rp_prove_run_pairs": (C.c_int, [C.c_void_p]),
    "zkp_rp_prove_run_responses": (C.c_int, [C.c_void_p, _u8p, C.c_int]),
    "zkp_rp_verify_run_with_challenge": (C.c_int, [C.c_void_p, _u8p, C.c_int]),
    "zkp_rp_prove_fetch": (C.c_int, [C.c_void_p, _u32p, _u32p, _u8p, _u8p, _u32p, _u32p]),
    "zkp_rangeproof_ni_verify": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, _u32p, _u32p, _u32p, _u32p, _u8p, _u32p, _u32p,
                                           _u8p, _u8p, _u8p]),
    "zkp_rp_verify_stage": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, _u32p, _u32p, _u32p, _u32p, _u8p, _u32p, _u32p]),
    "zkp_rp_verify_stage_from_prove": (C.c_int, [C.c_void_p, _u32p]),
    "zkp_rp_verify_run": (C.c_int, [C.c_void_p]),
    "zkp_rp_verify_fetch": (C.c_int, [C.c_void_p, _u8p, _u8p, _u8p]),
    "zkp_rp_verify_enc_count": (C.c_longlong, [C.c_void_p]),
    "zkp_correct_key_ni_verify": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _u32p, _u32p, _u8p, C.c_int, _u8p, _u32p]),
    "zkp_correct_key_ni_rho": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _u32p, _u8p, C.c_int, _u32p]),
    "zkp_ck_verify_stage": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _u32p, _u32p, _u8p, C.c_int]),
    "zkp_ck_verify_run": (C.c_int, [C.c_void_p]),
    "zkp_ck_verify_fetch": (C.c_int, [C.c_void_p, _u8p, _u32p]),
    "zkp_zero_prove": (C.c_int, [C.c_void_p, C.c_int] + [_u32p] * 5),
    "zkp_zero_verify": (C.c_int, [C.c_void_p, C.c_int] + [_u32p] * 3 + [_u8p]),
    "zkp_ciphertext_prove": (C.c_int, [C.c_void_p, C.c_int, C.c_int] + [_u32p] * 8),
    "zkp_ciphertext_verify": (C.c_int, [C.c_void_p, C.c_int, C.c_int] + [_u32p] * 4 + [_u8p]),
    "zkp_mul_prove": (C.c_int, [C.c_void_p, C.c_int] + [_u32p] * 15 + [_u8p]),
    "zkp_mul_verify": (C.c_int, [C.c_void_p, C.c_int] + [_u32p] * 8 + [_u8p, _u8p]),
    "zkp_verlin_prove": (C.c_int, [C.c_void_p, C.c_int, C.c_int] + [_u32p] * 16),
    "zkp_verlin_verify": (C.c_int, [C.c_void_p, C.c_int, C.c_int] + [_u32p] * 8 + [_u8p]),
    "zkp_verify_opening": (C.c_int, [C.c_void_p, C.c_int, C.c_int] + [_u32p] * 3 + [_u8p]),
    "zkp_dlog_prove": (C.c_int, [C.c_void_p, C.c_int, C.c_int] + [_u32p] * 4 + [C.c_int, _u32p, C.c_int, C.c_int, _u32p, _u32p, _u8p]),
    "zkp_dlog_verify": (C.c_int, [C.c_void_p, C.c_int, C.c_int] + [_u32p] * 5 + [C.c_int, _u8p, _u8p]),
    "zkp_correct_message_prove": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int] + [_u32p] * 10 + [_u8p]),
    "zkp_correct_message_verify": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int] + [_u32p] * 5 + [_u8p, _u8p]),
    "zkp_tune": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "zkp_imad_peak": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double)]),
    "zkp_enc_kernel_launches": (C.c_int, [C.c_void_p] + [C.POINTER(C.c_longlong)] * 2),
    "zkp_enc_executed_mads": (C.c_int, [C.c_void_p] + [C.POINTER(C.c_double)] * 2),
}

_lib = None


class ZkpError(RuntimeError):
    pass


def load():
    """Load libzkp_b200.so (built in-tree by `make lib` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ZkpError(f"{LIB_PATH} is not built (run `make lib`); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


# ---- limb helpers -----------------------------------------------------------
def to_limbs(x: int, limbs: int) -> np.ndarray:
    return np.frombuffer(int(x).to_bytes(limbs * 4, "little"), dtype="<u4").copy()


def from_limbs(a) -> int:
    return int.from_bytes(np.ascontiguousarray(a, dtype="<u4").tobytes(), "little")


def ints_to_limbs(xs, limbs: int) -> np.ndarray:
    """list (possibly nested) of non-negative ints -> uint32 array [..., limbs]."""
    arr = np.asarray(xs, dtype=object)
    flat = arr.reshape(-1)
    buf = b"".join(int(v).to_bytes(limbs * 4, "little") for v in flat)
    return np.frombuffer(buf, dtype="<u4").reshape(arr.shape + (limbs,)).copy()


def limbs_to_ints(a):
    a = np.ascontiguousarray(a, dtype="<u4")
    limbs = a.shape[-1]
    raw = a.tobytes()
    n = a.size // limbs
    vals = [int.from_bytes(raw[i * limbs * 4:(i + 1) * limbs * 4], "little") for i in range(n)]
    return np.asarray(vals + [None], dtype=object)[:-1].reshape(a.shape[:-1]).tolist() if a.ndim > 1 else vals[0]


def _p32(a):
    return None if a is None else a.ctypes.data_as(_u32p)


def _p8(a):
    return None if a is None else a.ctypes.data_as(_u8p)


def _c32(a):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return a


def _c8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


class Context:
    """One engine context = one GPU (zkp_ctx).  Single-threaded."""

    def __init__(self, device: int = 0, stream=None):
        self._lib = load()
        h = C.c_void_p()
        rc = self._lib.zkp_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != ZKP_OK:
            raise ZkpError(f"zkp_ctx_create(device={device}) failed with {rc}: no usable CUDA device (no CPU fallback)")
        self._h = h
        self.device = device
        self.n_limbs = 0

    def close(self):
        if getattr(self, "_h", None):
            self._lib.zkp_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != ZKP_OK:
            raise ZkpError(f"zkp error {rc}: {self._lib.zkp_last_error(self._h).decode()}")

    # -- misc
    @property
    def sm_count(self):
        return self._lib.zkp_sm_count(self._h)

    def sync(self):
        self._ck(self._lib.zkp_sync(self._h))

    def profile_enable(self, on=True):
        self._ck(self._lib.zkp_profile_enable(self._h, 1 if on else 0))

    def profile_reset(self):
        self._ck(self._lib.zkp_profile_reset(self._h))

    def profile_get(self, kernel):
        ms, n, u = C.c_double(), C.c_longlong(), C.c_double()
        self._ck(self._lib.zkp_profile_get(self._h, kernel, C.byref(ms), C.byref(n), C.byref(u)))
        return ms.value, n.value, u.value

    def tune(self, knob, value):
        """zkp_tune: TUNE_ENC_KERNEL (0 = K1m when the key qualifies, 1 = K1) / TUNE_JOBS_SHAPE (0 auto, 1 wide, 2 narrow lanes)"""
        self._ck(self._lib.zkp_tune(self._h, int(knob), int(value)))

    def imad_peak(self, variant=0):
        v = C.c_double()
        self._ck(self._lib.zkp_imad_peak(self._h, variant, C.byref(v)))
        return v.value

    def enc_kernel_launches(self):
        """launch counts of the encryption kernels on this context: {"k1m", "k1"}"""
        a, b = C.c_longlong(), C.c_longlong()
        self._ck(self._lib.zkp_enc_kernel_launches(self._h, C.byref(a), C.byref(b)))
        return {"k1m": a.value, "k1": b.value}

    def enc_executed_mads(self):
        """IMAD.WIDE one encryption executes under the current key: {"k1m", "k1"} (k1m = 0 when the key does not qualify)"""
        a, b = C.c_double(), C.c_double()
        self._ck(self._lib.zkp_enc_executed_mads(self._h, C.byref(a), C.byref(b)))
        return {"k1m": a.value, "k1": b.value}

    # -- key
    def set_key(self, n_limbs_arr):
        n = _c32(n_limbs_arr)
        self._ck(self._lib.zkp_set_key(self._h, _p32(n), n.shape[-1]))
        self.n_limbs = n.shape[-1]

    def set_modulus(self, mod, exp):
        mod, exp = _c32(mod), _c32(exp)
        self._ck(self._lib.zkp_set_modulus(self._h, _p32(mod), mod.shape[-1], _p32(exp), exp.shape[-1]))

    @property
    def nn_limbs(self):
        return self._lib.zkp_nn_limbs(self._h)

    # -- K1 / K2 / K3
    def modexp_shared(self, bases):
        bases = _c32(bases)
        batch, bl = bases.shape
        out = np.empty((batch, self.nn_limbs), dtype=np.uint32)
        self._ck(self._lib.zkp_modexp_shared(self._h, _p32(bases), bl, batch, _p32(out)))
        return out

    def paillier_enc(self, m, r):
        m, r = _c32(m), _c32(r)
        batch = m.shape[0]
        assert r.shape[0] == batch
        out = np.empty((batch, self.nn_limbs), dtype=np.uint32)
        self._ck(self._lib.zkp_paillier_enc(self._h, _p32(m), m.shape[1], _p32(r), r.shape[1], batch, _p32(out)))
        return out

    def modexp_var(self, bases, exps, mods, per=1, exp_bits=None, exp_per=None, mod_per=None):
        """out[j] = bases[j]^exps[j // exp_per] mod mods[j // mod_per]; `per` sets both."""
        bases, exps, mods = _c32(bases), _c32(exps), _c32(mods)
        exp_per = per if exp_per is None else exp_per
        mod_per = per if mod_per is None else mod_per
        batch, ml = bases.shape
        assert mods.shape[1] == ml and mods.shape[0] == (batch + mod_per - 1) // mod_per
        assert exps.shape[0] == (batch + exp_per - 1) // exp_per
        if exp_bits is None:
            exp_bits = 32 * exps.shape[1]
        out = np.empty((batch, ml), dtype=np.uint32)
        self._ck(self._lib.zkp_modexp_var(self._h, _p32(bases), _p32(exps), exps.shape[1], exp_bits, exp_per, _p32(mods), ml,
                                          mod_per, batch, _p32(out)))
        return out

    def modmul(self, a, b, which_nn=True, b_per=1):
        a, b = _c32(a), _c32(b)
        batch = a.shape[0]
        out = np.empty_like(a)
        self._ck(self._lib.zkp_modmul(self._h, 1 if which_nn else 0, _p32(a), _p32(b), b_per, batch, _p32(out)))
        return out

    # -- K4
    def sha256_transcript(self, items):
        items = _c32(items)
        batch, count, limbs = items.shape
        out = np.empty((batch, 32), dtype=np.uint8)
        self._ck(self._lib.zkp_sha256_transcript(self._h, _p32(items), limbs, count, batch, _p8(out)))
        return out

    # -- RangeProofNi
    def rp_prove_stage(self, ef, range_, x, r, w1, swap, r1, r2):
        range_, x, r, w1, r1, r2 = map(_c32, (range_, x, r, w1, r1, r2))
        swap = _c8(swap)
        batch, wl = range_.shape
        assert w1.shape == (batch, ef, wl) and swap.shape == (batch, ef)
        assert r1.shape == r2.shape == (batch, ef, self.n_limbs) and r.shape == (batch, self.n_limbs) and x.shape == (batch, wl)
        self._ck(self._lib.zkp_rp_prove_stage(self._h, batch, ef, wl, _p32(range_), _p32(x), _p32(r), _p32(w1), _p8(swap),
                                              _p32(r1), _p32(r2)))
        self._rp_shape = (batch, ef, wl)

    def rp_prove_run(self):
        self._ck(self._lib.zkp_rp_prove_run(self._h))

    def rp_prove_run_pairs(self):
        self._ck(self._lib.zkp_rp_prove_run_pairs(self._h))

    def rp_prove_run_responses(self, challenge=None):
        """challenge: uint8 [batch, nbytes] raw ChallengeBits of the interactive proof, or None for Fiat-Shamir."""
        if challenge is None:
            self._ck(self._lib.zkp_rp_prove_run_responses(self._h, None, 0))
        else:
            ch = _c8(challenge)
            assert ch.ndim == 2 and ch.shape[0] == self._rp_shape[0]
            self._ck(self._lib.zkp_rp_prove_run_responses(self._h, _p8(ch), ch.shape[1]))

    def rp_prove_fetch_pairs(self):
        batch, ef, wl = self._rp_shape
        nnl = self.nn_limbs
        c1, c2 = np.empty((batch, ef, nnl), np.uint32), np.empty((batch, ef, nnl), np.uint32)
        self._ck(self._lib.zkp_rp_prove_fetch(self._h, _p32(c1), _p32(c2), None, None, None, None))
        return c1, c2

    def rp_verify_run_with_challenge(self, challenge):
        ch = _c8(challenge)
        assert ch.ndim == 2 and ch.shape[0] == self._rpv_batch
        self._ck(self._lib.zkp_rp_verify_run_with_challenge(self._h, _p8(ch), ch.shape[1]))

    def rp_prove_fetch(self, want_pairs=True):
        batch, ef, wl = self._rp_shape
        nl, nnl = self.n_limbs, self.nn_limbs
        out = {
            "c1": np.empty((batch, ef, nnl), np.uint32) if want_pairs else None,
            "c2": np.empty((batch, ef, nnl), np.uint32) if want_pairs else None,
            "digest": np.empty((batch, 32), np.uint8),
            "kind": np.empty((batch, ef), np.uint8),
            "resp_w": np.empty((batch, ef, 2, wl), np.uint32),
            "resp_r": np.empty((batch, ef, 2, nl), np.uint32),
        }
        self._ck(self._lib.zkp_rp_prove_fetch(self._h, _p32(out["c1"]), _p32(out["c2"]), _p8(out["digest"]), _p8(out["kind"]),
                                              _p32(out["resp_w"]), _p32(out["resp_r"])))
        return out

    def rangeproof_ni_prove(self, ef, range_, x, r, w1, swap, r1, r2):
        self.rp_prove_stage(ef, range_, x, r, w1, swap, r1, r2)
        self.rp_prove_run()
        return self.rp_prove_fetch()

    def rp_verify_stage(self, ef, range_, cipher_x, c1, c2, kind, resp_w, resp_r):
        range_, cipher_x, c1, c2, resp_w, resp_r = map(_c32, (range_, cipher_x, c1, c2, resp_w, resp_r))
        kind = _c8(kind)
        batch, wl = range_.shape
        nl, nnl = self.n_limbs, self.nn_limbs
        assert c1.shape == c2.shape == (batch, ef, nnl) and cipher_x.shape == (batch, nnl)
        assert kind.shape == (batch, ef) and resp_w.shape == (batch, ef, 2, wl) and resp_r.shape == (batch, ef, 2, nl)
        self._ck(self._lib.zkp_rp_verify_stage(self._h, batch, ef, wl, _p32(range_), _p32(cipher_x), _p32(c1), _p32(c2),
                                               _p8(kind), _p32(resp_w), _p32(resp_r)))
        self._rpv_batch = batch

    def rp_verify_stage_from_prove(self, cipher_x):
        cipher_x = _c32(cipher_x)
        self._ck(self._lib.zkp_rp_verify_stage_from_prove(self._h, _p32(cipher_x)))
        self._rpv_batch = cipher_x.shape[0]

    def rp_verify_run(self):
        self._ck(self._lib.zkp_rp_verify_run(self._h))

    def rp_verify_fetch(self):
        b = self._rpv_batch
        accept, fault, digest = np.empty(b, np.uint8), np.empty(b, np.uint8), np.empty((b, 32), np.uint8)
        self._ck(self._lib.zkp_rp_verify_fetch(self._h, _p8(accept), _p8(fault), _p8(digest)))
        return accept, fault, digest

    def rp_verify_enc_count(self):
        return self._lib.zkp_rp_verify_enc_count(self._h)

    def rangeproof_ni_verify(self, ef, range_, cipher_x, c1, c2, kind, resp_w, resp_r):
        self.rp_verify_stage(ef, range_, cipher_x, c1, c2, kind, resp_w, resp_r)
        self.rp_verify_run()
        return self.rp_verify_fetch()

    # -- NiCorrectKeyProof
    def ck_verify_stage(self, n, sigma, salt: bytes):
        n, sigma = _c32(n), _c32(sigma)
        batch, nl = n.shape
        assert sigma.shape == (batch, CK_M2, nl)
        s = np.frombuffer(bytes(salt), dtype=np.uint8).copy() if len(salt) else np.zeros(1, np.uint8)
        self._ck(self._lib.zkp_ck_verify_stage(self._h, batch, nl, _p32(n), _p32(sigma), _p8(s), len(salt)))
        self._ck_shape = (batch, nl)

    def correct_key_ni_rho(self, n, salt: bytes):
        n = _c32(n)
        batch, nl = n.shape
        rho = np.empty((batch, CK_M2, nl), np.uint32)
        s = np.frombuffer(bytes(salt), dtype=np.uint8).copy() if len(salt) else np.zeros(1, np.uint8)
        self._ck(self._lib.zkp_correct_key_ni_rho(self._h, batch, nl, _p32(n), _p8(s), len(salt), _p32(rho)))
        return rho

    def ck_verify_run(self):
        self._ck(self._lib.zkp_ck_verify_run(self._h))

    def ck_verify_fetch(self, want_rho=False):
        batch, nl = self._ck_shape
        accept = np.empty(batch, np.uint8)
        rho = np.empty((batch, CK_M2, nl), np.uint32) if want_rho else None
        self._ck(self._lib.zkp_ck_verify_fetch(self._h, _p8(accept), _p32(rho)))
        return (accept, rho) if want_rho else accept

    def correct_key_ni_verify(self, n, sigma, salt: bytes, want_rho=False):
        self.ck_verify_stage(n, sigma, salt)
        self.ck_verify_run()
        return self.ck_verify_fetch(want_rho)

    # -- sigma protocols (rows: [batch][n_limbs] or [batch][nn_limbs]; z rows [batch][z_limbs])
    @property
    def z_limbs(self):
        return self.n_limbs + 12

    def _rows(self, arrs, widths):
        out = []
        batch = None
        for a, w in zip(arrs, widths):
            a = _c32(a)
            assert a.ndim == 2 and a.shape[1] == w, (a.shape, w)
            batch = a.shape[0] if batch is None else batch
            assert a.shape[0] == batch
            out.append(a)
        return batch, out

    def zero_prove(self, r, c, r_prime):
        nl, nnl = self.n_limbs, self.nn_limbs
        batch, (r, c, r_prime) = self._rows((r, c, r_prime), (nl, nnl, nl))
        z, a = np.empty((batch, nnl), np.uint32), np.empty((batch, nnl), np.uint32)
        self._ck(self._lib.zkp_zero_prove(self._h, batch, _p32(r), _p32(c), _p32(r_prime), _p32(z), _p32(a)))
        return z, a

    def zero_verify(self, c, z, a):
        nnl = self.nn_limbs
        batch, (c, z, a) = self._rows((c, z, a), (nnl, nnl, nnl))
        acc = np.empty(batch, np.uint8)
        self._ck(self._lib.zkp_zero_verify(self._h, batch, _p32(c), _p32(z), _p32(a), _p8(acc)))
        return acc

    def ciphertext_prove(self, x, r, c, x_prime, r_prime):
        nl, nnl, zl = self.n_limbs, self.nn_limbs, self.z_limbs
        batch, (x, r, c, x_prime, r_prime) = self._rows((x, r, c, x_prime, r_prime), (nl, nl, nnl, nl, nl))
        z1, z2, cp = np.empty((batch, zl), np.uint32), np.empty((batch, nnl), np.uint32), np.empty((batch, nnl), np.uint32)
        self._ck(self._lib.zkp_ciphertext_prove(self._h, batch, zl, _p32(x), _p32(r), _p32(c), _p32(x_prime), _p32(r_prime),
                                                _p32(z1), _p32(z2), _p32(cp)))
        return z1, z2, cp

    def ciphertext_verify(self, c, z1, z2, c_prime):
        nnl, zl = self.nn_limbs, self.z_limbs
        batch, (c, z1, z2, c_prime) = self._rows((c, z1, z2, c_prime), (nnl, zl, nnl, nnl))
        acc = np.empty(batch, np.uint8)
        self._ck(self._lib.zkp_ciphertext_verify(self._h, batch, zl, _p32(c), _p32(z1), _p32(z2), _p32(c_prime), _p8(acc)))
        return acc

    def mul_prove(self, a, b, r_a, r_b, r_c, e_a, e_b, e_c, d, r_d):
        nl, nnl = self.n_limbs, self.nn_limbs
        batch, ins = self._rows((a, b, r_a, r_b, r_c, e_a, e_b, e_c, d, r_d), (nl,) * 5 + (nnl,) * 3 + (nl, nl))
        f = np.empty((batch, nl), np.uint32)
        z1, z2, e_d, e_db = (np.empty((batch, nnl), np.uint32) for _ in range(4))
        fault = np.empty(batch, np.uint8)
        self._ck(self._lib.zkp_mul_prove(self._h, batch, *[_p32(v) for v in ins], _p32(f), _p32(z1), _p32(z2), _p32(e_d), _p32(e_db),
                                         _p8(fault)))
        return f, z1, z2, e_d, e_db, fault

    def mul_verify(self, e_a, e_b, e_c, f, z1, z2, e_d, e_db):
        nl, nnl = self.n_limbs, self.nn_limbs
        batch, ins = self._rows((e_a, e_b, e_c, f, z1, z2, e_d, e_db), (nnl, nnl, nnl, nl, nnl, nnl, nnl, nnl))
        acc, fault = np.empty(batch, np.uint8), np.empty(batch, np.uint8)
        self._ck(self._lib.zkp_mul_verify(self._h, batch, *[_p32(v) for v in ins], _p8(acc), _p8(fault)))
        return acc, fault

    def verlin_prove(self, x, x_prime, x_dp, r_x, c, c_prime, phi_x, a, a_prime, a_dp, r_a):
        nl, nnl, zl = self.n_limbs, self.nn_limbs, self.z_limbs
        batch, ins = self._rows((x, x_prime, x_dp, r_x, c, c_prime, phi_x, a, a_prime, a_dp, r_a), (nl,) * 4 + (nnl,) * 3 + (nl,) * 4)
        phi_a, r_z = np.empty((batch, nnl), np.uint32), np.empty((batch, nnl), np.uint32)
        z, zp, zdp = (np.empty((batch, zl), np.uint32) for _ in range(3))
        self._ck(self._lib.zkp_verlin_prove(self._h, batch, zl, *[_p32(v) for v in ins], _p32(phi_a), _p32(z), _p32(zp), _p32(zdp),
                                            _p32(r_z)))
        return phi_a, z, zp, zdp, r_z

    def verlin_verify(self, c, c_prime, phi_x, phi_a, z, z_prime, z_dp, r_z):
        nnl, zl = self.nn_limbs, self.z_limbs
        batch, ins = self._rows((c, c_prime, phi_x, phi_a, z, z_prime, z_dp, r_z), (nnl,) * 4 + (zl,) * 3 + (nnl,))
        acc = np.empty(batch, np.uint8)
        self._ck(self._lib.zkp_verlin_verify(self._h, batch, zl, *[_p32(v) for v in ins], _p8(acc)))
        return acc

    # -- the remaining public proofs
    def verify_opening(self, m, r, c):
        """CorrectOpening::verify_opening (correct_opening.rs:17-30): ok[b] = (c[b] == Enc(m[b], r[b]))."""
        m = _c32(m)
        batch, (m, r, c) = self._rows((m, r, c), (m.shape[1], self.n_limbs, self.nn_limbs))
        ok = np.empty(batch, np.uint8)
        self._ck(self._lib.zkp_verify_opening(self._h, batch, m.shape[1], _p32(m), _p32(r), _p32(c), _p8(ok)))
        return ok

    def dlog_prove(self, N, g, ni, secret, r, y_limbs):
        """CompositeDLogProof::prove (wi_dlog_proof.rs:46-65) -> (x, y, fault); rows [batch][n_limbs], one N per proof."""
        N = _c32(N)
        nl = N.shape[1]
        secret, r = _c32(secret), _c32(r)
        batch, (N, g, ni, secret, r) = self._rows((N, g, ni, secret, r), (nl, nl, nl, secret.shape[1], r.shape[1]))
        x, y, fault = np.empty((batch, nl), np.uint32), np.empty((batch, y_limbs), np.uint32), np.empty(batch, np.uint8)
        self._ck(self._lib.zkp_dlog_prove(self._h, batch, nl, _p32(N), _p32(g), _p32(ni), _p32(secret), secret.shape[1], _p32(r),
                                          r.shape[1], y_limbs, _p32(x), _p32(y), _p8(fault)))
        return x, y, fault

    def dlog_verify(self, N, g, ni, x, y):
        """CompositeDLogProof::verify (wi_dlog_proof.rs:66-91) -> (accept, fault)."""
        N, y = _c32(N), _c32(y)
        nl = N.shape[1]
        batch, (N, g, ni, x, y) = self._rows((N, g, ni, x, y), (nl, nl, nl, nl, y.shape[1]))
        acc, fault = np.empty(batch, np.uint8), np.empty(batch, np.uint8)
        self._ck(self._lib.zkp_dlog_verify(self._h, batch, nl, _p32(N), _p32(g), _p32(ni), _p32(x), _p32(y), y.shape[1], _p8(acc),
                                           _p8(fault)))
        return acc, fault

    def correct_message_prove(self, valid, msg, r, e_rand, z_rand, w):
        """CorrectMessageProof::prove (correct_message.rs:35-125).  valid [batch][M][ml], msg [batch][ml], r / w [batch][nl],
        e_rand [batch][M-1][8], z_rand [batch][M-1][nl] -> dict(ciphertext, e_vec, z_vec, a_vec, fault)."""
        valid, msg, r, w = _c32(valid), _c32(msg), _c32(r), _c32(w)
        batch, M, ml = valid.shape
        nl, nnl = self.n_limbs, self.nn_limbs
        e_rand = _c32(e_rand).reshape(batch, max(M - 1, 0), 8)
        z_rand = _c32(z_rand).reshape(batch, max(M - 1, 0), nl)
        assert msg.shape == (batch, ml) and r.shape == (batch, nl) and w.shape == (batch, nl)
        out = {"ciphertext": np.empty((batch, nnl), np.uint32), "e_vec": np.empty((batch, M, 8), np.uint32),
               "z_vec": np.empty((batch, M, nl), np.uint32), "a_vec": np.empty((batch, M, nnl), np.uint32), "fault": np.empty(batch, np.uint8)}
        self._ck(self._lib.zkp_correct_message_prove(self._h, batch, M, ml, _p32(valid), _p32(msg), _p32(r),
                                                     _p32(e_rand) if M > 1 else None, _p32(z_rand) if M > 1 else None, _p32(w),
                                                     _p32(out["ciphertext"]), _p32(out["e_vec"]), _p32(out["z_vec"]), _p32(out["a_vec"]),
                                                     _p8(out["fault"])))
        return out

    def correct_message_verify(self, ciphertext, valid, e_vec, z_vec, a_vec):
        """CorrectMessageProof::verify (correct_message.rs:126-161) -> (accept, fault)."""
        valid, e_vec = _c32(valid), _c32(e_vec)
        batch, M, ml = valid.shape
        nl, nnl = self.n_limbs, self.nn_limbs
        el = e_vec.shape[2]
        ciphertext, z_vec, a_vec = _c32(ciphertext), _c32(z_vec), _c32(a_vec)
        assert ciphertext.shape == (batch, nnl) and e_vec.shape == (batch, M, el) and z_vec.shape == (batch, M, nl) and a_vec.shape == (batch, M, nnl)
        acc, fault = np.empty(batch, np.uint8), np.empty(batch, np.uint8)
        self._ck(self._lib.zkp_correct_message_verify(self._h, batch, M, ml, el, _p32(ciphertext), _p32(valid), _p32(e_vec), _p32(z_vec),
                                                      _p32(a_vec), _p8(acc), _p8(fault)))
        return acc, fault
