"""ctypes access to the C++ host mirror (zk-paillier_b200/libzkp_host.so): JSON in, JSON out."""
import ctypes as C
import json
import os

from util import ROOT

_lib = None


def call(op, **req):
    global _lib
    if _lib is None:
        _lib = C.CDLL(os.path.join(ROOT, "zk-paillier_b200", "libzkp_host.so"))
        _lib.zkh_call.restype = C.c_void_p
        _lib.zkh_call.argtypes = [C.c_char_p, C.c_char_p]
        _lib.zkh_free.argtypes = [C.c_void_p]
    p = _lib.zkh_call(op.encode(), json.dumps(req).encode())
    try:
        return json.loads(C.string_at(p).decode())
    finally:
        _lib.zkh_free(p)


class Stream:
    """The byte stream handed to the host mirror as rng_hex, replayed on the Python side."""

    def __init__(self, data: bytes):
        self.data, self.pos = data, 0

    def __call__(self, n):
        out = self.data[self.pos:self.pos + n]
        assert len(out) == n, "stream exhausted"
        self.pos += n
        return out
