"""Seeded synthetic workloads for the parity tests and bench.py.

All randomness the reference would draw from OsRng (range_proof.rs:136-159) is generated here as
explicit arrays, from one numpy PCG64 stream, so the CUDA path and the CPU oracle see identical inputs.
Nothing in this module computes on big integers beyond sampling.
"""
import numpy as np

from .native import ints_to_limbs

DEFAULT_SEED = 0x5A4B50
RANGE_BITS = 256  # range_proof.rs:377, range_proof_ni.rs:133
W_LIMBS = 12      # 384 bits: holds x + w for the reject-path x in [100 q, 10000 q)


def _rand_int(rng, bits):
    nbytes = (bits + 7) // 8
    return int.from_bytes(rng.bytes(nbytes), "big") >> (8 * nbytes - bits)


def _rand_limbs_below_pow2(rng, shape, limbs, bits):
    """uniform integers in [0, 2^bits) as limb rows (fast path, no Python ints)."""
    a = np.frombuffer(rng.bytes(int(np.prod(shape)) * limbs * 4), dtype="<u4").reshape(tuple(shape) + (limbs,)).copy()
    full, rem = divmod(bits, 32)
    if full < limbs:
        a[..., full + (1 if rem else 0):] = 0
        if rem:
            a[..., full] &= (1 << rem) - 1
    return a


def rangeproof_batch(n: int, batch: int, ef: int = 128, seed: int = DEFAULT_SEED, reject_every: int = 0,
                     w_limbs: int = W_LIMBS):
    """Inputs of RangeProofNi::prove for `batch` proofs under the key n.

    Returns a dict of uint32 limb arrays (+ python ints for range/x/r) shaped as zkp_rangeproof_ni_prove
    expects.  Every `reject_every`-th proof (if > 0) uses x in [100 q, 10000 q) as the reference's negative
    test does (range_proof_ni.rs:185-188); the others use x in [0, q/3)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    nbits = n.bit_length()
    n_limbs = (nbits + 31) // 32
    n_limbs += (-n_limbs) % 4
    ranges, xs, w1 = [], [], []
    for b in range(batch):
        q = _rand_int(rng, RANGE_BITS) | (1 << (RANGE_BITS - 1))
        third = q // 3
        if reject_every and b % reject_every == reject_every - 1:
            x = 100 * q + _rand_int(rng, 300) % (9900 * q)
        else:
            x = _rand_int(rng, RANGE_BITS + 64) % third
        ranges.append(q)
        xs.append(x)
        w1.append([third + _rand_int(rng, RANGE_BITS + 64) % third for _ in range(ef)])  # [third, 2*third)
    swap = (np.frombuffer(rng.bytes(batch * ef), dtype=np.uint8) & 1).reshape(batch, ef).copy()
    # r, r1, r2 uniform below 2^(|n|-1) <= n (synthetic stand-in for sample_below(n))
    r = _rand_limbs_below_pow2(rng, (batch,), n_limbs, nbits - 1)
    r1 = _rand_limbs_below_pow2(rng, (batch, ef), n_limbs, nbits - 1)
    r2 = _rand_limbs_below_pow2(rng, (batch, ef), n_limbs, nbits - 1)
    return {
        "n_limbs": n_limbs,
        "w_limbs": w_limbs,
        "ef": ef,
        "range_int": ranges,
        "x_int": xs,
        "range": ints_to_limbs(ranges, w_limbs),
        "x": ints_to_limbs(xs, w_limbs),
        "x_n": ints_to_limbs(xs, n_limbs),
        "r": r,
        "w1": ints_to_limbs(w1, w_limbs),
        "swap": swap,
        "r1": r1,
        "r2": r2,
    }


def correct_key_batch(keys, batch: int, salt: bytes, make_sigma, n_limbs: int, bad_every: int = 0):
    """Inputs of NiCorrectKeyProof::verify for `batch` proofs; proof b uses keys[b % len(keys)] (p, q).

    `make_sigma(p, q, salt)` -> list of 11 ints (the oracle's NiCorrectKeyProof::proof; statement
    generation is not on the measured path).  Every `bad_every`-th proof gets sigma_3 + 1."""
    per_key = []
    for (p, q) in keys:
        per_key.append((p * q, make_sigma(p, q, salt)))
    ns, sig = [], []
    for b in range(batch):
        n, s = per_key[b % len(per_key)]
        s = list(s)
        if bad_every and b % bad_every == bad_every - 1:
            s[3] = (s[3] + 1) % n
        ns.append(n)
        sig.append(s)
    return {"n_int": ns, "sigma_int": sig, "n": ints_to_limbs(ns, n_limbs), "sigma": ints_to_limbs(sig, n_limbs)}


def correct_key_distinct(bits: int, batch: int, salt: bytes, seed: int = DEFAULT_SEED, bad_every: int = 0, device: int = 0, with_primes: bool = False):
    """Inputs of NiCorrectKeyProof::verify for `batch` proofs with a DISTINCT `bits`-bit modulus each (BASELINE.json configs[2]).

    Built on the GPU by the C++ host mirror (libzkp_host.so: Paillier::keypairs_batch - Miller-Rabin waves on the device - and
    NiCorrectKeyProof::proof_batch), reproducibly from `seed`; every `bad_every`-th proof has sigma_0 + 1.  Statement
    generation is not on the measured path.  Raises without a GPU, like everything else here."""
    import ctypes as C
    import os

    lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libzkp_host.so"))
    fn = lib.zkh_correct_key_workload
    fn.restype = C.c_int
    nl = bits // 32
    n = np.zeros((batch, nl), np.uint32)
    sigma = np.zeros((batch, 11, nl), np.uint32)
    pq = np.zeros((batch, 2, bits // 64), np.uint32) if with_primes else None
    err = C.create_string_buffer(512)
    sd = int(seed).to_bytes(8, "little") + b"correct_key_distinct"
    u32p = C.POINTER(C.c_uint32)
    rc = fn(C.c_int(device), C.c_int(bits), C.c_int(batch), sd, C.c_int(len(sd)), salt, C.c_int(len(salt)), C.c_int(bad_every),
            n.ctypes.data_as(u32p), sigma.ctypes.data_as(u32p), pq.ctypes.data_as(u32p) if with_primes else None, err, C.c_int(512))
    if rc != 0:
        raise RuntimeError("zkh_correct_key_workload: " + err.value.decode(errors="replace"))
    out = {"n": n, "sigma": sigma}
    if with_primes:
        out["pq"] = pq
    return out
