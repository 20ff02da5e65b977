"""CPU ORACLE (test infrastructure, NOT a product path).

A plain Python-int + hashlib restatement of the reference's hot path, line by line, with every
random draw turned into an explicit input (the reference draws from OsRng inside rayon closures and
has no seed hook: SURVEY.md section 7 "Determinism").  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module; the product (libzkp_b200.so
and the host mirror in zk-paillier_b200/) never does.

PARITY PINNING.  The reference (ZenGo-X/zk-paillier @ 6ffbef0, Rust) cannot be built here: no
cargo/rustc, and its arithmetic lives in two un-vendored crates, curv-kzen 0.10 (feature
rust-gmp-kzen -> GMP) and kzen-paillier 0.4.3 (Cargo.toml:29-36,41-42).  Its own tests hold NO golden
vectors (28 randomized prove->verify round trips).  Therefore:
  * pinned:    every modexp / mulmod / Enc output (unique canonical residue; cross-checked here between
               CPython pow(), GMP 6.3.0 mpz_powm -- the reference's real backend -- and OpenSSL BN_mod_exp),
               SHA-256 (FIPS 180-4 vectors), accept/reject on the reference's test classes, and the
               fixtures the reference ships (test primes, primorial P, SALT_STRING, 128 / 11 / 256).
  * UNPINNED ("parity unpinned"): the four rules recalled from the dependency crates' public source, each
               isolated in ONE function below so a single edit fixes parity if they are ever checked against
               a real cargo build: bigint_to_bytes, sample_bits, serde_bigint_native, serde_encryption_key.
               The check exists and waits for its input: rust/gen_vectors links the UNMODIFIED crate and writes
               tests/golden/reference_vectors.json; tests/test_reference_vectors.py compares this oracle (and the
               CUDA path) with that file and reports xfail "parity unpinned" while it is absent (no cargo here).

Citations are file:line into /root/reference/src.
"""
import hashlib
import json

# ----------------------------------------------------------------------------------------------
# constants (zkproofs/range_proof_ni.rs:23, range_proof.rs:30, correct_key_ni.rs:26-30)
SECURITY_PARAMETER = 128
STATISTICAL_ERROR_FACTOR = 40
M2 = 11
DIGEST_SIZE = 256
SALT_STRING = bytes([75, 90, 101, 110])  # "KZen"
ALPHA = 6370


def _primes_below(k):
    sieve = bytearray([1]) * k
    sieve[0:2] = b"\x00\x00"
    for i in range(2, int(k ** 0.5) + 1):
        if sieve[i]:
            sieve[i * i::i] = bytearray(len(sieve[i * i::i]))
    return [i for i in range(k) if sieve[i]]


SMALL_PRIMES = _primes_below(ALPHA)


def primorial():
    """P = product of all primes < 6370 (correct_key_ni.rs:25-26; equality with the literal is checked in tests)."""
    p = 1
    for q in SMALL_PRIMES:
        p *= q
    return p


# The reference's fixed 2048-bit test key (range_proof_ni.rs:142-143, range_proof.rs:380-381, benches/all.rs:74-75).
TEST_P = int(
    "148677972634832330983979593310074301486537017973460461278300587514468301043894574906886127642530475786889672304776052879927627556769456140664043088700743909632312483413393134504352834240399191134336344285483935856491230340093391784574980688823380828143810804684752914935441384845195613674104960646037368551517"
)
TEST_Q = int(
    "158741574437007245654463598139927898730476924736461654463975966787719309357536545869203069369466212089132653564188443272208127277664424448947476335413293018778018615899291704693105620242763173357203898195318179150836424196645745308205164116144020613415407736216097185962171301808761138424668335445923774195463"
)


class IncorrectProof(Exception):
    """zkproofs/errors.rs:5-13"""

    def __str__(self):
        return "given proof doesn't match a statement"


class ReferencePanic(Exception):
    """Raised where the reference would panic (assert_eq!, unwrap on None, index out of range)."""


# ----------------------------------------------------------------------------------------------
# the four RECALLED rules of the un-vendored crates (parity unpinned; see module docstring)
def bigint_to_bytes(x: int) -> bytes:
    """curv-kzen 0.10 BigInt::to_bytes on the GMP backend (RECALLED): mpz_export of the magnitude,
    big-endian, into a buffer of (sizeinbase(x,2)+7)/8 bytes => minimal length, no leading zero
    bytes, and zero -> the single byte 0x00."""
    if x < 0:
        raise ReferencePanic("negative BigInt on the hashed path")
    return x.to_bytes(max(1, (x.bit_length() + 7) // 8), "big")


def bigint_from_bytes(b: bytes) -> int:
    """BigInt::from_bytes = mpz_import big-endian (RECALLED)."""
    return int.from_bytes(b, "big")


def sample_bits(stream, bits: int) -> int:
    """curv-kzen BigInt::sample(bits) (RECALLED): ceil(bits/8) bytes from the RNG, big-endian,
    shifted right by 8*bytes - bits.  `stream(nbytes)` supplies the bytes."""
    nbytes = (bits + 7) // 8
    return int.from_bytes(stream(nbytes), "big") >> (8 * nbytes - bits)


def sample_below(stream, upper: int) -> int:
    """BigInt::sample_below: rejection loop on sample(bit_length(upper)) (RECALLED)."""
    bits = upper.bit_length()
    while True:
        v = sample_bits(stream, bits)
        if v < upper:
            return v


def sample_range(stream, lo: int, hi: int) -> int:
    """BigInt::sample_range(lo, hi) = lo + sample_below(hi - lo) (RECALLED)."""
    return lo + sample_below(stream, hi - lo)


def serde_bigint_native(x: int) -> str:
    """curv-kzen 0.10 BigInt's own Serialize for human-readable formats (RECALLED): lower-case hex of to_bytes()."""
    return bigint_to_bytes(x).hex()


def serde_bigint_native_parse(s: str) -> int:
    return int.from_bytes(bytes.fromhex(s), "big")


def serde_encryption_key(n: int) -> dict:
    """kzen-paillier 0.4.3 EncryptionKey serde (RECALLED): a minimal form holding only n as a decimal string."""
    return {"n": str(n)}


# ----------------------------------------------------------------------------------------------
# L2: transcript hash (zkproofs/utils.rs:9-22)
def compute_digest(items) -> int:
    h = hashlib.sha256()
    for v in items:
        h.update(bigint_to_bytes(v))
    return bigint_from_bytes(h.digest())


def transcript_bytes(items) -> bytes:
    return b"".join(bigint_to_bytes(v) for v in items)


# L1: kzen-paillier (RECALLED)
def paillier_encrypt(n: int, m: int, r: int) -> int:
    """Paillier::encrypt_with_chosen_randomness: rn = r^n mod nn; gm = (m*n + 1) % nn; c = gm*rn % nn."""
    nn = n * n
    rn = pow(r, n, nn)
    gm = (m * n + 1) % nn
    return (gm * rn) % nn


def paillier_mul(n: int, c: int, m: int) -> int:
    return pow(c, m, n * n)


def paillier_add(n: int, c1: int, c2: int) -> int:
    return (c1 * c2) % (n * n)


def paillier_open(p: int, q: int, c: int):
    """Decrypt and recover the randomness (used by tests only; correct_opening.rs:49-56)."""
    n = p * q
    nn = n * n
    lam = (p - 1) * (q - 1)
    u = pow(c, lam, nn)
    m = ((u - 1) // n) * pow(lam, -1, n) % n
    gm_inv = pow((1 + m * n) % nn, -1, nn)
    rn = c * gm_inv % nn
    r = extract_nroot(p, q, rn % n)
    return m, r


def extract_nroot(p: int, q: int, z: int) -> int:
    """kzen-paillier extract_nroot(dk, z) = z^(n^-1 mod phi) mod n via CRT (RECALLED); the result is the
    unique n-th root mod n, so the CRT route does not change the value."""
    n = p * q
    phi = (p - 1) * (q - 1)
    d = pow(n, -1, phi)
    zp = pow(z % p, d % (p - 1), p)
    zq = pow(z % q, d % (q - 1), q)
    h = (zq - zp) * pow(p, -1, q) % q
    return (zp + p * h) % n


# ----------------------------------------------------------------------------------------------
# RangeProof / RangeProofNi  (range_proof.rs, range_proof_ni.rs)
def challenge_bit(e_bytes: bytes, i: int) -> int:
    """BitVec::from_bytes(e)[i], MSB first within each byte (range_proof.rs:221,225,267,273).
    Index past the end panics in the reference."""
    if i // 8 >= len(e_bytes):
        raise ReferencePanic("challenge bit index out of range")
    return (e_bytes[i // 8] >> (7 - i % 8)) & 1


def generate_encrypted_pairs(n, range_, w1_samples, swap_bits, r1, r2):
    """range_proof.rs:128-193 with the draws made explicit:
    w1_samples[i] in [third, 2*third) (:136-139), swap_bits[i] = the coin of :146, r1/r2 in [0,n) (:151-159)."""
    ef = len(w1_samples)
    third = range_ // 3
    w1 = list(w1_samples)
    w2 = [x - third for x in w1]  # :141
    for i in range(ef):  # :144-149
        if swap_bits[i]:
            w1[i], w2[i] = w2[i], w1[i]
    c1 = [paillier_encrypt(n, w1[i], r1[i]) for i in range(ef)]  # :161-173
    c2 = [paillier_encrypt(n, w2[i], r2[i]) for i in range(ef)]  # :175-187
    return {"c1": c1, "c2": c2}, {"w1": w1, "w2": w2, "r1": list(r1), "r2": list(r2)}


def generate_proof(n, secret_x, secret_r, e_bytes, range_, data, ef):
    """range_proof.rs:210-252.  Responses are ('Open', w1, r1, w2, r2) or ('Mask', j, masked_x, masked_r)."""
    third = range_ // 3
    two_thirds = 2 * third
    out = []
    for i in range(ef):
        if not challenge_bit(e_bytes, i):
            out.append(("Open", data["w1"][i], data["r1"][i], data["w2"][i], data["r2"][i]))
        elif third < secret_x + data["w1"][i] < two_thirds:
            out.append(("Mask", 1, secret_x + data["w1"][i], secret_r * data["r1"][i] % n))
        else:
            out.append(("Mask", 2, secret_x + data["w2"][i], secret_r * data["r2"][i] % n))
    return out


def verifier_output_bits(n, e_bytes, pairs, proof, range_, cipher_x, ef):
    """range_proof.rs:254-348: the per-index verdicts."""
    nn = n * n
    third = range_ // 3
    two_thirds = 2 * third
    res = []
    for i in range(ef):
        ei = challenge_bit(e_bytes, i)
        if i >= len(proof):
            raise ReferencePanic("responses shorter than error_factor")
        resp = proof[i]
        if not ei and resp[0] == "Open":
            _, w1, r1, w2, r2 = resp
            ok = paillier_encrypt(n, w1, r1) == pairs["c1"][i] and paillier_encrypt(n, w2, r2) == pairs["c2"][i]
            flag = (w2 < third and third < w1 < two_thirds) or (w1 < third and third < w2 < two_thirds)
            res.append(bool(ok and flag))
        elif ei and resp[0] == "Mask":
            _, j, mx, mr = resp
            c = (pairs["c1"][i] if j == 1 else pairs["c2"][i]) * cipher_x % nn
            ok = c == paillier_encrypt(n, mx, mr)
            if mx < third or mx > two_thirds:
                ok = False
            res.append(bool(ok))
        else:
            res.append(False)
    return res


def verifier_output(n, e_bytes, pairs, proof, range_, cipher_x, ef):
    if not all(verifier_output_bits(n, e_bytes, pairs, proof, range_, cipher_x, ef)):  # :350-354
        raise IncorrectProof()


def range_challenge(n, pairs) -> bytes:
    """e = to_bytes(compute_digest([n] ++ c1 ++ c2))  (range_proof_ni.rs:58-61, 89-92)."""
    return bigint_to_bytes(compute_digest([n] + list(pairs["c1"]) + list(pairs["c2"])))


def range_digest32(n, pairs) -> bytes:
    """The raw 32-byte SHA-256 output of the same transcript (what the CUDA path reports)."""
    return hashlib.sha256(transcript_bytes([n] + list(pairs["c1"]) + list(pairs["c2"]))).digest()


class RangeProofNi:
    """range_proof_ni.rs:36-128."""

    def __init__(self, n, range_, ciphertext, encrypted_pairs, proof, error_factor):
        self.n, self.range, self.ciphertext = n, range_, ciphertext
        self.encrypted_pairs, self.proof, self.error_factor = encrypted_pairs, proof, error_factor

    @staticmethod
    def prove(n, range_, ciphertext, secret_x, secret_r, w1_samples, swap_bits, r1, r2):
        pairs, data = generate_encrypted_pairs(n, range_, w1_samples, swap_bits, r1, r2)
        e = range_challenge(n, pairs)
        proof = generate_proof(n, secret_x, secret_r, e, range_, data, len(w1_samples))
        return RangeProofNi(n, range_, ciphertext, pairs, proof, len(w1_samples))

    def verify(self, n, ciphertext):
        if n != self.n or ciphertext != self.ciphertext:  # assert_eq! :86,88
            raise ReferencePanic("ek / ciphertext mismatch")
        self.verify_self()

    def verify_self(self):
        e = range_challenge(self.n, self.encrypted_pairs)
        verifier_output(self.n, e, self.encrypted_pairs, self.proof, self.range, self.ciphertext, self.error_factor)

    # serde (range_proof_ni.rs:35-44 derive; range_proof.rs:31-81; serialize.rs)
    def to_json(self) -> str:
        def resp(r):
            if r[0] == "Open":
                return {"Open": {"w1": str(r[1]), "r1": str(r[2]), "w2": str(r[3]), "r2": str(r[4])}}
            return {"Mask": {"j": r[1], "masked_x": str(r[2]), "masked_r": str(r[3])}}

        return json.dumps(
            {
                "ek": serde_encryption_key(self.n),
                "range": serde_bigint_native(self.range),
                "ciphertext": serde_bigint_native(self.ciphertext),
                "encrypted_pairs": {"c1": [str(v) for v in self.encrypted_pairs["c1"]], "c2": [str(v) for v in self.encrypted_pairs["c2"]]},
                "proof": [resp(r) for r in self.proof],
                "error_factor": self.error_factor,
            },
            separators=(",", ":"),
        )

    @staticmethod
    def from_json(s: str):
        d = json.loads(s)

        def resp(o):
            if "Open" in o:
                v = o["Open"]
                return ("Open", int(v["w1"]), int(v["r1"]), int(v["w2"]), int(v["r2"]))
            v = o["Mask"]
            return ("Mask", int(v["j"]), int(v["masked_x"]), int(v["masked_r"]))

        pairs = {"c1": [int(v) for v in d["encrypted_pairs"]["c1"]], "c2": [int(v) for v in d["encrypted_pairs"]["c2"]]}
        return RangeProofNi(int(d["ek"]["n"]), serde_bigint_native_parse(d["range"]), serde_bigint_native_parse(d["ciphertext"]),
                            pairs, [resp(o) for o in d["proof"]], d["error_factor"])


# ----------------------------------------------------------------------------------------------
# NiCorrectKeyProof (correct_key_ni.rs:42-117)
def mask_generation(out_length: int, seed: int) -> int:
    """correct_key_ni.rs:105-117."""
    msklen = out_length // DIGEST_SIZE + 1
    acc = 0
    for j in range(msklen):
        acc += compute_digest([seed, j]) << (j * DIGEST_SIZE)
    return acc


def correct_key_rho(n: int, salt: bytes):
    """correct_key_ni.rs:74-86 (verify) == :44-63 (proof)."""
    key_length = n.bit_length()
    salt_bn = compute_digest([bigint_from_bytes(salt)])
    return [mask_generation(key_length, compute_digest([n, salt_bn, i])) % n for i in range(M2)]


class NiCorrectKeyProof:
    def __init__(self, sigma_vec):
        self.sigma_vec = list(sigma_vec)

    @staticmethod
    def proof(p: int, q: int, salt=None):
        """correct_key_ni.rs:42-71 (dk = {p, q})."""
        n = p * q
        rho = correct_key_rho(n, SALT_STRING if salt is None else salt)
        return NiCorrectKeyProof([extract_nroot(p, q, r) for r in rho])

    def verify(self, n: int, salt: bytes):
        """correct_key_ni.rs:73-100."""
        import math

        rho = correct_key_rho(n, salt)
        gcd_test = math.gcd(primorial(), n)
        if len(self.sigma_vec) < M2:
            raise ReferencePanic("sigma_vec shorter than M2")
        derived = [pow(self.sigma_vec[i], n, n) for i in range(M2)]
        if rho == derived and gcd_test == 1:
            return
        raise IncorrectProof()

    def to_json(self):
        return json.dumps({"sigma_vec": [str(s) for s in self.sigma_vec]}, separators=(",", ":"))

    @staticmethod
    def from_json(s):
        return NiCorrectKeyProof([int(v) for v in json.loads(s)["sigma_vec"]])


# ----------------------------------------------------------------------------------------------
# Sigma protocols.  Randomness explicit: r_prime etc. are arguments.
class ZeroProof:
    """zero_enc_proof.rs:26-94.  witness r, statement (n, c)."""

    def __init__(self, z, a):
        self.z, self.a = z, a

    @staticmethod
    def prove(r, n, c, r_prime):
        nn = n * n
        a = paillier_encrypt(n, 0, r_prime)  # :46-52
        e = compute_digest([n, c, a])  # :54-58
        z = r_prime * pow(r, e, nn) % nn  # :60-61 (mod_mul reduces operands first; same value)
        return ZeroProof(z, a)

    def verify(self, n, c):
        e = compute_digest([n, c, self.a])
        c_z = paillier_encrypt(n, 0, self.z)
        c_z_test = paillier_add(n, paillier_mul(n, c, e), self.a)
        if c_z != c_z_test:
            raise IncorrectProof()


class CiphertextProof:
    """correct_ciphertext.rs:22-98.  witness (x, r), statement (n, c)."""

    def __init__(self, z1, z2, c_prime):
        self.z1, self.z2, self.c_prime = z1, z2, c_prime

    @staticmethod
    def prove(x, r, n, c, x_prime, r_prime):
        nn = n * n
        c_prime = paillier_encrypt(n, x_prime, r_prime)
        e = compute_digest([n, c, c_prime])
        z1 = x_prime + x * e  # unreduced (:59)
        z2 = r_prime * pow(r, e, nn) % nn
        return CiphertextProof(z1, z2, c_prime)

    def verify(self, n, c):
        e = compute_digest([n, c, self.c_prime])
        c_z = paillier_encrypt(n, self.z1, self.z2)
        c_z_test = paillier_add(n, paillier_mul(n, c, e), self.c_prime)
        if c_z != c_z_test:
            raise IncorrectProof()


def _mod_inv(a, m):
    """BigInt::mod_inv -> Option; callers unwrap() (multiplication_proof.rs:96,137)."""
    try:
        return pow(a, -1, m)
    except ValueError:
        raise ReferencePanic("mod_inv of a non-invertible value (unwrap on None)")


class MulProof:
    """multiplication_proof.rs:32-145.  witness (a,b,c,r_a,r_b,r_c), statement (n, e_a, e_b, e_c)."""

    def __init__(self, f, z1, z2, e_d, e_db):
        self.f, self.z1, self.z2, self.e_d, self.e_db = f, z1, z2, e_d, e_db

    @staticmethod
    def prove(a, b, c, r_a, r_b, r_c, n, e_a, e_b, e_c, d, r_d):
        nn = n * n
        e_d = paillier_encrypt(n, d, r_d)
        r_db = r_d * r_b  # unreduced (:70)
        db = d * b  # unreduced (:71)
        e_db = paillier_encrypt(n, db, r_db)
        e = compute_digest([n, e_a, e_b, e_c, e_d, e_db])
        ea = e * a % n
        f = (ea + d) % n
        z1 = pow(r_a, e, nn) * r_d % nn
        r_b_f = pow(r_b, f, nn)
        r_c_e = pow(r_c, e, nn)
        inv = _mod_inv(r_db * r_c_e % nn, nn)
        z2 = r_b_f * inv % nn
        return MulProof(f, z1, z2, e_d, e_db)

    def verify(self, n, e_a, e_b, e_c):
        nn = n * n
        e = compute_digest([n, e_a, e_b, e_c, self.e_d, self.e_db])
        enc_f_z1 = paillier_encrypt(n, self.f, self.z1)
        enc_0_z2 = paillier_encrypt(n, 0, self.z2)
        lhs1 = pow(e_a, e, nn) * self.e_d % nn
        inv = _mod_inv(self.e_db * pow(e_c, e, nn) % nn, nn)
        lhs2 = pow(e_b, self.f, nn) * inv % nn
        if not (lhs1 == enc_f_z1 and lhs2 == enc_0_z2):
            raise IncorrectProof()


def gen_phi(n, c, c_prime, y, y_prime, y_dp, r_y):
    """verlin_proof.rs:138-165."""
    return paillier_add(n, paillier_add(n, paillier_mul(n, c, y), paillier_mul(n, c_prime, y_prime)), paillier_encrypt(n, y_dp, r_y))


class VerlinProof:
    """verlin_proof.rs:34-134.  witness (x, x', x'', r_x), statement (n, c, c', phi_x)."""

    def __init__(self, phi_a, z, z_prime, z_double_prime, r_z):
        self.phi_a, self.z, self.z_prime, self.z_double_prime, self.r_z = phi_a, z, z_prime, z_double_prime, r_z

    @staticmethod
    def prove(x, x_prime, x_dp, r_x, n, c, c_prime, phi_x, a, a_prime, a_dp, r_a):
        nn = n * n
        phi_a = gen_phi(n, c, c_prime, a, a_prime, a_dp, r_a)
        e = compute_digest([n, c, c_prime, phi_x, phi_a])
        z = x * e + a
        zp = x_prime * e + a_prime
        zdp = x_dp * e + a_dp
        r_z = pow(r_x, e, nn) * r_a % nn
        return VerlinProof(phi_a, z, zp, zdp, r_z)

    def verify(self, n, c, c_prime, phi_x):
        e = compute_digest([n, c, c_prime, phi_x, self.phi_a])
        rhs = paillier_add(n, paillier_mul(n, phi_x, e), self.phi_a)
        phi_z = gen_phi(n, c, c_prime, self.z, self.z_prime, self.z_double_prime, self.r_z)
        if phi_z != rhs:
            raise IncorrectProof()


# ----------------------------------------------------------------------------------------------
# The remaining public proofs
def verify_opening(n, m, r, c) -> bool:
    """CorrectOpening::verify_opening, correct_opening.rs:17-30."""
    return c == paillier_encrypt(n, m, r)


DLOG_K, DLOG_K_PRIME, DLOG_SAMPLE_S = 128, 128, 256  # wi_dlog_proof.rs:19-21


class CompositeDLogProof:
    """wi_dlog_proof.rs:28-91.  statement (N, g, ni), secret s with ni = g^-s mod N; r is the prover's sample below
    2^(K + K' + SAMPLE_S) (explicit here; the reference draws it with sample_below, :52-53)."""

    def __init__(self, x, y):
        self.x, self.y = x, y

    @staticmethod
    def prove(N, g, ni, secret, r):
        x = pow(g, r, N)  # :54
        e = compute_digest([x, g, N, ni])  # :55-60
        return CompositeDLogProof(x, r + e * secret)  # :61 (unreduced)

    def verify(self, N, g, ni):
        import math

        if not N > 2 ** DLOG_K:  # :68 assert!
            raise ReferencePanic("N <= 2^K")
        if math.gcd(g, N) != 1 or math.gcd(ni, N) != 1:  # :71-72 assert_eq!
            raise ReferencePanic("g or ni not in Z_N*")
        e = compute_digest([self.x, g, N, ni])  # :74-79
        if self.x != pow(g, self.y, N) * pow(ni, e, N) % N:  # :80-85
            raise IncorrectProof()

    def to_json(self):
        """#[derive(Serialize)] with curv's native BigInt serde (RECALLED: hex of to_bytes), wi_dlog_proof.rs:28-32."""
        return json.dumps({"x": serde_bigint_native(self.x), "y": serde_bigint_native(self.y)}, separators=(",", ":"))

    @staticmethod
    def from_json(s):
        d = json.loads(s)
        return CompositeDLogProof(serde_bigint_native_parse(d["x"]), serde_bigint_native_parse(d["y"]))


CM_B = 256  # correct_message.rs:19


class CorrectMessageProof:
    """correct_message.rs:25-162.  Randomness explicit: r (encryption), e_rand / z_rand (the M-1 simulated branches,
    sample(B) and sample_below(n)), w."""

    def __init__(self, e_vec, z_vec, a_vec, ciphertext, valid_messages, n):
        self.e_vec, self.z_vec, self.a_vec, self.ciphertext, self.valid_messages, self.n = e_vec, z_vec, a_vec, ciphertext, valid_messages, n

    @staticmethod
    def _u_vec(n, ciphertext, valid_messages):
        nn = n * n
        return [ciphertext * _mod_inv((m * n + 1) % nn, nn) % nn for m in valid_messages]  # :50-56 / :134-142

    @staticmethod
    def prove(n, valid_messages, message, r, e_rand, z_rand, w):
        nn, M = n * n, len(valid_messages)
        ciphertext = paillier_encrypt(n, message, r)  # :43-49
        u = CorrectMessageProof._u_vec(n, ciphertext, valid_messages)

        def rnd(vec, j):
            if j >= len(vec):
                raise ReferencePanic("index out of bounds: the message is not one of the valid messages")
            return vec[j]

        a_vec, j = [], 0
        for i in range(M):  # :66-83
            if valid_messages[i] == message:
                a_vec.append(pow(w, n, nn))
            else:
                zi_n = pow(rnd(z_rand, j), n, nn)
                ui_ei_inv = _mod_inv(pow(u[i], rnd(e_rand, j), nn), nn)
                j += 1
                a_vec.append(zi_n * ui_ei_inv % nn)
        two_b = 1 << CM_B
        chal = compute_digest(a_vec) % two_b  # :85-87
        ei = (chal - sum(e_rand[: M - 1]) % two_b) % two_b  # :88-91
        zi = w * pow(r, ei, n) % n  # :92-93
        e_vec, z_vec, j = [], [], 0
        for i in range(M):  # :95-121
            if valid_messages[i] == message:
                e_vec.append(ei)
                z_vec.append(zi)
            else:
                e_vec.append(rnd(e_rand, j))
                z_vec.append(rnd(z_rand, j))
                j += 1
        return CorrectMessageProof(e_vec, z_vec, a_vec, ciphertext, list(valid_messages), n)

    def verify(self):
        n = self.n
        nn, two_b = n * n, 1 << CM_B
        chal = compute_digest(self.a_vec) % two_b  # :128-129
        if chal != sum(self.e_vec) % two_b:  # :130-133 assert_eq!
            raise ReferencePanic("chal != ei_sum")
        u = CorrectMessageProof._u_vec(n, self.ciphertext, self.valid_messages)
        ok = [pow(u[i], self.e_vec[i], nn) * self.a_vec[i] % nn == pow(self.z_vec[i], n, nn) for i in range(len(u))]  # :143-150
        if not all(ok):
            raise IncorrectProof()


# ----------------------------------------------------------------------------------------------
# serde codecs of src/serialize.rs (decimal strings)
def serialize_bigint(x: int) -> str:
    """serialize.rs:9-11"""
    return str(x)


def deserialize_bigint(s: str) -> int:
    """serialize.rs:23-26 (from_str_radix(s, 10))"""
    if not isinstance(s, str) or not s.lstrip("-").isdigit():
        raise ValueError("invalid decimal bigint")
    return int(s)


def serialize_vecbigint(xs) -> list:
    """serialize.rs:42-48"""
    return [str(x) for x in xs]


def deserialize_vecbigint(seq) -> list:
    """serialize.rs:58-73; a malformed element panics (unwrap, :69)."""
    out = []
    for s in seq:
        if not isinstance(s, str) or not s.lstrip("-").isdigit():
            raise ReferencePanic("from_str_radix(...).unwrap() on a malformed decimal string")
        out.append(int(s))
    return out


# ----------------------------------------------------------------------------------------------
# CorrectKey, the interactive proof of correct_key.rs:64-172 (STATISTICAL_ERROR_FACTOR = 40, :26).
class CorrectKeyProveError(Exception):
    """correct_key.rs:174-184"""


class CorrectKey:
    @staticmethod
    def challenge(n, s, r):
        """correct_key.rs:65-107 with the draws s_i, r_i <- sample_below(n) made explicit.
        Returns (Challenge {sn, e, z}, VerificationAid {s_digest})."""
        sn = [pow(si, n, n) for si in s]
        rn = [pow(ri, n, n) for ri in r]
        e = compute_digest([n] + sn + rn)
        z = [ri * pow(si, e, n) % n for ri, si in zip(r, s)]
        return {"sn": sn, "e": e, "z": z}, {"s_digest": compute_digest(s)}

    @staticmethod
    def prove(p, q, ch):
        """correct_key.rs:109-162"""
        import math

        n = p * q
        if any(math.gcd(n, v) != 1 for v in ch["sn"]):
            raise CorrectKeyProveError("`challenge.sn[i]` isn't co-prime with `n`")
        if any(math.gcd(n, v) != 1 for v in ch["z"]):
            raise CorrectKeyProveError("`challenge.z[i]` isn't co-prime with `n`")
        phi = (q - 1) * (p - 1)
        phimine = phi - (ch["e"] % phi)
        rn = [pow(zi, n, n) * pow(sni, phimine, n) % n for zi, sni in zip(ch["z"], ch["sn"])]
        if any(math.gcd(n, v) != 1 for v in rn):
            raise CorrectKeyProveError("`rn[i]` isn't co-prime with `n`")
        if ch["e"] != compute_digest([n] + ch["sn"] + rn):
            raise CorrectKeyProveError("`challenge.e` wasn't computed correctly")
        return {"s_digest": compute_digest([extract_nroot(p, q, sni) for sni in ch["sn"]])}

    @staticmethod
    def verify(proof, va):
        """correct_key.rs:164-171"""
        if proof["s_digest"] != va["s_digest"]:
            raise IncorrectProof()

    @staticmethod
    def challenge_to_json(ch):
        return json.dumps({"sn": [str(v) for v in ch["sn"]], "e": str(ch["e"]), "z": [str(v) for v in ch["z"]]}, separators=(",", ":"))
