"""CPU suite: the C++ host mirror's own big-integer / codec layer (no GPU needed) against Python ints and the
oracle's sampling rule."""
import random

from hostlib import Stream, call
from util import po


def test_bigint_against_python_ints():
    rng = random.Random(1)
    cases = [(0, 1), (1, 1), (2**32 - 1, 2**32), (2**64, 2**32 - 1), (10**40, 10**20 + 1)]
    for _ in range(200):
        ab, bb = rng.choice([8, 64, 300, 2048, 4096]), rng.choice([8, 33, 64, 1024, 2047])
        cases.append((rng.getrandbits(ab), rng.getrandbits(bb) | 1))
    for a, b in cases:
        r = call("bigint.selftest", a=str(a), b=str(b))
        assert r["ok"], r
        assert int(r["sum"]) == a + b and int(r["prod"]) == a * b
        assert int(r["quot"]) == a // b and int(r["rem"]) == a % b
        if a >= b:
            assert int(r["diff"]) == a - b
        import math
        g = math.gcd(a, b)
        assert int(r["gcd"]) == g
        if g == 1 and b > 1:
            assert int(r["inv"]) == pow(a, -1, b)
        else:
            assert "inv" not in r or b == 1
        assert r["hex"] == po.serde_bigint_native(a) and r["bits"] == a.bit_length() and r["roundtrip"] is True
        assert int(r["or"]) == a | b and int(r["shl"]) == a << 37 and int(r["shr"]) == a >> 37 and r["mod_small"] == a % 4093


def test_knuth_division_corner_cases():
    # qhat over-estimation and add-back paths of algorithm D
    B = 2**32
    cases = [((B**4 - 1), (B**2 - 1)), (B**3, B**2 - 1), ((B - 1) * B**3, (B // 2) * B + 1), (B**5 - B**2, B**3 - 1),
             (0x7fffffff800000010000000000000000, 0x800000008000000200000005)]
    for a, b in cases:
        r = call("bigint.selftest", a=str(a), b=str(b))
        assert int(r["quot"]) == a // b and int(r["rem"]) == a % b


def test_sampling_rule_matches_oracle():
    data = random.Random(2).randbytes(4000)
    lo, hi = 12345, (1 << 255) // 3
    r = call("sample", lo=str(lo), hi=str(hi), count=20, rng_hex=data.hex())
    s = Stream(data)
    assert [int(v) for v in r["values"]] == [po.sample_range(s, lo, hi) for _ in range(20)]


def test_errors_are_reported_not_thrown():
    r = call("bigint.selftest", a="12x", b="1")
    assert r["ok"] is False and r["kind"] == "error"
    r = call("nope")
    assert r["ok"] is False


def test_bigint_on_adversarial_limb_patterns():
    """Operands built from the limbs that break multiword arithmetic - 0, 1, 2^31, 2^32 - 1, 2^32 - 2 and a random one -
    at every length up to 9 limbs: carries that ripple across the whole number, borrows out of zero limbs, the
    over-estimated quotient digits of algorithm D (divisor top limb 2^31 or 2^32 - 1, dividend top limbs equal to it)."""
    import math

    rng = random.Random(7)
    pat = [0, 1, 2**31, 2**32 - 1, 2**32 - 2]

    def build(nl):
        v = 0
        for _ in range(nl):
            v = (v << 32) | (rng.choice(pat) if rng.random() < 0.8 else rng.getrandbits(32))
        return v

    checked = 0
    for la in range(1, 10):
        for lb in range(1, 7):
            for _ in range(12):
                a, b = build(la), build(lb)
                if b == 0:
                    b = 1
                r = call("bigint.selftest", a=str(a), b=str(b))
                assert r["ok"], (a, b, r)
                assert int(r["sum"]) == a + b and int(r["prod"]) == a * b, (a, b)
                assert int(r["quot"]) == a // b and int(r["rem"]) == a % b, (a, b)
                if a >= b:
                    assert int(r["diff"]) == a - b, (a, b)
                g = math.gcd(a, b)
                assert int(r["gcd"]) == g, (a, b)
                if g == 1 and b > 1:
                    assert int(r["inv"]) == pow(a, -1, b), (a, b)
                assert r["hex"] == po.serde_bigint_native(a) and r["bits"] == a.bit_length() and r["roundtrip"] is True
                assert int(r["or"]) == a | b and int(r["shl"]) == a << 37 and int(r["shr"]) == a >> 37 and r["mod_small"] == a % 4093
                checked += 1
    assert checked == 9 * 6 * 12


def test_request_json_is_untrusted_input():
    """The request text is parsed by the mirror's own JSON reader (host/json.hpp): malformed, truncated, deeply nested or
    oddly escaped input is an error reply, never a crash or an exception across the C boundary."""
    import ctypes as C
    import json
    import os

    from util import ROOT

    call("bigint.selftest", a="1", b="1")  # loads the library
    import hostlib

    lib = hostlib._lib

    def raw(op, text):
        p = lib.zkh_call(op.encode(), text)
        try:
            return json.loads(C.string_at(p).decode())
        finally:
            lib.zkh_free(p)

    good = raw("bigint.selftest", b'{"a": "6", "b": "4", "extra": [1, -7, true, false, null, {"k": "\\u00e9\\n\\t\\"\\\\"}]}')
    assert good["ok"] and good["sum"] == "10"
    for bad in (b"", b"{", b'{"a": "6", "b": "4"', b'{"a": "6" "b": "4"}', b'{"a": "6", "b": "4"} trailing', b'{"a": "\\uZZZZ", "b": "1"}',
                b'{"a": "6", "b": "4", "x": ' + b"[" * 100000 + b"}", b'{"a": "6", "b": "4", "x": ' + b'{"y":' * 5000 + b"1" + b"}" * 5000 + b"}",
                b"\xff\xfe\x00garbage", b'{"a": 6, "b": 4}', b'["a", "b"]', b"null",
                b'{"a": "6", "b": "4", "x": 2.5e3}'):  # the wire format has integers only (u8 / usize): a fraction or exponent is refused
        r = raw("bigint.selftest", bad)
        assert r["ok"] is False and r.get("kind") == "error", (bad[:40], r)
    # a very long decimal string is data, not a parser problem
    big = "9" * 4000
    r = raw("bigint.selftest", json.dumps({"a": big, "b": "7"}).encode())
    assert r["ok"] and int(r["rem"]) == int(big) % 7
