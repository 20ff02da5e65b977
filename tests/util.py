"""Shared helpers for the test-suite (fixtures, oracle imports)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import c_oracle  # noqa: E402  (oracle/c_oracle.py)
import zkp_oracle as po  # noqa: E402  (oracle/zkp_oracle.py)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def keys(bits):
    """[(p, q)] fixtures: the reference's fixed test key first for 2048, then tests/golden/keys.json."""
    d = json.load(open(os.path.join(GOLDEN, "keys.json")))
    ks = [(int(k["p"]), int(k["q"])) for k in d[str(bits)]]
    if bits == 2048:
        ks.insert(0, (po.TEST_P, po.TEST_Q))
    return ks


def limbs_for(bits):
    l = (bits + 31) // 32
    return l + (-l) % 4
