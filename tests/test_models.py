"""CPU suite: executable models of the device algorithms, lane by lane and limb by limb, checked against Python
big-int arithmetic.  They were written BEFORE the kernels ran on a GPU and are kept as the specification of
  * the split-accumulator CIOS Montgomery step of mp_coop.cuh (cios_model.py),
  * the plain product on the same rows with the low half captured at lane 0 (mulwide_model.py),
  * Montgomery multiplication modulo n^2 in two-digit base-n form, the default encryption kernel (mont2d_model.py),
  * the symmetric squaring (each pair of lane blocks once) + reduction-only rows of K1m variant 9 (sqr_sym_model.py),
  * the pair rows of K2h's narrow-lane and latency layouts: two multiplier limbs and a two-limb quotient per step
    (cios_pair_model.py; every instantiated lane layout, initial accumulator, two-product form, the bounds of the hanging limbs),
  * the block fetch of the one-warp-per-transcript SHA-256 (K4w): byte offsets of minimal big-endian items, words by funnel
    shift of little-endian limbs or byte by byte across items and padding, schedule expanded per lane (sha_fetch_model.py)."""
import os
import runpy

import pytest

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "models")


@pytest.mark.parametrize("name", ["cios_model.py", "mulwide_model.py", "mont2d_model.py", "sqr_sym_model.py", "cios_pair_model.py", "sha_fetch_model.py"])
def test_model(name, capsys):
    runpy.run_path(os.path.join(HERE, name), run_name="__main__")
    out = capsys.readouterr().out
    assert "ok" in out
