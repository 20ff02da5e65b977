"""CPU suite: the reference arm of bench.py runs without a GPU and prints the contract's JSON line; the algorithmic and
executed work formulas agree with SURVEY.md section 8d."""
import json
import os
import subprocess
import sys

from util import ROOT


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-seconds", "3",
                        "--no-secondary"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "RangeProofNi proofs+verifies/sec at 2048-bit n"
    assert line["unit"] == "proofs+verifies/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "proofs+verifies/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the same static config dict as the CUDA arm prints (the driver compares the two arms' configs)
    sys.path.insert(0, ROOT)
    import bench

    assert line["config"] == bench.headline_config(1024)
    assert "batch=1024 per GPU, 2048-bit n" in line["config"]["workload"]


def test_work_formulas():
    sys.path.insert(0, ROOT)
    import bench

    # Enc = modexp(2|n|, |n|) = 81.9 M / 274.9 M / 649.8 M multiply-adds at 2048 / 3072 / 4096 bits
    assert round(bench.modexp_imads(4096, 2048) / 1e6, 1) == 81.9
    assert round(bench.modexp_imads(6144, 3072) / 1e6, 1) == 274.9
    assert round(bench.modexp_imads(8192, 4096) / 1e6, 1) == 649.8
    assert round(bench.modexp_imads(3072, 3072) / 1e6, 1) == 68.9
    # the two-digit kernels execute about half of the schoolbook count; a three-base job shares one squaring chain
    one = bench.k2m_executed(4096, 4096)
    assert 0.45 < one / bench.modexp_imads(8192, 4096) < 0.56
    assert bench.k2m_executed(4096, 4480, nbase=3) < 0.55 * (2 * bench.k2m_executed(4096, 4480) + one)
