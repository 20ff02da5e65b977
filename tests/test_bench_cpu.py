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


def _line(name):
    return json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])


def test_recorded_bench_lines_carry_every_key_of_the_contract():
    """The lines a B200 printed (profiles/, committed): every key the driver and the judge read is there, with the units and
    the accounting the task states (value / e2e / roofline / cpu_baseline / clocks / gpu_launches; e2e copies declared)."""
    sys.path.insert(0, ROOT)
    import bench

    for name, gpus in (("r02_bench_final.json", 1), ("r02_bench_2gpu.json", 2), ("r02_bench_4gpu.json", 4), ("r02_bench_8gpu.json", 8)):
        d = _line(name)
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                  "config", "e2e", "roofline", "clocks", "gpu_launches"):
            assert k in d, (name, k)
        assert d["metric"] == "RangeProofNi proofs+verifies/sec at 2048-bit n" and d["unit"] == "proofs+verifies/s"
        assert d["n_gpus"] == gpus and d["scaling"] == "weak" and d["higher_is_better"] is True and d["vs_baseline"] is None
        assert d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["config"] == bench.headline_config(1024)
        # value = statements proven AND verified by all ranks over the max-over-ranks time of the timed steps
        assert abs(d["value"] - 1024 * gpus / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6
        e = d["e2e"]
        assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 1e8 and e["d2h_bytes_per_step"] > 1e8 and 0.9 < e["value"] / d["value"] < 1.01
        r = d["roofline"]
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.8 < r["frac"] < 0.9 and r["traffic"] > 0
        assert 1.6 < r["algorithmic_ratio"] < 1.8          # SURVEY 8d's count over the same peak: the kernel executes about half of it
        c = d["clocks"]
        assert c["sm_mhz"] >= 0.95 * c["sm_max_mhz"] and not any("slowdown" in x for x in c["reasons"])
        sec = d["secondary"]
        assert {"correct_key_3072", "mul_verlin_4096"} <= set(sec)
        for v in sec.values():
            assert "error" not in v, (name, v)
            if "clocks" in v:
                assert not any("slowdown" in x for x in v["clocks"]["reasons"])
    one = _line("r02_bench_final.json")
    cb = one["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["unit"] == one["unit"] and "proofs" in cb["sample"] and 20 < one["value"] / cb["value"] < 200
    assert {"latency_one_proof", "zero_1024_cpu"} <= set(one["secondary"])
    ref = _line("r02_bench_reference_arm_final.json")
    assert ref["impl"] == "reference" and ref["config"] == one["config"] and ref["metric"] == one["metric"] and ref["unit"] == one["unit"]
    # weak scaling as recorded: 2 / 4 / 8 GPUs within 1 % of N times one GPU
    for name, gpus in (("r02_bench_2gpu.json", 2), ("r02_bench_4gpu.json", 4), ("r02_bench_8gpu.json", 8)):
        assert 0.99 < _line(name)["value"] / (gpus * one["value"]) < 1.01
