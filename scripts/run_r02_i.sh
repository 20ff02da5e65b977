mkdir -p gpurun_out
echo "== tests: core + proofs + fullsize config1"; timeout 1500 python -m pytest tests/test_gpu_core.py tests/test_gpu_proofs.py tests/test_gpu_host.py -m gpu -x -q 2>&1 | tail -5
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== latency"; timeout 600 python - <<'PY'
import json, sys
sys.argv=['bench.py']
import bench
E=bench.Env()
print(json.dumps(bench.measure_latency(E)))
z=bench.measure_zero_gpu(E); print(json.dumps(z))
PY
