mkdir -p gpurun_out
echo "== sigma tests (all layouts, both row forms)"; timeout 2000 python -m pytest tests/test_gpu_sigma.py tests/test_gpu_core.py tests/test_gpu_proofs.py -m gpu -x -q 2>&1 | tail -5
echo "== latency breakdown"; timeout 300 python scripts/latency_breakdown.py 2>&1 | tail -5
echo "== fill curve"; FILL_B=512,1536,3072 timeout 600 python scripts/fill_curve.py 4096 2>&1 | tail -10
echo "== sigma bench"; timeout 900 python bench.py --config sigma --no-cpu 2> gpurun_out/sigma.err | tee gpurun_out/bench_sigma_r02b.json | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step')}, {k:d['e2e'][k] for k in ('value','sequential','two_contexts')}, {k:d['roofline'][k] for k in ('frac','frac_two_contexts')})"; tail -3 gpurun_out/sigma.err
