mkdir -p gpurun_out
echo "== sigma tests"; timeout 1500 python -m pytest tests/test_gpu_sigma.py tests/test_gpu_more.py -m gpu -x -q 2>&1 | tail -5
echo "== config5 + host sigma"; timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_host.py -m gpu -x -q -k "config5 or sigma or mixed" 2>&1 | tail -4
echo "== sigma bench"; timeout 900 python bench.py --config sigma --no-cpu 2> gpurun_out/sigma.err | tee gpurun_out/bench_sigma_r02.json | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step')}, {k:d['e2e'][k] for k in ('value','sequential','two_contexts')}, {k:d['roofline'][k] for k in ('frac','frac_two_contexts','algorithmic_ratio')})"; tail -3 gpurun_out/sigma.err
echo "== sigma bench B=4096"; timeout 900 python bench.py --config sigma --no-cpu --batch 4096 --steps 2 2> gpurun_out/sigma2.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step')}, {k:d['e2e'][k] for k in ('value','sequential','two_contexts')}, {k:d['roofline'][k] for k in ('frac','frac_two_contexts','algorithmic_ratio')})"; tail -3 gpurun_out/sigma2.err
