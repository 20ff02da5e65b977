mkdir -p gpurun_out
echo "== new host tests"; timeout 900 python -m pytest tests/test_gpu_host.py -m gpu -x -q -k "mixed_keys or json_nesting or sigma_protocols or correct_key" 2>&1 | tail -6
echo "== ref vectors (mock) on gpu"; timeout 600 python -m pytest tests/test_reference_vectors.py -m gpu -x -q 2>&1 | tail -5
echo "== config2 distinct"; timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "config2" 2>&1 | tail -6
echo "== sigma bench"; timeout 900 python bench.py --config sigma --no-cpu 2> gpurun_out/sigma.err | tee gpurun_out/bench_sigma_r02.json | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e'], {k:d['roofline'][k] for k in ('frac','frac_two_contexts','algorithmic_ratio')})"; tail -3 gpurun_out/sigma.err
echo "== unroll variants (lab)"; for u in 1 2 4; do ZKP_B200_K2H_UNROLL=$u ZKP_B200_LIB=zk-paillier_b200/libzkp_b200_lab.so FILL_B=1536,3072 timeout 300 python scripts/fill_curve.py 4096 2>&1 | grep narrow; done
echo "== correct_key bench"; timeout 900 python bench.py --config correct_key 2> gpurun_out/ck.err | tee gpurun_out/bench_ck_r02.json | cut -c1-1800; tail -3 gpurun_out/ck.err
