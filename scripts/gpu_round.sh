#!/bin/bash
# One GPU session: parity tests, smoke, bench, ncu launch list and one full capture of the top kernel.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_round.sh [tag] [tests]
TAG=${1:-r01}
TESTS=${2:-tests}
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then echo "== pytest -m gpu ($TESTS)"; timeout 900 python -m pytest $TESTS -m gpu -x -q 2>&1 | tail -15; fi
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== bench"; timeout 600 python bench.py 2> gpurun_out/bench_${TAG}.err | tee gpurun_out/bench_${TAG}.json | cut -c1-1500
tail -5 gpurun_out/bench_${TAG}.err
if [ -n "$BIG_BATCH" ]; then echo "== bench at the per-GPU share of configs[3] (batch $BIG_BATCH)"; timeout 900 python bench.py --batch $BIG_BATCH --steps 1 --warmup 1 --e2e-steps 1 --no-cpu 2> gpurun_out/bench_${TAG}_b${BIG_BATCH}.err | tee gpurun_out/bench_${TAG}_b${BIG_BATCH}.json | cut -c1-400; fi
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
tail -3 gpurun_out/launches_${TAG}.csv | cut -c1-300
echo "== ncu full capture of the encryption kernel (K1m; K1 when ZKP_B200_ENC=k1)"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:enc2m|modexp_shared' -s 1 -c 2 -f -o gpurun_out/prof_enc_${TAG} \
    python bench.py --batch 148 --steps 1 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_full_${TAG}.log | cut -c1-300
ls -la gpurun_out
