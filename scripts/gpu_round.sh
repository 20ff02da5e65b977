#!/bin/bash
# One GPU session (round 2 recipe): parity tests, smoke, the bench line, the ncu launch list and one full capture per kernel
# family, digested ON THE BOX (gpurun copies back at most 64 MiB: the .ncu-rep files stay in /tmp there).
# Usage (from the repo root, under gpurun):  bash scripts/gpu_round.sh [tag] [tests]
TAG=${1:-r02}
TESTS=${2:-tests}
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then echo "== pytest -m gpu ($TESTS)"; timeout 2400 python -m pytest $TESTS -m gpu -x -q 2>&1 | tail -6; fi
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (as the driver runs it)"; SECONDS=0
timeout 1500 python bench.py --steps 20 --warmup 3 2> gpurun_out/bench_${TAG}.err | tee gpurun_out/bench_${TAG}.json | cut -c1-400; echo "wall ${SECONDS}s"; tail -3 gpurun_out/bench_${TAG}.err
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 20 --warmup 3 2>/dev/null | tee gpurun_out/bench_${TAG}_reference.json | cut -c1-200
for c in dlog correct_message; do timeout 600 python bench.py --config $c 2> /dev/null > gpurun_out/bench_${c}_${TAG}.json; done
echo "== ncu launch list of a bench step"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 --no-secondary > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
digest() { python scripts/ncu_digest.py /tmp/$1.ncu-rep gpurun_out/${TAG}_$2_ncu_digest.json "$3" > /dev/null; ncu -i /tmp/$1.ncu-rep --page details > gpurun_out/${TAG}_$2_ncu_details.txt 2>&1; }
echo "== ncu full: K1m, K2 (NiCorrectKey), K2h (MulProof + VerlinProof), K4w"
timeout 900 $NCU -k 'regex:enc2m' -s 1 -c 1 -o /tmp/p_k1m python bench.py --batch 148 --steps 1 --warmup 1 --no-cpu --e2e-steps 1 --no-secondary > /dev/null 2>&1
digest p_k1m k1m "K1m enc2m_kernel<8,8>, one wave (bench.py --batch 148)"
timeout 900 $NCU -k 'regex:modexp_var_kernel' -s 4 -c 1 -o /tmp/p_k2 python bench.py --config correct_key --batch 1024 --steps 1 --no-cpu > /dev/null 2>&1
digest p_k2 k2 "K2 modexp_var_kernel<8,12>, NiCorrectKeyProof verify x1024 at 3072 bits (distinct moduli)"
timeout 900 $NCU -k 'regex:modexp2m_jobs' -s 12 -c 4 -o /tmp/p_k2h python bench.py --config sigma --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
digest p_k2h k2h_sigma "K2h launches of one MulProof verify x512 + VerlinProof verify x512 step at 4096-bit n"
timeout 900 $NCU -k 'regex:sha256_transcript_warp' -s 2 -c 1 -o /tmp/p_k4w python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 --no-secondary > /dev/null 2>&1
digest p_k4w k4w "K4w sha256_transcript_warp_kernel, 1024 RangeProofNi transcripts of 131 kB"
ls -la gpurun_out | tail -20; du -sh gpurun_out
