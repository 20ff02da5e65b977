#!/bin/bash
# One `ncu --set full` capture per kernel family other than K1m (which gpu_round.sh captures): K2 (per-instance modulus,
# NiCorrectKeyProof), K2m (two-digit, per-job exponent; MulProof / VerlinProof), K4 (SHA-256 transcript), K3 (modmul select).
# Reports land in gpurun_out/; scripts/ncu_digest.py turns them into the committed digests under profiles/.
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:modexp_var_kernel -c 1 -o gpurun_out/prof_k2 python bench.py --config correct_key --batch 1024 --steps 1 --warmup 0 --e2e-steps 0 > gpurun_out/ncu_k2.log 2>&1
timeout 600 $NCU -k regex:modexp2m_var_kernel -s 2 -c 1 -o gpurun_out/prof_k2m python bench.py --config sigma --batch 512 --steps 1 --warmup 0 > gpurun_out/ncu_k2m.log 2>&1
timeout 600 $NCU -k regex:sha256_transcript -s 1 -c 1 -o gpurun_out/prof_k4 python bench.py --batch 1024 --steps 1 --warmup 1 --no-cpu --e2e-steps 0 > gpurun_out/ncu_k4.log 2>&1
timeout 600 $NCU -k regex:modmul_select -s 1 -c 1 -o gpurun_out/prof_k3 python bench.py --batch 1024 --steps 1 --warmup 1 --no-cpu --e2e-steps 0 > gpurun_out/ncu_k3.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/ncu_k2.log gpurun_out/ncu_k2m.log gpurun_out/ncu_k4.log gpurun_out/ncu_k3.log | cut -c1-200
