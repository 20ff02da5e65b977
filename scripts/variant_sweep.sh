#!/bin/bash
# K1m tuning sweep on one B200: every layout / loop-structure variant (modexp2m.cu: Enc2mConfig) at window 5, then window 6.
# The variants exist only in the lab build: run `make lab` first.
export ZKP_B200_LIB=zk-paillier_b200/libzkp_b200_lab.so
mkdir -p gpurun_out
rm -f gpurun_out/k1m_variants.jsonl
for v in 0 1 2 3 4 5 6 7 8; do
  ZKP_B200_K1M_VARIANT=$v timeout 120 python scripts/k1m_variants.py 2048 v$v 2>&1 | tail -2
done
for v in ${SWEEP_W6:-0 1 3 7}; do
  ZKP_B200_K1M_VARIANT=$v ZKP_B200_K1M_WINDOW=6 timeout 120 python scripts/k1m_variants.py 2048 v${v}w6 2>&1 | tail -2
done
