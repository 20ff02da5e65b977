mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
FILL_B=1536 timeout 900 $NCU -k regex:modexp2m_jobs -s 2 -c 2 -o gpurun_out/prof_k2h_b1536 python scripts/fill_curve.py 4096 > gpurun_out/ncu_k2h.log 2>&1
tail -3 gpurun_out/ncu_k2h.log
ls -la gpurun_out/prof_k2h*
