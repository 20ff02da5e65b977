mkdir -p gpurun_out
TAG=r02
rm -f gpurun_out/*.ncu-rep
echo "== ncu launch list of a bench step"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 --no-secondary > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
wc -l gpurun_out/launches_${TAG}.csv
echo "== ncu full: K1m"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:enc2m' -s 1 -c 1 -f -o /tmp/prof_k1m_${TAG} \
    python bench.py --batch 148 --steps 1 --warmup 1 --no-cpu --e2e-steps 1 --no-secondary > gpurun_out/ncu_k1m_${TAG}.log 2>&1
python scripts/ncu_digest.py /tmp/prof_k1m_${TAG}.ncu-rep gpurun_out/${TAG}_k1m_ncu_digest.json "K1m enc2m_kernel<8,8>, one wave (bench.py --batch 148)" > /dev/null
ncu -i /tmp/prof_k1m_${TAG}.ncu-rep --page details > gpurun_out/${TAG}_k1m_ncu_details.txt 2>&1
echo "== ncu full: K2h (sigma 512+512)"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:modexp2m_jobs' -s 12 -c 4 -f -o /tmp/prof_k2h_${TAG} \
    python bench.py --config sigma --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_k2h_${TAG}.log 2>&1
python scripts/ncu_digest.py /tmp/prof_k2h_${TAG}.ncu-rep gpurun_out/${TAG}_k2h_sigma_ncu_digest.json "K2h launches of one MulProof verify x512 + VerlinProof verify x512 step at 4096-bit n (short launch, long launch per call)" > /dev/null
ncu -i /tmp/prof_k2h_${TAG}.ncu-rep --page details > gpurun_out/${TAG}_k2h_sigma_ncu_details.txt 2>&1
echo "== ncu full: K4w"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:sha256_transcript_warp' -s 2 -c 1 -f -o /tmp/prof_k4w_${TAG} \
    python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 --no-secondary > gpurun_out/ncu_k4w_${TAG}.log 2>&1
python scripts/ncu_digest.py /tmp/prof_k4w_${TAG}.ncu-rep gpurun_out/${TAG}_k4w_ncu_digest.json "K4w sha256_transcript_warp_kernel, 1024 RangeProofNi transcripts of 131 kB" > /dev/null
ls -la gpurun_out | tail -12; du -sh gpurun_out
