mkdir -p gpurun_out
TAG=r02
echo "== sanitizers on the new kernels (K2h all layouts, multi-base, K4w, keygen)"
SUBSET='tests/test_gpu_sigma.py tests/test_gpu_proofs.py'
SEL='(1024 or sha256) and not 2048 and not 3072 and not 4096 and not large_batch'
for tool in memcheck racecheck synccheck initcheck; do
  extra=""; [ "$tool" = racecheck ] && extra="--racecheck-report all"; [ "$tool" = memcheck ] && extra="--leak-check full"
  timeout 1200 compute-sanitizer --tool $tool $extra --error-exitcode 77 --log-file gpurun_out/san_${tool}_tests_${TAG}b.log \
      python -m pytest $SUBSET -m gpu -q -x -k "$SEL" > gpurun_out/san_${tool}_tests_${TAG}b.out 2>&1
  echo "$tool exit $?"; tail -1 gpurun_out/san_${tool}_tests_${TAG}b.out; tail -2 gpurun_out/san_${tool}_tests_${TAG}b.log
done
echo "== ncu launch list of a bench step"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 --no-secondary > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
tail -2 gpurun_out/launches_${TAG}.csv | cut -c1-200
echo "== ncu full: K1m"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:enc2m' -s 1 -c 1 -f -o gpurun_out/prof_k1m_${TAG} \
    python bench.py --batch 148 --steps 1 --warmup 1 --no-cpu --e2e-steps 1 --no-secondary > gpurun_out/ncu_k1m_${TAG}.log 2>&1
echo "== ncu full: K2h (sigma 512+512) and K4w"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:modexp2m_jobs' -s 4 -c 4 -f -o gpurun_out/prof_k2h_${TAG} \
    python bench.py --config sigma --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_k2h_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:sha256_transcript_warp' -s 2 -c 1 -f -o gpurun_out/prof_k4w_${TAG} \
    python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 --no-secondary > gpurun_out/ncu_k4w_${TAG}.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
