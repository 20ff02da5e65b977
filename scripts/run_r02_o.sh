mkdir -p gpurun_out
TAG=r02c
SUBSET='tests/test_gpu_sigma.py tests/test_gpu_proofs.py tests/test_gpu_core.py'
SEL='(1024 or sha256 or two_digit) and not 2048 and not 2047 and not 3072 and not 4096 and not 1536 and not 6144 and not 8192 and not large_batch'
for tool in memcheck racecheck synccheck initcheck; do
  extra=""; [ "$tool" = racecheck ] && extra="--racecheck-report all"; [ "$tool" = memcheck ] && extra="--leak-check full"
  timeout 1200 compute-sanitizer --tool $tool $extra --error-exitcode 77 --log-file gpurun_out/san_${tool}_${TAG}.log \
      python -m pytest $SUBSET -m gpu -q -x -k "$SEL" > gpurun_out/san_${tool}_${TAG}.out 2>&1
  echo "$tool exit $?"; tail -1 gpurun_out/san_${tool}_${TAG}.out; grep -E "SUMMARY" gpurun_out/san_${tool}_${TAG}.log
done
