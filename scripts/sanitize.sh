#!/bin/bash
# compute-sanitizer over smoke() and a reduced -m gpu subset that launches every kernel family at small widths:
# K1m (narrow + WIDE entry), K1, K2m, K2, K3, K4, the binary extended-Euclid mod_inv, the RangeProofNi / NiCorrectKey /
# sigma-protocol glue. Usage (under gpurun, repo root):  bash scripts/sanitize.sh [tag] [tools]
# Logs land in gpurun_out/san_<tool>_<tag>.log; scripts/sanitize_digest.py turns them into profiles/<tag>_sanitizer.json.
TAG=${1:-r02}
TOOLS=${2:-"memcheck racecheck synccheck initcheck"}
SUBSET='tests/test_gpu_core.py tests/test_gpu_proofs.py tests/test_gpu_sigma.py tests/test_gpu_more.py'
SEL='(1024 or sha256 or small_exponents or leading_zero or edge_cases or two_phase) and not 8192 and not 6144 and not 4096 and not 3072 and not 2048'
mkdir -p gpurun_out
for tool in $TOOLS; do
  extra=""
  [ "$tool" = racecheck ] && extra="--racecheck-report all"
  [ "$tool" = memcheck ] && extra="--leak-check full"
  echo "== $tool: smoke()"
  timeout 900 compute-sanitizer --tool $tool $extra --error-exitcode 77 --log-file gpurun_out/san_${tool}_smoke_${TAG}.log \
      python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_${tool}_smoke_${TAG}.out 2>&1
  echo "exit $?"; tail -2 gpurun_out/san_${tool}_smoke_${TAG}.out; tail -3 gpurun_out/san_${tool}_smoke_${TAG}.log
  echo "== $tool: pytest subset"
  timeout 1500 compute-sanitizer --tool $tool $extra --error-exitcode 77 --log-file gpurun_out/san_${tool}_tests_${TAG}.log \
      python -m pytest $SUBSET -m gpu -q -x -k "$SEL" > gpurun_out/san_${tool}_tests_${TAG}.out 2>&1
  echo "exit $?"; tail -3 gpurun_out/san_${tool}_tests_${TAG}.out; tail -3 gpurun_out/san_${tool}_tests_${TAG}.log
done
