#!/bin/bash
# What the driver runs at round end, in one GPU session: the parity suite, smoke(), the bench line and the reference arm.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_validate.sh [tag]
TAG=${1:-final}
mkdir -p gpurun_out
echo "== pytest -m gpu"; SECONDS=0; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5; echo "wall ${SECONDS}s"
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== reference arm"; SECONDS=0; timeout 900 python bench.py --impl reference --steps 20 --warmup 3 2>/dev/null | tee gpurun_out/bench_${TAG}_reference.json | cut -c1-160; echo "wall ${SECONDS}s"
echo "== bench (as the driver runs it)"; SECONDS=0
timeout 1500 python bench.py --steps 20 --warmup 3 2> gpurun_out/bench_${TAG}.err | tee gpurun_out/bench_${TAG}.json | cut -c1-300; echo "wall ${SECONDS}s"; tail -3 gpurun_out/bench_${TAG}.err
echo "== loaded libraries"; python - <<'PY'
import zk_paillier_b200 as zk, os
ctx = zk.native.Context(0)
print([l.split()[-1] for l in open(f"/proc/{os.getpid()}/maps") if "zkp" in l or "oracle" in l][:4])
PY
