mkdir -p gpurun_out
echo "== full gpu suite"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench default (driver shape: 20 steps)"; /usr/bin/time -v timeout 1500 python bench.py --steps 20 --warmup 3 2> gpurun_out/bench_r02_final.err | tee gpurun_out/bench_r02_final.json | cut -c1-300; grep -E "Elapsed|Traceback|Error" gpurun_out/bench_r02_final.err | head
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 20 --warmup 3 2>/dev/null | tee gpurun_out/bench_r02_reference.json | cut -c1-200
