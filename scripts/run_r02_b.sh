mkdir -p gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests/test_gpu_sigma.py tests/test_gpu_more.py -m gpu -x -q 2>&1 | tail -5
echo "== config5"; timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k config5 2>&1 | tail -4
for sh in 0 1 2; do echo "== sigma bench shape $sh"; timeout 600 python bench.py --config sigma --jobs-shape $sh --no-cpu 2>&1 | tail -1 | cut -c1-330; done
echo "== fill curve"; FILL_B=148,512,1024,1536,3072 timeout 600 python scripts/fill_curve.py 4096 2>&1 | tail -20
