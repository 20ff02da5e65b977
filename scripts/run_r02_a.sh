mkdir -p gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests/test_gpu_sigma.py tests/test_gpu_more.py tests/test_gpu_core.py -m gpu -x -q 2>&1 | tail -8
echo "== config5"; timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k config5 2>&1 | tail -4
for sh in 0 1 2; do echo "== sigma bench shape $sh"; timeout 600 python bench.py --config sigma --jobs-shape $sh --no-cpu 2>&1 | tail -1 | cut -c1-400; done
echo "== fill curve"; timeout 600 python scripts/fill_curve.py 4096 2>&1 | tail -20
echo "== probes"; ZKP_B200_LIB=zk-paillier_b200/libzkp_b200_lab.so timeout 600 python scripts/imad_probes.py 2>&1 | tail -20
