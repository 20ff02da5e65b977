mkdir -p gpurun_out
echo "== full gpu suite"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench default"; timeout 1500 python bench.py 2> gpurun_out/bench_r02a.err | tee gpurun_out/bench_r02a.json | cut -c1-600; tail -5 gpurun_out/bench_r02a.err
echo "== nosub probe"; for lib in libzkp_b200_lab.so libzkp_b200_lab_nosub.so; do K1M_NO_CHECK=1 ZKP_B200_LIB=zk-paillier_b200/$lib timeout 200 python scripts/k1m_variants.py 2048 $lib 2>&1 | tail -1; done
