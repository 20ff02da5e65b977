mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q --durations=60 > gpurun_out/gpu_durations.txt 2>&1; tail -70 gpurun_out/gpu_durations.txt
