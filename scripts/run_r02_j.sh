mkdir -p gpurun_out
echo "== 2-GPU bench (torchrun, NCCL)"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 2> gpurun_out/bench_2gpu_r02.err | grep '^{' | tee gpurun_out/bench_2gpu_r02.json | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('headline', d['value'], 'e2e', d['e2e']['value'], 'n_gpus', d['n_gpus'])
for k,v in (d.get('secondary') or {}).items(): print(k, v.get('value'), (v.get('e2e') or {}).get('value'), v.get('n_gpus'))"
tail -3 gpurun_out/bench_2gpu_r02.err
echo "== 2-GPU reference arm"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 --ref-seconds 20 2>/dev/null | grep '^{' | cut -c1-300
