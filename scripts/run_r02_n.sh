mkdir -p gpurun_out
echo "== 2-GPU bench (solo-measured secondary entries, one reduction)"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 2> gpurun_out/bench_2gpu_r02.err | grep '^{' | tee gpurun_out/bench_2gpu_r02.json | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('headline', d['value'], 'e2e', d['e2e']['value'], 'n_gpus', d['n_gpus'])
for k,v in (d.get('secondary') or {}).items(): print(k, v.get('error') or (v.get('value'), (v.get('e2e') or {}).get('value'), v.get('n_gpus'), v.get('aggregation')))"
tail -3 gpurun_out/bench_2gpu_r02.err
echo "== 2-GPU --config sigma / correct_key"
for c in sigma correct_key; do timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --config $c --steps 3 --no-cpu 2>/dev/null | grep '^{' | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['metric'], d['n_gpus'], round(d['value']), round(d['e2e']['value']))"; done
