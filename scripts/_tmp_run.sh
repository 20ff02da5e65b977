mkdir -p gpurun_out
for nb in 3072:256 4096:128; do bits=${nb%%:*}; b=${nb##*:}; echo "== RangeProofNi at $bits bits, batch $b"; timeout 900 python bench.py --n-bits $bits --batch $b --steps 3 --warmup 3 --no-secondary 2> /dev/null | tee gpurun_out/bench_rp_${bits}_r02.json | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'frac', round(d['roofline']['frac'],3), 'cpu', round(d['cpu_baseline']['value'],2), d['clocks']['sm_mhz'], d['clocks']['reasons'])"; done
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 900 $NCU -k 'regex:modexp2m_jobs' -s 12 -c 4 -o /tmp/p_k2h python bench.py --config sigma --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
python scripts/ncu_digest.py /tmp/p_k2h.ncu-rep gpurun_out/r02_k2h_sigma_ncu_digest.json "K2h launches of one MulProof verify x512 + VerlinProof verify x512 step at 4096-bit n (final build: single rows in the MulProof launch, pair rows in the latency-bound VerlinProof launch)" | python -c "
import sys,json
d=json.loads(sys.stdin.read()); d=d if isinstance(d,list) else [d]
for k in d:
    m=k['metrics']; g=lambda n: m.get(n,{}).get('value')
    print(k['kernel'][:50], g('gpu__time_duration.sum'), g('sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed'), g('sm__issue_active.avg.pct_of_peak_sustained_elapsed'))"
