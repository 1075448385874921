#!/bin/bash
mkdir -p gpurun_out
python scripts/rts_bench.py 65536 16 2>&1 | tail -1
timeout 900 python bench.py --workload live_rts --rts-steps 1000 --steps 2 --no-cpu-baseline > gpurun_out/r02k_bench_rts.json 2> gpurun_out/r02k_bench_rts.err; tail -3 gpurun_out/r02k_bench_rts.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02k_bench_rts.json').read().strip().split('\n')[-1])
print('value', d['value'], 'ms_per_step', d['ms_per_step'], {k: (round(v,1) if isinstance(v,float) else v) for k,v in d['phases'].items()}, 'frac', d['roofline']['frac'])
PY
timeout 600 python bench.py --workload live_100k --steps 200 --warmup 5 --no-extras > gpurun_out/r02k_bench_live100k.json 2> gpurun_out/r02k_bench_live100k.err; tail -2 gpurun_out/r02k_bench_live100k.err
timeout 600 python bench.py --workload kinematic_16m --steps 100 --warmup 5 --no-extras > gpurun_out/r02k_bench_kin16m.json 2> gpurun_out/r02k_bench_kin16m.err; tail -2 gpurun_out/r02k_bench_kin16m.err
timeout 600 python bench.py --workload kinematic_1m --steps 200 --warmup 5 --no-extras > gpurun_out/r02k_bench_kin1m.json 2> gpurun_out/r02k_bench_kin1m.err; tail -2 gpurun_out/r02k_bench_kin1m.err
python - <<'PY'
import json
for f in ('live100k','kin16m','kin1m'):
    try:
        d=json.loads(open(f'gpurun_out/r02k_bench_{f}.json').read().strip().split('\n')[-1])
        print(f, 'value', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'cpu', d.get('cpu_baseline',{}).get('value'), 'sust', d.get('sustained',{}).get('value'))
    except Exception as e: print(f, 'ERR', e)
PY
