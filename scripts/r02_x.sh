#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02x_gpu_tests.txt 2>&1; tail -3 gpurun_out/r02x_gpu_tests.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02x_bench_default.json 2> gpurun_out/r02x_bench_default.err; tail -2 gpurun_out/r02x_bench_default.err
timeout 900 python bench.py --workload live_rts --rts-steps 10000 --rts-segment 100 --steps 1 --no-cpu-baseline > gpurun_out/r02x_bench_rts10k.json 2> gpurun_out/r02x_bench_rts10k.err; tail -2 gpurun_out/r02x_bench_rts10k.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02x_bench_default.json').read().strip().split('\n')[-1])
print('value', d['value'], d['per_kind_ms'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'sust', d['sustained']['value'], 'graph', d['cuda_graph_replay'].get('value'))
print('extras', {k: v.get('frac_of_peak', v.get('steps_per_s')) for k,v in d['extras'].items() if isinstance(v, dict)})
r=json.loads(open('gpurun_out/r02x_bench_rts10k.json').read().strip().split('\n')[-1])
print('rts10k value', r['value'], 'ms', r['ms_per_step'], {k: round(v,1) for k,v in r['phases'].items()}, r['config']['tiles_per_pass'], r['config']['tile_filters'])
PY
