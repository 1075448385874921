#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02u_gpu_tests.txt 2>&1; tail -4 gpurun_out/r02u_gpu_tests.txt
timeout 900 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r02u_bench_default.json 2> gpurun_out/r02u_bench_default.err; tail -2 gpurun_out/r02u_bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02u_bench_default.json').read().strip().split('\n')[-1])
print('value', d['value'], 'per_kind', d['per_kind_ms'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'sust', d['sustained']['value'], 'graph', d['cuda_graph_replay'].get('value'))
PY
