#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_parity_gpu.py tests/test_features_gpu.py -x -q -k "graph or fallback" > gpurun_out/r02m_tests.txt 2>&1; tail -4 gpurun_out/r02m_tests.txt
for w in kinematic_1m live_100k live_1m; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r02m_$w.json 2> gpurun_out/r02m_$w.err; tail -2 gpurun_out/r02m_$w.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02m_$w.json').read().strip().split('\n')[-1])
print('$w', 'value', d['value'], 'graph', d.get('cuda_graph_replay'), 'sust', d.get('sustained',{}).get('value'))
PY
done
