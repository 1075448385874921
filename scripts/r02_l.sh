#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02l_gpu_tests.txt 2>&1; tail -4 gpurun_out/r02l_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02l_bench_default.json 2> gpurun_out/r02l_bench_default.err; tail -3 gpurun_out/r02l_bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02l_bench_default.json').read().strip().split('\n')[-1])
print('value', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'single', d.get('single_filter_dropin',{}).get('us_per_predict_and_update_batch'), 'hostabi', d.get('e2e_stateless_host_c_abi',{}).get('value'))
print('extras', {k: v.get('frac_of_peak', v.get('steps_per_s', v)) for k,v in d['extras'].items()})
PY
