#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02j_gpu_tests.txt 2>&1; tail -4 gpurun_out/r02j_gpu_tests.txt
python scripts/msckf_only.py 10000 > gpurun_out/r02j_msckf_time.log 2>&1; tail -2 gpurun_out/r02j_msckf_time.log
{ for tool in memcheck racecheck synccheck; do echo "== compute-sanitizer --tool $tool python scripts/sanitize_smoke.py"; timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_smoke.py 2>&1 | grep -vE "^=========\s*$" | tail -8; done; } > gpurun_out/r02j_sanitizer.txt 2>&1
cat gpurun_out/r02j_sanitizer.txt | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02j_bench_default.json 2> gpurun_out/r02j_bench_default.err; tail -3 gpurun_out/r02j_bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02j_bench_default.json').read().strip().split('\n')[-1])
print('value', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'single', d.get('single_filter_dropin'))
print('cpu', {k:v for k,v in d['cpu_baseline'].items() if k not in ('what','sample')})
print('extras msckf', d['extras'].get('msckf_10k_feature_step'))
PY
