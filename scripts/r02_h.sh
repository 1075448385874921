#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_msckf_gpu.py -x -q > gpurun_out/r02h_tests.txt 2>&1; tail -4 gpurun_out/r02h_tests.txt
python -m pytest tests/test_parity_gpu.py -x -q -k msckf >> gpurun_out/r02h_tests.txt 2>&1; tail -2 gpurun_out/r02h_tests.txt
python scripts/msckf_only.py 10000 > gpurun_out/r02h_msckf_time.log 2>&1; cat gpurun_out/r02h_msckf_time.log
timeout 600 python bench.py --workload msckf_10k --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r02h_bench_msckf.json 2> gpurun_out/r02h_bench_msckf.err; tail -3 gpurun_out/r02h_bench_msckf.err; cut -c1-300 gpurun_out/r02h_bench_msckf.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ekf_step_cta -s 3 -c 1 -o gpurun_out/r02h_cta python bench.py --workload msckf_10k --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02h_cta.log 2>&1
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02h_ref.json 2> gpurun_out/r02h_ref.err; python -c "
import json; d=json.load(open('gpurun_out/r02h_ref.json')); print(d['value'], d['cpu_baseline']['threads'], d['cpu_baseline']['threads_pinned'], d['cpu_baseline']['thread_sweep_steps_per_s'])"
