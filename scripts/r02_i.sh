#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_features_gpu.py -x -q > gpurun_out/r02i_tests.txt 2>&1; tail -4 gpurun_out/r02i_tests.txt
timeout 600 python bench.py --workload msckf_10k --steps 50 --warmup 5 > gpurun_out/r02i_bench_msckf.json 2> gpurun_out/r02i_bench_msckf.err; tail -3 gpurun_out/r02i_bench_msckf.err; cut -c1-300 gpurun_out/r02i_bench_msckf.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/r02i_msckf_launches.csv python bench.py --workload msckf_10k --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r02i_launches.log 2>&1
grep -E "compute_pos|ekf_leaf|ekf_step_cta" gpurun_out/r02i_msckf_launches.csv | tail -6 | cut -c1-260
