#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_msckf_gpu.py tests/test_features_gpu.py -x -q > gpurun_out/r02g_tests.txt 2>&1; tail -6 gpurun_out/r02g_tests.txt
timeout 600 python bench.py --workload msckf_10k --steps 50 --warmup 5 > gpurun_out/r02g_bench_msckf.json 2> gpurun_out/r02g_bench_msckf.err; tail -3 gpurun_out/r02g_bench_msckf.err; cut -c1-1600 gpurun_out/r02g_bench_msckf.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ekf_step_cta -s 3 -c 1 -o gpurun_out/r02g_cta python bench.py --workload msckf_10k --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02g_cta.log 2>&1
timeout 900 python bench.py --workload live_rts --rts-steps 1000 --steps 2 > gpurun_out/r02g_bench_rts.json 2> gpurun_out/r02g_bench_rts.err; tail -3 gpurun_out/r02g_bench_rts.err; cut -c1-2800 gpurun_out/r02g_bench_rts.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02g_ref.json 2> gpurun_out/r02g_ref.err; cut -c1-1500 gpurun_out/r02g_ref.json
