#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02d_gpu_tests.txt 2>&1
tail -25 gpurun_out/r02d_gpu_tests.txt
python scripts/msckf_only.py 10000 > gpurun_out/r02d_msckf_time.log 2>&1; cat gpurun_out/r02d_msckf_time.log
timeout 600 python bench.py --workload msckf_10k --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02d_bench_msckf.json 2> gpurun_out/r02d_bench_msckf.err; tail -3 gpurun_out/r02d_bench_msckf.err; cut -c1-1500 gpurun_out/r02d_bench_msckf.json
timeout 900 python bench.py --workload live_rts --rts-steps 200 --steps 1 --no-cpu-baseline > gpurun_out/r02d_bench_rts.json 2> gpurun_out/r02d_bench_rts.err; tail -3 gpurun_out/r02d_bench_rts.err; cut -c1-2500 gpurun_out/r02d_bench_rts.json
