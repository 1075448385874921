#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02f_gpu_tests.txt 2>&1
tail -8 gpurun_out/r02f_gpu_tests.txt
timeout 600 python bench.py --workload msckf_10k --steps 50 --warmup 5 > gpurun_out/r02f_bench_msckf.json 2> gpurun_out/r02f_bench_msckf.err; tail -3 gpurun_out/r02f_bench_msckf.err; cut -c1-1800 gpurun_out/r02f_bench_msckf.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_bench_default.json 2> gpurun_out/r02f_bench_default.err; tail -3 gpurun_out/r02f_bench_default.err; cut -c1-3000 gpurun_out/r02f_bench_default.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --sustain 0 > gpurun_out/r02f_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ekf_step_pair -s 6 -c 2 -o gpurun_out/r02f_pair python bench.py --steps 4 --warmup 3 --no-extras --no-cpu-baseline --sustain 0 --e2e-steps 3 > gpurun_out/r02f_pair.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ekf_rts -s 0 -c 1 -o gpurun_out/r02f_rts python scripts/rts_bench.py 16384 12 > gpurun_out/r02f_rts.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:compute_pos -s 1 -c 1 -o gpurun_out/r02f_cpos python bench.py --workload msckf_10k --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_cpos.log 2>&1
ls -la gpurun_out | tail -12
