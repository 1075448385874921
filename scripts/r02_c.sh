#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_msckf_gpu.py -x -q > gpurun_out/r02c_msckf_test.txt 2>&1
tail -15 gpurun_out/r02c_msckf_test.txt
python scripts/msckf_only.py 10000 > gpurun_out/r02c_msckf_time.log 2>&1
cat gpurun_out/r02c_msckf_time.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ekf_step_cta -s 1 -c 1 -o gpurun_out/r02c_cta python scripts/msckf_only.py 4096 > gpurun_out/r02c_cta.log 2>&1
python -m pytest tests/test_parity_gpu.py -x -q -k "msckf" > gpurun_out/r02c_parity_msckf.txt 2>&1
tail -5 gpurun_out/r02c_parity_msckf.txt
