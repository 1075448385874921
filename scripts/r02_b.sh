#!/bin/bash
mkdir -p gpurun_out
python scripts/cpu_arm_sweep.py > gpurun_out/r02b_cpu_sweep.txt 2>&1
python -m pytest tests/test_features_gpu.py -x -q > gpurun_out/r02b_features_test.txt 2>&1
tail -5 gpurun_out/r02b_features_test.txt
cat gpurun_out/r02b_cpu_sweep.txt
