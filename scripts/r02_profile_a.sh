#!/bin/bash
# round 2, first GPU call: source-level ncu captures of the two secondary kernels as they stood at the end of round 1,
# plus the new in-place CPU reference arm on the box's host cores
mkdir -p gpurun_out
nproc > gpurun_out/r02a_nproc.txt; lscpu | head -25 >> gpurun_out/r02a_nproc.txt; numactl -H >> gpurun_out/r02a_nproc.txt 2>&1
nvidia-smi topo -m >> gpurun_out/r02a_nproc.txt 2>&1
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02a_ref.json 2> gpurun_out/r02a_ref.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ekf_step_cta -s 1 -c 1 -o gpurun_out/r02a_cta python scripts/msckf_only.py 4096 > gpurun_out/r02a_cta.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ekf_rts -s 0 -c 1 -o gpurun_out/r02a_rts python scripts/rts_bench.py 16384 6 > gpurun_out/r02a_rts.log 2>&1
python scripts/msckf_only.py 10000 > gpurun_out/r02a_msckf_time.log 2>&1
python scripts/rts_bench.py 65536 16 > gpurun_out/r02a_rts_time.log 2>&1
ls -la gpurun_out
