#!/bin/bash
mkdir -p gpurun_out
python scripts/dbg_msckf_bench.py > gpurun_out/r02e_dbg_msckf.txt 2>&1; head -40 gpurun_out/r02e_dbg_msckf.txt
python -m pytest tests/test_parity_gpu.py -x -q -k "checkpointed or two_passes or ill_conditioned or kinematic_sampled or rts or golden" > gpurun_out/r02e_tests.txt 2>&1; tail -15 gpurun_out/r02e_tests.txt
for d in rednose_b200/generated rednose_b200/generated_g_rts7; do echo -n "== $d  "; REDNOSE_B200_GENERATED_DIR=$PWD/$d python scripts/rts_bench.py 65536 16 2>&1 | tail -1; done
timeout 900 python bench.py --workload live_rts --rts-steps 200 --steps 1 --no-cpu-baseline > gpurun_out/r02e_bench_rts.json 2> gpurun_out/r02e_bench_rts.err; tail -3 gpurun_out/r02e_bench_rts.err; cut -c1-2500 gpurun_out/r02e_bench_rts.json
