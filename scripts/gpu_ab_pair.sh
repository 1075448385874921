set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for mode in pair single; do
  echo "== $mode"
  REDNOSE_B200_WARP_KERNEL=$mode python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extras --e2e-steps 5 2>&1 | tail -1 > gpurun_out/bench_r1h_$mode.json
  python -c "import sys,json; d=json.loads(open('gpurun_out/bench_r1h_$mode.json').read()); print('%.3e steps/s'%d['value'], d['per_kind_ms'], 'frac %.3f'%d['roofline']['frac'], 'e2e %.3e'%d['e2e']['value'], d['clocks'])"
done
ncu --set full --clock-control none --import-source on -k regex:ekf_step_pair -s 3 -c 2 -o gpurun_out/prof_r1h python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 0 > gpurun_out/b_ncu_r1h.log 2>&1
ncu -i gpurun_out/prof_r1h.ncu-rep --page raw --csv > gpurun_out/raw_r1h.csv 2>/dev/null
ls -la gpurun_out/ | tail -5
