#!/bin/bash
# bench every rednose_b200/generated* variant folder (built with different RNB_* knobs) back to back
for d in ${@:-rednose_b200/generated*}; do
  echo -n "== $d  "
  REDNOSE_B200_GENERATED_DIR=$PWD/$d python bench.py --steps 60 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 3 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.3e steps/s'%d['value'], d['per_kind_ms'], 'frac %.3f'%d['roofline']['frac'], 'e2e %.3e'%d['e2e']['value'], d['clocks'])"
done
