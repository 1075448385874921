#!/bin/bash
mkdir -p gpurun_out
for d in generated_g_late generated; do
  echo "== $d"
  REDNOSE_B200_GENERATED_DIR=$PWD/rednose_b200/$d python scripts/dbg_rts_race.py 250 fwd 2>&1 | tail -1 | cut -c1-200
  REDNOSE_B200_GENERATED_DIR=$PWD/rednose_b200/$d python bench.py --steps 40 --warmup 5 --no-extras --no-cpu-baseline --sustain 0 --e2e-steps 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('value', d['value'], d['per_kind_ms'])"
done
