#!/bin/bash
# time the RTS backward kernel for every rednose_b200/generated* variant folder
for d in ${@:-rednose_b200/generated*}; do
  echo -n "== $d  "
  REDNOSE_B200_GENERATED_DIR=$PWD/$d python scripts/rts_bench.py 65536 16 2>&1 | tail -1
done
