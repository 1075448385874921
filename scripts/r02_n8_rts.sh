#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}; T=${2:-1000}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload live_rts --rts-steps $T --steps ${3:-2} > gpurun_out/r02_n${N}_live_rts.json 2> gpurun_out/r02_n${N}_live_rts.err
tail -3 gpurun_out/r02_n${N}_live_rts.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02_n${N}_live_rts.json').read().strip().split('\n')[-1])
print('value', d['value'], 'ms_per_step', d['ms_per_step'], 'phases', {k: (round(v,1) if isinstance(v,float) else v) for k,v in d['phases'].items()}, 'frac', d['roofline']['frac'], 'numa', d.get('numa_node'))
print('cpu', d.get('cpu_baseline',{}).get('value'))
PY
cat /sys/fs/cgroup/cpu.max
