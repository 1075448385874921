#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-extras > gpurun_out/r02_n${N}_default.json 2> gpurun_out/r02_n${N}_default.err
tail -3 gpurun_out/r02_n${N}_default.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02_n${N}_default.json').read().strip().split('\n')[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'd2h GB/s/gpu', d['e2e']['d2h_GBps_per_gpu'], 'pose', d.get('e2e_pose_columns_only',{}).get('value'), 'numa', d.get('numa_node'), 'gather', d.get('final_gather_ms'), d.get('final_gather_P_ms'))
print('cpu', d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('threads'), d.get('cpu_baseline',{}).get('cgroup_cpu_quota'))
PY
