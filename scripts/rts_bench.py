#!/usr/bin/env python3
"""Time (and optionally profile) the batched RTS backward kernel on a recorded live history."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
from bench import make_problem, kind_schedule
from rednose_b200.batched import BatchedEKF
from rednose_b200.filters import ensure_generated
from rednose_b200.filters.live import LiveKalman

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = torch.device("cuda", 0)
d = ensure_generated(LiveKalman)
x0, P0, Q, pools, _, quat = make_problem("live", B, 11, d)
e = BatchedEKF(d, "live", Q, x0, P0, device=dev, quaternion_idxs=quat)
dpool = {k: (torch.as_tensor(z[0]).to(dev), torch.as_tensor(R[0]).to(dev)) for k, (z, R) in pools.items()}
hist = e.new_history(T)
sched = kind_schedule("live", T)
for k in range(T):
  zk, Rk = dpool[sched[k]]
  e.step_recorded(hist, sched[k], 0.01 * (k + 1), zk.clone(), Rk)
torch.cuda.synchronize()
for rep in range(2):
  t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  t0.record()
  xs, Ps = e.rts_smooth(hist, norm_quats=True, in_place=False)
  t1.record(); torch.cuda.synchronize()
  ms = t0.elapsed_time(t1) / (T - 1)
  print(f"RTS B={B} T={T}: {ms:.3f} ms per backward step, {B / ms * 1e3:.3e} steps/s, {12176 * B / ms * 1e-6:.1f} GB/s algorithmic, finite={bool(torch.isfinite(Ps).all())}")
