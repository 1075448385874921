#!/usr/bin/env python3
"""debug: find the filter / step where the 10 000-step smoother of rank 5's shard goes non-finite, and look at P_{k+1|k} there."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import make_problem, kind_schedule
from rednose_b200.batched import BatchedEKF
from rednose_b200.filters import ensure_generated
from rednose_b200.filters.live import LiveKalman
from rednose_b200.smoothing import CheckpointedSmoother
dev = torch.device("cuda", 0)
d = ensure_generated(LiveKalman)
B, T = 125000, 10000
x0, P0, Q, pools, (dim, edim), quat = make_problem("live", B, seed=1239, lib_dir=d)
zp = {k: torch.as_tensor(z).to(dev) for k, (z, _) in pools.items()}
Rk = {k: torch.as_tensor(R[0]).to(dev) for k, (_, R) in pools.items()}
sched = kind_schedule("live", T)
x0d, P0d = torch.as_tensor(x0).to(dev), torch.as_tensor(P0).to(dev).expand(B, -1, -1)
first = {}
def obs_fn(k, lo, hi): return 0.01 * (k + 1), sched[k], zp[sched[k]][k % 2][lo:hi].clone(), Rk[sched[k]]
def sink(lo, hi, k0, xs, Ps):
  fin = torch.isfinite(xs.sum(-1)) & torch.isfinite(Ps.flatten(2).sum(-1))   # [n, tile] (a NaN / inf anywhere poisons the sum)
  if not bool(fin.all()) and "k0" not in first:
    bad = (~fin).nonzero()
    kk = int(bad[:, 0].max())
    fl = bad[bad[:, 0] == kk][:, 1]
    first.update(k0=k0, step=k0 + kk, filters=(fl + lo).tolist()[:8], n=int((~fin).any(dim=0).sum()))
cs = CheckpointedSmoother(d, "live", Q, dim, edim, quaternion_idxs=quat, device=dev, segment=100, hbm_budget_bytes=70 << 30)
for rep in range(3):
  first.clear()
  cs.run(x0d, P0d, T, obs_fn, sink, norm_quats=True)
  torch.cuda.synchronize()
  print("rep", rep, "first failure (backward order):", dict(first))
first.clear()
