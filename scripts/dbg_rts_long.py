#!/usr/bin/env python3
"""debug: where does a 10 000-step live history go non-finite (seed of rank 5 in the 8-GPU run)?"""
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
B, T = 125000, int(sys.argv[1]) if len(sys.argv) > 1 else 10000
x0, P0, Q, pools, (dim, edim), quat = make_problem("live", B, seed=1239, lib_dir=d)
eng = BatchedEKF(d, "live", Q, x0, P0, device=dev, quaternion_idxs=quat)
zp = {k: torch.as_tensor(z).to(dev) for k, (z, _) in pools.items()}
Rk = {k: torch.as_tensor(R[0]).to(dev) for k, (_, R) in pools.items()}
sched = kind_schedule("live", T)
eng.filter_time = 0.0
bad_first = None
for k in range(T):
  eng.predict_and_update_batch(0.01 * (k + 1), sched[k], zp[sched[k]][k % 2].clone(), Rk[sched[k]])
  if (k + 1) % 500 == 0:
    fx = torch.isfinite(eng.x).all(dim=1) & torch.isfinite(eng.P).flatten(1).all(dim=1)
    dmin = float(torch.diagonal(eng.P, dim1=1, dim2=2).min())
    pmax = float(eng.P[fx].abs().max()) if fx.any() else float("nan")
    print(f"forward step {k+1}: non-finite filters {int((~fx).sum())}, min diag P {dmin:.3e}, max |P| {pmax:.3e}, max |x-x0| pos {float((eng.x[fx][:, :3] - torch.as_tensor(x0).to(dev)[fx][:, :3]).abs().max()):.3e}")
    if bad_first is None and (~fx).any():
      bad_first = (~fx).nonzero()[:5, 0].tolist(); print("first bad filters", bad_first)
torch.cuda.synchronize()
# smoother on a 4096-filter slice
n = 4096
x0d, P0d = torch.as_tensor(x0[:n]).to(dev), torch.as_tensor(P0).to(dev).expand(n, -1, -1)
res = {}
def obs_fn(k, lo, hi): return 0.01 * (k + 1), sched[k], zp[sched[k]][k % 2][lo:hi].clone(), Rk[sched[k]]
def sink(lo, hi, k0, xs, Ps):
  fin = torch.isfinite(xs).flatten(1).all(dim=1) & torch.isfinite(Ps).flatten(1).all(dim=1)
  if not bool(fin.all()):
    res.setdefault("bad_segments", []).append((k0, int((~fin).sum())))
cs = CheckpointedSmoother(d, "live", Q, dim, edim, quaternion_idxs=quat, device=dev, segment=50)
cs.run(x0d, P0d, T, obs_fn, sink, norm_quats=True)
print("smoother on 4096 filters: bad segments (k0, non-finite steps):", res.get("bad_segments", [])[-5:], "count", len(res.get("bad_segments", [])))
