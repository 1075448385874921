#!/usr/bin/env python3
"""debug: is the forward-with-history kernel or the RTS kernel non-deterministic (a race) at the 1e-9 per filter-step level?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import make_problem, kind_schedule
from rednose_b200.batched import BatchedEKF
from rednose_b200.filters import ensure_generated
from rednose_b200.filters.live import LiveKalman
dev = torch.device("cuda", 0)
d = ensure_generated(LiveKalman)
B, T, REPS = 62500, 101, int(sys.argv[1]) if len(sys.argv) > 1 else 200
x0, P0, Q, pools, (dim, edim), quat = make_problem("live", B, seed=1239, lib_dir=d)
zp = {k: torch.as_tensor(z).to(dev) for k, (z, _) in pools.items()}
Rk = {k: torch.as_tensor(R[0]).to(dev) for k, (_, R) in pools.items()}
sched = kind_schedule("live", T)
e = BatchedEKF(d, "live", Q, x0, P0, device=dev, quaternion_idxs=quat)
xs0, Ps0 = e.x.clone(), e.P.clone()
hist = e.new_history(T)
def forward():
  e.x.copy_(xs0); e.P.copy_(Ps0); e.filter_time = 0.0; hist.n = 0
  for k in range(T):
    e.step_recorded(hist, sched[k], 0.01 * (k + 1), zp[sched[k]][k % 2].clone(), Rk[sched[k]])
forward()
ref = [t.clone() for t in (hist.x_pred, hist.P_pred, hist.x_filt, hist.P_filt)]
bad_f = 0
locs = {}
for r in range(REPS):
  forward()
  if not torch.equal(ref[3], hist.P_filt):
    diff = (ref[3] != hist.P_filt).nonzero()
    bad_f += 1
    key = tuple(diff[0].tolist()[:2])
    locs[key] = locs.get(key, 0) + 1
print("forward-with-history: runs whose P_filt history differs from the first run:", bad_f, "of", REPS, "; first-difference (step, filter) -> count:", dict(list(locs.items())[:8]))
if len(sys.argv) > 2 and sys.argv[2] == "fwd":
  sys.exit(0)
for t, a in zip((hist.x_pred, hist.P_pred, hist.x_filt, hist.P_filt), ref):
  t.copy_(a)
tx, tP = ref[2][T - 1].clone(), ref[3][T - 1].clone()      # any finite terminal will do
xs_ref = torch.empty_like(hist.x_filt); Ps_ref = torch.empty_like(hist.P_filt)
e.rts_smooth(hist, norm_quats=True, out=(xs_ref, Ps_ref), terminal=(tx, tP), k0=100)
xs = torch.empty_like(hist.x_filt); Ps = torch.empty_like(hist.P_filt)
bad_r = 0
for r in range(REPS):
  e.rts_smooth(hist, norm_quats=True, out=(xs, Ps), terminal=(tx, tP), k0=100)
  okx, okP = torch.equal(xs[:T - 1], xs_ref[:T - 1]), torch.equal(Ps[:T - 1], Ps_ref[:T - 1])
  if not (okx and okP):
    bad_r += 1
    diff = (Ps[:T - 1] != Ps_ref[:T - 1]).nonzero()
    fin = bool(torch.isfinite(Ps[:T - 1]).all())
    print(f"rts rep {r}: differs (x ok {okx}, P ok {okP}, finite {fin}); {diff.shape[0]} elements; first {diff[0].tolist() if diff.numel() else None}; steps touched {sorted(set(diff[:, 0].tolist()))[:6]} filters {sorted(set(diff[:, 1].tolist()))[:6]}")
print("rts mismatching runs:", bad_r, "of", REPS, " reference finite:", bool(torch.isfinite(Ps_ref[:T - 1]).all()))
