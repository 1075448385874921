#!/usr/bin/env python3
"""Tiny end-to-end exercise of every kernel family, meant to run under compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rednose_b200.batched import BatchedEKF
from rednose_b200.filters import ensure_generated
from rednose_b200.filters.kinematic import KinematicKalman
from rednose_b200.filters.live import LiveKalman
from rednose_b200.filters.msckf import MsckfKalman
from rednose_b200.scheduler import RaggedScheduler
from tests.util import kinematic_batch, live_batch, msckf_batch

d = ensure_generated(LiveKalman); ensure_generated(KinematicKalman); ensure_generated(MsckfKalman)
x, P, Q, z, R = kinematic_batch(257)
e = BatchedEKF(d, "kinematic", Q, x, P); e.step(1, 0.01, z, R)
B = 45
x, P, Q = live_batch(B, seed=1)
e = BatchedEKF(d, "live", Q, x, P, quaternion_idxs=[3])
hist = e.new_history(5)
rng = np.random.default_rng(0)
for k, kind in enumerate([12, 4, 10, 3, 13]):
  m = 1 if kind == 3 else 3
  e.step_recorded(hist, kind, 0.01 * (k + 1), rng.normal(size=(B, m)) * 0.01 + (x[:, :3] if kind == 12 else 0.0)[..., :m] if kind == 12 else rng.normal(size=(B, m)) * 0.01, np.eye(m) * 0.1)
e.rts_smooth(hist, norm_quats=True)
e.step(4, 0.01, rng.normal(size=(B, 2, 3)) * 0.01, np.tile(np.eye(3) * 0.1, (B, 2, 1, 1)))   # two observations per predict
sch = RaggedScheduler(e)
sch.tick(np.arange(0, B, 2), 0.2, np.full((B + 1) // 2, 4), {4: rng.normal(size=((B + 1) // 2, 3)) * 0.01}, {4: np.eye(3) * 0.1})
xm, Pm, Qm, point = msckf_batch(5, seed=2)
em = BatchedEKF(d, "msckf", Qm, xm, Pm, quaternion_idxs=[3])
em.step(12, 0.01, xm[:, :3] + rng.normal(size=(5, 3)), np.eye(3) * 25.0)
em.update(17, rng.normal(size=(5, 20)) * 0.01, np.eye(20) * 1e-4, ea=point)
em.augment()
torch.cuda.synchronize()
assert torch.isfinite(e.x).all() and torch.isfinite(em.P).all()
print("sanitize smoke done")
