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
# round 2: fused augment in the CTA kernel, gated feature update, segment RTS, MSCKF front-end, single-filter entry points
zf = rng.normal(size=(5, 20)); zf[0] *= 100.0
em.step(17, 0.01, zf * 0.01, np.eye(20) * 1e-4, ea=point, augment=True)
em.step(12, 0.01, rng.normal(size=(5, 2, 3)) + xm[:, None, :3], np.tile(np.eye(3) * 25.0, (5, 2, 1, 1)), augment=True)
hist2 = e.new_history(4)
for k in range(4):
  e.step_recorded(hist2, 4, 1.0 + 0.01 * (k + 1), rng.normal(size=(B, 3)) * 0.01, np.eye(3) * 0.1)
tx, tP = e.x.clone(), e.P.clone()
e.rts_smooth(hist2, norm_quats=True, in_place=True, terminal=(tx, tP), k0=7)
from rednose_b200.features import FeatureFrontend, to_c_matrix
from tests.test_features_cpu import synth_frame, synth_tracks
fe = FeatureFrontend(10)
to_c, poses, img, _ = synth_tracks(70, seed=3, noise=1e-3)
fe.compute_pos_batch(to_c, torch.as_tensor(poses).cuda(), torch.as_tensor(img).cuda(), fallback_depth=30.0)
fe.compute_pos(to_c, poses[0], img[0])
tr, ft, em_ = zip(*[synth_frame(300, 120, 40 + s_, True, s_ == 1) for s_ in range(3)])
fe.merge_features_batch(torch.as_tensor(np.stack(tr)).cuda(), torch.as_tensor(np.stack(ft)).cuda(), torch.as_tensor(np.stack(em_)).cuda())
fe.sane_batch(torch.as_tensor(np.stack(tr)[0]).cuda())
from rednose_b200.ekf_sym import EKF_sym
kf = EKF_sym(d, "live", Q, x[0], P[0], 23, 22, quaternion_idxs=[3])
kf.predict_and_update_batch(0.01, 4, rng.normal(size=(1, 3)) * 0.01, np.eye(3)[None] * 0.1)
torch.cuda.synchronize()
assert torch.isfinite(e.x).all() and torch.isfinite(em.P).all()
print("sanitize smoke done")
