#!/usr/bin/env python3
"""Kinematic (thread-per-filter) fused step on 16M filters: timing harness for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rednose_b200.batched import BatchedEKF
from rednose_b200.filters import ensure_generated
from rednose_b200.filters.kinematic import KinematicKalman as F
B = 1 << 24
d = ensure_generated(F)
rng = np.random.default_rng(0)
e = BatchedEKF(d, "kinematic", F.Q, np.tile(F.initial_x, (B, 1)) + rng.normal(size=(B, 2)), np.diag(F.initial_P_diag))
z = torch.as_tensor(rng.normal(0, 0.1, (B, 1))).cuda()
R = torch.as_tensor(np.tile(np.array([[0.01]]), (B, 1, 1))).cuda()
dt = torch.full((B,), 0.01, dtype=torch.float64, device="cuda")
for it in range(5):
  t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  zz = z.clone()
  t0.record(); e.step(1, dt, zz, R); t1.record(); torch.cuda.synchronize()
  print(f"kinematic step B={B}: {t0.elapsed_time(t1):.4f} ms -> {128 * B / t0.elapsed_time(t1) * 1e-6:.0f} GB/s")
