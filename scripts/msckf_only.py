#!/usr/bin/env python3
"""MSCKF (CTA-per-filter) fused feature step only: timing harness for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rednose_b200.batched import BatchedEKF
from rednose_b200.filters import ensure_generated
from rednose_b200.filters.msckf import MsckfKalman
from tests.util import msckf_batch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
d = ensure_generated(MsckfKalman)
x, P, Q, point = msckf_batch(64, seed=2)
reps = (B + 63) // 64
x, P, point = np.tile(x, (reps, 1))[:B], np.tile(P, (reps, 1, 1))[:B], np.tile(point, (reps, 1))[:B]
e = BatchedEKF(d, "msckf", Q, x, P, quaternion_idxs=[3] + [26 + 7 * c for c in range(10)])
rng = np.random.default_rng(0)
z = torch.as_tensor(rng.normal(size=(B, 20)) * 1e-3).cuda()
R = torch.as_tensor(np.eye(20) * 1e-6).cuda()
ea = torch.as_tensor(point).cuda()
for it in range(4):
  t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  t0.record(); e.step(17, 0.01, z.clone(), R, ea=ea); t1.record(); torch.cuda.synchronize()
  print(f"msckf fused feature step B={B}: {t0.elapsed_time(t1):.3f} ms")
