#!/usr/bin/env python3
"""debug: where does the msckf_10k workload first produce a non-finite value?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
class A: pass
# re-create run_msckf's setup through a tiny copy of its code path
from rednose_b200.batched import BatchedEKF
from rednose_b200.features import FeatureFrontend, to_c_matrix
from rednose_b200.filters import ensure_generated
from rednose_b200.filters.msckf import DIM, EDIM, MsckfKalman
dev = torch.device("cuda", 0)
d = ensure_generated(MsckfKalman); fe = FeatureFrontend(10)
B = 10000
g = torch.Generator(device=dev); g.manual_seed(77)
f64 = dict(dtype=torch.float64, device=dev)
def quat2rot_t(q):
  w, x, y, z = q.unbind(-1)
  return torch.stack([w*w+x*x-y*y-z*z, 2*(x*y-w*z), 2*(w*y+x*z), 2*(x*y+w*z), w*w-x*x+y*y-z*z, 2*(y*z-w*x), 2*(x*z-w*y), 2*(w*x+y*z), w*w-x*x-y*y+z*z], -1).reshape(q.shape[:-1] + (3, 3))
dt, speed = 0.05, 10.0
x0 = torch.as_tensor(MsckfKalman.initial_x).to(dev).repeat(B, 1)
q = torch.randn(B, 4, generator=g, **f64); q = q / q.norm(dim=1, keepdim=True)
Rm = quat2rot_t(q)
x0[:, 0:3] += torch.randn(B, 3, generator=g, **f64) * 100.0
x0[:, 3:7] = q
x0[:, 7:10] = Rm[:, :, 0] * speed
for c in range(10):
  o = 23 + 7 * c
  x0[:, o:o + 3] = x0[:, 0:3] - Rm[:, :, 0] * (speed * dt) * (10 - c)
  x0[:, o + 3:o + 7] = q
pd = np.concatenate([[25.0] * 3 + [0.05**2] * 3 + [1.0] * 3 + [0.1**2] * 3 + [0.01**2] * 3 + [0.01**2] + [0.5**2] * 3 + [0.01**2] * 3] + [[1.0] * 3 + [0.02**2] * 3] * 10)
eng = BatchedEKF(d, "msckf", MsckfKalman.Q, x0, np.diag(pd), device=dev, quaternion_idxs=[3] + [26 + 7 * c for c in range(10)])
sigma = 1e-3
Rk = torch.eye(20, **f64) * sigma**2
to_c = torch.as_tensor(to_c_matrix().reshape(9)).to(dev)
for step in range(30):
  clones = eng.x[:, 23:].reshape(B, 10, 7)
  Rl = quat2rot_t(clones[:, 9, 3:7])
  local = torch.stack([torch.rand(B, generator=g, **f64) * 35 + 15, torch.rand(B, generator=g, **f64) * 10 - 5, torch.rand(B, generator=g, **f64) * 6 - 3], 1)
  point = clones[:, 9, 0:3] + torch.einsum('bij,bj->bi', Rl, local)
  pc = torch.einsum('bcji,bcj->bci', quat2rot_t(clones[:, :, 3:7]), point[:, None, :] - clones[:, :, 0:3])
  z = torch.stack([pc[:, :, 1] / pc[:, :, 0], pc[:, :, 2] / pc[:, :, 0]], -1).reshape(B, 20)
  noise = torch.randn(B, 20, generator=g, **f64) * sigma
  out = torch.rand(B, generator=g, device=dev) < 0.05
  noise[out] *= 50.0
  z = (z + noise).contiguous()
  poses = eng.x[:, 23:].contiguous()
  pos, param, iters = fe.compute_pos_batch(to_c, poses, z)
  err = (pos - point).norm(dim=1)
  bad = ~torch.isfinite(pos).all(dim=1)
  print(f"step {step}: pc.x min {float(pc[:,:,0].min()):.2f}  tri err median {float(err[~bad].median()):.3f} max {float(err[~bad].max()):.1f}  nonfinite pos {int(bad.sum())} (outliers among them {int((bad & out).sum())})  iters max {int(iters.max())} mean {float(iters.double().mean()):.2f} n30 {int((iters==30).sum())}")
  eng.step(17, dt, z, Rk, ea=pos)
  fx, fP = torch.isfinite(eng.x).all(dim=1), torch.isfinite(eng.P).flatten(1).all(dim=1)
  print(f"        after step: nonfinite x {int((~fx).sum())} P {int((~fP).sum())}  (of which had bad pos {int((~fx & bad).sum())})  min diag P {float(torch.diagonal(eng.P, dim1=1, dim2=2).min()):.3e}")
  eng.augment()
  if (~fx).any() and step > 3:
    break
