"""Shared helpers for the parity tests: oracle access (test infrastructure) and synthetic inputs."""

import numpy as np

from oracle.handle import Oracle  # noqa: F401

LIVE_KINDS = {3: 1, 4: 3, 9: 3, 10: 3, 12: 3, 13: 3, 14: 3, 19: 3}  # kind -> ZDIM (gen/live.cpp:1780-1802)
LIVE_R = {3: [0.2**2], 4: [0.025**2] * 3, 9: [0.00025**2] * 3, 10: [0.5**2] * 3, 12: [5.0**2] * 3,
          13: [0.1**2] * 3, 14: [0.05**2] * 3, 19: [0.05**2] * 3}


def rel_err(a, b):
  """Per-array max-norm relative error (hard part 4)."""
  a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
  denom = np.max(np.abs(b))
  return float(np.max(np.abs(a - b)) / (denom if denom > 0 else 1.0))


def live_batch(B, seed=0, well_conditioned=True):
  """Random but physically plausible live_kf states / covariances."""
  from rednose_b200.filters.live import LiveKalman
  rng = np.random.default_rng(seed)
  x = np.tile(LiveKalman.initial_x, (B, 1))
  x[:, 0:3] += rng.normal(0, 100.0, (B, 3))
  q = rng.normal(size=(B, 4))
  x[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
  x[:, 7:10] = rng.normal(0, 5.0, (B, 3))
  x[:, 10:13] = rng.normal(0, 0.1, (B, 3))
  x[:, 13:16] = rng.normal(0, 0.01, (B, 3))
  x[:, 16] = 1.0 + rng.normal(0, 0.01, B)
  x[:, 17:20] = rng.normal(0, 1.0, (B, 3))
  x[:, 20:23] = rng.normal(0, 0.01, (B, 3))
  if well_conditioned:
    s = np.sqrt(np.array([25.0] * 3 + [0.05**2] * 3 + [1.0] * 3 + [0.1**2] * 3 + [0.01**2] * 3 + [0.01**2] + [0.5**2] * 3 + [0.01**2] * 3))
  else:
    s = np.sqrt(LiveKalman.initial_P_diag)
  L = np.eye(22)[None] + 0.2 * np.tril(rng.normal(size=(B, 22, 22)), -1)
  L = s[None, :, None] * L
  P = L @ np.transpose(L, (0, 2, 1))
  P = 0.5 * (P + np.transpose(P, (0, 2, 1)))
  return x, P, LiveKalman.Q.copy()


def live_obs(oracle, kind, x, seed=1, noise_scale=1.0):
  """z = h_kind(x) + N(0, R) with the default R of the kind; uses the oracle's leaf h (reference-generated C)."""
  rng = np.random.default_rng(seed + kind)
  m = LIVE_KINDS[kind]
  B = x.shape[0]
  R = np.tile(np.diag(LIVE_R[kind]), (B, 1, 1))
  z = np.zeros((B, m))
  dummy = np.zeros(1)
  for b in range(B):
    oracle.leaf(f"h_{kind}", np.ascontiguousarray(x[b]), dummy, z[b])
  z += noise_scale * rng.normal(size=(B, m)) * np.sqrt(np.array(LIVE_R[kind]))[None, :]
  return z, R


def kinematic_batch(B, seed=0):
  from rednose_b200.filters.kinematic import KinematicKalman
  rng = np.random.default_rng(seed)
  x = np.tile(KinematicKalman.initial_x, (B, 1)) + rng.normal(size=(B, 2))
  L = np.eye(2)[None] + 0.3 * np.tril(rng.normal(size=(B, 2, 2)), -1)
  P = L @ np.transpose(L, (0, 2, 1))
  z = x[:, :1] + rng.normal(0, 0.1, (B, 1))
  R = np.tile(np.array([[0.1**2]]), (B, 1, 1))
  return x, P, KinematicKalman.Q.copy(), z, R


def msckf_batch(B, seed=0, outlier_frac=0.0):
  """Synthetic MSCKF states: live main state, 10 clones = the main pose displaced along a short track,
  one 3-D point 10-50 m ahead seen from every clone."""
  from rednose_b200.filters.msckf import DIM, DIM_AUGMENT, EDIM, N_CLONES, MsckfKalman
  from rednose_b200.filters.live import DIM_STATE
  from rednose_b200.geometry import quat2rot
  rng = np.random.default_rng(seed)
  xm, _, _ = live_batch(B, seed=seed)
  x = np.zeros((B, DIM))
  x[:, :DIM_STATE] = xm
  Rm = quat2rot(xm[:, 3:7])                       # device -> ecef
  fwd = Rm[:, :, 0]
  for c in range(N_CLONES):
    o = DIM_STATE + c * DIM_AUGMENT
    x[:, o:o + 3] = xm[:, 0:3] - fwd * 0.5 * (N_CLONES - c) + rng.normal(0, 0.02, (B, 3))
    q = xm[:, 3:7] + rng.normal(0, 0.002, (B, 4))
    x[:, o + 3:o + 7] = q / np.linalg.norm(q, axis=1, keepdims=True)
  s = np.sqrt(np.concatenate([[25.0] * 3 + [0.05**2] * 3 + [1.0] * 3 + [0.1**2] * 3 + [0.01**2] * 3 + [0.01**2] + [0.5**2] * 3 + [0.01**2] * 3]
                             + [[1.0] * 3 + [0.02**2] * 3] * N_CLONES))
  L = np.eye(EDIM)[None] + 0.05 * np.tril(rng.normal(size=(B, EDIM, EDIM)), -1)
  L = s[None, :, None] * L
  P = L @ np.transpose(L, (0, 2, 1))
  P = 0.5 * (P + np.transpose(P, (0, 2, 1)))
  local = np.stack([rng.uniform(10, 50, B), rng.uniform(-5, 5, B), rng.uniform(-3, 3, B)], 1)
  point = xm[:, 0:3] + np.einsum('bij,bj->bi', Rm, local)
  return x, P, MsckfKalman.Q.copy(), point


def msckf_feature_obs(oracle, x, point, seed=1, sigma=1e-3, outlier_frac=0.0):
  rng = np.random.default_rng(seed)
  B = x.shape[0]
  z = np.zeros((B, 20))
  for b in range(B):
    oracle.leaf("h_17", np.ascontiguousarray(x[b]), np.ascontiguousarray(point[b]), z[b])
  noise = rng.normal(0, sigma, (B, 20))
  out = rng.random(B) < outlier_frac
  noise[out] *= 50.0
  R = np.tile(np.eye(20) * sigma**2, (B, 1, 1))
  return z + noise, R, out
