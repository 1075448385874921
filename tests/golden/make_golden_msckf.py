#!/usr/bin/env python3
"""Golden vectors for the MSCKF pieces, from the REFERENCE ITSELF (run here, where /root/reference is mounted).

Same recipe as make_golden.py: the reference's own Python driver `EKF_sym` (rednose/helpers/ekf_sym.py, imported
unmodified from /root/reference) on the reference-generated leaf C of oracle/_ref/libmsckf.so, switched to its numpy
maths: the block predict of `_predict_python` (ekf_sym.py:541-557), the left-null-space projection via SVD `null()`
and the Mahalanobis gate of `_update_python` (:575-603), `augment` (:365-391).  x and P do not depend on the basis the
projection picks, so they pin this repository's Householder projection (and the oracle's restated full-pivot-LU
kernel) to reference code; the projected innovation itself is basis dependent and is not stored.

  python tests/golden/make_golden_msckf.py        ->  tests/golden/msckf_reference.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def main():
  sys.path.insert(0, REPO)
  from oracle import build_ref
  build_ref.build("msckf", "rednose_b200.filters.msckf:MsckfKalman")
  from rednose_b200.filters.msckf import DIM_AUGMENT, DIM_AUGMENT_ERR, N_CLONES, MsckfKalman
  from rednose_b200.filters.live import DIM_STATE, DIM_STATE_ERR
  from tests.util import LIVE_R, msckf_batch
  QUATS = [3] + [DIM_STATE + 3 + 7 * c for c in range(N_CLONES)]
  sys.path.insert(0, REF)
  for m in [k for k in sys.modules if k == "rednose" or k.startswith("rednose.")]:
    del sys.modules[m]
  from rednose.helpers.ekf_sym import EKF_sym
  import rednose
  assert os.path.realpath(rednose.__file__).startswith(REF), rednose.__file__

  NF = 2
  x0, P0, Q, point = msckf_batch(NF, seed=77)
  FEAT = int(MsckfKalman.feature_kind)
  # (kind, augment after the update)
  plan = [(12, False), (FEAT, False), (4, True), (10, False), (FEAT, True), (12, False), (FEAT, False)]
  ts = 0.01 * np.arange(1, len(plan) + 1)
  rng = np.random.default_rng(77)
  out = dict(x0=x0, P0=P0, Q=Q, point=point, kinds=np.array([k for k, _ in plan]), augment=np.array([a for _, a in plan]), t=ts)
  for b in range(NF):
    kf = EKF_sym(build_ref.OUT, "msckf", Q, x0[b], P0[b], DIM_STATE, DIM_STATE_ERR, N=N_CLONES, dim_augment=DIM_AUGMENT,
                 dim_augment_err=DIM_AUGMENT_ERR, maha_test_kinds=[FEAT], quaternion_idxs=QUATS)
    kf._predict = kf._predict_python      # the reference's numpy maths
    kf._update = kf._update_python
    zs, Rs, xs, Ps = [], [], [], []
    for k, (kind, aug) in enumerate(plan):
      if kind == FEAT:
        m, ea, sig2 = 2 * N_CLONES, np.ascontiguousarray(point[b]), 1e-6
      else:
        m, ea, sig2 = 3, np.zeros(1), None
      hz = np.zeros((m, 1))
      kf.hs[kind](kf.x, ea, hz)
      Rd = np.full(m, sig2) if sig2 else np.array(LIVE_R[kind])
      z = hz[:, 0] + rng.normal(size=m) * np.sqrt(Rd)
      est = kf.predict_and_update_batch(ts[k], kind, z[None, :], np.diag(Rd)[None], extra_args=[ea] if kind == FEAT else [[]], augment=bool(aug))
      assert est is not None
      zp, Rp = np.zeros(2 * N_CLONES), np.zeros(2 * N_CLONES)
      zp[:m], Rp[:m] = z, Rd
      zs.append(zp); Rs.append(Rp); xs.append(kf.state().copy()); Ps.append(kf.covs().copy())
    out[f"z{b}"], out[f"Rdiag{b}"], out[f"xk{b}"], out[f"Pk{b}"] = np.array(zs), np.array(Rs), np.array(xs), np.array(Ps)
  path = os.path.join(HERE, "msckf_reference.npz")
  np.savez_compressed(path, **out)
  print("wrote", path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
  main()
