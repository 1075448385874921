#!/usr/bin/env python3
"""Generate golden vectors from the REFERENCE ITSELF (run here, where /root/reference is mounted).

The reference's own Python driver `EKF_sym` (rednose/helpers/ekf_sym.py, imported unmodified from
/root/reference) is pointed at oracle/_ref/liblive.so for its leaf functions (reference-generated C) and
switched to its in-repo numpy maths `_predict_python` / `_update_python` (ekf_sym.py:346-349,533-624), so
every number below is produced by reference code only: sympy-generated f/F/h/H/H_mod/err/inv_err +
the reference's numpy predict/update/rts_smooth.  Output: tests/golden/live_reference.npz.

  python tests/golden/make_golden.py
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
  build_ref.build("live")
  build_ref.build("kinematic")
  # import the reference package, not this repo's compatibility shim
  sys.path.insert(0, REF)
  for m in [k for k in sys.modules if k == "rednose" or k.startswith("rednose.")]:
    del sys.modules[m]
  from rednose.helpers.ekf_sym import EKF_sym
  import rednose
  assert os.path.realpath(rednose.__file__).startswith(REF), rednose.__file__
  from tests.util import LIVE_KINDS, LIVE_R, live_batch

  rng = np.random.default_rng(2024)
  NF, T = 2, 40
  x0, P0, Q = live_batch(NF, seed=2024)
  kinds = [12] + [int(rng.choice([4, 10, 3, 13, 9, 14, 19, 12])) for _ in range(T - 1)]
  ts = 0.01 * np.arange(1, T + 1)
  out = dict(x0=x0, P0=P0, Q=Q, kinds=np.array(kinds), t=ts)
  for b in range(NF):
    kf = EKF_sym(build_ref.OUT, "live", Q, x0[b], P0[b], 23, 22, quaternion_idxs=[3])
    kf._predict = kf._predict_python   # the reference's numpy maths
    kf._update = kf._update_python
    zs, Rs, ys, xk, Pk, xkm1, Pkm1, estimates = [], [], [], [], [], [], [], []
    for k in range(T):
      kind = kinds[k]
      m = LIVE_KINDS[kind]
      hz = np.zeros((m, 1))
      kf.hs[kind](kf.x, np.zeros(1), hz)
      z = hz[:, 0] + rng.normal(size=m) * np.sqrt(np.array(LIVE_R[kind]))
      R = np.diag(LIVE_R[kind])
      zpad, Rpad = np.zeros(3), np.zeros((3, 3))
      zpad[:m], Rpad[:m, :m] = z, R
      est = kf.predict_and_update_batch(ts[k], kind, z[None, :], R[None, :, :])
      estimates.append(est)
      zs.append(zpad); Rs.append(Rpad)
      ypad = np.zeros(3); ypad[:m] = np.asarray(est[6][0]).ravel()
      ys.append(ypad)
      xkm1.append(est[0].copy()); xk.append(est[1].copy()); Pkm1.append(est[2].copy()); Pk.append(est[3].copy())
    out[f"z{b}"], out[f"R{b}"], out[f"y{b}"] = np.array(zs), np.array(Rs), np.array(ys)
    out[f"x_pred{b}"], out[f"x_filt{b}"] = np.array(xkm1), np.array(xk)
    out[f"P_pred{b}"], out[f"P_filt{b}"] = np.array(Pkm1), np.array(Pk)
    xs, Ps = kf.rts_smooth(estimates, norm_quats=True)   # mutates `estimates` in place, hence the copies above
    out[f"xs{b}"], out[f"Ps{b}"] = np.array(xs), np.array(Ps)
  # kinematic known-answer values straight from the reference test file (examples/test_kinematic_kf.py:52-55)
  out["kinematic_golden"] = np.array([-0.010866289677966417, 0.04477103863330089, -0.8553720537261753, 0.6695762270974388])
  np.savez_compressed(os.path.join(HERE, "live_reference.npz"), **out)
  print("wrote", os.path.join(HERE, "live_reference.npz"), {k: v.shape for k, v in out.items() if k.endswith("0")})


if __name__ == "__main__":
  main()
