"""CPU tests: the ORACLE (restated ekf_c.c + reference-generated leaf C, and the restated RTS) against
golden vectors produced by the reference's own Python maths (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from tests.util import LIVE_KINDS, Oracle, rel_err

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "live_reference.npz")


@pytest.fixture(scope="module")
def gold():
  return np.load(GOLD)


def test_oracle_matches_reference_python_maths_on_live(oracle_dir, gold):
  """Forward filter, 40 steps over all 8 kinds: restated C core == reference numpy predict/update."""
  o = Oracle(oracle_dir, "live")
  Q, kinds, ts = gold["Q"], gold["kinds"], gold["t"]
  for b in range(2):
    x, P = gold["x0"][b:b + 1].copy(), gold["P0"][b:b + 1].copy()
    t_prev = ts[0]
    for k, kind in enumerate(kinds):
      m = LIVE_KINDS[int(kind)]
      z, R = gold[f"z{b}"][k, :m][None], gold[f"R{b}"][k, :m, :m][None]
      # python-driver semantics (ekf_sym.py:505-522): predict, update, then normalise
      x, P, y = o.batch_step(int(kind), x, P, Q, ts[k] - t_prev, z, R, quat_idxs=[3], flags=2, nthreads=1)
      t_prev = ts[k]
      assert rel_err(x[0], gold[f"x_filt{b}"][k]) < 1e-10, (b, k, kind)
      assert rel_err(P[0], gold[f"P_filt{b}"][k]) < 1e-9, (b, k, kind)
      assert rel_err(y[0], gold[f"y{b}"][k, :m]) < 1e-7 or np.max(np.abs(y[0] - gold[f"y{b}"][k, :m])) < 1e-9, (b, k, kind)


def test_restated_rts_matches_reference_rts(oracle_dir, gold):
  from oracle.rts_numpy import rts_smooth
  o = Oracle(oracle_dir, "live")
  for b in range(2):
    xs, Ps = rts_smooth(o, gold[f"x_pred{b}"], gold[f"x_filt{b}"], gold[f"P_pred{b}"], gold[f"P_filt{b}"], gold["t"], 23, 22, norm_quats=True)
    assert rel_err(xs, gold[f"xs{b}"]) < 1e-12 and rel_err(Ps, gold[f"Ps{b}"]) < 1e-12
