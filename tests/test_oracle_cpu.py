"""CPU tests (no GPU): pin the ORACLE to the reference's golden vectors, and exercise the host-side
driver logic (time handling, rewind / fast-forward) by pointing the drivers at the oracle's CPU
build of the same C-ABI.  Nothing here touches the product's compute path."""
import numpy as np
import pytest

from rednose_b200.ekf_sym import EKF_sym
from rednose_b200.ekf_sym_pyx import EKF_sym_pyx

# reference known-answer values: examples/test_kinematic_kf.py:52-55
GOLDEN = (-0.010866289677966417, 0.04477103863330089, -0.8553720537261753, 0.6695762270974388)
Q = np.diag([0.1**2, 2.0**2])
X0 = np.array([0.5, 0.0])
P0 = np.diag([1.0, 1.0])


def run_kinematic_procedure(kf):
  """The exact procedure of examples/test_kinematic_kf.py:10-48 (legacy seed 0, measurement drawn before each update)."""
  np.random.seed(0)
  dt = 0.01
  ts = np.arange(0, 5, step=dt)
  x = 0.0
  for t, v in zip(ts, np.sin(ts * 5)):
    meas = np.random.normal(x, 0.1)
    kf.predict_and_update_batch(t, 1, np.atleast_2d([meas]), np.array([[[0.1**2]]]))
    x += v * dt
  s, P = kf.state(), kf.covs()
  return s[0], np.sqrt(P[0, 0]), s[1], np.sqrt(P[1, 1])


@pytest.mark.parametrize("cls", [EKF_sym, EKF_sym_pyx])
def test_oracle_reproduces_reference_golden(oracle_dir, cls):
  kf = cls(oracle_dir, "kinematic", Q, X0, P0, 2, 2)
  got = run_kinematic_procedure(kf)
  for g, want in zip(got, GOLDEN):
    assert abs(g - want) < 5e-8  # assertAlmostEqual, 7 places, like the reference test
    assert abs(g - want) < 1e-13  # and in fact to rounding


def test_compare_procedure_rewind(oracle_dir):
  """examples/test_compare.py:86-120: both drivers agree after every step with observations 20 <-> 40 swapped."""
  np.random.seed(0)
  a = EKF_sym_pyx(oracle_dir, "compare", Q, X0, P0, 2, 2)
  b = EKF_sym(oracle_dir, "compare", Q, X0, P0, 2, 2)
  dt = 0.01
  ts = np.arange(0, 5, step=dt)
  xs = np.empty(ts.shape)
  x = 0.0
  for i, v in enumerate(np.sin(ts * 5)):
    xs[i] = x
    x += v * dt
  ts[20], ts[40] = ts[40], ts[20]
  xs[20], xs[40] = xs[40], xs[20]
  n_rewinds = 0
  for t, x in zip(ts, xs):
    z = np.array([[np.random.normal(x, 0.1)]])
    R = np.array([[[0.1**2]]])
    n_rewinds += int(a.get_filter_time() is not None and not np.isnan(a.get_filter_time()) and t < a.get_filter_time())
    a.predict_and_update_batch(t, 1, z, R)
    b.predict_and_update_batch(t, 1, z, R)
    assert abs(a.get_filter_time() - b.get_filter_time()) < 1e-7
    assert np.allclose(a.state(), b.state())
    assert np.allclose(a.covs(), b.covs())
  assert n_rewinds == 20


def test_too_old_observation_is_dropped(oracle_dir):
  kf = EKF_sym(oracle_dir, "kinematic", Q, X0, P0, 2, 2, max_rewind_age=0.05)
  z, R = np.array([[0.0]]), np.array([[[0.01]]])
  for t in np.arange(0, 0.5, 0.01):
    assert kf.predict_and_update_batch(t, 1, z, R) is not None
  before = kf.state().copy()
  assert kf.predict_and_update_batch(0.1, 1, z, R) is None  # older than max_rewind_age
  assert np.array_equal(before, kf.state())
  assert kf.predict_and_update_batch(0.47, 1, z, R) is not None  # within the window: rewinds
  assert abs(kf.get_filter_time() - 0.49) < 1e-12


def test_full_pivot_kernel_and_solve(oracle_dir):
  """Sanity of the restated FullPivLU through the live oracle: update twice with the same data is deterministic
  and P stays symmetric to rounding."""
  from tests.util import Oracle, live_batch, live_obs
  o = Oracle(oracle_dir, "live")
  x, P, Qm = live_batch(8, seed=3)
  z, R = live_obs(o, 4, x)
  x1, P1, y1 = o.batch_step(4, x, P, Qm, 0.01, z, R, nthreads=1)
  x2, P2, y2 = o.batch_step(4, x, P, Qm, 0.01, z, R, nthreads=4)
  assert np.array_equal(x1, x2) and np.array_equal(P1, P2) and np.array_equal(y1, y2)
  assert np.max(np.abs(P1 - np.transpose(P1, (0, 2, 1)))) / np.max(np.abs(P1)) < 1e-12
