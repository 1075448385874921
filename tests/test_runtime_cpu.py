"""CPU tests of librednose_b200.so: plugin registry (rednose/helpers/ekf_load.cc) and native driver plumbing."""
import numpy as np

from rednose_b200.ekf_sym_pyx import EKF_sym_pyx, runtime


def test_registry_load_lookup(oracle_dir):
  rt = runtime()
  assert rt.rednose_b200_load_and_register(oracle_dir.encode(), b"kinematic") == 0
  assert rt.rednose_b200_load_and_register(oracle_dir.encode(), b"kinematic") == 0   # idempotent (ekf_load.cc:23-25)
  assert rt.rednose_b200_lookup(b"kinematic")
  assert not rt.rednose_b200_lookup(b"no_such_filter")
  assert rt.rednose_b200_load_and_register(oracle_dir.encode(), b"no_such_filter") != 0


def test_driver_state_views_and_time(oracle_dir):
  Q, x0, P0 = np.diag([0.01, 4.0]), np.array([0.5, 0.0]), np.eye(2)
  kf = EKF_sym_pyx(oracle_dir, "kinematic", Q, x0, P0, 2, 2)
  assert np.isnan(kf.get_filter_time()) and kf.filter_time is None      # ekf_sym.cc:42
  kf.x[1, 0] = 3.0                                                          # write-through view
  assert kf.state()[1] == 3.0
  kf.predict(0.5)                                                           # first predict only sets the time (dt = 0)
  assert kf.get_filter_time() == 0.5 and kf.state()[0] == 0.5
  kf.predict(1.0)
  assert abs(kf.state()[0] - (0.5 + 0.5 * 3.0)) < 1e-15
  kf.init_state(np.array([1.0, 2.0]), np.eye(2) * 4.0, 7.0)
  assert kf.get_filter_time() == 7.0 and np.array_equal(kf.covs(), np.eye(2) * 4.0)
  r = kf.predict_and_update_batch(7.1, 1, np.array([[1.2]]), np.array([[[0.01]]]))
  assert len(r) == 9 and r[4] == 7.1 and r[5] == 1 and r[6][0].shape == (1,)
  try:
    kf.predict_and_update_batch(7.2, 99, np.array([[1.2]]), np.array([[[0.01]]]))
    raise AssertionError("unknown kind must raise")
  except KeyError:
    pass


def test_runtime_exports_every_symbol_declared_in_the_public_header():
  """include/rednose_b200.h: every function prototype must resolve in librednose_b200.so."""
  import re
  from rednose_b200.build import INCLUDE_DIR
  rt = runtime()
  with open(f"{INCLUDE_DIR}/rednose_b200.h", encoding="utf-8") as f:
    text = f.read()
  names = re.findall(r"^(?:void|int|double|const rednose_ekf_desc) \*?(rednose_\w+)\(", text, flags=re.M)
  assert len(names) >= 18, names
  for n in names:
    assert hasattr(rt, n), n


def test_product_fails_loudly_without_a_gpu(gen_dir):
  """No CPU fallback: on a box without CUDA the single-filter C-ABI latches a CUDA error and the binding raises."""
  import pytest
  import torch
  if torch.cuda.is_available():
    pytest.skip("a GPU is present")
  from rednose_b200.ekf_sym import EKF_sym
  from rednose_b200.batched import BatchedEKF
  kf = EKF_sym(gen_dir, "kinematic", np.diag([0.01, 4.0]), np.array([0.5, 0.0]), np.eye(2), 2, 2)
  with pytest.raises(RuntimeError, match="CUDA error"):
    kf.predict_and_update_batch(0.0, 1, np.array([[0.1]]), np.array([[[0.01]]]))
  with pytest.raises(RuntimeError, match="no CPU fallback"):
    BatchedEKF(gen_dir, "kinematic", np.diag([0.01, 4.0]), np.array([0.5, 0.0]), np.eye(2), batch=4)


def test_augment_native_driver_equals_python_driver_and_reference_formula(oracle_dir):
  """MSCKF clone-window shift (ekf_sym.py:365-391): the C++ driver, the Python driver and the selection-matrix
  formula of the reference agree bit for bit; augment=True inside predict_and_update_batch shifts after the update."""
  import os
  import pytest
  from oracle import build_ref
  if build_ref.reference_available():
    build_ref.build("msckf", "rednose_b200.filters.msckf:MsckfKalman")
  if not os.path.exists(os.path.join(oracle_dir, "libmsckf.so")):
    pytest.skip("oracle/_ref/libmsckf.so not built")
  from rednose_b200.ekf_sym import EKF_sym
  from rednose_b200.filters.msckf import MsckfKalman as F
  from tests.util import msckf_batch
  x, P, Q, _ = msckf_batch(1, seed=4)
  kw = dict(N=10, dim_augment=7, dim_augment_err=6, quaternion_idxs=[3])
  a = EKF_sym_pyx(oracle_dir, "msckf", Q, x[0], P[0], 23, 22, **kw)
  b = EKF_sym(oracle_dir, "msckf", Q, x[0], P[0], 23, 22, **kw)
  a.set_filter_time(1.5); b.set_filter_time(1.5)
  a.augment(); b.augment()
  d1, d2, d3, d4, n = 23, 22, 7, 6, 82
  xr = x[0].copy()
  xr[d1:-d3] = x[0][d1 + d3:]
  xr[-d3:] = x[0][:d3]
  T = np.zeros((n, n - d4))
  T[:-d4, :] = np.eye(n - d4)
  T[-d4:, :d4] = np.eye(d4)
  Pr = T @ np.delete(np.delete(P[0], np.s_[d2:d2 + d4], axis=1), np.s_[d2:d2 + d4], axis=0) @ T.T
  for kf in (a, b):
    assert np.array_equal(kf.state(), xr) and np.array_equal(kf.covs(), Pr)
    assert kf.get_augment_times()[-1] == 1.5 and len(kf.get_augment_times()) == 10
  # augment=True through the batch call: update first, then shift (ekf_sym.py:525-526)
  z, R = x[0][None, :3] + 0.1, np.diag([25.0] * 3)[None]
  ra = a.predict_and_update_batch(1.6, 12, z, R, augment=True)
  rb = b.predict_and_update_batch(1.6, 12, z, R, augment=True)
  assert np.allclose(ra[1], rb[1], rtol=0, atol=1e-9) and np.allclose(a.state(), b.state(), rtol=0, atol=1e-9)
  assert np.allclose(a.state()[-7:], a.state()[:7]) and np.allclose(a.covs(), b.covs(), rtol=1e-12, atol=1e-12)


def test_python_maha_test_follows_the_reference_formula(oracle_dir):
  """EKF_sym.maha_test (ekf_sym.py:626-649) on the CPU oracle library."""
  from rednose_b200.chi2 import chi2_ppf
  from rednose_b200.ekf_sym import EKF_sym
  from rednose_b200.filters.live import LiveKalman as F
  from tests.util import Oracle, live_batch, live_obs
  o = Oracle(oracle_dir, "live")
  x, P, Q = live_batch(6, seed=8)
  kf = EKF_sym(oracle_dir, "live", Q, x[0], P[0], 23, 22)
  z, R = live_obs(o, 4, x, seed=3, noise_scale=4.0)
  for b in range(6):
    h, H, Hm = np.zeros(3), np.zeros(3 * 23), np.zeros(23 * 22)
    xb = np.ascontiguousarray(x[b])
    o.leaf("h_4", xb, np.zeros(1), h); o.leaf("H_4", xb, np.zeros(1), H); o.leaf("H_mod_fun", xb, Hm)
    He = H.reshape(3, 23) @ Hm.reshape(23, 22)
    y = z[b] - h
    d = y @ np.linalg.inv(He @ P[b] @ He.T + R[b]) @ y
    assert kf.maha_test(xb, P[b], 4, z[b], R[b]) == bool(d <= chi2_ppf(0.95, 3))


def test_two_builds_of_one_filter_coexist_per_directory(oracle_dir, tmp_path):
  """The reference's registry is keyed by name (ekf_load.cc:13-25: the second library of a name is never loaded);
  here a plugin is found by (directory, name), so a driver created on directory B runs B's library even when a
  filter of the same name from directory A is already loaded (CUDA library next to a CPU build in one process)."""
  import ctypes
  import shutil
  rt = runtime()
  rt.rednose_b200_lookup_in.restype = ctypes.c_void_p
  rt.rednose_b200_lookup_in.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
  other = str(tmp_path)
  for f in ("libkinematic.so", "kinematic.h"):
    shutil.copy(f"{oracle_dir}/{f}", f"{other}/{f}")
  assert rt.rednose_b200_load_and_register(oracle_dir.encode(), b"kinematic") == 0
  assert rt.rednose_b200_load_and_register(other.encode(), b"kinematic") == 0
  a, b = rt.rednose_b200_lookup_in(oracle_dir.encode(), b"kinematic"), rt.rednose_b200_lookup_in(other.encode(), b"kinematic")
  assert a and b and a != b                                       # two distinct plugin descriptors
  assert rt.rednose_b200_lookup(b"kinematic") in (a, b)            # name-only lookup: first registered
  assert not rt.rednose_b200_lookup_in(b"/nonexistent", b"kinematic")
  Q, x0, P0 = np.diag([0.01, 4.0]), np.array([0.5, 0.0]), np.eye(2)
  k1, k2 = EKF_sym_pyx(oracle_dir, "kinematic", Q, x0, P0, 2, 2), EKF_sym_pyx(other, "kinematic", Q, x0, P0, 2, 2)
  for kf in (k1, k2):
    kf.predict_and_update_batch(0.1, 1, np.array([[1.2]]), np.array([[[0.01]]]))
  assert np.array_equal(k1.state(), k2.state())


def test_empty_observation_batch_predicts_and_checkpoints(oracle_dir):
  """n = 0 observations (KalmanFilter.predict_and_observe with no data; ekf_sym.cc:158-194 loops zero times):
  the driver predicts to t, returns the predicted state as both x_{k|k-1} and x_{k|k}, and keeps a checkpoint."""
  Q, x0, P0 = np.diag([0.01, 4.0]), np.array([0.5, 2.0]), np.eye(2)
  kf = EKF_sym_pyx(oracle_dir, "kinematic", Q, x0, P0, 2, 2)
  kf.predict_and_update_batch(1.0, 1, np.array([[0.6]]), np.array([[[0.01]]]))
  before = kf.state().copy()
  r = kf.predict_and_update_batch(1.5, 1, [], [])
  assert r is not None and r[6] == [] and np.array_equal(r[0], r[1]) and np.array_equal(r[2], r[3])
  assert abs(kf.state()[0] - (before[0] + 0.5 * before[1])) < 1e-12 and kf.get_filter_time() == 1.5


def test_covariance_view_survives_augment(oracle_dir):
  """The .P view points into the driver's storage; augment() (ekf_sym.py:365-391) must rewrite that storage in place."""
  import os
  import pytest
  from oracle import build_ref
  if not os.path.exists(os.path.join(build_ref.OUT, "libmsckf.so")):
    pytest.skip("oracle/_ref/libmsckf.so not built")
  from rednose_b200.filters.msckf import MsckfKalman
  kf = MsckfKalman(build_ref.OUT).filter
  rng = np.random.default_rng(0)
  L = np.tril(rng.normal(size=(82, 82))) * 0.01 + np.eye(82)
  kf.init_state(MsckfKalman.initial_x, L @ L.T, 0.0)
  view = kf.P
  want_new_clone = kf.covs()[:6, :6].copy()
  kf.augment()
  assert np.array_equal(view, kf.covs()) and np.array_equal(view[-6:, -6:], want_new_clone)
