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
