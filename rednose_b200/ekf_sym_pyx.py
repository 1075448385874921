"""``EKF_sym_pyx``: the class the examples import (examples/kinematic_kf.py:12, live_kf.py:12).

In the reference this is a Cython wrapper (rednose/helpers/ekf_sym_pyx.pyx) around the C++ driver
``EKFSym`` (rednose/helpers/ekf_sym.cc).  Here it is a ctypes binding over the native driver in
``librednose_b200.so`` (csrc/runtime.cc), which owns x, P, the filter time and the rewind ring and calls
the filter library's C-ABI (CUDA kernels) for every predict / update.

The class exposes the UNION of the Cython API (ekf_sym_pyx.pyx:113-192) and of what examples/live_kf.py
touches on the Python driver: ``.x`` as a live (DIM, 1) view, ``.P``, ``.filter_time``, ``rts_smooth``,
``augment``, ``get_augment_times``, ``maha_test`` -- the reference's Cython class leaves the last four as
``NotImplementedError`` (:182-192), so live_kf.py only ever ran against the Python driver there.

C++-driver behaviour kept: quaternions are normalised after the predict inside
predict_and_update_batch (ekf_sym.cc:162 -> :207); an unset filter time reads back as NaN (:42).
"""
import ctypes
import os

import numpy as np

from rednose_b200 import build
from rednose_b200.ekf_sym import EKF_sym

_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_int_p = ctypes.POINTER(ctypes.c_int)
_runtime = None


def runtime():
  """Load (building it if needed) librednose_b200.so; RTLD_GLOBAL so filter libraries can self-register."""
  global _runtime  # pylint: disable=global-statement
  if _runtime is None:
    path = build.compile_runtime()
    rt = ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
    rt.rednose_ekfsym_create.restype = ctypes.c_void_p
    rt.rednose_ekfsym_create.argtypes = [ctypes.c_char_p, ctypes.c_char_p, _c_double_p, _c_double_p, _c_double_p] + [ctypes.c_int] * 7 + \
                                        [_c_int_p, ctypes.c_int, _c_int_p, ctypes.c_int, ctypes.c_double]
    rt.rednose_ekfsym_destroy.argtypes = [ctypes.c_void_p]
    rt.rednose_ekfsym_init_state.argtypes = [ctypes.c_void_p, _c_double_p, _c_double_p, ctypes.c_double]
    for fn in ("x_ptr", "P_ptr"):
      getattr(rt, f"rednose_ekfsym_{fn}").restype = _c_double_p
      getattr(rt, f"rednose_ekfsym_{fn}").argtypes = [ctypes.c_void_p]
    rt.rednose_ekfsym_get_filter_time.restype = ctypes.c_double
    rt.rednose_ekfsym_get_filter_time.argtypes = [ctypes.c_void_p]
    rt.rednose_ekfsym_set_filter_time.argtypes = [ctypes.c_void_p, ctypes.c_double]
    for fn in ("reset_rewind", "normalize_quaternions", "augment"):
      getattr(rt, f"rednose_ekfsym_{fn}").argtypes = [ctypes.c_void_p]
    rt.rednose_ekfsym_rewind_depth.argtypes = [ctypes.c_void_p]
    rt.rednose_ekfsym_get_augment_times.argtypes = [ctypes.c_void_p, _c_double_p]
    rt.rednose_ekfsym_set_global.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_double]
    rt.rednose_ekfsym_predict.argtypes = [ctypes.c_void_p, ctypes.c_double]
    rt.rednose_ekfsym_predict_and_update_batch.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_int, _c_double_p, _c_double_p, _c_double_p,
                                                           ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int] + [_c_double_p] * 5
    rt.rednose_b200_lookup.restype = ctypes.c_void_p
    rt.rednose_b200_lookup.argtypes = [ctypes.c_char_p]
    rt.rednose_b200_load_and_register.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
    _runtime = rt
  return _runtime


def _dp(a):
  return a.ctypes.data_as(_c_double_p)


class EKF_sym_pyx(EKF_sym):
  def __init__(self, gen_dir, name, Q, x_initial, P_initial, dim_main, dim_main_err, N=0, dim_augment=0,  # pylint: disable=dangerous-default-value
               dim_augment_err=0, maha_test_kinds=[], quaternion_idxs=[], global_vars=[], max_rewind_age=1.0, logger=None):
    self._rt = runtime()
    self._h = None
    Qc = np.ascontiguousarray(Q, dtype=np.float64)
    x0 = np.ascontiguousarray(np.asarray(x_initial, dtype=np.float64).reshape(-1))
    P0 = np.ascontiguousarray(P_initial, dtype=np.float64)
    mk = (ctypes.c_int * max(1, len(maha_test_kinds)))(*maha_test_kinds)
    qi = (ctypes.c_int * max(1, len(quaternion_idxs)))(*quaternion_idxs)
    h = self._rt.rednose_ekfsym_create(os.fsencode(gen_dir), name.encode(), _dp(Qc), _dp(x0), _dp(P0), x0.shape[0], P0.shape[0],
                                       dim_main, dim_main_err, N, dim_augment, dim_augment_err, mk, len(maha_test_kinds),
                                       qi, len(quaternion_idxs), float(max_rewind_age))
    if not h:
      raise RuntimeError(f"could not create the native driver for filter '{name}' in {gen_dir}")
    self._h = ctypes.c_void_p(h)
    # leaf functions / rts_smooth / maha_test come from the Python driver, bound to the same library;
    # its x / P / filter_time attributes are redirected to the native driver below
    super().__init__(gen_dir, name, Qc, x0, P0, dim_main, dim_main_err, N, dim_augment, dim_augment_err,
                     maha_test_kinds, quaternion_idxs, global_vars, max_rewind_age, logger)

  def __del__(self):
    if getattr(self, "_h", None):
      self._rt.rednose_ekfsym_destroy(self._h)
      self._h = None

  # ---- state lives in the native driver; these are live views ----
  @property
  def x(self):
    return np.ctypeslib.as_array(self._rt.rednose_ekfsym_x_ptr(self._h), shape=(self.dim_x, 1))

  @x.setter
  def x(self, v):
    if self._h is not None and hasattr(self, "dim_x"):
      self.x[:] = np.asarray(v, dtype=np.float64).reshape(self.dim_x, 1)

  @property
  def P(self):
    return np.ctypeslib.as_array(self._rt.rednose_ekfsym_P_ptr(self._h), shape=(self.dim_err, self.dim_err))

  @P.setter
  def P(self, v):
    if self._h is not None and hasattr(self, "dim_err"):
      self.P[:] = np.asarray(v, dtype=np.float64)

  @property
  def filter_time(self):
    t = self._rt.rednose_ekfsym_get_filter_time(self._h)
    return None if np.isnan(t) else t

  @filter_time.setter
  def filter_time(self, t):
    if self._h is not None:
      self._rt.rednose_ekfsym_set_filter_time(self._h, np.nan if t is None else float(t))

  def init_state(self, state, covs, filter_time):
    if self._h is None:
      return
    s = np.ascontiguousarray(np.asarray(state, dtype=np.float64).reshape(-1))
    c = np.ascontiguousarray(covs, dtype=np.float64)
    self._rt.rednose_ekfsym_init_state(self._h, _dp(s), _dp(c), np.nan if filter_time is None else float(filter_time))

  def state(self):
    return np.array(self.x).flatten()

  def covs(self):
    return np.array(self.P)

  def get_filter_time(self):
    return self._rt.rednose_ekfsym_get_filter_time(self._h)  # NaN when unset, like the C++ driver

  def set_filter_time(self, t):
    self._rt.rednose_ekfsym_set_filter_time(self._h, float(t))

  def reset_rewind(self):
    self._rt.rednose_ekfsym_reset_rewind(self._h)

  def normalize_quaternions(self):
    self._rt.rednose_ekfsym_normalize_quaternions(self._h)

  def set_global(self, global_var, val):
    if self._rt.rednose_ekfsym_set_global(self._h, str(global_var).encode(), float(val)) != 0:
      raise KeyError(global_var)

  def augment(self):
    self._rt.rednose_ekfsym_augment(self._h)

  def get_augment_times(self):
    out = np.zeros(max(1, self.N))
    self._rt.rednose_ekfsym_get_augment_times(self._h, _dp(out))
    return list(out[:self.N])

  def predict(self, t):
    self._rt.rednose_ekfsym_predict(self._h, float(t))
    self._raise_on_cuda("predict")

  def _raise_on_cuda(self, what):
    from rednose_b200.loader import raise_on_cuda_error
    raise_on_cuda_error(self._lib, self.name, what)

  def predict_and_update_batch(self, t, kind, z, R, extra_args=[[]], augment=False):  # pylint: disable=dangerous-default-value
    n = len(z)
    if n == 0:
      # an empty observation batch is legal (KalmanFilter.predict_and_observe with no data, ekf_sym.cc:158-194 with n = 0):
      # predict to t, checkpoint, return the predicted state twice
      zc, zdim, Rc = np.zeros((0, 1)), 0, np.zeros((0, 1, 1))
    else:
      zc = np.ascontiguousarray(np.asarray([np.asarray(zi, dtype=np.float64).reshape(-1) for zi in z], dtype=np.float64).reshape(n, -1))
      zdim = zc.shape[1]
      Rc = np.ascontiguousarray(np.asarray(R, dtype=np.float64).reshape(n, zdim, zdim))
    ea_rows = [np.asarray(extra_args[i] if i < len(extra_args) else [], dtype=np.float64).reshape(-1) for i in range(n)]
    eadim = ea_rows[0].shape[0] if n else 0
    eac = np.ascontiguousarray(np.asarray(ea_rows, dtype=np.float64).reshape(n, eadim)) if eadim else np.zeros(1)
    xk1, xk = np.empty(self.dim_x), np.empty(self.dim_x)
    Pk1, Pk = np.empty((self.dim_err, self.dim_err)), np.empty((self.dim_err, self.dim_err))
    y = np.zeros((max(n, 1), max(zdim, 1)))
    rc = self._rt.rednose_ekfsym_predict_and_update_batch(self._h, float(t), int(kind), _dp(zc), _dp(Rc), _dp(eac), n, zdim, eadim,
                                                          int(bool(augment)), _dp(xk1), _dp(xk), _dp(Pk1), _dp(Pk), _dp(y))
    if rc < 0:
      raise KeyError(kind)
    self._raise_on_cuda(f"update_{kind}")
    if rc == 0:
      return None
    ydim = zdim - eadim if (self.msckf and kind in self.feature_track_kinds) else zdim
    return xk1, xk, Pk1, Pk, t, kind, [y[i, :ydim].copy() for i in range(n)], z, extra_args
