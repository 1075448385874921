"""``EKF_sym_pyx``: the class the examples import (examples/kinematic_kf.py:12, live_kf.py:12).

In the reference this is a Cython wrapper (rednose/helpers/ekf_sym_pyx.pyx) around the C++
driver ``EKFSym`` (rednose/helpers/ekf_sym.cc).  Here it exposes the UNION of that API and of
what examples/live_kf.py touches on the Python driver (``.x``, ``.filter_time``, ``rts_smooth``,
``augment``, ``maha_test`` -- methods the Cython class leaves as NotImplementedError,
ekf_sym_pyx.pyx:182-192).  Behavioural differences of the C++ driver that are kept:

  * quaternions are normalised after the predict inside predict_and_update_batch
    (ekf_sym.cc:162 -> :207), which the Python driver does not do (ekf_sym.py:508);
  * an unset filter time reads back as NaN (ekf_sym.cc:42) rather than None;
  * ``augment=True`` is refused (``assert(!augment)``, ekf_sym.cc:186) unless the filter is MSCKF,
    where this class falls back to the Python driver's augment().
"""
import numpy as np

from rednose_b200.ekf_sym import EKF_sym


class EKF_sym_pyx(EKF_sym):
  def __init__(self, gen_dir, name, Q, x_initial, P_initial, dim_main, dim_main_err, N=0, dim_augment=0,  # pylint: disable=dangerous-default-value
               dim_augment_err=0, maha_test_kinds=[], quaternion_idxs=[], global_vars=[], max_rewind_age=1.0, logger=None):
    super().__init__(gen_dir, name, np.asarray(Q, dtype=np.float64), np.asarray(x_initial, dtype=np.float64),
                     np.asarray(P_initial, dtype=np.float64), dim_main, dim_main_err, N, dim_augment, dim_augment_err,
                     maha_test_kinds, quaternion_idxs, global_vars, max_rewind_age, logger)

  def get_filter_time(self):
    return np.nan if self.filter_time is None else self.filter_time

  def _predict_and_update_batch(self, t, kind, z, R, extra_args, augment=False):
    assert len(z) == len(R)
    if self.filter_time is None:
      self.filter_time = t
    dt = t - self.filter_time
    assert dt >= 0.0
    self.x, self.P = self._predict(self.x, self.P, dt)
    self.normalize_quaternions()          # ekf_sym.cc:207
    self.filter_time = t
    xk_km1, Pk_km1 = np.copy(self.x).flatten(), np.copy(self.P)
    y = []
    for i in range(len(z)):
      z_i = np.array(z[i], dtype=np.float64, order='C')
      R_i = np.array(R[i], dtype=np.float64, order='C')
      ea_i = np.array(extra_args[i] if i < len(extra_args) else [], dtype=np.float64)
      assert z_i.shape[0] == R_i.shape[0] == R_i.shape[1]
      self.x, self.P, y_i = self._update(self.x, self.P, kind, z_i, R_i, extra_args=ea_i)
      self.normalize_quaternions()        # ekf_sym.cc:213
      y.append(y_i)
    xk_k, Pk_k = np.copy(self.x).flatten(), np.copy(self.P)
    if augment:
      assert self.msckf, "augment requires an MSCKF filter"
      self.augment()
    self.checkpoint((t, kind, z, R, extra_args))
    return xk_km1, xk_k, Pk_km1, Pk_k, t, kind, y, z, extra_args
