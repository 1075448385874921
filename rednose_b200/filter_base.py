"""Base class for user filter definitions (mirror of rednose/helpers/kalmanfilter.py:6-52).

A subclass sets ``name``, ``initial_x``, ``initial_P_diag``, ``Q``, ``obs_noise`` and
creates ``self.filter`` (an EKF_sym / EKF_sym_pyx); this class only forwards.
"""
from typing import Any

import numpy as np


class KalmanFilter:
  name = "<name>"
  initial_x: np.ndarray = np.zeros(0)
  initial_P_diag: np.ndarray = np.zeros(0)
  Q: np.ndarray = np.zeros((0, 0))
  obs_noise: dict[int, Any] = {}
  filter = None  # set by the subclass constructor

  x = property(lambda self: self.filter.state())
  t = property(lambda self: self.filter.get_filter_time())
  P = property(lambda self: self.filter.covs())

  def init_state(self, state, covs_diag=None, covs=None, filter_time=None):
    # precedence: explicit diagonal, explicit matrix, keep current covariance
    if covs_diag is not None:
      covs = np.diag(covs_diag)
    elif covs is None:
      covs = self.filter.covs()
    self.filter.init_state(state, covs, filter_time)

  def get_R(self, kind, n):
    # the per-kind noise matrix repeated n times -> [n, m, m]
    noise = np.asarray(self.obs_noise[kind])
    return np.broadcast_to(noise, (n,) + noise.shape).copy()

  def predict_and_observe(self, t, kind, data, R=None):
    if len(data) > 0:
      data = np.atleast_2d(data)
    if R is None:
      R = self.get_R(kind, len(data))
    return self.filter.predict_and_update_batch(t, kind, data, R)
