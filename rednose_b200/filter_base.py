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


class BatchedKalmanFilter:
  """The same front-end for a batch of filters on the GPU: ``self.filter`` is a rednose_b200.batched.BatchedEKF.

  ``predict_and_observe(t, kind, data[B, m])`` = the batch-wide counterpart of KalmanFilter.predict_and_observe
  (rednose/helpers/kalmanfilter.py:45-52); the per-kind noise of ``obs_noise`` is passed once and shared by the
  whole batch instead of being replicated n times by get_R (:37-43).
  """
  obs_noise: dict[int, Any] = {}
  filter = None

  x = property(lambda self: self.filter.state())
  t = property(lambda self: self.filter.filter_time)
  P = property(lambda self: self.filter.covs())

  def init_state(self, state, covs_diag=None, covs=None, filter_time=None):
    if covs_diag is not None:
      covs = np.diag(covs_diag)
    elif covs is None:
      covs = self.filter.covs()
    self.filter.init_state(state, covs, filter_time)

  def get_R(self, kind, n=None):
    return np.asarray(self.obs_noise[kind], dtype=np.float64)   # [m, m], shared by the batch

  def predict_and_observe(self, t, kind, data, R=None):
    return self.filter.predict_and_update_batch(t, kind, data, self.get_R(kind) if R is None else R)

  def maha_test(self, kind, data, R=None, maha_thresh=0.95):
    return self.filter.maha_test(kind, data, self.get_R(kind) if R is None else R, maha_thresh=maha_thresh)
