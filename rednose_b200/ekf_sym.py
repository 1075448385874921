"""Single-filter Python driver over a generated CUDA filter library.

API mirror of the reference class ``EKF_sym`` (rednose/helpers/ekf_sym.py:220-690): same
constructor, same methods, same return tuples, so user code such as examples/live_kf.py runs
unchanged.  What differs is underneath: every numeric call (``<name>_predict``,
``<name>_update_<kind>``, leaf functions) executes CUDA kernels on the GPU through the
cffi-loaded C-ABI; there is no numpy or CPU implementation of the filter maths here (the
reference keeps ``_predict_python`` / ``_update_python`` alternates, :533-624 -- their
restatement lives under oracle/ as test infrastructure only).

Driver semantics kept from the reference:
  * time handling / dt assertion                      ekf_sym.py:452-462, 501-509
  * out-of-order observations: rewind + fast-forward  ekf_sym.py:464-482, 418-450 (ring of 512)
  * per-observation canonicalisation and quaternion normalisation after each update (:514-522)
  * the 9-tuple (xk_km1, xk_k, Pk_km1, Pk_k, t, kind, y, z, extra_args)  (:531)
  * rts_smooth (:651-690), maha_test (:626-649), augment (:365-391)
"""
from __future__ import annotations

import logging
from bisect import bisect_right

import numpy as np

from rednose_b200.chi2 import chi2_ppf
from rednose_b200.loader import load_code, raise_on_cuda_error
from rednose_b200.codegen import gen_code  # noqa: F401  (re-exported: `from ...ekf_sym import gen_code`)

REWIND_TO_KEEP = 512


def solve(a, b):
  a = np.asarray(a)
  if a.shape == (1, 1):
    return b / a[0, 0]
  return np.linalg.solve(a, b)


def null(H, eps=1e-12):
  """Orthonormal basis of the null space of H (columns), via SVD."""
  _, s, vh = np.linalg.svd(H)
  n_extra = max(0, H.shape[1] - s.shape[0])
  mask = np.concatenate([s <= eps, np.ones(n_extra, dtype=bool)])
  return vh[mask].T


class _RewindBuffer:
  """Checkpoints (filter_time, x, P, observation) for late-observation handling."""

  def __init__(self):
    self.clear()

  def clear(self):
    self.t, self.states, self.obs = [], [], []

  def push(self, t, x, P, obs):
    self.t.append(t)
    self.states.append((np.copy(x), np.copy(P)))
    self.obs.append(obs)
    del self.t[:-REWIND_TO_KEEP], self.states[:-REWIND_TO_KEEP], self.obs[:-REWIND_TO_KEEP]

  def too_old(self, t, max_age):
    return len(self.t) == 0 or t < self.t[0] or t < self.t[-1] - max_age

  def rewind_to(self, t):
    """Drop everything newer than t; return (time, x, P) to restore and the dropped observations."""
    idx = bisect_right(self.t, t)
    assert self.t[idx - 1] <= t and self.t[idx] > t
    restore = (self.t[idx - 1],) + self.states[idx - 1]
    replay = self.obs[idx:]
    del self.t[idx:], self.states[idx:], self.obs[idx:]
    return restore, replay


class EKF_sym:
  def __init__(self, folder, name, Q, x_initial, P_initial, dim_main, dim_main_err,  # pylint: disable=dangerous-default-value
               N=0, dim_augment=0, dim_augment_err=0, maha_test_kinds=[], quaternion_idxs=[], global_vars=None,
               max_rewind_age=1.0, logger=logging):
    self.name = name
    self.msckf = N > 0
    self.N, self.dim_augment, self.dim_augment_err = N, dim_augment, dim_augment_err
    self.dim_main, self.dim_main_err = dim_main, dim_main_err
    self.logger = logger if logger is not None else logging

    x_initial = np.asarray(x_initial).reshape((-1, 1))
    self.dim_x, self.dim_err = x_initial.shape[0], P_initial.shape[0]
    assert dim_main + dim_augment * N == self.dim_x
    assert dim_main_err + dim_augment_err * N == self.dim_err
    assert Q.shape == P_initial.shape

    self.maha_test_kinds = maha_test_kinds
    self.quaternion_idxs = quaternion_idxs
    self.Q = np.ascontiguousarray(Q, dtype=np.float64)
    self.max_rewind_age = max_rewind_age
    self._rewind = _RewindBuffer()
    self.init_state(x_initial, P_initial, None)

    self._ffi, self._lib = load_code(folder, name)
    self._bind(global_vars)

  # ------------------------------------------------------------------ binding ---
  def _ptr(self, arr):
    return self._ffi.cast("double *", arr.ctypes.data)

  def _bind(self, global_vars):
    lib, name = self._lib, self.name
    prefix_h, prefix_he = f"{name}_h_", f"{name}_He_"
    syms = dir(lib)
    kinds = [int(s[len(prefix_h):]) for s in syms if s.startswith(prefix_h)]
    self.feature_track_kinds = [int(s[len(prefix_he):]) for s in syms if s.startswith(prefix_he)]

    def checked(fn_name):
      fn = getattr(lib, f"{name}_{fn_name}")

      def call(*args):
        fn(*args)
        raise_on_cuda_error(lib, name, fn_name)
      return call

    def leaf(fn_name, scalar_middle=False):
      fn = checked(fn_name)
      if scalar_middle:
        return lambda a, s, out: fn(self._ptr(a), float(s), self._ptr(out))
      return lambda *arrs: fn(*[self._ptr(a) for a in arrs])

    self.f = leaf("f_fun", scalar_middle=True)
    self.F = leaf("F_fun", scalar_middle=True)
    self.err_function = leaf("err_fun")
    self.inv_err_function = leaf("inv_err_fun")
    self.H_mod = leaf("H_mod_fun")
    self.hs = {k: leaf(f"h_{k}") for k in kinds}
    self.Hs = {k: leaf(f"H_{k}") for k in kinds}
    self.Hes = {k: leaf(f"He_{k}") for k in kinds if self.msckf and k in self.feature_track_kinds}
    self.set_globals = {g: getattr(lib, f"{name}_set_{g}") for g in (global_vars or [])}

    predict_fn = checked("predict")

    def _predict(x, P, dt):
      predict_fn(self._ptr(x), self._ptr(P), self._ptr(self.Q), float(dt))
      return x, P
    self._predict = _predict

    update_fns = {k: checked(f"update_{k}") for k in kinds}

    def _update(x, P, kind, z, R, extra_args=()):
      ea = np.ascontiguousarray(extra_args, dtype=np.float64)
      update_fns[kind](self._ptr(x), self._ptr(P), self._ptr(z), self._ptr(R), self._ptr(ea))
      # the library overwrites z with the innovation (ekf_c.c:120); feature kinds return the projected part
      y = z[:-len(ea)] if (self.msckf and kind in self.feature_track_kinds) else z
      return x, P, y
    self._update = _update

  # -------------------------------------------------------------------- state ---
  def init_state(self, state, covs, filter_time):
    self.x = np.array(np.asarray(state).reshape((-1, 1)), dtype=np.float64)
    self.P = np.array(covs, dtype=np.float64)
    self.filter_time = filter_time
    self.augment_times = [0] * self.N
    self._rewind.clear()

  def reset_rewind(self):
    self._rewind.clear()

  def state(self):
    return np.array(self.x).flatten()

  def covs(self):
    return self.P

  def set_filter_time(self, t):
    self.filter_time = t

  def get_filter_time(self):
    return self.filter_time

  def get_augment_times(self):
    return self.augment_times

  def set_global(self, global_var, val):
    self.set_globals[global_var](val)

  def normalize_slice(self, start, end_ex):
    self.x[start:end_ex] /= np.linalg.norm(self.x[start:end_ex])

  def normalize_quaternions(self):
    for idx in self.quaternion_idxs:
      self.normalize_slice(idx, idx + 4)

  # rewind-buffer views kept for callers that poke at them like the reference's lists
  rewind_t = property(lambda self: self._rewind.t)
  rewind_states = property(lambda self: self._rewind.states)
  rewind_obscache = property(lambda self: self._rewind.obs)

  def augment(self):
    """Shift the clone window by one: oldest clone out, copy of the first dim_augment main states in."""
    assert self.msckf
    d1, d2, d3, d4 = self.dim_main, self.dim_main_err, self.dim_augment, self.dim_augment_err
    self.x[d1:-d3] = self.x[d1 + d3:]
    self.x[-d3:] = self.x[:d3]
    assert self.x.shape == (self.dim_x, 1) and self.P.shape == (self.dim_err, self.dim_err)
    # covariance: delete the oldest clone's rows/cols, then append rows/cols of the cloned main block.
    # Equivalent to the reference's T P_reduced T^T with a selection matrix T (ekf_sym.py:381-388).
    keep = np.r_[0:d2, d2 + d4:self.dim_err]
    src = np.r_[keep, 0:d4]
    self.P = np.ascontiguousarray(self.P[np.ix_(src, src)])
    self.augment_times = self.augment_times[1:] + [self.filter_time]

  # ------------------------------------------------------------------- rewind ---
  def rewind(self, t):
    (self.filter_time, x, P), replay = self._rewind.rewind_to(t)
    self.x[:] = x
    self.P[:] = P
    return replay

  def checkpoint(self, obs):
    self._rewind.push(self.filter_time, self.x, self.P, obs)

  # ------------------------------------------------------------------ filtering ---
  def predict(self, t):
    if self.filter_time is None:
      self.filter_time = t
    dt = t - self.filter_time
    assert dt >= 0
    self.x, self.P = self._predict(self.x, self.P, dt)
    self.normalize_quaternions()
    self.filter_time = t

  def predict_and_update_batch(self, t, kind, z, R, extra_args=[[]], augment=False):  # pylint: disable=dangerous-default-value
    replay = []
    if self.filter_time is not None and t < self.filter_time:
      if self._rewind.too_old(t, self.max_rewind_age):
        self.logger.error(f"observation too old at {t:.3f} with filter at {self.filter_time:.3f}, ignoring")
        return None
      replay = self.rewind(t)
    ret = self._predict_and_update_batch(t, kind, z, R, extra_args, augment)
    for late in replay:  # fast-forward through what was rewound over
      self._predict_and_update_batch(*late)
    return ret

  def _predict_and_update_batch(self, t, kind, z, R, extra_args, augment=False):
    """Predict to time t, then apply the n observations z[n, dim_z] (noise R[n, dim_z, dim_z]) of one kind."""
    assert z.shape[0] == R.shape[0] and z.shape[1] == R.shape[1] == R.shape[2]
    if self.filter_time is None:
      self.filter_time = t
    dt = t - self.filter_time
    assert dt >= 0
    self.x, self.P = self._predict(self.x, self.P, dt)
    self.filter_time = t
    xk_km1, Pk_km1 = np.copy(self.x).flatten(), np.copy(self.P)

    y = []
    for i in range(len(z)):
      # user data: canonicalise to fresh contiguous float64 buffers (the library writes into z_i)
      z_i = np.array(z[i], dtype=np.float64, order='F')
      R_i = np.array(R[i], dtype=np.float64, order='F')
      ea_i = np.array(extra_args[i], dtype=np.float64, order='F')
      self.x, self.P, y_i = self._update(self.x, self.P, kind, z_i, R_i, extra_args=ea_i)
      self.normalize_quaternions()
      y.append(y_i)
    xk_k, Pk_k = np.copy(self.x).flatten(), np.copy(self.P)

    if augment:
      self.augment()
    self.checkpoint((t, kind, z, R, extra_args))
    return xk_km1, xk_k, Pk_km1, Pk_k, t, kind, y, z, extra_args

  # ------------------------------------------------------------------ queries ---
  def maha_test(self, x, P, kind, z, R, extra_args=[], maha_thresh=0.95):  # pylint: disable=dangerous-default-value
    """True if the observation passes the chi2 gate at `maha_thresh` (ekf_sym.py:626-649)."""
    z = np.asarray(z, dtype=np.float64).reshape((-1, 1))
    ea = np.ascontiguousarray(extra_args, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    h = np.zeros(z.shape)
    H = np.zeros((z.shape[0], self.dim_x))
    self.hs[kind](x, ea, h)
    self.Hs[kind](x, ea, H)
    y = z - h
    H_mod = np.zeros((x.shape[0], P.shape[0]))
    self.H_mod(x, H_mod)
    H = H @ H_mod
    S_inv = np.linalg.inv(H @ P @ H.T + R)
    return not bool((y.T @ S_inv @ y).item() > chi2_ppf(maha_thresh, y.shape[0]))

  def rts_smooth(self, estimates, norm_quats=False):
    """Rauch-Tung-Striebel backward pass over the tuples returned by predict_and_update_batch.

    Follows ekf_sym.py:651-690 including its quirks: starts from the PREDICTED last state, works
    in place on the arrays inside `estimates`, smooths only the main block, and normalises the
    hard-coded quaternion slice 3:7 when `norm_quats`.
    """
    xk_n, Pk_n = estimates[-1][0], estimates[-1][2]
    Fk_1 = np.zeros(Pk_n.shape, dtype=np.float64)
    d1, d2 = self.dim_main, self.dim_main_err
    xs, Ps = [xk_n], [Pk_n]
    for k in range(len(estimates) - 2, -1, -1):
      xk1_n, Pk1_n = xk_n, Pk_n
      if norm_quats:
        xk1_n[3:7] /= np.linalg.norm(xk1_n[3:7])
      xk1_k, _, Pk1_k, _, t2 = estimates[k + 1][:5]
      _, xk_k, _, Pk_k, t1 = estimates[k][:5]
      self.F(xk_k, t2 - t1, Fk_1)
      Ck = np.linalg.solve(Pk1_k[:d2, :d2], Fk_1[:d2, :d2].dot(Pk_k[:d2, :d2].T)).T
      xk_n = xk_k
      delta_x = np.zeros((Pk_n.shape[0], 1), dtype=np.float64)
      self.inv_err_function(xk1_k, xk1_n, delta_x)
      delta_x[:d2] = Ck.dot(delta_x[:d2])
      x_new = np.zeros((xk_n.shape[0], 1), dtype=np.float64)
      self.err_function(xk_k, delta_x, x_new)
      xk_n[:d1] = x_new[:d1, 0]
      Pk_n = Pk_k
      Pk_n[:d2, :d2] = Pk_k[:d2, :d2] + Ck.dot(Pk1_n[:d2, :d2] - Pk1_k[:d2, :d2]).dot(Ck.T)
      xs.append(xk_n)
      Ps.append(Pk_n)
    return np.flipud(np.vstack(xs)), np.stack(Ps, 0)[::-1]
