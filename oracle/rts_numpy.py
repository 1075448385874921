"""ORACLE -- test infrastructure only.  numpy restatement of the reference's RTS smoother
(rednose/helpers/ekf_sym.py:651-690), one filter at a time, with the leaf functions F_fun / err_fun /
inv_err_fun taken from the oracle library (reference-generated C).  Quirks kept on purpose:
starts from the predicted last state (:658-659), works in place on the inputs (:678,684-686),
normalises the hard-coded quaternion slice 3:7 of x_{k+1|N} when norm_quats (:666-667)."""
import numpy as np


def rts_smooth(oracle, x_pred, x_filt, P_pred, P_filt, t, dim_main, dim_main_err, norm_quats=False):
  """x_*: [T, DIM], P_*: [T, EDIM, EDIM], t: [T] of ONE filter.  Returns (xs [T, DIM], Ps [T, EDIM, EDIM])."""
  x_pred, x_filt, P_pred, P_filt = (np.array(a, dtype=np.float64) for a in (x_pred, x_filt, P_pred, P_filt))
  T = x_pred.shape[0]
  d1, d2 = dim_main, dim_main_err
  xk_n = x_pred[-1]
  Pk_n = P_pred[-1]
  Fk_1 = np.zeros(Pk_n.shape)
  xs, Ps = [xk_n], [Pk_n]
  for k in range(T - 2, -1, -1):
    xk1_n = xk_n
    if norm_quats:
      xk1_n[3:7] /= np.linalg.norm(xk1_n[3:7])
    Pk1_n = Pk_n
    xk1_k, Pk1_k = x_pred[k + 1], P_pred[k + 1]
    xk_k, Pk_k = x_filt[k], P_filt[k]
    dt = t[k + 1] - t[k]
    oracle.leaf("F_fun", np.ascontiguousarray(xk_k), float(dt), Fk_1)
    Ck = np.linalg.solve(Pk1_k[:d2, :d2], Fk_1[:d2, :d2].dot(Pk_k[:d2, :d2].T)).T
    xk_n = xk_k
    delta_x = np.zeros(Pk_n.shape[0])
    oracle.leaf("inv_err_fun", np.ascontiguousarray(xk1_k), np.ascontiguousarray(xk1_n), delta_x)
    delta_x[:d2] = Ck.dot(delta_x[:d2])
    x_new = np.zeros(xk_n.shape[0])
    oracle.leaf("err_fun", np.ascontiguousarray(xk_k), delta_x, x_new)
    xk_n[:d1] = x_new[:d1]
    Pk_n = Pk_k
    Pk_n[:d2, :d2] = Pk_k[:d2, :d2] + Ck.dot(Pk1_n[:d2, :d2] - Pk1_k[:d2, :d2]).dot(Ck.T)
    xs.append(xk_n)
    Ps.append(Pk_n)
  return np.flipud(np.vstack(xs)), np.stack(Ps, 0)[::-1]
