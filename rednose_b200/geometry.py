"""Rotation helpers that filter definitions import (surface of
rednose/helpers/sympy_helpers.py:1-120; live_kf.py:9 uses euler_rotate,
quat_matrix_r, quat_rotate).  Symbolic builders return sympy matrices, numeric
ones numpy arrays.  Conventions (kept identical to the reference so generated
models agree): quaternions are [w, x, y, z]; eulers are (roll, pitch, yaw) with
R = Rz(yaw) Ry(pitch) Rx(roll); the symbolic quat_rotate(q) and the numeric
quat2rot(q) are the same matrix (tests/test_support_cpu.py, which also checks the
symbolic builders against the reference's own helpers where it is mounted).
"""
import numpy as np
import sympy as sp


# ---------------------------------------------------------------- numeric ---
def quat2rot(quats):
  q = np.asarray(quats, dtype=float)
  single = q.ndim < 2
  q = np.atleast_2d(q)
  w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
  R = np.empty((q.shape[0], 3, 3))
  R[:, 0, 0] = w * w + x * x - y * y - z * z
  R[:, 1, 1] = w * w - x * x + y * y - z * z
  R[:, 2, 2] = w * w - x * x - y * y + z * z
  R[:, 0, 1], R[:, 1, 0] = 2 * (x * y - w * z), 2 * (x * y + w * z)
  R[:, 0, 2], R[:, 2, 0] = 2 * (w * y + x * z), 2 * (x * z - w * y)
  R[:, 1, 2], R[:, 2, 1] = 2 * (y * z - w * x), 2 * (w * x + y * z)
  return R[0] if single else R


rotations_from_quats = quat2rot


def euler2quat(eulers):
  e = np.asarray(eulers, dtype=float)
  out_shape = (-1, 4) if e.ndim > 1 else (4,)
  e = np.atleast_2d(e)
  cg, sg = np.cos(e[:, 0] / 2), np.sin(e[:, 0] / 2)
  ct, st = np.cos(e[:, 1] / 2), np.sin(e[:, 1] / 2)
  cp, sp_ = np.cos(e[:, 2] / 2), np.sin(e[:, 2] / 2)
  q = np.stack([cg * ct * cp + sg * st * sp_,
                sg * ct * cp - cg * st * sp_,
                cg * st * cp + sg * ct * sp_,
                cg * ct * sp_ - sg * st * cp], axis=1)
  q[q[:, 0] < 0] *= -1  # canonical sign: w >= 0
  return q.reshape(out_shape)


def euler2rot(eulers):
  return quat2rot(euler2quat(eulers))


def rot_matrix(roll, pitch, yaw):
  def axis_rot(a, i, j):
    m = np.eye(3)
    m[i, i] = m[j, j] = np.cos(a)
    m[i, j], m[j, i] = -np.sin(a), np.sin(a)
    return m
  # Rx uses (1,2), Ry uses (2,0) so that the sign of sin lands as [c 0 s; 0 1 0; -s 0 c]
  return axis_rot(yaw, 0, 1) @ axis_rot(pitch, 2, 0) @ axis_rot(roll, 1, 2)


# --------------------------------------------------------------- symbolic ---
def cross(v):
  return sp.Matrix([[0, -v[2], v[1]],
                    [v[2], 0, -v[0]],
                    [-v[1], v[0], 0]])


def rot_to_euler(R):
  return sp.Matrix([sp.atan2(R[2, 1], R[2, 2]), sp.asin(-R[2, 0]), sp.atan2(R[1, 0], R[0, 0])])


def _sym_axis_rot(a, i, j):
  m = sp.eye(3)
  m[i, i] = m[j, j] = sp.cos(a)
  m[i, j], m[j, i] = -sp.sin(a), sp.sin(a)
  return m


def euler_rotate(roll, pitch, yaw):
  return _sym_axis_rot(yaw, 0, 1) * _sym_axis_rot(pitch, 2, 0) * _sym_axis_rot(roll, 1, 2)


def quat_rotate(q0, q1, q2, q3):
  sq = [q0**2, q1**2, q2**2, q3**2]
  body_to_world_T = sp.Matrix([
    [sq[0] + sq[1] - sq[2] - sq[3], 2 * (q1 * q2 + q0 * q3), 2 * (q1 * q3 - q0 * q2)],
    [2 * (q1 * q2 - q0 * q3), sq[0] - sq[1] + sq[2] - sq[3], 2 * (q2 * q3 + q0 * q1)],
    [2 * (q1 * q3 + q0 * q2), 2 * (q2 * q3 - q0 * q1), sq[0] - sq[1] - sq[2] + sq[3]]])
  return body_to_world_T.T


def _quat_matrix(p, sign):
  # left (sign=+1) / right (sign=-1) Hamilton product matrices
  s = sign
  return sp.Matrix([[p[0], -p[1], -p[2], -p[3]],
                    [p[1], p[0], -s * p[3], s * p[2]],
                    [p[2], s * p[3], p[0], -s * p[1]],
                    [p[3], -s * p[2], s * p[1], p[0]]])


def quat_matrix_l(p):
  return _quat_matrix(p, 1)


def quat_matrix_r(p):
  return _quat_matrix(p, -1)
