"""1-D constant-velocity filter: state [position, velocity], one position observation.

Same model as the reference example (examples/kinematic_kf.py:31-69): DIM = EDIM = 2,
x0 = [0.5, 0], P0 = I, Q = diag(0.1^2, 2^2), R_position = 0.1^2, Euler step
position += dt * velocity.
"""
import sys

import numpy as np

from rednose_b200.filter_base import KalmanFilter


class ObservationKind:
  UNKNOWN = 0
  NO_OBSERVATION = 1
  POSITION = 1
  names = ['Unknown', 'No observation', 'Position']

  @classmethod
  def to_string(cls, kind):
    return cls.names[kind]


class States:
  POSITION = slice(0, 1)
  VELOCITY = slice(1, 2)


class KinematicKalman(KalmanFilter):
  name = 'kinematic'
  initial_x = np.array([0.5, 0.0])
  initial_P_diag = np.array([1.0, 1.0])
  Q = np.diag([0.1**2, 2.0**2])
  obs_noise = {ObservationKind.POSITION: np.array([[0.1**2]])}

  @staticmethod
  def generate_code(generated_dir, name=None):
    import sympy as sp
    from rednose_b200.codegen import gen_code
    n = KinematicKalman.initial_x.shape[0]
    state_sym = sp.MatrixSymbol('state', n, 1)
    dt = sp.Symbol('dt')
    pos, vel = state_sym[0, 0], state_sym[1, 0]
    f_sym = sp.Matrix([pos + dt * vel, vel])
    obs_eqs = [[sp.Matrix([pos]), ObservationKind.POSITION, None]]
    gen_code(generated_dir, name or KinematicKalman.name, f_sym, dt, state_sym, obs_eqs, n, n)

  def __init__(self, generated_dir, filter_cls=None):
    if filter_cls is None:
      from rednose_b200.ekf_sym_pyx import EKF_sym_pyx as filter_cls
    n = self.initial_x.shape[0]
    self.filter = filter_cls(generated_dir, self.name, self.Q, self.initial_x, np.diag(self.initial_P_diag), n, n)


if __name__ == "__main__":
  KinematicKalman.generate_code(sys.argv[2])
