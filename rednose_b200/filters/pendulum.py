"""Small test filter exercising the generator features the shipped examples do not use:
``global_vars`` (runtime scalars settable with ``<name>_set_<var>``, rednose/helpers/ekf_sym.py:129-132,166-171),
``extra_routines`` (additional exported sympy functions, ekf_sym.py:94-95) and an observation kind with
extra arguments but no null-space projection, Mahalanobis gated.

State [theta, omega, bias] of a damped pendulum: theta' = omega, omega' = -(g / L) sin(theta) - c omega,
with the gravity ``g`` and damping ``c`` as global variables.  Kind 1: angle + bias (m = 1).  Kind 2: position of
the bob relative to a pivot given as extra_args (m = 2), gated.
"""
import sys

import numpy as np


class PendulumKalman:
  name = 'pendulum'
  initial_x = np.array([0.3, 0.0, 0.01])
  initial_P_diag = np.array([0.1**2, 0.5**2, 0.05**2])
  Q = np.diag([1e-4, 1e-2, 1e-6])
  LENGTH = 1.5
  global_var_names = ['grav', 'damp']

  @staticmethod
  def symbolic_model():
    import sympy as sp
    state_sym = sp.MatrixSymbol('state', 3, 1)
    theta, omega, bias = state_sym[0, 0], state_sym[1, 0], state_sym[2, 0]
    dt = sp.Symbol('dt')
    grav, damp = sp.Symbol('grav'), sp.Symbol('damp')
    L = PendulumKalman.LENGTH
    f_sym = sp.Matrix([theta + dt * omega, omega + dt * (-(grav / L) * sp.sin(theta) - damp * omega), bias])
    pivot = sp.MatrixSymbol('pivot', 2, 1)
    obs_eqs = [[sp.Matrix([theta + bias]), 1, None],
               [sp.Matrix([pivot[0, 0] + L * sp.sin(theta), pivot[1, 0] - L * sp.cos(theta)]), 2, pivot]]
    energy = sp.Matrix([sp.Rational(1, 2) * (L * omega)**2 + grav * L * (1 - sp.cos(theta))])
    extra_routines = [('energy', energy, [state_sym])]
    return dict(f_sym=f_sym, dt_sym=dt, x_sym=state_sym, obs_eqs=obs_eqs, dim_x=3, dim_err=3,
                maha_test_kinds=[2], global_vars=[grav, damp], extra_routines=extra_routines)

  @staticmethod
  def generate_code(generated_dir, name=None):
    from rednose_b200.codegen import gen_code
    gen_code(generated_dir, name or PendulumKalman.name, **PendulumKalman.symbolic_model())


if __name__ == "__main__":
  PendulumKalman.generate_code(sys.argv[2])
