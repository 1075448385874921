"""Synthetic MSCKF: the live_kf main state augmented with N = 10 cloned camera poses.

BASELINE.json config 5 ("MSCKF augmented state, 10 cloned camera poses, n ~ 90").  The reference ships
no MSCKF example; this one follows what its generator and driver assume (rednose/helpers/ekf_sym.py:57-66,
365-391): a clone is the FIRST dim_augment = 7 main states (ECEF position + attitude quaternion, 6 error
states), clones sit behind the main state, and the feature-track kind takes the 3-D point as extra_args and is
projected on the left null space of He = dh/d(point).

  DIM  = 23 + 10 * 7 = 93      EDIM = 22 + 10 * 6 = 82      MEDIM = 22
  kind 17 FEATURE_TRACK_TEST: z = 10 x (u, v) normalised image coordinates of one point seen from the 10
  clones (ZDIM 20, EADIM 3, 17 after projection), Mahalanobis gated.
"""
import sys

import numpy as np

from rednose_b200.filters.live import DIM_STATE, DIM_STATE_ERR, LiveKalman, ObservationKind

N_CLONES = 10
DIM_AUGMENT, DIM_AUGMENT_ERR = 7, 6
DIM = DIM_STATE + N_CLONES * DIM_AUGMENT
EDIM = DIM_STATE_ERR + N_CLONES * DIM_AUGMENT_ERR


class MsckfKalman:
  name = 'msckf'
  N = N_CLONES
  feature_kind = ObservationKind.FEATURE_TRACK_TEST

  initial_x = np.concatenate([LiveKalman.initial_x] + [LiveKalman.initial_x[:DIM_AUGMENT]] * N_CLONES)
  initial_P_diag = np.concatenate([LiveKalman.initial_P_diag] + [LiveKalman.initial_P_diag[:DIM_AUGMENT_ERR]] * N_CLONES)
  Q = np.diag(np.concatenate([np.diag(LiveKalman.Q), np.zeros(N_CLONES * DIM_AUGMENT_ERR)]))

  @staticmethod
  def symbolic_model():
    import sympy as sp
    from rednose_b200.geometry import quat_matrix_r, quat_rotate
    main = LiveKalman.symbolic_model()
    n, ne = DIM, EDIM
    state_sym = sp.MatrixSymbol('state', n, 1)
    st = sp.Matrix(state_sym)
    err_sym = sp.MatrixSymbol('state_err', ne, 1)
    er = sp.Matrix(err_sym)
    nom_x, true_x = sp.MatrixSymbol('nom_x', n, 1), sp.MatrixSymbol('true_x', n, 1)
    delta_x = sp.MatrixSymbol('delta_x', ne, 1)
    nom, tru, dl = sp.Matrix(nom_x), sp.Matrix(true_x), sp.Matrix(delta_x)

    def lift(expr, pairs):
      """Re-express a main-model expression on the augmented symbols (same leading indices)."""
      rep = {}
      for old, new in pairs:
        for i in range(old.shape[0]):
          rep[old[i, 0]] = new[i, 0]
      return expr.xreplace(rep)

    m_state, m_err = main['x_sym'], main['eskf_params'][4]
    m_inject, m_nom, m_delta = main['eskf_params'][0]
    m_invert, _, m_true = main['eskf_params'][1]
    pairs_f = [(m_state, state_sym), (m_err, err_sym)]

    f_sym = sp.Matrix(st)  # clones are static
    f_sym[:DIM_STATE, :] = lift(sp.Matrix(main['f_sym']), pairs_f)
    f_err_sym = sp.Matrix(er)
    f_err_sym[:DIM_STATE_ERR, :] = lift(sp.Matrix(main['eskf_params'][3]), pairs_f)

    H_mod = sp.zeros(n, ne)
    H_mod[:DIM_STATE, :DIM_STATE_ERR] = lift(sp.Matrix(main['eskf_params'][2]), pairs_f)
    inject = sp.zeros(n, 1)
    inject[:DIM_STATE, :] = lift(sp.Matrix(m_inject), [(m_nom, nom_x), (m_delta, delta_x)])
    invert = sp.zeros(ne, 1)
    invert[:DIM_STATE_ERR, :] = lift(sp.Matrix(m_invert), [(m_nom, nom_x), (m_true, true_x)])
    for c in range(N_CLONES):
      o, oe = DIM_STATE + c * DIM_AUGMENT, DIM_STATE_ERR + c * DIM_AUGMENT_ERR
      q = st[o + 3:o + 7, :]
      H_mod[o:o + 3, oe:oe + 3] = sp.eye(3)
      H_mod[o + 3:o + 7, oe + 3:oe + 6] = sp.Rational(1, 2) * quat_matrix_r(q)[:, 1:]
      inject[o:o + 3, :] = nom[o:o + 3, :] + dl[oe:oe + 3, :]
      dq = sp.Matrix([1] + list(sp.Rational(1, 2) * dl[oe + 3:oe + 6, :]))
      inject[o + 3:o + 7, :] = quat_matrix_r(nom[o + 3:o + 7, :]) * dq
      invert[oe:oe + 3, :] = tru[o:o + 3, :] - nom[o:o + 3, :]
      invert[oe + 3:oe + 6, :] = 2 * (quat_matrix_r(nom[o + 3:o + 7, :]).T * tru[o + 3:o + 7, :])[1:, :]
    eskf_params = [[inject, nom_x, delta_x], [invert, nom_x, true_x], H_mod, f_err_sym, err_sym]

    # observations: the live kinds on the main state + the feature track through all clones
    obs_eqs = [[lift(sp.Matrix(h), pairs_f), kind, ea] for h, kind, ea in main['obs_eqs']]
    point = sp.MatrixSymbol('point', 3, 1)
    rows = []
    for c in range(N_CLONES):
      o = DIM_STATE + c * DIM_AUGMENT
      pc = quat_rotate(*st[o + 3:o + 7, :]).T * (sp.Matrix(point) - st[o:o + 3, :])  # point in the clone's device frame
      rows += [pc[1] / pc[0], pc[2] / pc[0]]
    obs_eqs.append([sp.Matrix(rows), MsckfKalman.feature_kind, point])
    msckf_params = [DIM_STATE, DIM_AUGMENT, DIM_STATE_ERR, DIM_AUGMENT_ERR, N_CLONES, [MsckfKalman.feature_kind]]
    return dict(f_sym=f_sym, dt_sym=main['dt_sym'], x_sym=state_sym, obs_eqs=obs_eqs, dim_x=n, dim_err=ne,
                eskf_params=eskf_params, msckf_params=msckf_params, maha_test_kinds=[MsckfKalman.feature_kind])

  @staticmethod
  def generate_code(generated_dir, name=None):
    from rednose_b200.codegen import gen_code
    gen_code(generated_dir, name or MsckfKalman.name, **MsckfKalman.symbolic_model())

  def __init__(self, generated_dir, filter_cls=None):
    if filter_cls is None:
      from rednose_b200.ekf_sym_pyx import EKF_sym_pyx as filter_cls
    self.filter = filter_cls(generated_dir, self.name, self.Q, self.initial_x, np.diag(self.initial_P_diag),
                             DIM_STATE, DIM_STATE_ERR, N=N_CLONES, dim_augment=DIM_AUGMENT, dim_augment_err=DIM_AUGMENT_ERR,
                             maha_test_kinds=[self.feature_kind], quaternion_idxs=[3] + [DIM_STATE + 3 + 7 * c for c in range(N_CLONES)])


if __name__ == "__main__":
  MsckfKalman.generate_code(sys.argv[2])
