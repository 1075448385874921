"""Error-state localisation filter (quaternion attitude): DIM 23 / EDIM 22, 8 observation kinds.

Same model as the reference example (examples/live_kf.py:72-246); the layout tables below
replace its slice classes but produce identical symbolic expressions:

  state        pos(3) quat(4) vel(3) omega(3) gyro_bias(3) odo_scale(1) accel(3) imu_angles(3)
  error state  pos(3) rot(3)  vel(3) omega(3) gyro_bias(3) odo_scale(1) accel(3) imu_angles(3)

Process model (live_kf.py:148-168): Euler step of  pos' = vel,  quat' = 1/2 Omega(omega) quat,
vel' = R(quat) accel.  Error dynamics (:170-183), H_mod (:185-190), inject / invert (:192-211),
observation models (:219-244).
"""
import sys

import numpy as np

from rednose_b200.loader import KalmanError

EARTH_GM = 3.986005e14  # m^3/s^2


class ObservationKind:
  UNKNOWN = 0
  NO_OBSERVATION = 1
  GPS_NED = 2
  ODOMETRIC_SPEED = 3
  PHONE_GYRO = 4
  GPS_VEL = 5
  PSEUDORANGE_GPS = 6
  PSEUDORANGE_RATE_GPS = 7
  SPEED = 8
  NO_ROT = 9
  PHONE_ACCEL = 10
  ORB_POINT = 11
  ECEF_POS = 12
  CAMERA_ODO_TRANSLATION = 13
  CAMERA_ODO_ROTATION = 14
  ORB_FEATURES = 15
  MSCKF_TEST = 16
  FEATURE_TRACK_TEST = 17
  LANE_PT = 18
  IMU_FRAME = 19
  PSEUDORANGE_GLONASS = 20
  PSEUDORANGE_RATE_GLONASS = 21
  PSEUDORANGE = 22
  PSEUDORANGE_RATE = 23


# (name, width in the state, width in the error state)
_LAYOUT = [('ECEF_POS', 3, 3), ('ECEF_ORIENTATION', 4, 3), ('ECEF_VELOCITY', 3, 3), ('ANGULAR_VELOCITY', 3, 3),
           ('GYRO_BIAS', 3, 3), ('ODO_SCALE', 1, 1), ('ACCELERATION', 3, 3), ('IMU_OFFSET', 3, 3)]


class States:
  pass


_o = _e = 0
for _n, _w, _we in _LAYOUT:
  setattr(States, _n, slice(_o, _o + _w))
  setattr(States, _n + '_ERR', slice(_e, _e + _we))
  _o, _e = _o + _w, _e + _we
DIM_STATE, DIM_STATE_ERR = _o, _e


class LiveKalman:
  name = 'live'

  initial_x = np.array([-2.7e6, 4.2e6, 3.8e6,   1, 0, 0, 0,   0, 0, 0,   0, 0, 0,   0, 0, 0,   1,   0, 0, 0,   0, 0, 0], dtype=float)
  initial_P_diag = np.concatenate([[10000.0**2] * 3, [10.0**2] * 3, [10.0**2] * 3, [1.0**2] * 3,
                                   [0.05**2] * 3, [0.02**2], [1.0**2] * 3, [0.01**2] * 3])
  Q = np.diag(np.concatenate([[0.03**2] * 3, [0.0] * 3, [0.0] * 3, [0.1**2] * 3, [(0.005 / 100)**2] * 3,
                              [(0.02 / 100)**2], [3.0**2] * 3, [(0.05 / 60)**2] * 3]))

  obs_noise_diag = {ObservationKind.ODOMETRIC_SPEED: [0.2**2],
                    ObservationKind.PHONE_GYRO: [0.025**2] * 3,
                    ObservationKind.PHONE_ACCEL: [0.5**2] * 3,
                    ObservationKind.CAMERA_ODO_ROTATION: [0.05**2] * 3,
                    ObservationKind.IMU_FRAME: [0.05**2] * 3,
                    ObservationKind.NO_ROT: [0.00025**2] * 3,
                    ObservationKind.ECEF_POS: [5.0**2] * 3}

  @staticmethod
  def symbolic_model():
    """Return the arguments of gen_code for this model (everything but folder/name)."""
    import sympy as sp
    from rednose_b200.geometry import euler_rotate, quat_matrix_r, quat_rotate
    S = States
    n, ne = DIM_STATE, DIM_STATE_ERR
    state_sym = sp.MatrixSymbol('state', n, 1)
    st = sp.Matrix(state_sym)
    pos, q, v = st[S.ECEF_POS, :], st[S.ECEF_ORIENTATION, :], st[S.ECEF_VELOCITY, :]
    omega, gyro_bias = st[S.ANGULAR_VELOCITY, :], st[S.GYRO_BIAS, :]
    odo_scale, accel, imu_angles = st[S.ODO_SCALE, :][0, 0], st[S.ACCELERATION, :], st[S.IMU_OFFSET, :]
    dt = sp.Symbol('dt')
    R_q = quat_rotate(*q)

    # quaternion kinematics: q' = 1/2 Omega(omega) q
    wr, wp, wy = omega
    Omega = sp.Rational(1, 2) * sp.Matrix([[0, -wr, -wp, -wy], [wr, 0, wy, -wp], [wp, -wy, 0, wr], [wy, wp, -wr, 0]])
    xdot = sp.zeros(n, 1)
    xdot[S.ECEF_POS, :] = v
    xdot[S.ECEF_ORIENTATION, :] = Omega * q
    xdot[S.ECEF_VELOCITY, :] = R_q * accel
    f_sym = st + dt * xdot

    # error-state dynamics
    err_sym = sp.MatrixSymbol('state_err', ne, 1)
    er = sp.Matrix(err_sym)
    R_err = euler_rotate(*er[S.ECEF_ORIENTATION_ERR, :])
    edot = sp.zeros(ne, 1)
    edot[S.ECEF_POS_ERR, :] = er[S.ECEF_VELOCITY_ERR, :]
    edot[S.ECEF_ORIENTATION_ERR, :] = R_err * R_q * (omega + er[S.ANGULAR_VELOCITY_ERR, :])
    edot[S.ECEF_VELOCITY_ERR, :] = R_err * R_q * (accel + er[S.ACCELERATION_ERR, :])
    f_err_sym = er + dt * edot

    # d state / d error-state
    H_mod = sp.zeros(n, ne)
    H_mod[S.ECEF_POS, S.ECEF_POS_ERR] = sp.eye(3)
    H_mod[S.ECEF_ORIENTATION, S.ECEF_ORIENTATION_ERR] = sp.Rational(1, 2) * quat_matrix_r(q)[:, 1:]
    H_mod[S.ECEF_ORIENTATION.stop:, S.ECEF_ORIENTATION_ERR.stop:] = sp.eye(n - S.ECEF_ORIENTATION.stop)

    # inject: true = nom [+] delta ; invert: delta = true [-] nom
    nom_x = sp.MatrixSymbol('nom_x', n, 1)
    true_x = sp.MatrixSymbol('true_x', n, 1)
    delta_x = sp.MatrixSymbol('delta_x', ne, 1)
    nom, tru, dl = sp.Matrix(nom_x), sp.Matrix(true_x), sp.Matrix(delta_x)
    inject = sp.zeros(n, 1)
    inject[S.ECEF_POS, :] = nom[S.ECEF_POS, :] + dl[S.ECEF_POS_ERR, :]
    dq = sp.Matrix([1] + list(sp.Rational(1, 2) * dl[S.ECEF_ORIENTATION_ERR, :]))
    inject[S.ECEF_ORIENTATION, :] = quat_matrix_r(nom[S.ECEF_ORIENTATION, :]) * dq
    inject[S.ECEF_ORIENTATION.stop:, :] = nom[S.ECEF_ORIENTATION.stop:, :] + dl[S.ECEF_ORIENTATION_ERR.stop:, :]
    invert = sp.zeros(ne, 1)
    invert[S.ECEF_POS_ERR, :] = tru[S.ECEF_POS, :] - nom[S.ECEF_POS, :]
    dq_inv = quat_matrix_r(nom[S.ECEF_ORIENTATION, :]).T * tru[S.ECEF_ORIENTATION, :]
    invert[S.ECEF_ORIENTATION_ERR, :] = 2 * dq_inv[1:, :]
    invert[S.ECEF_ORIENTATION_ERR.stop:, :] = tru[S.ECEF_ORIENTATION.stop:, :] - nom[S.ECEF_ORIENTATION.stop:, :]
    eskf_params = [[inject, nom_x, delta_x], [invert, nom_x, true_x], H_mod, f_err_sym, err_sym]

    # observation models
    R_imu = euler_rotate(*imu_angles)
    r2 = pos[0]**2 + pos[1]**2 + pos[2]**2
    gravity = R_q.T * ((EARTH_GM / (r2**sp.Rational(3, 2))) * pos)
    speed = sp.sqrt(v[0]**2 + v[1]**2 + v[2]**2)
    K = ObservationKind
    obs_eqs = [[sp.Matrix([speed * odo_scale]), K.ODOMETRIC_SPEED, None],
               [R_imu * (omega + gyro_bias), K.PHONE_GYRO, None],
               [sp.Matrix(omega), K.NO_ROT, None],
               [R_imu * (gravity + accel), K.PHONE_ACCEL, None],
               [sp.Matrix(pos), K.ECEF_POS, None],
               [R_q.T * v, K.CAMERA_ODO_TRANSLATION, None],
               [sp.Matrix(omega), K.CAMERA_ODO_ROTATION, None],
               [sp.Matrix(imu_angles), K.IMU_FRAME, None]]
    return dict(f_sym=f_sym, dt_sym=dt, x_sym=state_sym, obs_eqs=obs_eqs, dim_x=n, dim_err=ne, eskf_params=eskf_params)

  @staticmethod
  def generate_code(generated_dir, name=None):
    from rednose_b200.codegen import gen_code
    gen_code(generated_dir, name or LiveKalman.name, **LiveKalman.symbolic_model())

  def __init__(self, generated_dir, filter_cls=None):
    if filter_cls is None:
      from rednose_b200.ekf_sym_pyx import EKF_sym_pyx as filter_cls
    self.dim_state, self.dim_state_err = DIM_STATE, DIM_STATE_ERR
    self.obs_noise = {k: np.diag(d) for k, d in self.obs_noise_diag.items()}
    self.filter = filter_cls(generated_dir, self.name, self.Q, self.initial_x, np.diag(self.initial_P_diag),
                             self.dim_state, self.dim_state_err)

  x = property(lambda self: self.filter.state())
  t = property(lambda self: self.filter.filter_time)
  P = property(lambda self: self.filter.covs())

  def rts_smooth(self, estimates):
    return self.filter.rts_smooth(estimates, norm_quats=True)

  def init_state(self, state, covs_diag=None, covs=None, filter_time=None):
    if covs_diag is not None:
      covs = np.diag(covs_diag)
    elif covs is None:
      covs = self.filter.covs()
    self.filter.init_state(state, covs, filter_time)

  def get_R(self, kind, n):
    return np.broadcast_to(self.obs_noise[kind], (n,) + self.obs_noise[kind].shape).copy()

  def predict_and_observe(self, t, kind, data):
    if len(data) > 0:
      data = np.atleast_2d(data)
    K = ObservationKind
    if kind in (K.CAMERA_ODO_TRANSLATION, K.CAMERA_ODO_ROTATION):
      # columns 3: carry the per-observation standard deviations
      z = data[:, :3]
      R = np.stack([np.diag(row[3:]**2) for row in data]) if len(data) else np.zeros((0, 3, 3))
    elif kind == K.ODOMETRIC_SPEED:
      z = np.array(data)
      R = np.full((len(data), 1, 1), 0.2**2)
    else:
      z, R = data, self.get_R(kind, len(data))
    r = self.filter.predict_and_update_batch(t, kind, z, R)

    # keep the attitude quaternion normalised; refuse to continue if it degenerated
    quat_norm = np.linalg.norm(self.filter.x[States.ECEF_ORIENTATION, 0])
    if not 0.1 < quat_norm < 10:
      raise KalmanError("Kalman filter quaternions unstable")
    self.filter.x[States.ECEF_ORIENTATION, 0] = self.filter.x[States.ECEF_ORIENTATION, 0] / quat_norm
    return r


if __name__ == "__main__":
  LiveKalman.generate_code(sys.argv[2])
