from rednose_b200.ekf_sym import EKF_sym, gen_code, null, solve  # noqa: F401
