from rednose_b200.ekf_sym_pyx import EKF_sym_pyx  # noqa: F401
