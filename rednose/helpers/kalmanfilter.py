from rednose_b200.filter_base import KalmanFilter  # noqa: F401
