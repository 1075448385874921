from rednose_b200.kalmanfilter import KalmanFilter  # noqa: F401
