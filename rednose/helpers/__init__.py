from rednose_b200.loader import KalmanError, TEMPLATE_DIR, load_code, write_code  # noqa: F401
