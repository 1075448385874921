from rednose_b200.chi2 import chi2_ppf  # noqa: F401
