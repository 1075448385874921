"""Chi-square quantiles used for Mahalanobis gating thresholds.

Behavioural mirror of the reference lookup (rednose/helpers/chi2_lookup.py:15-18):
the quantile is a *linear interpolation* over the probability grid 0.01 .. 0.98
(step 0.01) of exact chi2 quantiles, not the exact ppf.  The reference ships the
grid as a 200x98 .npy table; here each row is recomputed on demand with scipy
and memoised, which yields the same float64 values (checked in
tests/test_support_cpu.py against the reference's shipped table).
"""
from functools import lru_cache

import numpy as np

_P_GRID = np.arange(.01, .99, .01)


@lru_cache(maxsize=None)
def _row(dim: int) -> np.ndarray:
  from scipy.stats import chi2
  if dim <= 0:
    return np.zeros_like(_P_GRID)
  return chi2.ppf(_P_GRID, dim)


def chi2_ppf(p, dim):
  return np.interp(p, _P_GRID, _row(int(dim)))
