"""Filter definitions shipped with the engine (the workloads BASELINE.json names).

They are written against the same public surface user code uses (``gen_code``,
``EKF_sym_pyx``, ``KalmanFilter``) and are mathematically identical to the reference's
``examples/kinematic_kf.py`` / ``examples/live_kf.py``; tests/test_dropin.py checks that the
reference's unmodified example files generate the same model through this package.
"""
import os

from rednose_b200.build import GENERATED_DIR


def ensure_generated(filter_cls, folder=None, force=False):
  """Generate + compile the filter's library if it is missing or stale; return the folder."""
  from rednose_b200 import build
  folder = folder or GENERATED_DIR
  name = filter_cls.name
  lib = os.path.join(folder, f"lib{name}.so")
  src = os.path.join(folder, f"{name}.cu")
  if force or not os.path.exists(src) or not os.path.exists(os.path.join(folder, f"{name}.h")):
    filter_cls.generate_code(folder)
  elif not os.path.exists(lib) or not build._newer(lib, [src] + build.csrc_sources()):
    build.compile_filter(folder, name)
  return folder
