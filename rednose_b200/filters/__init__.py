"""Filter definitions shipped with the engine (the workloads BASELINE.json names).

They are written against the same public surface user code uses (``gen_code``,
``EKF_sym_pyx``, ``KalmanFilter``) and are mathematically identical to the reference's
``examples/kinematic_kf.py`` / ``examples/live_kf.py``; tests/test_dropin.py checks that the
reference's unmodified example files generate the same model through this package.
"""
import os

from rednose_b200.build import GENERATED_DIR


def _generator_newer_than(src, filter_cls):
  """True if the code generator or the filter definition changed after `src` was written (stale generated source)."""
  import inspect
  import rednose_b200.codegen as cg
  import rednose_b200.codegen.symbolic as sym
  deps = [cg.__file__, sym.__file__]
  for cls in inspect.getmro(filter_cls):
    try:
      deps.append(inspect.getsourcefile(cls))
    except TypeError:
      pass
  if filter_cls.__module__.endswith(".msckf"):
    import rednose_b200.filters.live as live
    deps.append(live.__file__)
  t = os.path.getmtime(src)
  return any(d and os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def ensure_generated(filter_cls, folder=None, force=False):
  """Generate + compile the filter's library if it is missing or stale; return the folder."""
  from rednose_b200 import build
  folder = folder or GENERATED_DIR
  name = filter_cls.name
  lib = os.path.join(folder, f"lib{name}.so")
  src = os.path.join(folder, f"{name}.cu")
  if force or not os.path.exists(src) or not os.path.exists(os.path.join(folder, f"{name}.h")) or _generator_newer_than(src, filter_cls):
    filter_cls.generate_code(folder)
  elif not os.path.exists(lib) or not build._newer(lib, [src] + build.csrc_sources()):
    build.compile_filter(folder, name)
  return folder
