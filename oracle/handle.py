"""ORACLE -- test infrastructure only.  cffi handle on oracle/_ref/lib<name>.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import os

import numpy as np
from cffi import FFI


class Oracle:
  """cffi handle on oracle/_ref/lib<name>.so (reference leaf C + restated ekf_c.c)."""

  def __init__(self, folder, name):
    self.name = name
    with open(os.path.join(folder, f"{name}.h"), encoding="utf-8") as f:
      protos = [ln for ln in f.read().split("\n") if ln.startswith("void ")]
    self.ffi = FFI()
    self.ffi.cdef("\n".join(protos))
    self.lib = self.ffi.dlopen(os.path.join(folder, f"lib{name}.so"))

  def p(self, a):
    return self.ffi.cast("double *", a.ctypes.data) if a is not None else self.ffi.NULL

  def batch_step(self, kind, x, P, Q, dt, z, R, ea=None, quat_idxs=(), flags=0, nthreads=8):  # noqa: E501
    """In place on copies; returns (x, P, y)."""
    x, P, z, R = (np.array(a, dtype=np.float64, order='C') for a in (x, P, z, R))
    Q = np.ascontiguousarray(Q, dtype=np.float64)
    B = x.shape[0]
    zdim = z.shape[-1]
    dt_arr = np.ascontiguousarray(dt, dtype=np.float64) if np.ndim(dt) else None
    qi = self.ffi.new("int[]", list(quat_idxs) or [0])
    ea_c = np.ascontiguousarray(ea, dtype=np.float64) if ea is not None else None
    getattr(self.lib, f"{self.name}_oracle_batch_step")(
      int(kind), self.p(x), self.p(P), self.p(Q), self.ffi.cast("const double *", dt_arr.ctypes.data) if dt_arr is not None else self.ffi.NULL,
      0.0 if dt_arr is not None else float(dt), self.p(z), self.p(R), self.p(ea_c), zdim, ea_c.shape[-1] if ea_c is not None else 0,
      B, nthreads, qi, len(quat_idxs), flags)
    return x, P, z

  def predict(self, x, P, Q, dt):
    x, P = np.array(x, dtype=np.float64), np.array(P, dtype=np.float64)
    Q = np.ascontiguousarray(Q, dtype=np.float64)
    for b in range(x.shape[0]):
      getattr(self.lib, f"{self.name}_predict")(self.p(x[b]), self.p(P[b]), self.p(Q), float(dt if np.ndim(dt) == 0 else dt[b]))
    return x, P

  def update(self, kind, x, P, z, R, ea=None):
    x, P, z, R = (np.array(a, dtype=np.float64) for a in (x, P, z, R))
    dummy = np.zeros(1)
    for b in range(x.shape[0]):
      getattr(self.lib, f"{self.name}_update_{kind}")(self.p(x[b]), self.p(P[b]), self.p(z[b]), self.p(R[b]), self.p(ea[b] if ea is not None else dummy))
    return x, P, z

  def leaf(self, fn, *args):
    getattr(self.lib, f"{self.name}_{fn}")(*[self.p(a) if isinstance(a, np.ndarray) else a for a in args])


