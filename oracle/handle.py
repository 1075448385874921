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




class Arena:
  """Persistent, in-place state of B filters on the host for the benchmark's reference arm: one allocation first-touched
  and stepped by a pool of pinned worker threads (oracle/batch_runner.inc); a step copies no arrays."""

  def __init__(self, oracle, B, nthreads=None, pin=True):
    self.o, self.B = oracle, int(B)
    self.nthreads = int(nthreads or len(os.sched_getaffinity(0)))
    self._f = lambda s: getattr(oracle.lib, f"{oracle.name}_oracle_arena_{s}")
    self.h = self._f("create")(self.B, self.nthreads, 1 if pin else 0)
    if self.h == oracle.ffi.NULL:
      raise MemoryError("oracle arena")
    info = oracle.ffi.new("long long[4]")
    self._f("info")(self.h, info)
    self.pinned, self.bytes_resident = int(info[2]), int(info[3])
    self._keep = []

  def load(self, x, P):
    """x [B, DIM] or [DIM] (broadcast); P [B, EDIM, EDIM] or [EDIM, EDIM] (broadcast).  Copied once, by the workers."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    P = np.ascontiguousarray(P, dtype=np.float64)
    self.dim, self.edim = x.shape[-1], P.shape[-1]
    self._f("load")(self.h, self.o.p(x), self.dim if x.ndim == 2 else 0, self.o.p(P), self.edim * self.edim if P.ndim == 3 else 0)

  def step(self, kind, Q, dt, z, R, ea=None, y=None, quat_idxs=(), flags=0):
    """One in-place predict + update_<kind> over every filter.  z [B, zdim], R [B, zdim, zdim] or [zdim, zdim] must be
    C-contiguous float64 (no conversion copies are made here); y, if given, receives the innovations."""
    ffi = self.o.ffi
    for a in (z, R, Q):
      assert a.dtype == np.float64 and a.flags.c_contiguous
    zdim = z.shape[-1]
    qi = ffi.new("int[]", list(quat_idxs) or [0])
    dt_arr = dt if np.ndim(dt) else None
    cp = lambda a: ffi.cast("const double *", a.ctypes.data) if a is not None else ffi.NULL
    self._f("step")(self.h, int(kind), cp(Q), cp(dt_arr), 0.0 if dt_arr is not None else float(dt), cp(z), self.o.p(y),
                    cp(R), zdim * zdim if R.ndim == 3 else 0, cp(ea), zdim, ea.shape[-1] if ea is not None else 0, qi, len(quat_idxs), flags)

  def read(self, b0=0, b1=None):
    b1 = self.B if b1 is None else b1
    x = np.empty((b1 - b0, self.dim)); P = np.empty((b1 - b0, self.edim, self.edim))
    self._f("read")(self.h, b0, b1, self.o.p(x), self.o.p(P))
    return x, P

  def close(self):
    if self.h is not None:
      self._f("destroy")(self.h)
      self.h = None

  def __del__(self):
    try:
      self.close()
    except Exception:  # pylint: disable=broad-except
      pass
