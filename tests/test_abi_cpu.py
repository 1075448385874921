"""CPU tests (no GPU): the generated CUDA libraries build for sm_100a, load, and export every
symbol their headers declare; the plugin descriptor is readable.  No compute calls."""
import ctypes
import os
import re

import pytest


def _protos(header_path):
  with open(header_path, encoding="utf-8") as f:
    return [ln for ln in f.read().split("\n") if re.match(r"^(void|int|void\*) \w+\(", ln)]


@pytest.mark.parametrize("name", ["kinematic", "live"])
def test_library_exports_every_declared_symbol(gen_dir, name):
  lib = ctypes.CDLL(os.path.join(gen_dir, f"lib{name}.so"))
  protos = _protos(os.path.join(gen_dir, f"{name}.h"))
  assert len(protos) > 10
  for p in protos:
    sym = re.match(r"^(?:void|int|void\*) (\w+)\(", p).group(1)
    assert hasattr(lib, sym), sym


@pytest.mark.parametrize("name,dims,kinds", [("kinematic", (2, 2, 2), [1]), ("live", (23, 22, 22), [3, 4, 9, 10, 12, 13, 14, 19])])
def test_reference_symbol_set_and_descriptor(gen_dir, name, dims, kinds):
  """The reference's C symbol set (rednose/helpers/ekf_sym.py:149-171) is present under the same names."""
  lib = ctypes.CDLL(os.path.join(gen_dir, f"lib{name}.so"))
  for s in ["predict", "f_fun", "F_fun", "err_fun", "inv_err_fun", "H_mod_fun"]:
    assert hasattr(lib, f"{name}_{s}")
  for k in kinds:
    for s in ("update", "h", "H", "batch_update", "batch_step", "host_step"):
      assert hasattr(lib, f"{name}_{s}_{k}")

  class Desc(ctypes.Structure):
    _fields_ = [("abi_version", ctypes.c_int), ("name", ctypes.c_char_p), ("dim", ctypes.c_int), ("edim", ctypes.c_int),
                ("medim", ctypes.c_int), ("n_kinds", ctypes.c_int), ("kinds", ctypes.POINTER(ctypes.c_int))]
  lib.ekf_get.restype = ctypes.POINTER(Desc)
  d = lib.ekf_get().contents
  assert d.abi_version == 1 and d.name.decode() == name
  assert (d.dim, d.edim, d.medim) == dims
  assert [d.kinds[i] for i in range(d.n_kinds)] == kinds


def test_header_is_parseable_by_the_reference_loader_rule(gen_dir):
  """rednose/helpers/__init__.py:27 keeps only lines starting with 'void ' and feeds them to cffi.cdef."""
  from cffi import FFI
  with open(os.path.join(gen_dir, "live.h"), encoding="utf-8") as f:
    header = "\n".join(ln for ln in f.read().split("\n") if ln.startswith("void "))
  ffi = FFI()
  ffi.cdef(header)
  lib = ffi.dlopen(os.path.join(gen_dir, "liblive.so"))
  assert lib.live_update_12 and lib.live_batch_step_12


def test_include_header_compiles_as_c(tmp_path):
  import subprocess
  from rednose_b200.build import INCLUDE_DIR
  src = tmp_path / "t.c"
  src.write_text('#include "rednose_b200.h"\nint main(void){ rednose_ekf_desc d; (void)d; return 0; }\n')
  subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", f"-I{INCLUDE_DIR}", "-c", str(src), "-o", str(tmp_path / "t.o")], check=True)


def test_generator_features_symbols(gen_dir):
  """global_vars -> <name>_set_<var>, extra_routines -> <name>_<routine> (ekf_sym.py:94-95,166-171)."""
  from rednose_b200.filters import ensure_generated
  from rednose_b200.filters.pendulum import PendulumKalman
  d = ensure_generated(PendulumKalman)
  lib = ctypes.CDLL(os.path.join(d, "libpendulum.so"))
  for sym in ("pendulum_set_grav", "pendulum_set_damp", "pendulum_energy", "pendulum_update_2", "pendulum_batch_step_2"):
    assert hasattr(lib, sym)
