#!/usr/bin/env python3
"""ORACLE build recipe -- test infrastructure only (see oracle/README.md).

Builds ``oracle/_ref/lib<name>.so`` + ``oracle/_ref/<name>.h``: a CPU library exporting the
reference's exact C symbol set (rednose/helpers/ekf_sym.py:149-171), assembled from

  1. the leaf functions and C-ABI wrappers emitted by the reference's OWN, UNMODIFIED generator
     (``gen_code``, rednose/helpers/ekf_sym.py:29-217), run from /root/reference in a subprocess;
     its output is written only under oracle/_ref/ (git-ignored, never committed), and
  2. oracle/ekf_oracle_core.h, the Eigen-free restatement of rednose/templates/ekf_c.c that
     replaces the Eigen template the generator pastes in (Eigen is absent from this image), and
  3. oracle/batch_runner.inc: a threaded loop over independent filters used as the CPU baseline.

The reference sources are read where they lie; nothing is copied into the repository.
Usage:  python oracle/build_ref.py [kinematic live compare ...]
"""
import os
import subprocess
import sys
import textwrap

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("REDNOSE_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")

# reference example scripts that act as generators: python <script> <target> <dir>
# (site_scons/site_tools/rednose_filter.py:12-13)
EXAMPLE_SCRIPTS = {
  "kinematic": "examples/kinematic_kf.py",
  "live": "examples/live_kf.py",
  "compare": "examples/test_compare.py",
}


def reference_available():
  return os.path.exists(os.path.join(REF, "rednose", "helpers", "ekf_sym.py"))


def run_reference_generator(name, gen_dir, model_spec=None):
  """Produce <gen_dir>/<name>.cpp/.h with the reference generator (unmodified)."""
  os.makedirs(gen_dir, exist_ok=True)
  env = dict(os.environ, PYTHONPATH=REF + os.pathsep + REPO)
  if model_spec is None:
    script = os.path.join(REF, EXAMPLE_SCRIPTS[name])
    subprocess.run([sys.executable, script, name, gen_dir], check=True, env=env, cwd="/tmp")
  else:
    # a filter defined in this repository (e.g. the synthetic MSCKF model): feed its symbolic
    # definition to the reference's gen_code
    mod, attr = model_spec.split(":")
    code = textwrap.dedent(f"""
      import importlib, sys
      from rednose.helpers.ekf_sym import gen_code      # resolves to the reference (first on PYTHONPATH)
      import rednose, os
      assert os.path.realpath(rednose.__file__).startswith(os.path.realpath({REF!r})), rednose.__file__
      cls = getattr(importlib.import_module({mod!r}), {attr!r})
      kw = cls.symbolic_model()
      gen_code({gen_dir!r}, {name!r}, kw.pop('f_sym'), kw.pop('dt_sym'), kw.pop('x_sym'), kw.pop('obs_eqs'),
               kw.pop('dim_x'), kw.pop('dim_err'), **kw)
    """)
    subprocess.run([sys.executable, "-c", code], check=True, env=env, cwd="/tmp")


def assemble(name, gen_dir):
  """Splice the generated .cpp around oracle/ekf_oracle_core.h; return the path of the new TU."""
  with open(os.path.join(gen_dir, f"{name}.cpp"), encoding="utf-8") as f:
    src = f.read()
  eigen_at = src.index("#include <eigen3/Eigen/Dense>")
  head = src[:eigen_at]
  # drop the include of <name>.h (it pulls rednose/helpers/ekf.h -> Eigen)
  head = "\n".join(ln for ln in head.split("\n") if not ln.startswith("#include"))
  # the pasted template ends where post_code starts: "\n}\n" closing the anonymous namespace,
  # followed by 'extern "C" {' (ekf_sym.py:134-135)
  post_at = src.index('extern "C" {', eigen_at)
  ns_close = src.rindex("}", eigen_at, post_at)
  post = src[ns_close:]
  post = post[:post.index("const EKF ")]  # the C++ plugin struct needs ekf.h (Eigen); not part of the C-ABI
  kinds = [int(ln.split("MAHA_THRESH_")[1].split(" ")[0]) for ln in head.split("\n") if "const static double MAHA_THRESH_" in ln]
  with open(os.path.join(HERE, "batch_runner.inc"), encoding="utf-8") as f:
    runner = f.read()
  dispatch = "\n".join(f"    case {k}: {name}_update_{k}(x, P, z, R, ea); break;" for k in kinds)
  import re
  upd = re.findall(r"void " + name + r"_update_(\d+)\(double \*in_x.*?\{\s*update<(\d+), 3, (\d)>\(in_x, in_P, h_\d+, H_\d+, (\w+),", post, flags=re.S)
  dims = {k: int(re.search(r"#define " + k + r" (\d+)", head).group(1)) for k in ("DIM", "EDIM", "MEDIM")}
  kl = [int(u[0]) for u in upd]
  desc = ["", "// plugin descriptor over the CPU functions, so the native driver (runtime.cc) can be tested without a GPU",
          f'#include "{os.path.join(REPO, "include", "rednose_b200.h")}"',
          "namespace {",
          f"const int okinds_[] = {{ {', '.join(map(str, kl))} }};",
          f"const int ozdims_[] = {{ {', '.join(u[1] for u in upd)} }};",
          f"const int oeadims_[] = {{ {', '.join('3' if u[3] != 'NULL' else '0' for u in upd)} }};",
          f"const int ofeat_[] = {{ {', '.join('1' if u[3] != 'NULL' else '0' for u in upd)} }};",
          f"const int omaha_[] = {{ {', '.join(u[2] for u in upd)} }};",
          f"const rednose_leaf3_fn ohs_[] = {{ {', '.join(f'{name}_h_{k}' for k in kl)} }};",
          f"const rednose_leaf3_fn oHs_[] = {{ {', '.join(f'{name}_H_{k}' for k in kl)} }};",
          f"const rednose_leaf3_fn oHes_[] = {{ {', '.join((f'{name}_He_{u[0]}' if u[3] != 'NULL' else 'nullptr') for u in upd)} }};",
          f"const rednose_update_fn oupd_[] = {{ {', '.join(f'{name}_update_{k}' for k in kl)} }};",
          "const char* const onone_[] = { nullptr }; const rednose_set_fn osets_[] = { nullptr }; void* const oext_[] = { nullptr };",
          f'const rednose_ekf_desc odesc_ = {{ 1, "{name}", {dims["DIM"]}, {dims["EDIM"]}, {dims["MEDIM"]}, {len(kl)}, okinds_, ozdims_, oeadims_, ofeat_, omaha_,',
          f"  {name}_f_fun, {name}_F_fun, {name}_err_fun, {name}_inv_err_fun, {name}_H_mod_fun, {name}_predict, ohs_, oHs_, oHes_, oupd_,",
          "  0, onone_, osets_, 0, onone_, oext_, nullptr, nullptr, nullptr, nullptr, nullptr };",
          "}", 'extern "C" void* ekf_get() { return (void*)&odesc_; }', ""]
  dispatch_arena = "\n".join(f"          case {k}: {name}_update_{k}(xb, Pb, z, R, ea); break;" for k in kinds)
  runner = runner.replace("@DISPATCH_ARENA@", dispatch_arena).replace("@NAME@", name).replace("@DISPATCH@", dispatch) + "\n".join(desc)
  tu = (f"// assembled by oracle/build_ref.py from the reference generator's output -- not committed\n"
        f"#include <math.h>\n#include <string.h>\n#include <stddef.h>\n#include <vector>\n#include <thread>\n#include <cmath>\n{head}\n"
        f"#include \"{os.path.join(HERE, 'ekf_oracle_core.h')}\"\n{post}\n{runner}\n")
  path = os.path.join(gen_dir, f"{name}_oracle.cpp")
  with open(path, "w", encoding="utf-8") as f:
    f.write(tu)
  return path


def build(name, model_spec=None, force=False):
  lib = os.path.join(OUT, f"lib{name}.so")
  core = os.path.join(HERE, "ekf_oracle_core.h")
  runner = os.path.join(HERE, "batch_runner.inc")
  if not force and os.path.exists(lib) and all(os.path.getmtime(p) <= os.path.getmtime(lib) for p in (core, runner, __file__)):
    return lib
  if not reference_available():
    raise RuntimeError(f"{REF} not present: oracle/_ref can only be (re)built where the reference is mounted")
  gen_dir = os.path.join(OUT, "gen")
  run_reference_generator(name, gen_dir, model_spec)
  tu = assemble(name, gen_dir)
  # reference flags: -O2 -g -fPIC -std=c++1z (SConstruct:25-38)
  cmd = ["g++", "-O2", "-g", "-fPIC", "-std=c++17", "-shared", "-pthread", "-o", lib, tu]
  subprocess.run(cmd, check=True)
  # header for the cffi loader (the reference's own header, as generated)
  with open(os.path.join(gen_dir, f"{name}.h"), encoding="utf-8") as f:
    hdr = f.read()
  hdr += (f"\nvoid {name}_oracle_batch_step(int kind, double *x, double *P, double *Q, const double *dt_arr, double dt, "
          f"double *z, double *R, double *ea, int zdim, int eadim, long long B, int nthreads, const int *quat_idxs, int n_quat, int flags);\n"
          f"void *{name}_oracle_arena_create(long long B, int nthreads, int pin);\n"
          f"void {name}_oracle_arena_destroy(void *h);\n"
          f"void {name}_oracle_arena_info(void *h, long long *out);\n"
          f"void {name}_oracle_arena_load(void *h, const double *x, long long x_stride, const double *P, long long P_stride);\n"
          f"void {name}_oracle_arena_step(void *h, int kind, const double *Q, const double *dt_arr, double dt, const double *z, double *y, "
          f"const double *R, long long R_stride, const double *ea, int zdim, int eadim, const int *quat_idxs, int n_quat, int flags);\n"
          f"void {name}_oracle_arena_read(void *h, long long b0, long long b1, double *x, double *P);\n")
  with open(os.path.join(OUT, f"{name}.h"), "w", encoding="utf-8") as f:
    f.write(hdr)
  return lib


def build_features(K=10, force=False):
  """oracle/_ref/libfeatures_<K>.so: oracle/features_oracle.h around res_fun / jac_fun printed by the reference's own
  sympy_into_c (run from /root/reference in a subprocess) from rednose_b200.features.residual_sym(K)."""
  name = f"features_{K}"
  lib = os.path.join(OUT, f"lib{name}.so")
  core = os.path.join(HERE, "features_oracle.h")
  if not force and os.path.exists(lib) and all(os.path.getmtime(p) <= os.path.getmtime(lib) for p in (core, __file__)):
    return lib
  if not reference_available():
    raise RuntimeError(f"{REF} not present: oracle/_ref can only be (re)built where the reference is mounted")
  gen_dir = os.path.join(OUT, "gen")
  os.makedirs(gen_dir, exist_ok=True)
  env = dict(os.environ, PYTHONPATH=REF + os.pathsep + REPO)
  leaf = os.path.join(gen_dir, f"{name}_leaf.c")
  code = textwrap.dedent(f"""
    import os, rednose
    assert os.path.realpath(rednose.__file__).startswith(os.path.realpath({REF!r})), rednose.__file__
    from rednose.helpers.sympy_helpers import sympy_into_c      # the reference's printer, unmodified
    from rednose_b200.features import residual_sym
    res, jac, args = residual_sym({K})
    header, code = sympy_into_c([('res_fun', res, args), ('jac_fun', jac, args)])
    open({leaf!r}, 'w').write(code)
  """)
  subprocess.run([sys.executable, "-c", code], check=True, env=env, cwd="/tmp")
  with open(leaf, encoding="utf-8") as f:
    leaf_c = f.read()
  tu = (f"// assembled by oracle/build_ref.py -- not committed\n#include <math.h>\n#include <string.h>\n#define KDIM {K}\n#define K {K}\n"
        f"extern \"C\" {{\n{leaf_c}\n}}\n#include \"{core}\"\n")
  path = os.path.join(gen_dir, f"{name}_oracle.cpp")
  with open(path, "w", encoding="utf-8") as f:
    f.write(tu)
  subprocess.run(["g++", "-O2", "-g", "-fPIC", "-std=c++17", "-shared", "-o", lib, path], check=True)
  with open(os.path.join(OUT, f"{name}.h"), "w", encoding="utf-8") as f:
    f.write("void compute_pos(double *to_c, double *poses, double *img_positions, double *param, double *pos);\n"
            "void res_fun(double *abr, double *poses, double *img_positions, double *out);\n"
            "void jac_fun(double *abr, double *poses, double *img_positions, double *out);\n"
            "void merge_features(double *tracks, double *features, long long *empty_idxs);\n"
            "void merge_features_n(double *tracks, double *features, long long *empty_idxs, int n_features, int n_tracks);\n"
            "int sane(double *track);\n")
  return lib


if __name__ == "__main__":
  names = sys.argv[1:] or ["kinematic", "live", "compare"]
  for n in names:
    spec = None
    if ":" in n and "=" in n:
      n, spec = n.split("=", 1)
    print(build(n, spec, force=True))
