"""CPU unit tests of the sympy -> CUDA layer (rednose_b200/codegen): printer rules, structural sparsity,
CSE emission, and the structure the generator derives for the live model (hard part 1)."""
import re

import numpy as np
import sympy as sp

from rednose_b200.codegen.symbolic import CudaPrinter, cse_block, is_structural_zero, normalise, sparse_pattern


def test_printer_expands_powers_instead_of_calling_pow():
  x = sp.MatrixSymbol('state', 3, 1)
  pr = CudaPrinter({'state': 's'}, {})
  a = x[0, 0]
  import math
  ev = lambda code, v: eval(code, {"s": [v, 0.0, 0.0], "sqrt": math.sqrt})   # the emitted C is valid Python here
  assert "pow" not in pr.doprint(a**2) and ev(pr.doprint(a**2), 3.0) == 9.0
  assert "pow" not in pr.doprint(a**-1) and ev(pr.doprint(a**-1), 4.0) == 0.25
  assert abs(ev(pr.doprint(a**sp.Rational(5, 2)), 4.0) - 32.0) < 1e-12
  assert "sqrt(s[0])" in pr.doprint(a**sp.Rational(3, 2)) and "pow" not in pr.doprint(a**sp.Rational(3, 2))
  assert "pow" not in pr.doprint((a**2 + x[1, 0]**2)**sp.Float(1.5))       # Float exponents too (live_kf.py:221)
  assert pr.doprint(a**sp.Rational(-3, 2)).startswith("(1.0/(")
  assert "pow(" in pr.doprint(a**sp.Symbol('n'))                             # genuinely symbolic exponents still work
  assert pr.doprint(sp.Integer(2) * a).startswith("2.0*")                   # no integer arithmetic in emitted code


def test_float_zeros_are_structural_zeros():
  """Matrices built from numpy arrays are full of Float(0.0), which sympy >= 1.13 does not equate with 0."""
  m = sp.Matrix(np.zeros((2, 2)))
  x = sp.Symbol('x')
  m[0, 1] = 1.0 * x
  m[1, 0] = x * 2 - 2 * x
  assert [(i, j) for i, j, _ in sparse_pattern(m)] == [(0, 1)]
  assert is_structural_zero(normalise(sp.Float(0.0))) and not is_structural_zero(x)


def test_cse_block_shares_subexpressions_across_outputs():
  x = sp.MatrixSymbol('state', 2, 1)
  e = sp.sin(x[0, 0] + x[1, 0])
  code = cse_block([("out[0]", e * 2), ("out[1]", e * e + 1)], CudaPrinter({'state': 'state'}, {}))
  assert code.count("sin(") == 1 and "const double _c0" in code


def test_live_model_structure(gen_dir):
  """F = I + dt*A with 33 value slots in 9 rows; per-kind H_err non-zeros."""
  with open(f"{gen_dir}/live.cu", encoding="utf-8") as f:
    src = f.read()
  assert "static constexpr int NF = 33, NFROWS = 9;" in src and "FROW_MASK = 0x1ffu" in src
  nh = dict((int(k), int(n)) for k, n in re.findall(r"static constexpr int KIND = (\d+), ZDIM = \d+, YDIM = \d+, EADIM = \d+, NH = (\d+);", src))
  assert nh == {3: 4, 4: 26, 9: 3, 10: 35, 12: 3, 13: 18, 14: 3, 19: 3}
  assert "pow(" not in src                                                    # every power was expanded
  # Mahalanobis thresholds: chi2 0.95 quantiles baked per kind (ekf_sym.py:144)
  t3 = float(re.search(r"KIND = 12.*?MAHA_THRESH = ([0-9.e+-]+);", src, flags=re.S).group(1))
  assert abs(t3 - 7.814727903251177) < 1e-9
