"""sympy -> CUDA device-function text.

Replaces the reference's ``sympy_into_c`` (rednose/helpers/sympy_helpers.py:122-162,
a thin wrapper over ``sympy.utilities.codegen`` that emits dense, CSE-free C99 with
``pow(x, 2)``).  Differences that matter on the GPU:

* common-subexpression elimination across ALL outputs of a routine (``sympy.cse``),
* integer / half-integer powers are expanded to products and ``sqrt`` (nvcc does not
  lower ``pow(x, 2)``; it calls the generic pow routine),
* structural zeros are never stored: callers ask for *value slots* of the non-zero
  pattern and get straight-line code that consumes them.
"""
from __future__ import annotations

import sympy as sp
from sympy.printing.c import C99CodePrinter


def normalise(expr):
  """Make structural zeros/ones visible: Float(0.0)/Float(1.0)/Float(2.0) -> Integer.

  Filter definitions are built from numpy arrays (examples/live_kf.py:161,180,187),
  so matrices are full of ``Float(0.0)`` which sympy >= 1.13 no longer equates with 0.
  """
  # compare as Python floats: Float(0.0) == 0 is False in sympy >= 1.13
  reps = {f: sp.Integer(int(f)) for f in expr.atoms(sp.Float) if abs(float(f)) <= 16.0 and float(f) == int(float(f))}  # small structural constants only
  return expr.xreplace(reps) if reps else expr


def _probe_value(e):
  """Numeric value of e at a fixed pseudo-random point (None if it cannot be evaluated)."""
  from sympy.matrices.expressions.matexpr import MatrixElement
  reps = {}
  for a in e.atoms(MatrixElement) | e.atoms(sp.Symbol):
    if isinstance(a, sp.Symbol) and any(a in m.free_symbols for m in e.atoms(MatrixElement)):
      continue  # the MatrixSymbol itself shows up as a free symbol of its elements
    h = (hash(str(a)) % 9973) / 9973.0
    reps[a] = sp.Float(0.37 + 0.91 * h)          # away from 0, 1 and the singular points of the usual models
  try:
    v = complex(sp.N(e.xreplace(reps), 20))
  except (TypeError, ValueError):
    return None
  return abs(v)


def is_structural_zero(e, expand_limit=400):
  """True if e is identically zero.  Exact zeros are caught directly; anything else is first probed numerically
  (a non-zero value at a generic point proves e != 0 and skips the expensive expand), and only expressions
  that evaluate to ~0 are expanded symbolically to confirm the cancellation."""
  if e == 0 or e.is_zero:
    return True
  if e.is_number:
    return float(e) == 0.0
  v = _probe_value(e)
  if v is not None and v > 1e-9:
    return False
  if sp.count_ops(e) <= expand_limit:
    return sp.expand(e) == 0
  return False


class CudaPrinter(C99CodePrinter):
  """Prints scalar sympy expressions as CUDA C (double precision)."""

  def __init__(self, array_names: dict[str, str], scalar_names: dict[str, str]):
    super().__init__({'contract': False})
    self._arrays = array_names    # MatrixSymbol name -> C pointer name
    self._scalars = scalar_names  # Symbol name -> C expression

  def _print_MatrixElement(self, expr):
    parent = expr.parent
    cname = self._arrays.get(str(parent.name) if hasattr(parent, 'name') else str(parent))
    if cname is None:
      raise KeyError(f"matrix symbol {parent} is not an argument of this routine")
    idx = int(expr.i) * int(parent.shape[1]) + int(expr.j)
    return f"{cname}[{idx}]"

  def _print_Symbol(self, expr):
    return self._scalars.get(expr.name, expr.name)

  def _mul_chain(self, base: str, n: int) -> str:
    return "(" + "*".join([base] * n) + ")"

  def _print_Pow(self, expr):
    base, exp = expr.base, expr.exp
    if exp.is_number:
      e2 = sp.nsimplify(2 * exp) if not exp.is_Rational else 2 * exp
      if e2.is_Integer and abs(int(e2)) <= 16:
        twice = int(e2)
        b = self.parenthesize(base, 1000)  # always parenthesised unless atomic
        n, half = divmod(abs(twice), 2)
        parts = []
        if n == 1:
          parts.append(b)
        elif n > 1:
          parts.append(self._mul_chain(b, n))
        if half:
          parts.append(f"sqrt({self._print(base)})")
        body = "*".join(parts) if parts else "1.0"
        if twice < 0:
          return f"(1.0/({body}))"
        return body if len(parts) == 1 else f"({body})"
    return f"pow({self._print(base)}, {self._print(exp)})"

  def _print_Integer(self, expr):
    # keep arithmetic in double: "2" -> "2.0" avoids int/int surprises in emitted code
    return f"{int(expr)}.0"

  def _print_Rational(self, expr):
    return f"({int(expr.p)}.0/{int(expr.q)}.0)"


def cse_block(outputs: list[tuple[str, sp.Expr]], printer: CudaPrinter, tmp_prefix: str = "_c", indent: str = "    ") -> str:
  """Emit ``const double _cK = ...;`` temporaries followed by ``target = expr;`` lines."""
  if not outputs:
    return ""
  exprs = [normalise(sp.sympify(e)) for _, e in outputs]
  repl, reduced = sp.cse(exprs, symbols=sp.numbered_symbols(tmp_prefix), optimizations='basic', order='none')
  lines = []
  for sym, e in repl:
    lines.append(f"{indent}const double {sym} = {printer.doprint(e)};")
  for (target, _), e in zip(outputs, reduced):
    lines.append(f"{indent}{target} = {printer.doprint(e)};")
  return "\n".join(lines) + "\n"


def sparse_pattern(mat: sp.Matrix):
  """[(i, j, expr)] of the structurally non-zero entries of a matrix, row-major."""
  out = []
  for i in range(mat.shape[0]):
    for j in range(mat.shape[1]):
      e = normalise(mat[i, j])
      if not is_structural_zero(e):
        out.append((i, j, e))
  return out
