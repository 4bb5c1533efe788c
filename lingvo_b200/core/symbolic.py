"""Symbolic dims (reference `core/symbolic.py`): sympy-backed when available."""

import contextlib
import threading

try:
  import sympy
  _HAS_SYMPY = True
except Exception:  # pylint: disable=broad-except
  sympy = None
  _HAS_SYMPY = False


class _Ctx(threading.local):

  def __init__(self):
    super().__init__()
    self.stack = []


_VALUES = _Ctx()


def Symbol(name):
  if not _HAS_SYMPY:
    raise RuntimeError('sympy is required for symbolic dims')
  return sympy.Symbol(name)


def IsSymbol(x):
  return _HAS_SYMPY and isinstance(x, sympy.Symbol)


def IsExpr(x):
  return _HAS_SYMPY and isinstance(x, sympy.Expr) and not x.is_number


@contextlib.contextmanager
def SymbolToValueMap(symbol_type, values):
  _VALUES.stack.append((symbol_type, dict(values)))
  try:
    yield
  finally:
    _VALUES.stack.pop()


STATIC_VALUES = 'static'
TENSOR_VALUES = 'tensor'


def EvalExpr(value_type, x=None):
  """EvalExpr(type, expr) or EvalExpr(expr, bindings)."""
  if x is None or isinstance(x, dict):
    expr, bindings = value_type, (x or {})
  else:
    expr = x
    bindings = {}
    for t, vals in _VALUES.stack:
      if t == value_type:
        bindings.update(vals)
  if not IsExpr(expr):
    return expr
  out = expr.subs(bindings)
  return int(out) if out.is_number and out == int(out) else (
      float(out) if out.is_number else out)


def ToStatic(expr):
  if IsExpr(expr):
    return EvalExpr(STATIC_VALUES, expr)
  return expr


def ToTensor(expr):
  if IsExpr(expr):
    return EvalExpr(TENSOR_VALUES, expr)
  return expr
