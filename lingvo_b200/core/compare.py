"""Metric comparison helpers (ref `lingvo/core/compare.py`)."""


def Maximize(x, y):
  """True iff x is better than y when larger is better."""
  return x > y


def Minimize(x, y):
  return x < y


class Comparator:
  """`Comparator(minimize)(new, best)` → True if `new` improves on `best`."""

  def __init__(self, minimize=True, tolerance=0.0):
    self._minimize = minimize
    self._tol = tolerance

  def __call__(self, new, best):
    if best is None:
      return True
    return new < best - self._tol if self._minimize else new > best + self._tol
