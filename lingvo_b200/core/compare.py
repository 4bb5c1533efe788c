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


# -- readable NestedMap assertions (ref compare.py:57-86) -----------------------------------
def assertNestedMapEqual(self, expected, actual):  # pylint: disable=invalid-name
  """Fails with a per-key diff of the two maps' DebugStrings instead of one long repr.
  `self` is a unittest.TestCase (or None: plain AssertionError with a unified diff)."""
  from lingvo_b200.core.nested_map import NestedMap  # pylint: disable=g-import-not-at-top
  if not hasattr(expected, 'DebugString'):
    expected = NestedMap(expected)
  a, b = expected.DebugString(), actual.DebugString()
  if self is not None and hasattr(self, 'assertMultiLineEqual'):
    return self.assertMultiLineEqual(a, b)
  if a != b:
    import difflib  # pylint: disable=g-import-not-at-top
    raise AssertionError('\n' + '\n'.join(difflib.ndiff(a.splitlines(), b.splitlines())))


import unittest as _unittest  # pylint: disable=g-import-not-at-top,g-bad-import-order


class NestedMapAssertions(_unittest.TestCase):
  """Mix into a TestCase to get `self.assertNestedMapEqual(expected, actual)`."""

  def assertNestedMapEqual(self, *args, **kwargs):  # pylint: disable=invalid-name
    return assertNestedMapEqual(self, *args, **kwargs)
