"""Lightweight symbolic tensor shape for `FPropMeta` (reference `core/tshape.py`)."""

import functools
import operator

from lingvo_b200.core import symbolic


class Shape:
  """A tensor shape whose dims may be ints or symbolic expressions."""

  def __init__(self, dims):
    self._shape = [symbolic.ToStatic(d) if not symbolic.IsExpr(d) else d
                   for d in list(dims)]

  def __repr__(self):
    return 'Shape(%r)' % (self._shape,)

  @property
  def rank(self):
    return len(self._shape)

  def __len__(self):
    return len(self._shape)

  def __getitem__(self, key):
    if isinstance(key, slice):
      return Shape(self._shape[key])
    return self._shape[key]

  def __iter__(self):
    return iter(self._shape)

  def __add__(self, other):
    other = list(other) if not isinstance(other, Shape) else other._shape
    return Shape(self._shape + list(other))

  def __radd__(self, other):
    return Shape(list(other) + self._shape)

  def __eq__(self, other):
    return list(self) == list(other)

  def ToTensorShape(self):
    return [symbolic.ToStatic(d) for d in self._shape]

  def num_elements(self):  # pylint: disable=invalid-name
    return functools.reduce(operator.mul, self._shape, 1)

  def size(self):  # pylint: disable=invalid-name
    return self.num_elements()

  def Subs(self, bindings):
    return Shape([symbolic.EvalExpr(d, bindings) if symbolic.IsExpr(d) else d
                  for d in self._shape])
