"""EGDD: exponentiated-gradient delta-delta optimizer (ref `lingvo/core/egdd.py`).

Per-variable (or per-dimension) learning-rate gains updated multiplicatively from
the agreement between the current gradient and a momentum of past gradients:
  gain ← clip(gain · exp(μ · sign-agreement), [gain_min, gain_max])
  w ← w − lr · lr_scale · gain · m,     m ← β·m + (1−β)·g
"""

import torch

from lingvo_b200.core import optimizer


class EGDD(optimizer.Base):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('momentum', 0.9, 'Momentum β.')
    p.Define('beta', 0.9, 'EMA of the gain signal.')
    p.Define('gain_learning_rate', 0.01, 'μ: learning rate of the log-gain.')
    p.Define('scale_learning_rate', 0.001, 'Learning rate of the global lr scale.')
    p.Define('initial_gain', 1.0, 'Initial gain.')
    p.Define('min_gain', 1e-2, 'Min gain.')
    p.Define('max_gain', 1e2, 'Max gain.')
    p.Define('initial_scale', 1.0, 'Initial lr scale.')
    p.Define('min_scale', 1e-1, 'Min lr scale.')
    p.Define('max_scale', 1e1, 'Max lr scale.')
    p.Define('use_directions', True, 'Use sign(g) instead of g for the gain signal.')
    p.Define('use_signs', True, 'Apply sign(m) updates.')
    return p

  def _Update(self, lr, variables, grads):
    p = self.params
    for v, g in zip(variables, grads):
      g = g.to(v.dtype)
      m = self._Slot(v, 'momentum')
      gain = self._Slot(v, 'gain', init=p.initial_gain)
      lr_scale = self._Slot(v, 'lr_scale', init=p.initial_scale, shape=[])
      gbar = self._Slot(v, 'gbar')
      sig = torch.sign(g) if p.use_directions else g
      agree = torch.sign(m) * sig
      gbar.mul_(p.beta).add_(agree, alpha=1 - p.beta)
      gain.mul_(torch.exp(p.gain_learning_rate * gbar)).clamp_(p.min_gain, p.max_gain)
      lr_scale.mul_(torch.exp(p.scale_learning_rate * agree.mean())).clamp_(p.min_scale,
                                                                           p.max_scale)
      m.mul_(p.momentum).add_(g, alpha=1 - p.momentum)
      step = torch.sign(m) if p.use_signs else m
      v.data.add_(-float(lr) * lr_scale * gain * step)
