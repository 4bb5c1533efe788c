"""GradDrop (ref `lingvo/core/graddrop.py`): sign-consistency gradient masking.

Identity in the forward pass. In backward, given per-loss gradients g_k w.r.t. this
layer's output, computes the positive-sign purity P = ½(1 + Σg_k / Σ|g_k|), samples
U ~ Uniform, and keeps only positive components where P > U and only negative
components where P < U. Usage: y = graddrop.FProp(theta, x); losses = [...];
`SetLosses([(loss, leak_ratio), …])` is emulated by `CombineLossGrads`.
"""

import torch

from lingvo_b200.core import base_layer


class GradDrop(base_layer.BaseLayer):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('keep_prob_function', 'linear', 'linear | sigmoid.')
    p.Define('keep_prob_function_scale', 1.0, 'Scale of the keep-prob function.')
    p.Define('use_input_sign_only', True, 'Multiply grads by sign(input) (batch-separated).')
    p.Define('keep_gradnorm_constant', True, 'Rescale to keep the gradient norm.')
    p.Define('marginalize_batch_dim', True, 'Sum grads over the batch before P.')
    p.Define('epsilon', 1e-7, 'Numerical epsilon.')
    p.Define('random_seed', None, 'Seed.') if 'random_seed' not in p else None
    return p

  def FProp(self, theta, x):
    self._x = x
    return x

  def CombineLossGrads(self, grads, leak_ratios=None):
    """grads: list of dLoss_k/dx (same shape as x) → masked combined gradient."""
    p = self.params
    x = self._x.detach()
    sign_in = torch.sign(x) if p.use_input_sign_only else torch.ones_like(x)
    gs = [g * sign_in for g in grads]
    tot = sum(gs)
    if p.marginalize_batch_dim:
      num = sum(g.sum(0, keepdim=True) for g in gs)
      den = sum(g.abs().sum(0, keepdim=True) for g in gs)
    else:
      num, den = tot, sum(g.abs() for g in gs)
    purity = 0.5 * (1.0 + num / (den + p.epsilon))
    if p.keep_prob_function == 'sigmoid':
      purity = torch.sigmoid(p.keep_prob_function_scale * (purity - 0.5) * 8)
    else:
      purity = (p.keep_prob_function_scale * (purity - 0.5) + 0.5).clamp(0, 1)
    u = torch.rand_like(purity)
    keep_pos, keep_neg = (purity > u), (purity < u)
    out = 0
    leak_ratios = leak_ratios or [0.0] * len(grads)
    for g, gg, leak in zip(grads, gs, leak_ratios):
      mask = (keep_pos & (gg > 0)) | (keep_neg & (gg < 0))
      out = out + leak * g + (1 - leak) * g * mask.to(g.dtype)
    if p.keep_gradnorm_constant:
      out = out * (sum(grads).norm() / out.norm().clamp_min(p.epsilon))
    return out
