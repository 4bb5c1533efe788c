"""RevNet blocks (ref `lingvo/core/reversible_layers.py`).

y1 = x1 + F(x2); y2 = x2 + G(y1). Activations are NOT stored: the backward pass
reconstructs x from y (x2 = y2 − G(y1); x1 = y1 − F(x2)) inside a custom
autograd Function, so activation memory is O(1) in depth (ref :27-140).
"""

from __future__ import annotations

import torch

from lingvo_b200.core import base_layer
from lingvo_b200.core.nested_map import NestedMap


class _RevFn(torch.autograd.Function):

  @staticmethod
  def forward(ctx, f, g, n_f, x1, x2, *flat_theta):
    ctx.f, ctx.g, ctx.n_f = f, g, n_f
    with torch.no_grad():
      tf, tg = flat_theta[:n_f], flat_theta[n_f:]
      y1 = x1 + f(tf, x2)
      y2 = x2 + g(tg, y1)
    ctx.save_for_backward(y1, y2, *flat_theta)
    return y1, y2

  @staticmethod
  def backward(ctx, dy1, dy2):
    y1, y2, *flat_theta = ctx.saved_tensors
    n_f = ctx.n_f
    tf = [t.detach().requires_grad_(t.requires_grad) for t in flat_theta[:n_f]]
    tg = [t.detach().requires_grad_(t.requires_grad) for t in flat_theta[n_f:]]
    with torch.enable_grad():
      y1_ = y1.detach().requires_grad_(True)
      gy1 = ctx.g(tg, y1_)
      x2 = (y2 - gy1).detach()
      grads_g = torch.autograd.grad(gy1, [y1_] + [t for t in tg if t.requires_grad], dy2,
                                    allow_unused=True)
      dy1_total = dy1 + (grads_g[0] if grads_g[0] is not None else 0)
      x2_ = x2.requires_grad_(True)
      fx2 = ctx.f(tf, x2_)
      grads_f = torch.autograd.grad(fx2, [x2_] + [t for t in tf if t.requires_grad], dy1_total,
                                    allow_unused=True)
      dx2 = dy2 + (grads_f[0] if grads_f[0] is not None else 0)
      dx1 = dy1_total

    def _Fill(ts, gs):
      it = iter(gs)
      return [next(it) if t.requires_grad else None for t in ts]
    return (None, None, None, dx1, dx2, *_Fill(tf, grads_f[1:]), *_Fill(tg, grads_g[1:]))


class RevNetLayer(base_layer.BaseLayer):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('f_params', None, 'Layer params for F.')
    p.Define('g_params', None, 'Layer params for G.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    self.CreateChild('f_block', p.f_params)
    self.CreateChild('g_block', p.g_params)

  def FProp(self, theta, inputs):
    """inputs: NestedMap(split1, split2) → NestedMap(split1, split2)."""
    tf_flat = theta.f_block.Flatten()
    tg_flat = theta.g_block.Flatten()

    def f(flat, x):
      return self.f_block.FProp(theta.f_block.Pack(list(flat)), x)

    def g(flat, x):
      return self.g_block.FProp(theta.g_block.Pack(list(flat)), x)
    y1, y2 = _RevFn.apply(f, g, len(tf_flat), inputs.split1, inputs.split2, *tf_flat, *tg_flat)
    return NestedMap(split1=y1, split2=y2)

  def ReverseAndGrad(self, theta, outputs):
    """Reconstructs the inputs from the outputs (for tests / inspection)."""
    with torch.no_grad():
      x2 = outputs.split2 - self.g_block.FProp(theta.g_block, outputs.split1)
      x1 = outputs.split1 - self.f_block.FProp(theta.f_block, x2)
    return NestedMap(split1=x1, split2=x2)


class StackedRevNetLayer(base_layer.BaseLayer):
  """Sequence of RevNetLayers (ref :143)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('sub_layer_params', [], 'List of RevNetLayer params.')
    p.Define('custom_gradient', True, 'Kept for parity (always reconstructing).')
    return p

  def __init__(self, params):
    super().__init__(params)
    self.CreateChildren('sub_layers', list(self.params.sub_layer_params))

  def FProp(self, theta, inputs):
    x = inputs
    for i, l in enumerate(self.sub_layers):
      x = l.FProp(theta.sub_layers[i], x)
    return x
