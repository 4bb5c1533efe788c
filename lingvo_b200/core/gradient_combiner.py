"""Gradient combination for multi-loss training (ref `lingvo/core/gradient_combiner.py`)."""
import torch

from lingvo_b200.core import base_layer
from lingvo_b200.core.nested_map import NestedMap


class GradientCombiner(base_layer.BaseLayer):
  """Combine(vmap, {loss_name: NestedMap(loss_metric, grads)}) → combined grads."""

  def Combine(self, vmap, losses_and_gradients):
    raise NotImplementedError(type(self))


class SumCombiner(GradientCombiner):
  """Weighted sum (the default linear aggregation)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('weights', None, 'loss name → weight (default 1).')
    return p

  def Combine(self, vmap, losses_and_gradients):
    w = self.params.weights or {}
    names = sorted(losses_and_gradients)
    flats = [losses_and_gradients[n].grads.Flatten() for n in names]
    out = []
    for i in range(len(flats[0])):
      gs = [w.get(n, 1.0) * f[i] for n, f in zip(names, flats) if f[i] is not None]
      out.append(sum(gs) if gs else None)
    return vmap.Pack(out)


class PCGradCombiner(GradientCombiner):
  """Gradient surgery (arXiv 2001.06782): project away conflicting components."""

  def Combine(self, vmap, losses_and_gradients):
    names = sorted(losses_and_gradients)
    flats = [[g if g is not None else torch.zeros_like(v) for g, v in
              zip(losses_and_gradients[n].grads.Flatten(), vmap.Flatten())] for n in names]
    vecs = [torch.cat([g.reshape(-1).float() for g in f]) for f in flats]
    proj = [v.clone() for v in vecs]
    for i in range(len(vecs)):
      for j in range(len(vecs)):
        if i == j:
          continue
        dot = torch.dot(proj[i], vecs[j])
        if dot < 0:
          proj[i] = proj[i] - dot / vecs[j].pow(2).sum().clamp_min(1e-12) * vecs[j]
    total = sum(proj)
    out, off = [], 0
    for v in vmap.Flatten():
      n = v.numel()
      out.append(total[off:off + n].reshape(v.shape).to(v.dtype))
      off += n
    return vmap.Pack(out)
