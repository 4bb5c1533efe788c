"""Gradient combination for multi-loss training (ref `lingvo/core/gradient_combiner.py`).

`Combine(vmap, {loss_name: NestedMap(loss_metric=…, grads=NestedMap of VarGrad)})` →
`(NestedMap of VarGrad, eval_metrics)`; the learner calls it when `loss_name` is a list
(`learner.py:_ComputeLossesAndGradients`). Per-loss gradient maps may miss variables the
loss does not depend on.
"""
import torch

from lingvo_b200.core import base_layer
from lingvo_b200.core import py_utils
from lingvo_b200.core.nested_map import NestedMap


def _Aligned(vmap, entry):
  """Gradient tensors of one loss, aligned with `vmap.FlattenItems()` (None where absent)."""
  by_key = {}
  for k, vg in entry.grads.FlattenItems():
    by_key[k] = vg.grad if isinstance(vg, py_utils.VarGrad) else vg
  return [by_key.get(k) for k, _ in vmap.FlattenItems()]


def _Pack(vmap, grads):
  out = NestedMap()
  for (k, v), g in zip(vmap.FlattenItems(), grads):
    if g is not None:
      out.Set(k, py_utils.VarGrad(v, g))
  return out


class GradientCombiner(base_layer.BaseLayer):
  """Base class: see module docstring for the contract."""

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
    per_loss = [_Aligned(vmap, losses_and_gradients[n]) for n in names]
    out = []
    for i in range(len(per_loss[0])):
      gs = [w.get(n, 1.0) * f[i] for n, f in zip(names, per_loss) if f[i] is not None]
      out.append(sum(gs) if gs else None)
    return _Pack(vmap, out), {}


class PCGradCombiner(GradientCombiner):
  """Gradient surgery (arXiv 2001.06782): each task gradient is projected onto the normal
  plane of every other task gradient it conflicts with (negative inner product), then the
  projected gradients are summed."""

  def Combine(self, vmap, losses_and_gradients):
    names = sorted(losses_and_gradients)
    variables = [v for _, v in vmap.FlattenItems()]
    per_loss = [[g if g is not None else torch.zeros_like(v)
                 for g, v in zip(_Aligned(vmap, losses_and_gradients[n]), variables)]
                for n in names]
    vecs = [torch.cat([g.reshape(-1).float() for g in f]) for f in per_loss]
    proj = [v.clone() for v in vecs]
    conflicts = 0
    for i in range(len(vecs)):
      for j in range(len(vecs)):
        if i == j:
          continue
        dot = torch.dot(proj[i], vecs[j])
        if dot < 0:
          conflicts += 1
          proj[i] = proj[i] - dot / vecs[j].pow(2).sum().clamp_min(1e-12) * vecs[j]
    total = sum(proj)
    out, off = [], 0
    for v in variables:
      n = v.numel()
      out.append(total[off:off + n].reshape(v.shape).to(v.dtype))
      off += n
    metrics = {'pcgrad_conflicts': (torch.tensor(float(conflicts)), torch.tensor(1.0))}
    return _Pack(vmap, out), metrics
