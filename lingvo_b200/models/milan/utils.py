"""Generic helpers for the dual-encoder code (ref `lingvo/tasks/milan/utils.py`)."""

from __future__ import annotations

import torch

from lingvo_b200.core import base_layer
from lingvo_b200.core.nested_map import NestedMap


def BatchMajorToTimeMajor(tensor):
  """[B, T, …] → [T, B, …] (ref :28)."""
  return tensor.transpose(0, 1)


class _FnLayer(base_layer.BaseLayer):
  """A parameter-free layer around a python callable."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('fn', None, 'Callable applied to the FProp arguments.')
    return p

  def FProp(self, theta, *args, **kwargs):
    return self.params.fn(*args, **kwargs)


def MakeFnLayer(fn, name=None):
  """Params of a layer whose FProp is `fn` (ref :34)."""
  return _FnLayer.Params().Set(fn=fn, name=name or getattr(fn, '__name__', 'fn_layer'))


def GetFromNestedMapOrDie(nested_map, key):
  """`a.b.c` lookup with a helpful error (ref :65)."""
  value = nested_map
  for part in key.split('.'):
    if not hasattr(value, 'get') or part not in value:
      raise ValueError('Could not find "%s" in %s' % (key, list(NestedMap(nested_map).FlattenItems())[:8]))
    value = value[part]
  return value


class Selector:
  """Picks / renames parts of a NestedMap (ref :81).

  `spec` is a key string (`'a.b'` → that value), or a dict `{new_name: key_or_spec}`
  (→ NestedMap of the selected values, recursively)."""

  def __init__(self, spec):
    if not isinstance(spec, (str, dict)):
      raise ValueError('Selector spec must be a str or dict, got %r' % type(spec))
    self._spec = spec

  def __call__(self, nested_map):
    def _Select(spec):
      if isinstance(spec, str):
        return GetFromNestedMapOrDie(nested_map, spec)
      return NestedMap({k: _Select(v) for k, v in spec.items()})
    return _Select(self._spec)


def InferBatchSize(batch) -> int:
  """Common leading dim of every tensor in `batch` (ref :130)."""
  sizes = {int(t.shape[0]) for t in NestedMap(batch).Flatten() if hasattr(t, 'shape') and len(t.shape)}
  if len(sizes) != 1:
    raise ValueError('Inconsistent or missing batch sizes: %s' % sorted(sizes))
  return sizes.pop()


def ResolveBatchDim(shape, batch_size):
  """Replaces an unknown (None / -1) leading dim with `batch_size` (ref :152)."""
  shape = list(shape)
  if shape and shape[0] in (None, -1):
    shape[0] = batch_size
  return torch.Size(shape)


class BatchFlattener:
  """Flattens / restores leading batch dims `[B, N, …] ↔ [B·N, …]` (ref :158)."""

  def __init__(self, batch_shape):
    self._batch_shape = list(batch_shape)
    if sum(1 for d in self._batch_shape if d in (None, -1)) > 1:
      raise ValueError('At most one unknown batch dim: %s' % self._batch_shape)

  @property
  def batch_shape(self):
    return self._batch_shape

  def Flatten(self, tensors):
    n = len(self._batch_shape)
    return NestedMap(x=tensors).Transform(lambda t: t.reshape((-1,) + tuple(t.shape[n:]))).x \
        if not isinstance(tensors, torch.Tensor) else tensors.reshape((-1,) + tuple(tensors.shape[n:]))

  def Unflatten(self, flat_tensors):
    shape = [(-1 if d in (None, -1) else d) for d in self._batch_shape]
    def _One(t):
      return t.reshape(tuple(shape) + tuple(t.shape[1:]))
    if isinstance(flat_tensors, torch.Tensor):
      return _One(flat_tensors)
    return NestedMap(x=flat_tensors).Transform(_One).x


def FlattenBatch(tensors, batch_shape):
  return BatchFlattener(batch_shape).Flatten(tensors)


def UnflattenBatch(flat_tensors, batch_shape):
  return BatchFlattener(batch_shape).Unflatten(flat_tensors)


def CollectRegularizationLosses(layer):
  """Regularisation terms exposed by sub-layers through a `losses` attribute (ref :256)."""
  out = []
  def _Visit(l):
    extra = getattr(l, 'losses', None)
    if extra:
      out.extend(extra if isinstance(extra, (list, tuple)) else [extra])
    for c in getattr(l, 'children', {}).values():
      for x in (c if isinstance(c, (list, tuple)) else [c]):
        if isinstance(x, base_layer.BaseLayer):
          _Visit(x)
  _Visit(layer)
  return out


def PadOrTrimDimension(tensor, new_size, axis, pad_value=0):
  """Pads (at the end) or truncates `axis` to `new_size` (ref :282)."""
  cur = tensor.shape[axis]
  if cur >= new_size:
    return tensor.narrow(axis, 0, new_size)
  pad_shape = list(tensor.shape)
  pad_shape[axis] = new_size - cur
  pad = torch.full(pad_shape, pad_value, dtype=tensor.dtype, device=tensor.device)
  return torch.cat([tensor, pad], axis)
