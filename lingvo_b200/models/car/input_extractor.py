"""Field extractors (ref `lingvo/tasks/car/input_extractor.py`).

An extractor declares which `tf.Example` features it needs (`FeatureMap`: name →
`(shape, dtype)` for fixed-length or `(None, dtype)` for variable-length features), turns
the parsed features of ONE record into a NestedMap of fixed-shape numpy arrays
(`_Extract`), and may veto the record (`Filter` → bucket key; ≥ `BUCKET_UPPER_BOUND`
drops it). Several extractors are composed by `base_extractor._BaseExtractor`.
"""

from __future__ import annotations

import numpy as np

from lingvo_b200.core import base_layer
from lingvo_b200.core.nested_map import NestedMap

BUCKET_UPPER_BOUND = 9999


class FieldsExtractor(base_layer.BaseLayer):
  """ref :34."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.name = cls.__name__
    return p

  def FeatureMap(self):
    """{feature name: (shape or None, numpy dtype or bytes)}."""
    raise NotImplementedError()

  def ContextMap(self):
    """SequenceExample context features (none by default)."""
    return {}

  def _Extract(self, features):
    raise NotImplementedError()

  def Extract(self, features):
    out = self._Extract(features)
    shapes = self.Shape()
    assert sorted(out.keys()) == sorted(shapes.keys()), '%s vs. %s' % (
        sorted(out.keys()), sorted(shapes.keys()))
    for (k, v), (_, s) in zip(sorted(out.FlattenItems()), sorted(shapes.FlattenItems())):
      if isinstance(v, np.ndarray) and s is not None:
        ok = len(v.shape) == len(s) and all(b is None or a == b for a, b in zip(v.shape, s))
        assert ok, '%s: %s vs. %s' % (k, v.shape, s)
    return out

  def _ExtractBatch(self, features):
    """Default batched extraction: per-example loop + stack."""
    n = len(next(iter(features.values())))
    outs = [self._Extract({k: v[i] for k, v in features.items()}) for i in range(n)]
    return outs[0].Pack([np.stack(vs) for vs in zip(*[o.Flatten() for o in outs])])

  def ExtractBatch(self, features):
    return self._ExtractBatch(features)

  def Filter(self, outputs):
    """Bucket key of the example; return `BUCKET_UPPER_BOUND` to drop it."""
    del outputs
    return 1

  def FilterBatch(self, outputs):
    """Batched filter: returns `outputs` with unwanted rows removed (default: keep)."""
    return outputs

  def Shape(self):
    """NestedMap of output shapes (tuples), without the batch dim."""
    raise NotImplementedError()

  def DType(self):
    """NestedMap of numpy dtypes."""
    raise NotImplementedError()


class NestedFieldsExtractor(FieldsExtractor):
  """Runs several child extractors and nests their outputs under their names (ref :200)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('extractors', None, 'hyperparams.Params of name → extractor params.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self._names = []
    for name, ep in self.params.extractors.IterParams():
      self.CreateChild(name, ep)
      self._names.append(name)

  def FeatureMap(self):
    out = {}
    for n in self._names:
      out.update(self.children[n].FeatureMap())
    return out

  def _Extract(self, features):
    return NestedMap({n: self.children[n].Extract(features) for n in self._names})

  def Filter(self, outputs):
    return max(self.children[n].Filter(outputs[n]) for n in self._names)

  def Shape(self):
    return NestedMap({n: self.children[n].Shape() for n in self._names})

  def DType(self):
    return NestedMap({n: self.children[n].DType() for n in self._names})


class LaserExtractor(FieldsExtractor):
  """Interface of laser extractors (ref :241). Output:

    points_xyz      [max_num_points, 3]
    points_feature  [max_num_points, num_features]
    points_padding  [max_num_points]   1.0 = padded

  Subclasses implement `_Extract`; `PadOrTrim` brings a variable point count to the
  static `max_num_points` (random subset when there are too many)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('max_num_points', None, 'Points kept per example (None: variable).')
    p.Define('num_features', 1, 'Features per laser point.')
    return p

  def PadOrTrim(self, xyz, feature, rng=None):
    p = self.params
    n = len(xyz)
    m = p.max_num_points or n
    if n > m:
      idx = (rng or np.random).permutation(n)[:m]
      idx.sort()
      xyz, feature, n = xyz[idx], feature[idx], m
    pad = np.ones(m, np.float32)
    pad[:n] = 0.0
    out_xyz = np.zeros((m, 3), np.float32)
    out_xyz[:n] = xyz
    out_f = np.zeros((m, p.num_features), np.float32)
    out_f[:n] = feature.reshape(n, -1)[:, :p.num_features]
    return NestedMap(points_xyz=out_xyz, points_feature=out_f, points_padding=pad)

  def Shape(self):
    p = self.params
    m = p.max_num_points
    return NestedMap(points_xyz=(m, 3), points_feature=(m, p.num_features),
                     points_padding=(m,))

  def DType(self):
    return NestedMap(points_xyz=np.float32, points_feature=np.float32,
                     points_padding=np.float32)
