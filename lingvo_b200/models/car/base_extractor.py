"""Record → extractors → preprocessors → batch (ref `lingvo/tasks/car/base_extractor.py`).

`_BaseExtractor` is an input generator assembled from named `FieldsExtractor`s and an
ordered list of `Preprocessor`s. Each record is parsed once (`tf.Example` wire format, in
the batcher's worker threads), handed to every extractor, dropped if any extractor's
`Filter` says so, then run through the preprocessors; the native batcher stacks the
resulting fixed-shape examples.
"""

from __future__ import annotations

import numpy as np
import torch

from lingvo_b200.core import base_input_generator
from lingvo_b200.core import hyperparams
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.car import input_extractor
from lingvo_b200.utils import tf_example

BUCKET_UPPER_BOUND = input_extractor.BUCKET_UPPER_BOUND


def _ParseExampleDefaults(features, feature_map):
  """Applies shapes / dtypes / defaults of a FeatureMap to a parsed example."""
  out = {}
  for name, (shape, dtype) in feature_map.items():
    if name not in features:
      if shape is None:
        out[name] = [] if dtype is bytes else np.zeros((0,), dtype)
        continue
      raise KeyError('feature "%s" missing (record has %s)' % (name, sorted(features)[:20]))
    v = features[name]
    if dtype is bytes:
      v = list(v)
      out[name] = v[0] if shape == () else v
    else:
      v = np.asarray(v, dtype)
      out[name] = v.reshape(shape) if shape is not None else v
  return out


class _BaseExtractor(base_input_generator.BaseInputGeneratorFromFiles):
  """ref :59."""

  @classmethod
  def Params(cls, extractors=None):
    p = super().Params()
    p.Define('extractors', extractors or hyperparams.Params(),
             'hyperparams.Params of name → FieldsExtractor params.')
    p.Define('preprocessors', hyperparams.Params(), 'name → Preprocessor params.')
    p.Define('preprocessors_order', [], 'Names of `preprocessors` in execution order.')
    p.Define('record_type', 'EXAMPLE', "'EXAMPLE' (tf.Example) or 'TEXT'.")
    p.Define('batched_input', False, 'Records hold whole batches (ExtractBatch path).')
    p.batch_size = 64
    p.file_parallelism = 128
    p.file_buffer_size = 128
    p.file_random_seed = 0
    p.bucket_upper_bound = [BUCKET_UPPER_BOUND - 1]
    p.bucket_batch_limit = [64]
    return p

  def __init__(self, params):
    params = params.Copy()
    if params.batch_size:
      params.bucket_batch_limit = [params.batch_size] * len(params.bucket_upper_bound)
    super().__init__(params)
    p = self.params
    self._extractors = NestedMap()
    for name, ep in p.extractors.IterParams():
      name = name.replace('.', '_')
      self.CreateChild(name, ep)
      self._extractors[name] = self.children[name]
    pre = dict(p.preprocessors.IterParams())
    if not set(p.preprocessors_order).issubset(pre):
      raise ValueError('preprocessor_order specifies keys which were not found in '
                       'preprocessors. preprocessors_order={} preprocessors keys={}'.format(
                           p.preprocessors_order, list(pre)))
    self.CreateChildren('preprocessors', [pre[k] for k in p.preprocessors_order])

  # -- schema -----------------------------------------------------------------------
  def FeatureMap(self):
    out = {}
    for e in self._extractors.values():
      out.update(e.FeatureMap())
    return out

  def ContextMap(self):
    out = {}
    for e in self._extractors.values():
      out.update(e.ContextMap())
    return out

  def Shape(self):
    shapes = self._extractors.Transform(lambda e: e.Shape())
    for pre in self.preprocessors:
      shapes = pre.TransformShapes(shapes)
    return shapes

  def DType(self):
    dtypes = self._extractors.Transform(lambda e: e.DType())
    for pre in self.preprocessors:
      dtypes = pre.TransformDTypes(dtypes)
    return dtypes

  @property
  def class_names(self):
    raise NotImplementedError('Return a list of class names strings.')

  # -- per-record pipeline ----------------------------------------------------------
  def ExtractUsingExtractors(self, record):
    """→ (bucket, NestedMap of per-extractor outputs) for one serialised record."""
    p = self.params
    if p.record_type == 'TEXT':
      features = {'line': record}
    else:
      features = _ParseExampleDefaults(tf_example.ParseExample(record), self.FeatureMap())
    return self.ProcessFeatures(features)

  def ProcessFeatures(self, features):
    buckets, extracted = [], NestedMap()
    for name, e in self._extractors.items():
      keys = set(e.FeatureMap()) | set(e.ContextMap())
      sub = features if self.params.record_type == 'TEXT' else {
          k: v for k, v in features.items() if k in keys}
      try:
        out = e.Extract(sub)
      except Exception as exc:   # pylint: disable=broad-except
        raise RuntimeError('Failed running extractor %s: %r' % (e.params.name, exc)) from exc
      extracted[name] = out
      buckets.append(int(e.Filter(out)))
    bucket = max(buckets) if buckets else 1
    if bucket >= BUCKET_UPPER_BOUND:
      return bucket, None
    as_t = lambda v: torch.from_numpy(np.ascontiguousarray(v)) if isinstance(v, np.ndarray) else v
    feats = extracted.Transform(as_t)
    for pre in self.preprocessors:
      feats = pre.TransformFeatures(feats)
    return bucket, feats

  def ProcessRecord(self, record, source_id=0):
    bucket, feats = self.ExtractUsingExtractors(record)
    if feats is None:
      return None
    to_np = lambda v: v.numpy() if isinstance(v, torch.Tensor) else v
    return feats.Transform(to_np), bucket

  def GetCpuPassthroughKeys(self):
    """Keys of string outputs that stay on the host."""
    keys = []
    for k, dt in self.DType().FlattenItems():
      if dt is bytes or dt is str:
        keys.append(k)
    return keys

  def NestedMapFromBatchedOutputs(self, outputs):
    return outputs if isinstance(outputs, NestedMap) else NestedMap(outputs)
