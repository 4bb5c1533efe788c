"""Milan input generator (ref `lingvo/tasks/milan/input_generator.py:26`).

Wraps a user `dataset_fn(batch_size=…, **kwargs)` returning an iterator of batched
feature dicts (e.g. `DatasetSpec.Read`). `features_to_read` filters features by regex;
`preprocessors` maps a feature name to a layer applied to that feature (e.g. the image
preprocessor). String features stay on the host (`cpu_passthrough_keys`).
"""

from __future__ import annotations

import re

import numpy as np
import torch

from lingvo_b200.core import base_input_generator
from lingvo_b200.core.nested_map import NestedMap


class MilanInputGenerator(base_input_generator.BaseInputGenerator):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('dataset_fn', None, 'Callable(batch_size=…, **dataset_fn_kwargs) → iterator.')
    p.Define('dataset_fn_kwargs', {}, 'Extra kwargs for dataset_fn (no batch_size).')
    p.Define('features_to_read', [], 'Regexes of feature names to keep (empty: all).')
    p.Define('preprocessors', {}, 'feature name → layer params.')
    p.Define('preprocess_parallelism', 1, 'Kept for parity.')
    p.Define('drop_string_features', False, 'Drop bytes/str features after preprocessing.')
    p.name = 'milan_input_generator'
    p.batch_size = 32
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    if 'batch_size' in p.dataset_fn_kwargs:
      raise ValueError('dataset_fn_kwargs may not contain "batch_size".')
    if not isinstance(p.features_to_read, (tuple, list, type(None))):
      raise ValueError('Expected sequence type for "features_to_read"; got {}'.format(
          type(p.features_to_read)))
    self._pre_names = []
    if p.preprocessors:
      self._pre_names = sorted(p.preprocessors)
      self.CreateChildren('preprocessors', [
          p.preprocessors[n].Copy().Set(name='pre_%d' % i)
          for i, n in enumerate(self._pre_names)])
    self._iter = None

  def _FilterFeaturesByName(self, features):
    p = self.params
    if not p.features_to_read:
      return features
    rx = re.compile('({})'.format('|'.join(p.features_to_read)))
    return NestedMap({k: v for k, v in features.items() if rx.match(k)})

  def _PreprocessInputBatch(self, input_batch, do_eval=False):
    del do_eval
    p = self.params
    batch = self._FilterFeaturesByName(NestedMap(dict(input_batch)))
    for i, name in enumerate(self._pre_names):
      batch[name] = self.preprocessors[i].FProp(None, batch[name])
    out = NestedMap()
    for k, v in batch.items():
      if isinstance(v, np.ndarray) and v.dtype.kind in 'biuf':
        out[k] = torch.from_numpy(np.ascontiguousarray(v))
      elif isinstance(v, torch.Tensor):
        out[k] = v
      elif not p.drop_string_features:
        out[k] = v
    return out

  def GetPreprocessedInputBatch(self):
    p = self.params
    if self._iter is None:
      self._iter = iter(p.dataset_fn(batch_size=self.InfeedBatchSize(), **p.dataset_fn_kwargs))
    try:
      raw = next(self._iter)
    except StopIteration:
      self._iter = None
      raise
    return self._PreprocessInputBatch(raw, self.do_eval)

  def Reset(self, sess=None):
    self._iter = None
