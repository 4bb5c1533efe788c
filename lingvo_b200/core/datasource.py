"""Data sources (ref `lingvo/core/datasource.py`).

A `DataSource` produces batches for an input generator:
  * `SimpleDataSource` (ref :85) — file pattern(s) (+weights) handed to the input
    generator's `_DataSourceFromFilePattern` (native yielder + batcher).
  * `CrossBatchMixingDataSource` (ref :194) — each batch comes from one of several
    sub-sources, chosen by weight.
  * `CurriculumDataSource` (ref :253) — switches sub-sources at global-step
    boundaries.
  * `PrefixedDataSource` (ref :325) — file patterns relative to a directory.
  * Iterator-style sources (`IteratorDataSource`, `BatchBySequenceLength`,
    `MixerDataSource`, `PrefetchDataSource`) cover the reference's `TFDataset*`
    family (ref :351-780) with plain Python iterables instead of `tf.data`.
"""

from __future__ import annotations

import os
import queue
import threading

import numpy as np

from lingvo_b200.core import base_layer
from lingvo_b200.core import py_utils
from lingvo_b200.core.nested_map import NestedMap


class DataSource(base_layer.BaseLayer):
  """Base (ref :38)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.name = 'datasource'
    return p

  def __init__(self, params):
    super().__init__(params)
    self._input_generator = None

  def SetInputGenerator(self, input_generator):
    self._input_generator = input_generator
    for child in self.children.Flatten() if hasattr(self.children, 'Flatten') else []:
      if isinstance(child, DataSource):
        child.SetInputGenerator(input_generator)

  def Initialize(self, sess=None):
    del sess

  def Reset(self, sess=None):
    del sess

  def GetNext(self):
    raise NotImplementedError

  def GetMeta(self):
    return NestedMap()


class SimpleDataSource(DataSource):
  """File pattern(s) read through the owning input generator (ref :85)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('file_pattern', '', 'str, list of str, or list of (pattern, weight).')
    p.Define('weights', None, 'Weights matching a list file_pattern.')
    p.Define('bprop_variable_filters', None, 'Per-source variable filters (kept for parity).')
    p.Define('file_type', '', 'Prepended as `type:` when the pattern has no type.')
    p.Define('pass_weights_by_param', False, 'Kept for parity.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self._stream = None

  def _Patterns(self):
    p = self.params
    fp = p.file_pattern

    def _Typed(x):
      return x if (':' in x.split('/')[0] or not p.file_type) else '%s:%s' % (p.file_type, x)
    if isinstance(fp, str):
      return _Typed(fp), None
    pats, weights = [], []
    for item in fp:
      if isinstance(item, (list, tuple)):
        pats.append(_Typed(item[0]))
        weights.append(float(item[1]))
      else:
        pats.append(_Typed(item))
    if not weights:
      weights = list(p.weights) if p.weights else [1.0] * len(pats)
    return pats, weights

  def GetNext(self):
    if self._stream is None:
      pats, weights = self._Patterns()
      kwargs = {}
      if weights is not None:
        kwargs['input_source_weights'] = weights
      self._stream = self._input_generator._DataSourceFromFilePattern(  # pylint: disable=protected-access
          pats, **kwargs)
    nxt = self._stream
    return nxt() if callable(nxt) else next(nxt)

  def Reset(self, sess=None):
    self._stream = None

  def GetMeta(self):
    p = self.params
    ret = NestedMap()
    if p.bprop_variable_filters:
      ret.bprop_variable_filters = p.bprop_variable_filters
    return ret


class PrefixedDataSource(SimpleDataSource):
  """Patterns are relative to `file_pattern_prefix` (ref :325)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('file_pattern_prefix', '', 'Directory prefix.')
    return p

  def _Patterns(self):
    p = self.params
    pats, weights = super()._Patterns()

    def _Pre(x):
      if ':' in x.split('/')[0]:
        t, rest = x.split(':', 1)
        return '%s:%s' % (t, ','.join(os.path.join(p.file_pattern_prefix, r)
                                      for r in rest.split(',')))
      return ','.join(os.path.join(p.file_pattern_prefix, r) for r in x.split(','))
    if isinstance(pats, str):
      return _Pre(pats), weights
    return [_Pre(x) for x in pats], weights


class CrossBatchMixingDataSource(DataSource):
  """Every batch is drawn from one sub-source, chosen with probability ∝ weight
  (ref :194)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('sub', None, 'List of DataSource params.')
    p.Define('weights', None, 'List of weights (or schedule layers).')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    assert p.sub and len(p.sub) == len(p.weights)
    self.CreateChildren('sub', list(p.sub))
    self._rng = np.random.RandomState(p.random_seed)

  def SetInputGenerator(self, input_generator):
    self._input_generator = input_generator
    for s in self.sub:
      s.SetInputGenerator(input_generator)

  def _Weights(self):
    ws = []
    for w in self.params.weights:
      ws.append(float(w.Value()) if hasattr(w, 'Value') else float(w))
    ws = np.asarray(ws, np.float64)
    return ws / ws.sum()

  def GetNext(self):
    i = int(self._rng.choice(len(self.sub), p=self._Weights()))
    batch = self.sub[i].GetNext()
    if isinstance(batch, NestedMap):
      batch.source_selected = np.asarray([i], np.int32)
    return batch

  def Reset(self, sess=None):
    for s in self.sub:
      s.Reset()


class CurriculumDataSource(DataSource):
  """Sub-source k is active for global_step in [boundaries[k-1], boundaries[k])
  (ref :253)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('sub', None, 'List of DataSource params.')
    p.Define('boundaries', None, 'Global-step boundaries, len(sub) - 1 of them.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    assert len(p.boundaries) == len(p.sub) - 1
    assert list(p.boundaries) == sorted(p.boundaries)
    self.CreateChildren('sub', list(p.sub))

  def SetInputGenerator(self, input_generator):
    self._input_generator = input_generator
    for s in self.sub:
      s.SetInputGenerator(input_generator)

  def GetNext(self):
    step = int(py_utils.GetGlobalStep())
    k = int(np.searchsorted(np.asarray(self.params.boundaries), step, side='right'))
    return self.sub[k].GetNext()


# ------------------------------------------------------------------ iterator sources --
class IteratorDataSource(DataSource):
  """Wraps `iter_fn()` → iterable of example NestedMaps (the `TFDatasetFnInput` /
  `TFDatasetAdaptor` role, ref :429, :558)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('iter_fn', None, 'Callable returning an iterable of examples/batches.')
    p.Define('repeat', True, 'Restart the iterator when exhausted.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self._it = None

  def GetNext(self):
    if self._it is None:
      self._it = iter(self.params.iter_fn())
    try:
      return next(self._it)
    except StopIteration:
      if not self.params.repeat:
        raise
      self._it = iter(self.params.iter_fn())
      return next(self._it)

  def Reset(self, sess=None):
    self._it = None


class _Transform(DataSource):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('sub', None, 'Upstream DataSource params.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self.CreateChild('sub', self.params.sub)

  def SetInputGenerator(self, input_generator):
    self._input_generator = input_generator
    self.sub.SetInputGenerator(input_generator)

  def Reset(self, sess=None):
    self.sub.Reset()


class CustomTransform(_Transform):
  """Applies `fn(example)` to every upstream element (ref :473)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('fn', None, 'Callable example → example (None drops it).')
    return p

  def GetNext(self):
    while True:
      out = self.params.fn(self.sub.GetNext())
      if out is not None:
        return out


class BatchBySequenceLength(_Transform):
  """Buckets upstream *examples* by `seqlen_fn(example)` and emits padded batches
  (ref `TFDatasetBatchBySequenceLength` :595)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('seqlen_fn', None, 'example → int length.')
    p.Define('bucket_upper_bound', [], 'Upper bounds.')
    p.Define('bucket_batch_limit', [], 'Batch sizes.')
    p.Define('require_sequential_order', False, 'Kept for parity.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self._buckets = [[] for _ in self.params.bucket_upper_bound]

  def _Merge(self, examples, keys):
    flat = [e.Flatten() for e in examples]
    outs = []
    for j in range(len(flat[0])):
      arrs = [np.asarray(f[j]) for f in flat]
      shape = np.max([a.shape for a in arrs], axis=0) if arrs[0].ndim else ()
      out = np.zeros((len(arrs),) + tuple(int(s) for s in shape), arrs[0].dtype)
      for i, a in enumerate(arrs):
        out[(i,) + tuple(slice(0, s) for s in a.shape)] = a
      outs.append(out)
    batch = examples[0].Pack(outs)
    batch.bucket_keys = np.asarray(keys, np.int32)
    return batch

  def GetNext(self):
    p = self.params
    while True:
      ex = self.sub.GetNext()
      n = int(p.seqlen_fn(ex))
      k = int(np.searchsorted(np.asarray(p.bucket_upper_bound), n, side='left'))
      if k >= len(self._buckets):
        continue
      self._buckets[k].append((ex, n))
      if len(self._buckets[k]) >= p.bucket_batch_limit[k]:
        items, self._buckets[k] = self._buckets[k], []
        return self._Merge([e for e, _ in items], [m for _, m in items])


class MixerDataSource(DataSource):
  """Element-level weighted mix of sub-sources, tagging `source_id`
  (ref `TFDatasetMixer` :707)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('sub', None, 'List of DataSource params.')
    p.Define('weights', None, 'Sampling weights.')
    p.Define('broadcast_dataset_structures', False, 'Kept for parity.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    self.CreateChildren('sub', list(p.sub))
    w = np.asarray(p.weights or [1.0] * len(p.sub), np.float64)
    self._w = w / w.sum()
    self._rng = np.random.RandomState(p.random_seed)

  def SetInputGenerator(self, input_generator):
    self._input_generator = input_generator
    for s in self.sub:
      s.SetInputGenerator(input_generator)

  def GetNext(self):
    i = int(self._rng.choice(len(self.sub), p=self._w))
    ex = self.sub[i].GetNext()
    if isinstance(ex, NestedMap):
      ex.source_id = np.int32(i)
    return ex


class PrefetchDataSource(_Transform):
  """Background-thread prefetch of `buffer_size` elements (ref `TFDatasetPrefetch` :695)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('buffer_size', 2, 'Elements kept ready.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self._q = None
    self._thread = None

  def _Run(self):
    while True:
      try:
        self._q.put(('ok', self.sub.GetNext()))
      except StopIteration:
        self._q.put(('eof', None))
        return
      except Exception as e:  # pylint: disable=broad-except
        self._q.put(('err', e))
        return

  def GetNext(self):
    if self._q is None:
      self._q = queue.Queue(maxsize=max(self.params.buffer_size, 1))
      self._thread = threading.Thread(target=self._Run, daemon=True)
      self._thread.start()
    kind, val = self._q.get()
    if kind == 'ok':
      return val
    if kind == 'eof':
      raise StopIteration()
    raise val
