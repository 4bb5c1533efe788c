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
    self._ig_ref = [None]

  # The owning input generator is a back-reference, not a child layer: it is kept in a
  # one-element holder so that the layer tree's "every BaseLayer attribute is a registered
  # child" check (and theta / vars traversals) do not see it.
  @property
  def _input_generator(self):
    return self._ig_ref[0]

  @_input_generator.setter
  def _input_generator(self, value):
    self._ig_ref[0] = value

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
    p.Define('source_id_offset', 0, 'Added to the `source_id` of every batch: gives the '
             'sub-sources of a cross-batch mixer distinct ids (ref :104).')
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
    batch = nxt() if callable(nxt) else next(nxt)
    off = self.params.source_id_offset
    if off and isinstance(batch, NestedMap) and 'source_id' in batch:
      batch.source_id = batch.source_id + off
    return batch

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
    p.Define('bprop_variable_filters', None,
             'One variable-name regex per sub-source: a batch drawn from source i only '
             'updates the variables matching filter i (read by the learner via GetMeta).')
    return p

  def GetMeta(self):
    ret = NestedMap()
    if self.params.bprop_variable_filters:
      ret.bprop_variable_filters = list(self.params.bprop_variable_filters)
    return ret

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    assert p.sub and len(p.sub) == len(p.weights)
    assert not p.bprop_variable_filters or len(p.bprop_variable_filters) == len(p.sub)
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


# =========================================================================================
# The `tf.data`-style family (ref :351-900) on a small lazy dataset algebra.
# =========================================================================================
class Dataset:
  """A re-iterable, lazily evaluated stream of elements — the slice of `tf.data.Dataset`
  the reference's `TFDataset*` sources rely on. Every transformation returns a new `Dataset`;
  `iter(ds)` starts a fresh pass, so `Reset()` is just taking a new iterator. Elements are
  example / batch `NestedMap`s of numpy arrays (host memory: the device prefetcher moves
  assembled batches to the GPU later)."""

  def __init__(self, make_iter):
    self._make_iter = make_iter

  def __iter__(self):
    return iter(self._make_iter())

  # -- constructors ------------------------------------------------------------------
  @classmethod
  def FromGenerator(cls, fn):
    return cls(fn)

  @classmethod
  def FromElements(cls, elements):
    elements = list(elements)
    return cls(lambda: iter(elements))

  @classmethod
  def SampleFrom(cls, datasets, weights=None, seed=None):
    """Element-level weighted mix; ends when every source is exhausted."""
    w = np.asarray(weights if weights is not None else [1.0] * len(datasets), np.float64)

    def Gen():
      rng = np.random.RandomState(seed)
      its = [iter(d) for d in datasets]
      alive = w.copy()
      while alive.sum() > 0:
        i = int(rng.choice(len(its), p=alive / alive.sum()))
        try:
          yield next(its[i])
        except StopIteration:
          alive[i] = 0.0
    return cls(Gen)

  # -- transformations ---------------------------------------------------------------
  def map(self, fn):   # pylint: disable=invalid-name
    return Dataset(lambda: (fn(x) for x in self))

  def filter(self, pred):   # pylint: disable=invalid-name
    return Dataset(lambda: (x for x in self if pred(x)))

  def take(self, n):   # pylint: disable=invalid-name
    def Gen():
      for i, x in enumerate(self):
        if i >= n:
          return
        yield x
    return Dataset(Gen)

  def repeat(self, count=None):   # pylint: disable=invalid-name
    def Gen():
      k = 0
      while count is None or k < count:
        empty = True
        for x in self:
          empty = False
          yield x
        if empty:
          return
        k += 1
    return Dataset(Gen)

  def concatenate(self, other):   # pylint: disable=invalid-name
    def Gen():
      yield from self
      yield from other
    return Dataset(Gen)

  def shard(self, num_shards, index):   # pylint: disable=invalid-name
    return Dataset(lambda: (x for i, x in enumerate(self) if i % num_shards == index))

  def shuffle(self, buffer_size, seed=None):   # pylint: disable=invalid-name
    """Buffer shuffle, reshuffled on every pass (`reshuffle_each_iteration=True`)."""
    epoch = [0]

    def Gen():
      rng = np.random.RandomState(None if seed is None else seed + epoch[0])
      epoch[0] += 1
      buf = []
      for x in self:
        if len(buf) < buffer_size:
          buf.append(x)
          continue
        j = rng.randint(len(buf))
        buf[j], x = x, buf[j]
        yield x
      rng.shuffle(buf)
      yield from buf
    return Dataset(Gen)

  def prefetch(self, buffer_size):   # pylint: disable=invalid-name
    """Produces elements on a background thread, `buffer_size` ahead of the consumer."""
    def Gen():
      q = queue.Queue(maxsize=max(int(buffer_size), 1))
      done = object()

      def Run():
        try:
          for x in self:
            q.put(('ok', x))
          q.put(('eof', done))
        except Exception as e:  # pylint: disable=broad-except
          q.put(('err', e))

      threading.Thread(target=Run, daemon=True).start()
      while True:
        kind, val = q.get()
        if kind == 'ok':
          yield val
        elif kind == 'eof':
          return
        else:
          raise val
    return Dataset(Gen)

  def bucket_by_sequence_length(self, length_fn, boundaries, batch_sizes, padded_shapes=None,   # pylint: disable=invalid-name
                                padding_values=None, pad_to_bucket_boundary=True,
                                drop_remainder=False):
    """Groups examples into length buckets (`length <= boundaries[k]`), emits a padded batch
    when a bucket reaches its batch size; leftovers are flushed at the end unless
    `drop_remainder`. Variable dims (None in `padded_shapes`, or all of them without shapes)
    are padded to the bucket boundary or to the longest example of the batch."""
    def Merge(examples, bound):
      flat_keys = [k for k, _ in examples[0].FlattenItems()]
      out = NestedMap()
      for k in flat_keys:
        arrs = [np.asarray(e.GetItem(k)) for e in examples]
        spec = padded_shapes.GetItem(k) if padded_shapes is not None else None
        nd = arrs[0].ndim
        tgt = []
        for d in range(nd):
          longest = max(a.shape[d] for a in arrs)
          want = None if spec is None else spec[d]
          if want is None:
            want = max(bound, longest) if (pad_to_bucket_boundary and d == 0) else longest
          tgt.append(int(want))
        fill = 0 if padding_values is None else padding_values.GetItem(k)
        batch = np.full((len(arrs),) + tuple(tgt), fill, arrs[0].dtype)
        for i, a in enumerate(arrs):
          batch[(i,) + tuple(slice(0, s) for s in a.shape)] = a
        out.Set(k, batch)
      return out

    def Gen():
      buckets = [[] for _ in boundaries]
      for ex in self:
        n = int(length_fn(ex))
        k = int(np.searchsorted(np.asarray(boundaries), n, side='left'))
        if k >= len(buckets):
          continue                                  # longer than the last bucket: dropped
        buckets[k].append(ex)
        if len(buckets[k]) >= batch_sizes[k]:
          items, buckets[k] = buckets[k], []
          yield Merge(items, boundaries[k])
      if not drop_remainder:
        for k, items in enumerate(buckets):
          if items:
            yield Merge(items, boundaries[k])
    return Dataset(Gen)


def _PeekFirst(dataset):
  """(first element or None, equivalent dataset). Sources adapted from stateful `GetNext`
  objects are not re-iterable, so the peeked pass is handed on — first element re-attached —
  as the returned dataset's first pass; later passes iterate `dataset` afresh."""
  it = iter(dataset)
  first = next(it, None)
  pending = [it]

  def Gen():
    if pending:
      rest = pending.pop()
      if first is not None:
        yield first
      yield from rest
    else:
      yield from dataset
  return first, Dataset(Gen)


class RepeatSentinelError(RuntimeError):
  """A `RepeatableTFDatasetTransform` reached the end of an epoch (message carries the
  reference's `REPEAT_SENTINEL_` marker, which eval loops look for)."""


class TFDatasetSource(DataSource):
  """Base of the dataset-backed sources (ref :351): subclasses build a `Dataset` in
  `GetDataset()`; this class owns the iterator, per-host sharding and Reset."""

  def __init__(self, params):
    super().__init__(params)
    self._dataset = {}
    self._iterator = {}

  @property
  def num_hosts(self):
    """Input replicas (one per training process when inputs are sharded by rank)."""
    ig = self._input_generator
    if ig is not None and getattr(ig.params, 'use_per_host_infeed', False):
      try:
        from lingvo_b200.core import generic_input   # pylint: disable=g-import-not-at-top
        return max(int(generic_input.ReplicaInfo()[0]), 1)
      except Exception:  # pylint: disable=broad-except
        return 1
    return 1

  @property
  def host_id(self):
    if self.num_hosts > 1:
      from lingvo_b200.core import generic_input   # pylint: disable=g-import-not-at-top
      return int(generic_input.ReplicaInfo()[1])
    return 0

  def GetDataset(self) -> Dataset:
    raise NotImplementedError()

  def _InitIterator(self):
    if self.host_id in self._dataset:
      return
    with py_utils.GlobalStepContext(None):          # datasets never capture the step
      ds = self.GetDataset()
    self._dataset[self.host_id] = ds
    self._iterator[self.host_id] = iter(ds)

  def Reset(self, sess=None):
    self._iterator = {k: iter(ds) for k, ds in self._dataset.items()}
    super().Reset(sess)

  def GetNext(self):
    self._InitIterator()
    return next(self._iterator[self.host_id])


class TFDatasetAdaptor(TFDatasetSource):
  """Presents any `DataSource` as an (endless) dataset (ref :429)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('sub', None, 'A DataSource to adapt.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self.CreateChild('sub', self.params.sub)

  def SetInputGenerator(self, input_generator):
    self._input_generator = input_generator
    self.sub.SetInputGenerator(input_generator)

  def GetDataset(self):
    def Gen():
      while True:
        try:
          yield self.sub.GetNext()
        except StopIteration:
          return
    return Dataset(Gen)

  def Reset(self, sess=None):
    self.sub.Reset()
    super().Reset(sess)


class TFDatasetTransform(TFDatasetSource):
  """Transforms the dataset of a child source (ref :446); non-dataset children are adapted."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('sub', None, 'A DataSource; wrapped in TFDatasetAdaptor if it is not '
             'dataset-backed.')
    return p

  def __init__(self, params):
    super().__init__(params)
    ds = self.params.sub
    if not issubclass(ds.cls, TFDatasetSource):
      ds = TFDatasetAdaptor.Params().Set(sub=ds)
    self.CreateChild('sub', ds)

  def SetInputGenerator(self, input_generator):
    self._input_generator = input_generator
    self.sub.SetInputGenerator(input_generator)

  def GetDataset(self):
    return self.Transform(self.sub.GetDataset())

  def Transform(self, dataset):
    raise NotImplementedError()

  def Reset(self, sess=None):
    self.sub.Reset()
    super().Reset(sess)


class CustomTFDatasetTransform(TFDatasetTransform):
  """Transforms with a method of the input generator: `fn(dataset, **kwargs) → dataset`
  (ref :473)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('fn', '', 'Name of the input-generator method to call.')
    p.Define('kwargs', None, 'Keyword arguments for fn.')
    return p

  def Transform(self, dataset):
    fn = getattr(self._input_generator, self.params.fn)
    return fn(dataset, **(self.params.kwargs or {}))


class RepeatableTFDatasetTransform(TFDatasetTransform):
  """Repeat policy owned by the input generator (ref :494): `repeat_steps` replays the first
  N batches forever; `repeat_with_sentinel` appends one all-zero batch whose `sentinel_key`
  holds `sentinel_value` after every epoch and repeats — `GetNext` raises
  `RepeatSentinelError('REPEAT_SENTINEL_')` on it, which is how eval loops find the end of an
  epoch without recreating the pipeline."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('sentinel_key', 'bucket_keys', 'Key overwritten in the sentinel batch.')
    p.Define('sentinel_value', -1, 'Impossible value marking the sentinel batch.')
    return p

  def GetDataset(self):
    p = self.params
    ig = self._input_generator
    self._repeat_steps = getattr(ig.params, 'repeat_steps', None) if ig is not None else None
    self._repeat_with_sentinel = (getattr(ig.params, 'repeat_with_sentinel', None)
                                  if ig is not None else None)
    ds = super().GetDataset()
    if self._repeat_steps:
      return ds.take(self._repeat_steps).repeat()
    if self._repeat_with_sentinel:
      def Gen():
        while True:
          last = None
          for x in ds:
            last = x
            yield x
          if last is None:
            return
          sentinel = last.Transform(lambda a: np.zeros_like(np.asarray(a)))
          sentinel.Set(p.sentinel_key, np.full_like(np.asarray(last.GetItem(p.sentinel_key)),
                                                    p.sentinel_value))
          yield sentinel
      return Dataset(Gen)
    return ds

  def GetNext(self):
    batch = super().GetNext()
    if self._repeat_with_sentinel and not self._repeat_steps:
      if np.any(np.asarray(batch.GetItem(self.params.sentinel_key)) ==
                self.params.sentinel_value):
        raise RepeatSentinelError('REPEAT_SENTINEL_')
    return batch

  def Transform(self, dataset):
    return dataset


class TFDatasetFnInput(TFDatasetSource):
  """Loads a dataset with a method of the input generator (ref :558): shuffled with a
  buffer unless sequential order is required, repeated unless evaluating."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('load_fn', 'LoadDataset', 'Input-generator method returning a Dataset (or any '
             're-iterable).')
    p.Define('kwargs', None, 'Keyword arguments for load_fn.')
    p.Define('shuffle_buffer_size', None, 'Records buffered for shuffling.')
    return p

  def __init__(self, params):
    super().__init__(params)
    if (not self.params.shuffle_buffer_size and
        not self.cluster.require_sequential_input_order):
      raise ValueError('shuffle_buffer_size must be set.')

  def GetDataset(self):
    p = self.params
    ds = getattr(self._input_generator, p.load_fn)(**(p.kwargs or {}))
    if not isinstance(ds, Dataset):
      src = ds
      ds = Dataset(lambda: iter(src() if callable(src) else src))
    if self.num_hosts > 1:
      ds = ds.shard(self.num_hosts, self.host_id)
    if not self.cluster.require_sequential_input_order:
      ds = ds.shuffle(p.shuffle_buffer_size, seed=p.random_seed)
    if not self.do_eval:
      ds = ds.repeat()
    return ds


class TFDatasetBatchBySequenceLength(TFDatasetTransform):
  """Batches unbatched examples by length buckets (ref :595). The input generator supplies
  `seqlen_fn(example)`, `input_shape_fn(key)` (None dims are padded to the bucket boundary)
  and `input_padding_fn(key, spec)`; each example gains `bucket_keys`."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('seqlen_fn', 'GetSequenceLength', 'example → sequence length.')
    p.Define('input_shape_fn', '_InputShape', 'tensor name → shape (None = variable).')
    p.Define('input_padding_fn', '_InputPaddingValue', '(name, spec) → padding value.')
    p.Define('bucket_upper_bound', [], 'Sorted bucket upper bounds; longer examples are '
             'filtered out.')
    p.Define('bucket_batch_limit', [], 'Batch size per bucket.')
    return p

  def Transform(self, dataset):
    p = self.params
    ig = self._input_generator
    seqlen_fn = getattr(ig, p.seqlen_fn)

    def SetBucketKeys(example):
      example.bucket_keys = np.int32(seqlen_fn(example))
      return example

    dataset = dataset.map(SetBucketKeys).filter(
        lambda x: x.bucket_keys <= p.bucket_upper_bound[-1])
    shape_fn = getattr(ig, p.input_shape_fn, None)
    pad_fn = getattr(ig, p.input_padding_fn, None)
    first, dataset = _PeekFirst(dataset)
    padded_shapes = padding_values = None
    if first is not None and shape_fn is not None:
      padded_shapes = NestedMap()
      for k, v in first.FlattenItems():
        padded_shapes.Set(k, tuple(shape_fn(k)) if np.asarray(v).ndim else ())
    if first is not None and pad_fn is not None:
      padding_values = NestedMap()
      for k, v in first.FlattenItems():
        padding_values.Set(k, pad_fn(k, np.asarray(v)))
    return dataset.bucket_by_sequence_length(
        lambda x: x.bucket_keys, list(p.bucket_upper_bound), list(p.bucket_batch_limit),
        padded_shapes=padded_shapes, padding_values=padding_values,
        pad_to_bucket_boundary=True,
        drop_remainder=not self.cluster.require_sequential_input_order)


class TFDatasetPrefetch(TFDatasetTransform):
  """Background prefetch of `buffer_size` elements (ref :688)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('buffer_size', 1, 'Prefetch buffer size.')
    return p

  def Transform(self, dataset):
    return dataset.prefetch(self.params.buffer_size)


class TFDatasetMixer(TFDatasetSource):
  """Element-level weighted mix of several dataset sources, tagging `source_id`
  (ref :699). `broadcast_dataset_structures` adds keys a source lacks as zeros of the
  shape / dtype another source uses (unknown dims become 1)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('sub', None, 'A list of TFDatasetSource params to mix.')
    p.Define('weights', None, 'Sampling weight of each source.')
    p.Define('broadcast_dataset_structures', False,
             'Make the element structures of the sources compatible.')
    return p

  def __init__(self, params):
    super().__init__(params)
    subs = []
    for sp in self.params.sub:
      subs.append(sp if issubclass(sp.cls, TFDatasetSource)
                  else TFDatasetAdaptor.Params().Set(sub=sp))
    self.CreateChildren('sub', subs)

  def SetInputGenerator(self, input_generator):
    self._input_generator = input_generator
    for s in self.sub:
      s.SetInputGenerator(input_generator)

  def GetDataset(self):
    p = self.params
    datasets = [s.GetDataset() for s in self.sub]

    def SetSourceId(i):
      def Fn(element):
        element.source_id = np.int32(i)
        return element
      return Fn

    datasets = [d.map(SetSourceId(i)) for i, d in enumerate(datasets)]
    if len(datasets) == 1:
      return datasets[0]
    if p.broadcast_dataset_structures:
      expected = {}
      for i, d in enumerate(datasets):
        first, datasets[i] = _PeekFirst(d)
        if first is None:
          continue
        for k, v in first.FlattenItems():
          a = np.asarray(v)
          if k in expected and expected[k][1] != a.dtype:
            raise ValueError('Incompatible dataset specs for key %s: %s vs %s' %
                             (k, expected[k][1], a.dtype))
          expected.setdefault(k, (tuple(1 for _ in a.shape), a.dtype))

      def Broadcast(element):
        for k, (shape, dtype) in expected.items():
          if not element.Has(k):
            element.Set(k, np.zeros(shape, dtype))
        return element

      datasets = [d.map(Broadcast) for d in datasets]
    return Dataset.SampleFrom(datasets, p.weights, p.random_seed)

  def Reset(self, sess=None):
    for s in self.sub:
      s.Reset()
    super().Reset(sess)


class TFDataServiceSource(TFDatasetTransform):
  """Input processing on a pool of background workers (ref `TFDataServiceSource` :868 — the
  tf.data service moves input pre-processing off the trainer's critical path).

  Here the "service" is a pool of `num_workers` threads, each running its own pass over a
  disjoint shard of the sub-dataset (`distributed_epoch`-style: every element is produced
  exactly once per epoch) and feeding one bounded queue; the trainer thread only dequeues.
  The native record yielders / batcher underneath release the GIL, so the workers scale on
  the host cores while the GPU step runs."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('num_workers', 4, 'Background workers.')
    p.Define('buffer_size', 8, 'Elements buffered between workers and the consumer.')
    p.Define('bucket_upper_bound', None, 'Kept for parity (bucketing happens upstream).')
    return p

  def Transform(self, dataset):
    p = self.params
    n = max(int(p.num_workers), 1)

    def Gen():
      q = queue.Queue(maxsize=max(int(p.buffer_size), 1))
      live = [n]
      lock = threading.Lock()

      def Work(i):
        try:
          for x in dataset.shard(n, i):
            q.put(('ok', x))
        except Exception as e:  # pylint: disable=broad-except
          q.put(('err', e))
        finally:
          with lock:
            live[0] -= 1
            if live[0] == 0:
              q.put(('eof', None))

      for i in range(n):
        threading.Thread(target=Work, args=(i,), daemon=True).start()
      while True:
        kind, val = q.get()
        if kind == 'ok':
          yield val
        elif kind == 'eof':
          return
        else:
          raise val
    return Dataset(Gen)
