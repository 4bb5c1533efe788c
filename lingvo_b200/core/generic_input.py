"""GenericInput: records → processor → bucketed, padded batches
(ref `lingvo/core/generic_input.py`, native `generic_input_op_kernels.cc`,
`record_batcher.cc`, `record_yielder.cc`).

    batcher = GenericInput(
        file_pattern='tfrecord:/data/train-*', processor=fn,
        bucket_upper_bound=[32, 64], bucket_batch_limit=[16, 8], ...)
    tensors, bucket_keys = batcher.GetNext()

`processor(record: bytes[, source_id]) -> (NestedMap | list of np arrays,
bucket_key:int)` or `None` to drop the example. The heavy lifting (file
globbing, shuffled multi-threaded reading, epochs, bucketing, padding+stacking)
is native C++ (`ops/csrc_host`); the processor runs under the GIL on the
batcher's worker threads.
"""

from __future__ import annotations

import atexit
import inspect
import weakref
from typing import Callable, List, Optional, Sequence

import numpy as np

from lingvo_b200 import ops
from lingvo_b200.core.nested_map import NestedMap


def ReplicaInfo():
  """(num_input_replicas, input_replica_id) of this process: one data-parallel rank per
  GPU under torchrun, so every rank reads a disjoint shard of the records (reference
  `record_yielder.h:84-85`; there the trainer passes its replica index explicitly)."""
  try:
    import torch.distributed as dist  # pylint: disable=g-import-not-at-top
    if dist.is_available() and dist.is_initialized():
      return dist.get_world_size(), dist.get_rank()
  except Exception:  # pylint: disable=broad-except
    pass
  return 1, 0


def MakeYielder(file_pattern, file_random_seed=0, file_buffer_size=10000,
                file_parallelism=4, repeat_count=-1, require_sequential_order=False,
                input_source_weights: Optional[Sequence[float]] = None,
                num_input_replicas: Optional[int] = None,
                input_replica_id: Optional[int] = None,
                file_buffer_size_in_seconds: float = 0.0):
  """Yielder for a `type:glob` pattern or a list of patterns (weighted mix).
  `num_input_replicas` / `input_replica_id` default to this process's data-parallel
  coordinates (`ReplicaInfo`)."""
  h = ops.host()
  if num_input_replicas is None or input_replica_id is None:
    num_input_replicas, input_replica_id = ReplicaInfo()
  if isinstance(file_pattern, (list, tuple)):
    pats = list(file_pattern)
    weights = list(input_source_weights) if input_source_weights else [1.0] * len(pats)
    if pats and isinstance(pats[0], (list, tuple)):      # [(pattern, weight), …]
      weights = [w for _, w in pats]
      pats = [p for p, _ in pats]
    kids = [MakeYielder(p, file_random_seed + i if file_random_seed else 0,
                        file_buffer_size, file_parallelism, repeat_count,
                        require_sequential_order, None, num_input_replicas,
                        input_replica_id, file_buffer_size_in_seconds)
            for i, p in enumerate(pats)]
    # every replica must draw its own mixing sequence, or all would pick the same source
    mix_seed = file_random_seed * 1000003 + input_replica_id + 1 if file_random_seed else 0
    return h.weighted_mix_record_yielder(kids, weights, mix_seed)
  if require_sequential_order:
    return h.sequential_record_yielder(
        file_pattern, repeat_count if repeat_count > 0 else -1, 0,
        num_input_replicas, input_replica_id)
  return h.basic_record_yielder(
      file_pattern, seed=file_random_seed, bufsize=file_buffer_size,
      parallelism=file_parallelism,
      num_epochs=repeat_count if repeat_count > 0 else 0,
      num_input_replicas=num_input_replicas, input_replica_id=input_replica_id,
      bufsize_in_seconds=float(file_buffer_size_in_seconds or 0.0))


# Worker threads of the native batcher call back into Python (the record processor). A worker
# that is still alive when the interpreter finalises gets killed while it waits for the GIL,
# which aborts the process ("terminate called without an active exception"). Every live
# pipeline is therefore closed from an atexit hook, i.e. *before* finalisation starts.
_LIVE_PIPELINES = weakref.WeakSet()


def _CloseLivePipelines():
  for gi in list(_LIVE_PIPELINES):
    try:
      gi.Close()
    except Exception:  # pylint: disable=broad-except
      pass


atexit.register(_CloseLivePipelines)


class GenericInput:
  """Bucketing batcher over a record yielder."""

  def __init__(self, processor: Callable, file_pattern=None, yielder=None,
               bucket_upper_bound=(1 << 30,), bucket_batch_limit=(1,),
               file_random_seed=0, file_buffer_size=10000, file_parallelism=4,
               num_threads=4, flush_every_n=0, repeat_count=-1,
               require_sequential_order=False, input_source_weights=None,
               bucket_adjust_every_n=0, fatal_errors=None,
               dynamic_padding_dimensions=None, dynamic_padding_constants=None,
               num_input_replicas=None, input_replica_id=None,
               file_buffer_size_in_seconds=0.0):
    """`bucket_adjust_every_n`: native `BucketAdjuster` re-optimises the bucket bounds.
    `fatal_errors`: None ⇒ any processor exception aborts; a list ⇒ only matching ones do,
    the rest skip the record. `dynamic_padding_constants`: per flattened tensor slot pad
    value (list, or a NestedMap / dict matching the sample); `dynamic_padding_dimensions`
    is accepted for parity (every dimension is padded to the batch maximum)."""
    self._h = ops.host()
    del dynamic_padding_dimensions
    self._yielder = yielder or MakeYielder(
        file_pattern, file_random_seed, file_buffer_size, file_parallelism,
        repeat_count, require_sequential_order, input_source_weights,
        num_input_replicas, input_replica_id, file_buffer_size_in_seconds)
    pad_values = []
    if dynamic_padding_constants is not None:
      if isinstance(dynamic_padding_constants, NestedMap):
        pad_values = [float(v) for v in dynamic_padding_constants.Flatten()]
      elif isinstance(dynamic_padding_constants, dict):
        pad_values = [float(v) for v in NestedMap(dynamic_padding_constants).Flatten()]
      else:
        pad_values = [float(v) for v in dynamic_padding_constants]
    self._template = None
    takes_source = len(inspect.signature(processor).parameters) >= 2

    def _Proc(record, source_id):
      out = processor(record, source_id) if takes_source else processor(record)
      if out is None:
        return None
      tensors, key = out
      if isinstance(tensors, NestedMap):
        if self._template is None:
          self._template = tensors.Transform(lambda _: None)
        flat = tensors.Flatten()
      else:
        flat = list(tensors)
      return int(key), [np.asarray(t) for t in flat]

    self._batcher = self._h.RecordBatcher(
        self._yielder, _Proc, [int(b) for b in bucket_upper_bound],
        [int(b) for b in bucket_batch_limit],
        1 if require_sequential_order else num_threads, flush_every_n,
        int(bucket_adjust_every_n or 0),
        None if fatal_errors is None else [str(e) for e in fatal_errors], pad_values)
    _LIVE_PIPELINES.add(self)

  def GetNext(self):
    """→ (NestedMap or list of batched np arrays, bucket_keys [n])."""
    keys, outs = self._batcher.get_next()
    outs = list(outs)
    if self._template is not None:
      return self._template.Pack(outs), keys
    return outs, keys

  def __iter__(self):
    while True:
      try:
        yield self.GetNext()
      except StopIteration:
        return

  @property
  def records_skipped(self):
    return self._batcher.records_skipped

  @property
  def records_failed(self):
    return self._batcher.records_failed

  @property
  def bucket_upper_bound(self):
    """Current bounds (they move when `bucket_adjust_every_n` is set)."""
    return list(self._batcher.bucket_upper_bound)

  def Close(self):
    self._batcher.close()


# -- keyed pipelines + replicated input (ref generic_input.py:28-47, 211-400) -----------------
_GENERIC_CACHE_V2 = {}
_ALLOW_V2_IN_EAGER = [True]


def SetAllowGenericInputV2InEager(allowed=True):
  _ALLOW_V2_IN_EAGER[0] = bool(allowed)


def IsGenericInputV2AllwedInEager():   # (sic) the reference's spelling
  return _ALLOW_V2_IN_EAGER[0]


def GenericInputV2Create(processor, **kwargs):
  """Creates — or returns the cached — pipeline for `generic_input_v2_key`. The key makes the
  pipeline a process-level resource: re-building the input generator (e.g. eval after train,
  program re-instantiation) keeps reading where the previous owner stopped (ref :211).
  Returns (resource, out_types, output_tmpl); `GenericInputV2GetNext` reads from it."""
  key = kwargs.pop('generic_input_v2_key', None)
  if key is None:
    raise RuntimeError('GenericInputV2Create needs a `generic_input_v2_key` identifying the '
                       'pipeline (any hashable, e.g. the input generator\'s path).')
  if key not in _GENERIC_CACHE_V2:
    _GENERIC_CACHE_V2[key] = GenericInput(processor, **kwargs)
  resource = _GENERIC_CACHE_V2[key]
  return resource, None, None


def GenericInputV2GetNext(resource, out_types=None, output_tmpl=None):
  del out_types, output_tmpl
  return resource.GetNext()


def ResetGenericInputV2Cache(key=None):
  """Closes and forgets one keyed pipeline (or all)."""
  for k in ([key] if key is not None else list(_GENERIC_CACHE_V2)):
    gi = _GENERIC_CACHE_V2.pop(k, None)
    if gi is not None:
      gi.Close()


class ReplicatedGenericInput:
  """`num_replicas` independent pipelines — replica i reads input shard i of `num_replicas`
  — whose batches are concatenated along the batch dim: one process feeding several
  model replicas (ref :319). All bucket batch limits must be equal so every replica
  contributes the same batch size."""

  def __init__(self, processor, num_replicas, replica_device_fn=None, **kwargs):
    del replica_device_fn                      # host pipelines: no device placement needed
    limits = kwargs.get('bucket_batch_limit')
    if num_replicas > 1 and limits:
      assert all(b == max(limits) for b in limits), limits
    key = kwargs.pop('generic_input_v2_key', None)
    self._replicas = []
    for i in range(num_replicas):
      kw = dict(kwargs, num_input_replicas=num_replicas, input_replica_id=i)
      if key is not None:
        self._replicas.append(GenericInputV2Create(processor, generic_input_v2_key=(key, i),
                                                   **kw)[0])
      else:
        self._replicas.append(GenericInput(processor, **kw))

  def GetNext(self):
    """All replicas are read for the same step; a batch is only as long as the bucket the
    replicas agree on (shorter ones are padded up to the longest)."""
    outs = [r.GetNext() for r in self._replicas]
    batches, keys = zip(*outs)
    is_map = isinstance(batches[0], NestedMap)
    flats = [b.Flatten() if is_map else list(b) for b in batches]
    merged = []
    for parts in zip(*flats):
      parts = [np.asarray(x) for x in parts]
      if parts[0].ndim > 1:
        width = [max(x.shape[d] for x in parts) for d in range(1, parts[0].ndim)]
        parts = [np.pad(x, [(0, 0)] + [(0, w - s) for w, s in zip(width, x.shape[1:])])
                 for x in parts]
      merged.append(np.concatenate(parts, 0))
    out = batches[0].Pack(merged) if is_map else merged
    return out, np.concatenate([np.asarray(k) for k in keys], 0)

  def Close(self):
    for r in self._replicas:
      r.Close()
