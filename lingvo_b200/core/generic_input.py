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

import inspect
from typing import Callable, List, Optional, Sequence

import numpy as np

from lingvo_b200 import ops
from lingvo_b200.core.nested_map import NestedMap


def MakeYielder(file_pattern, file_random_seed=0, file_buffer_size=10000,
                file_parallelism=4, repeat_count=-1, require_sequential_order=False,
                input_source_weights: Optional[Sequence[float]] = None):
  """Yielder for a `type:glob` pattern or a list of patterns (weighted mix)."""
  h = ops.host()
  if isinstance(file_pattern, (list, tuple)):
    pats = list(file_pattern)
    weights = list(input_source_weights) if input_source_weights else [1.0] * len(pats)
    if pats and isinstance(pats[0], (list, tuple)):      # [(pattern, weight), …]
      weights = [w for _, w in pats]
      pats = [p for p, _ in pats]
    kids = [MakeYielder(p, file_random_seed + i if file_random_seed else 0,
                        file_buffer_size, file_parallelism, repeat_count,
                        require_sequential_order) for i, p in enumerate(pats)]
    return h.weighted_mix_record_yielder(kids, weights, file_random_seed)
  if require_sequential_order:
    return h.sequential_record_yielder(file_pattern,
                                       repeat_count if repeat_count > 0 else -1)
  return h.basic_record_yielder(
      file_pattern, seed=file_random_seed, bufsize=file_buffer_size,
      parallelism=file_parallelism,
      num_epochs=repeat_count if repeat_count > 0 else 0)


class GenericInput:
  """Bucketing batcher over a record yielder."""

  def __init__(self, processor: Callable, file_pattern=None, yielder=None,
               bucket_upper_bound=(1 << 30,), bucket_batch_limit=(1,),
               file_random_seed=0, file_buffer_size=10000, file_parallelism=4,
               num_threads=4, flush_every_n=0, repeat_count=-1,
               require_sequential_order=False, input_source_weights=None):
    self._h = ops.host()
    self._yielder = yielder or MakeYielder(
        file_pattern, file_random_seed, file_buffer_size, file_parallelism,
        repeat_count, require_sequential_order, input_source_weights)
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
        1 if require_sequential_order else num_threads, flush_every_n)

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

  def Close(self):
    self._batcher.close()
