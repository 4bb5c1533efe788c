"""Round-2 native input features: replica sharding of the yielders
(reference record_yielder.h:84-85), `BucketAdjuster` (record_batcher.h:72), `fatal_errors`,
dynamic padding constants (generic_input_op_kernels.cc:143-145), hardware CRC32C."""

import os
import tempfile

import numpy as np
import pytest

from lingvo_b200 import ops
from lingvo_b200.core import generic_input
from lingvo_b200.utils import tfrecord


def _WriteShards(n_files, per_file):
  d = tempfile.mkdtemp()
  k = 0
  for f in range(n_files):
    w = tfrecord.TFRecordWriter(os.path.join(d, 'data-%05d-of-%05d' % (f, n_files)))
    for _ in range(per_file):
      w.write(b'%06d' % k)
      k += 1
    w.close()
  return 'tfrecord:' + os.path.join(d, 'data-*'), k


def _Drain(y):
  out = []
  while True:
    r = y.next()
    if r is None:
      return out
    out.append(int(r[0]))


@pytest.mark.parametrize('n_files,replicas', [(8, 4), (2, 4), (5, 3)])
def test_replicas_read_disjoint_shards_that_cover_the_data(n_files, replicas):
  pattern, total = _WriteShards(n_files, 37)
  h = ops.host()
  seen = []
  for r in range(replicas):
    y = h.basic_record_yielder(pattern, seed=11, bufsize=64, parallelism=3, num_epochs=1,
                               num_input_replicas=replicas, input_replica_id=r)
    seen.append(_Drain(y))
  flat = sorted(x for s in seen for x in s)
  assert flat == list(range(total))                    # disjoint and complete
  sizes = [len(s) for s in seen]
  assert max(sizes) - min(sizes) <= 37                 # balanced to within one file
  # sequential (eval) yielder: round-robin by record
  seq = [_Drain(h.sequential_record_yielder(pattern, 1, 0, replicas, r)) for r in range(replicas)]
  assert sorted(x for s in seq for x in s) == list(range(total))
  assert seq[0][:3] == [0, replicas, 2 * replicas]


def test_bucket_adjuster_minimises_padding_cost():
  h = ops.host()
  adj = h.BucketAdjuster(100, 3)
  rng = np.random.RandomState(0)
  keys = np.concatenate([rng.randint(5, 9, 500), rng.randint(40, 44, 300),
                         rng.randint(95, 100, 50)])
  for k in keys:
    adj.increment_histogram(int(k))
  bounds = adj.adjust_buckets([33, 66, 100])
  assert bounds[-1] == 100 and bounds == sorted(bounds)
  assert bounds[0] == 8 and bounds[1] == 43            # tight around the two clusters

  def cost(bs):
    return sum(next(b for b in bs if k <= b) for k in keys)
  assert cost(bounds) < 0.6 * cost([33, 66, 100])


def test_batcher_adjusts_buckets_skips_nonfatal_errors_and_pads_with_constants():
  pattern, total = _WriteShards(2, 300)

  def proc(rec):
    k = int(rec)
    if k % 50 == 7:
      raise ValueError('corrupt example %d' % k)
    n = 3 + (k % 5) if k % 2 else 20 + (k % 4)
    return [np.full([n], 1, np.int32), np.full([n], 0.5, np.float32)], n

  gi = generic_input.GenericInput(
      proc, file_pattern=pattern, bucket_upper_bound=[12, 24, 48],
      bucket_batch_limit=[8, 8, 8], file_random_seed=3, num_threads=2, repeat_count=1,
      bucket_adjust_every_n=100, fatal_errors=['FATAL'],
      dynamic_padding_constants=[-1, 0.0], num_input_replicas=1, input_replica_id=0)
  n_rows = 0
  for (ids, vals), keys in gi:
    assert ids.shape == vals.shape
    for row, k in zip(ids, keys):
      assert (row[:k] == 1).all() and (row[k:] == -1).all()      # slot-0 pad constant
    n_rows += ids.shape[0]
  assert gi.records_failed == total // 50                       # non-fatal ⇒ skipped, counted
  assert n_rows == total - gi.records_failed
  assert gi.bucket_upper_bound[-1] == 48 and gi.bucket_upper_bound[0] == 7


def test_fatal_error_substring_aborts_the_pipeline():
  pattern, _ = _WriteShards(1, 20)

  def proc(rec):
    if int(rec) == 5:
      raise RuntimeError('FATAL: schema mismatch')
    return [np.zeros([2], np.int32)], 2

  gi = generic_input.GenericInput(proc, file_pattern=pattern, bucket_upper_bound=[4],
                                  bucket_batch_limit=[4], repeat_count=1, num_threads=1,
                                  require_sequential_order=True, fatal_errors=['FATAL'],
                                  num_input_replicas=1, input_replica_id=0)
  with pytest.raises(RuntimeError, match='schema mismatch'):
    for _ in gi:
      pass


def test_crc32c_hardware_path_matches_table_implementation():
  h = ops.host()
  rng = np.random.RandomState(1)
  for n in (0, 1, 9, 8191, 8192, 24576, 24577, 70001):
    d = rng.randint(0, 256, n, dtype=np.uint8)
    c = 0xFFFFFFFF
    for b in d.tobytes():
      c = tfrecord._TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    want = c ^ 0xFFFFFFFF
    assert h.crc32c_buffer(d, 0) == want
    if n > 20:
      assert h.crc32c_buffer(d[n // 3:], h.crc32c_buffer(d[:n // 3], 0)) == want
  assert tfrecord.crc32c(b'123456789') == 0xE3069283


def test_keyed_and_replicated_generic_input():
  pattern, total = _WriteShards(4, 40)

  def proc(rec):
    k = int(rec)
    n = 2 + k % 3
    return generic_input.NestedMap(ids=np.full([n], k, np.int32)), n

  kw = dict(file_pattern=pattern, bucket_upper_bound=[8], bucket_batch_limit=[4],
            repeat_count=1, file_buffer_size=1, file_parallelism=1, num_threads=1,
            require_sequential_order=True)
  with pytest.raises(RuntimeError):
    generic_input.GenericInputV2Create(proc, **kw)
  res, _, _ = generic_input.GenericInputV2Create(proc, generic_input_v2_key='eval', **kw)
  first, _ = generic_input.GenericInputV2GetNext(res)
  res2, _, _ = generic_input.GenericInputV2Create(proc, generic_input_v2_key='eval', **kw)
  assert res2 is res                                   # same key → same live pipeline
  second, _ = generic_input.GenericInputV2GetNext(res2)
  assert set(first.ids[:, 0].tolist()).isdisjoint(second.ids[:, 0].tolist())
  generic_input.ResetGenericInputV2Cache('eval')
  res3, _, _ = generic_input.GenericInputV2Create(proc, generic_input_v2_key='eval', **kw)
  again, _ = generic_input.GenericInputV2GetNext(res3)
  assert again.ids[:, 0].tolist() == first.ids[:, 0].tolist()   # a fresh pipeline restarts
  generic_input.ResetGenericInputV2Cache()
  assert generic_input.IsGenericInputV2AllwedInEager()

  rep = generic_input.ReplicatedGenericInput(proc, 2, lambda i: 'cpu:%d' % i, **kw)
  seen = []
  try:
    while True:
      batch, keys = rep.GetNext()
      assert batch.ids.shape[0] == 8 and keys.shape == (8,)
      seen += batch.ids[:, 0].tolist()
      # rows 0-3 come from replica 0, rows 4-7 from replica 1: disjoint input shards
      assert set(batch.ids[:4, 0].tolist()).isdisjoint(batch.ids[4:, 0].tolist())
  except StopIteration:
    pass
  rep.Close()
  assert len(seen) == len(set(seen)) and len(seen) >= total - 8
  with pytest.raises(AssertionError):
    generic_input.ReplicatedGenericInput(proc, 2, None, **dict(
        kw, bucket_upper_bound=[4, 8], bucket_batch_limit=[8, 4]))
