"""Parallel map/shard helpers used by data-prep tools (ref `lingvo/tools/beam_utils.py`,
an Apache Beam wrapper there; a process pool here)."""
import concurrent.futures
import os


def BeamInit():
  pass


def GetPipelineRoot(options=None):
  return LocalPipeline(options)


def GetReader(file_type, file_pattern, value_coder=None):
  from lingvo_b200 import ops  # pylint: disable=g-import-not-at-top
  del value_coder

  def _Read():
    y = ops.host().sequential_record_yielder('%s:%s' % (file_type, file_pattern), 1)
    while True:
      r = y.next()
      if r is None:
        return
      yield r[0]
  return _Read


def GetWriter(file_type, file_pattern, value_coder=None, num_shards=1):
  from lingvo_b200 import ops  # pylint: disable=g-import-not-at-top
  del value_coder
  assert file_type == 'tfrecord'

  def _Write(records):
    ws = [ops.host().TFRecordWriter('%s-%05d-of-%05d' % (file_pattern, i, num_shards))
          for i in range(num_shards)]
    for i, r in enumerate(records):
      ws[i % num_shards].write(r)
    for w in ws:
      w.close()
  return _Write


class LocalPipeline:
  """`with GetPipelineRoot() as p: p.Map(fn, items)` – ordered parallel map."""

  def __init__(self, options=None):
    self._workers = (options or {}).get('workers', os.cpu_count() or 1)

  def __enter__(self):
    self._pool = concurrent.futures.ProcessPoolExecutor(self._workers)
    return self

  def __exit__(self, *a):
    self._pool.shutdown()

  def Map(self, fn, items, chunksize=16):
    return list(self._pool.map(fn, items, chunksize=chunksize))
