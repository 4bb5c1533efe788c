"""Converts Waymo Open Dataset TFRecords of `Frame` protos to `tf.Example` TFRecords (ref
`lingvo/tasks/car/waymo/tools/generate_waymo_tf.py`).

  python -m lingvo_b200.models.car.waymo.tools.generate_waymo_tf \
      --input_file_pattern='/data/waymo/training/segment-*.tfrecord' \
      --output_filebase=/data/waymo/train.tfr --num_shards=1000
"""

import argparse
import glob
import multiprocessing
import os
import sys

from lingvo_b200 import ops
from lingvo_b200.models.car.waymo.tools import waymo_proto_to_tfe


def _Job(job):
  shard, n_shards, base, files = job
  conv = waymo_proto_to_tfe.WaymoOpenDatasetConverter()
  w = ops.host().TFRecordWriter('%s-%05d-of-%05d' % (base, shard, n_shards))
  n = sum(conv.ConvertFile(f, w) for f in files)
  w.close()
  return n


def main(argv=None):
  ap = argparse.ArgumentParser()
  ap.add_argument('--input_file_pattern', required=True)
  ap.add_argument('--output_filebase', required=True)
  ap.add_argument('--num_shards', type=int, default=1000)
  ap.add_argument('--workers', type=int, default=max(1, (os.cpu_count() or 2) // 2))
  a = ap.parse_args(argv)
  files = sorted(glob.glob(a.input_file_pattern))
  if not files:
    raise FileNotFoundError(a.input_file_pattern)
  n_shards = min(a.num_shards, len(files))
  jobs = [(s, n_shards, a.output_filebase, files[s::n_shards]) for s in range(n_shards)]
  with multiprocessing.Pool(a.workers) as pool:
    counts = pool.map(_Job, jobs)
  print('converted %d frames into %d shards' % (sum(counts), n_shards))
  return 0


if __name__ == '__main__':
  sys.exit(main())
