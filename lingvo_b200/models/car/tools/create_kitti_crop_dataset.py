"""Builds a KITTI *classification* dataset: one example per labelled object holding the
laser points inside its (slightly inflated) 3-D box (ref
`lingvo/tasks/car/tools/create_kitti_crop_dataset.py`; the reference runs this on Beam —
here shards are processed by a multiprocessing pool).

  python -m lingvo_b200.models.car.tools.create_kitti_crop_dataset \
      --input_file_pattern='/data/kitti_object_3dop_train.tfrecord-*' \
      --output_filebase=/data/kitti_crops_train --num_shards=16
"""

from __future__ import annotations

import argparse
import glob
import multiprocessing
import os
import sys

import numpy as np

from lingvo_b200 import ops
from lingvo_b200.utils import tf_example

MAX_POINTS = 1024
BOX_INFLATION = 0.1     # metres added on each side


def _GetFilteredBoundingBoxData(example):
  """Objects with 3-D info and a real class → (boxes [K,7], labels, difficulties-inputs)."""
  has3d = np.asarray(example['object/has_3d_info'], bool)
  labels = [l.decode() if isinstance(l, bytes) else l for l in example['object/label']]
  keep = has3d & np.asarray([l != 'DontCare' for l in labels])
  boxes = np.concatenate([np.asarray(example['object/velo/bbox/xyz']).reshape(-1, 3),
                          np.asarray(example['object/velo/bbox/dim_xyz']).reshape(-1, 3),
                          np.asarray(example['object/velo/bbox/phi']).reshape(-1, 1)], 1)
  return boxes[keep], [l for l, k in zip(labels, keep) if k], np.nonzero(keep)[0]


def CropObjects(record):
  """One scene record → list of serialized crop examples."""
  ex = tf_example.ParseExample(record)
  xyz = np.asarray(ex['pointcloud/xyz'], np.float32).reshape(-1, 3)
  refl = np.asarray(ex['pointcloud/reflectance'], np.float32).reshape(-1, 1)
  boxes, labels, idx = _GetFilteredBoundingBoxData(ex)
  sid = ex['image/source_id'][0]
  out = []
  for b, (box, label, j) in enumerate(zip(boxes, labels, idx)):
    c, s = np.cos(box[6]), np.sin(box[6])
    rel = xyz - box[:3]
    lx, ly = rel[:, 0] * c + rel[:, 1] * s, -rel[:, 0] * s + rel[:, 1] * c
    half = box[3:6] / 2 + BOX_INFLATION
    inside = (np.abs(lx) <= half[0]) & (np.abs(ly) <= half[1]) & (np.abs(rel[:, 2]) <= half[2])
    pts = np.stack([lx, ly, rel[:, 2]], 1)[inside][:MAX_POINTS]     # object-centric coordinates
    if not len(pts):
      continue
    out.append(tf_example.MakeExample({
        'source_id': [sid], 'object_index': np.asarray([int(j)]), 'label': [label.encode()],
        'bbox_3d': box.astype(np.float32), 'num_points': np.asarray([len(pts)]),
        'points_xyz': pts.astype(np.float32).reshape(-1),
        'points_feature': refl[inside][:MAX_POINTS].reshape(-1),
        'occlusion': np.asarray([int(ex['object/occlusion'][j])]),
        'truncation': np.asarray([float(ex['object/truncation'][j])], np.float32)}))
    del b
  return out


class _ProcessShard:
  """Crops every scene of one input file into one output shard (ref :146)."""

  def __init__(self, output_filebase, num_shards):
    self._base, self._n = output_filebase, num_shards

  def __call__(self, job):
    shard, path = job
    y = ops.host().sequential_record_yielder('tfrecord:' + path, 1)
    w = ops.host().TFRecordWriter('%s-%05d-of-%05d' % (self._base, shard, self._n))
    n = 0
    while True:
      rec = y.next()
      if rec is None:
        break
      for ex in CropObjects(rec[0]):
        w.write(ex)
        n += 1
    w.close()
    return n


def main(argv=None):
  ap = argparse.ArgumentParser()
  ap.add_argument('--input_file_pattern', required=True)
  ap.add_argument('--output_filebase', required=True)
  ap.add_argument('--num_shards', type=int, default=0, help='default: one per input file')
  ap.add_argument('--workers', type=int, default=max(1, (os.cpu_count() or 2) // 2))
  a = ap.parse_args(argv)
  files = sorted(glob.glob(a.input_file_pattern))
  if not files:
    raise FileNotFoundError(a.input_file_pattern)
  n_shards = a.num_shards or len(files)
  jobs = [(i % n_shards, f) for i, f in enumerate(files)]
  fn = _ProcessShard(a.output_filebase, n_shards)
  if a.workers > 1:
    with multiprocessing.Pool(a.workers) as pool:
      counts = pool.map(fn, jobs)
  else:
    counts = [fn(j) for j in jobs]
  print('wrote %d object crops' % sum(counts))
  return 0


if __name__ == '__main__':
  sys.exit(main())
