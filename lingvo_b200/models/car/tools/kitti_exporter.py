"""Raw KITTI → TFRecords of `tf.Example` (ref `lingvo/tasks/car/tools/kitti_exporter.py`).

  python -m lingvo_b200.models.car.tools.kitti_exporter \
      --kitti_object_dir=/data/kitti/object --split=train|val|test \
      --split_file=ImageSets/train.txt --tfrecord_path=/out/kitti_object_3dop_train.tfrecord \
      --num_shards=100

The output schema is the one `kitti_input_generator` reads.
"""

from __future__ import annotations

import argparse
import os
import sys

import numpy as np

from lingvo_b200 import ops
from lingvo_b200.models.car.tools import kitti_data
from lingvo_b200.utils import tf_example


def _ReadFrame(root_dir, subdir, frame, with_labels):
  base = os.path.join(root_dir, subdir)
  velo = kitti_data.LoadVeloBinFile(os.path.join(base, 'velodyne', frame + '.bin'))
  calib = kitti_data.LoadCalibrationFile(os.path.join(base, 'calib', frame + '.txt'))
  img_path = os.path.join(base, 'image_2', frame + '.png')
  with open(img_path, 'rb') as f:
    img = f.read()
  objects = []
  if with_labels:
    objects = kitti_data.AnnotateKITTIObjectsWithBBox3D(
        kitti_data.LoadLabelFile(os.path.join(base, 'label_2', frame + '.txt')), calib)
  return velo, calib, img, objects


def _PngSize(data):
  """(width, height) from a PNG header."""
  return int.from_bytes(data[16:20], 'big'), int.from_bytes(data[20:24], 'big')


def MakeExample(frame, velo, calib, img, objects):
  w, h = _PngSize(img)
  f32 = lambda x: np.asarray(x, np.float32).reshape(-1)
  feats = {
      'image/source_id': [frame.encode()], 'image/encoded': [img], 'image/format': [b'png'],
      'image/height': np.asarray([h]), 'image/width': np.asarray([w]),
      'pointcloud/xyz': f32(velo['xyz']), 'pointcloud/reflectance': f32(velo['reflectance']),
      'transform/velo_to_image_plane': f32(kitti_data.VeloToImagePlaneTransformation(calib)),
      'transform/velo_to_camera': f32(kitti_data.VeloToCameraTransformation(calib)),
      'transform/camera_to_velo': f32(kitti_data.CameraToVeloTransformation(calib)),
      'object/label': [o['type'].encode() for o in objects],
      'object/has_3d_info': np.asarray([int(o['has_3d_info']) for o in objects], np.int64),
      'object/occlusion': np.asarray([o['occluded'] for o in objects], np.int64),
      'object/truncation': f32([o['truncated'] for o in objects]),
      'object/image/bbox/xmin': f32([o['bbox'][0] for o in objects]),
      'object/image/bbox/ymin': f32([o['bbox'][1] for o in objects]),
      'object/image/bbox/xmax': f32([o['bbox'][2] for o in objects]),
      'object/image/bbox/ymax': f32([o['bbox'][3] for o in objects]),
      'object/velo/bbox/xyz': f32([o['bbox3d'][:3] for o in objects]),
      'object/velo/bbox/dim_xyz': f32([o['bbox3d'][3:6] for o in objects]),
      'object/velo/bbox/phi': f32([o['bbox3d'][6] for o in objects]),
  }
  return tf_example.MakeExample(feats)


def _ExportObjectDatasetToTFRecord(root_dir, split, split_file, tfrecord_path, num_shards):
  with open(split_file, encoding='utf-8') as f:
    frames = [l.strip() for l in f if l.strip()]
  subdir = 'testing' if split == 'test' else 'training'
  writers = [ops.host().TFRecordWriter('%s-%05d-of-%05d' % (tfrecord_path, i, num_shards))
             for i in range(num_shards)]
  for n, frame in enumerate(frames):
    velo, calib, img, objects = _ReadFrame(root_dir, subdir, frame, split != 'test')
    writers[n % num_shards].write(MakeExample(frame, velo, calib, img, objects))
  for w in writers:
    w.close()
  return len(frames)


def main(argv=None):
  ap = argparse.ArgumentParser()
  ap.add_argument('--kitti_object_dir', required=True)
  ap.add_argument('--split', default='train', choices=['train', 'val', 'trainval', 'test'])
  ap.add_argument('--split_file', required=True)
  ap.add_argument('--tfrecord_path', required=True)
  ap.add_argument('--num_shards', type=int, default=100)
  a = ap.parse_args(argv)
  n = _ExportObjectDatasetToTFRecord(a.kitti_object_dir, a.split, a.split_file, a.tfrecord_path,
                                     a.num_shards)
  print('exported %d frames' % n)
  return 0


if __name__ == '__main__':
  sys.exit(main())
