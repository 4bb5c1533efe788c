"""Decoder outputs → KITTI detection text files for the official evaluation server (ref
`lingvo/tasks/car/tools/export_kitti_detection.py`).

  python -m lingvo_b200.models.car.tools.export_kitti_detection \
      --decoder_path=/logdir/decoder_test/decoder_out_000010000 --calib_dir=/kitti/testing/calib \
      --output_dir=/tmp/kitti_submission --class_names=Car,Pedestrian,Cyclist

The decoder dump is the pickle written by `DecodeProgram` / the decoder runner: a list of
(key, dict) or dicts holding `source_id`, `bboxes [N,7]`, `scores [N]`, `class_ids [N]`.
"""

from __future__ import annotations

import argparse
import os
import pickle
import sys

import numpy as np

from lingvo_b200.models.car.tools import kitti_data


def LoadCalibData(fname):
  calib = kitti_data.LoadCalibrationFile(fname)
  return dict(velo_to_cam=kitti_data.VeloToCameraTransformation(calib), P2=calib['P2'])


def ExtractNpContent(np_dict, calib):
  """→ list of per-detection dicts in KITTI conventions."""
  out = []
  v2c = calib['velo_to_cam']
  for box, score, cls in zip(np.asarray(np_dict['bboxes']), np.asarray(np_dict['scores']),
                             np.asarray(np_dict['class_ids'])):
    if score <= 0:
      continue
    loc, dims, rot_y = kitti_data.BBox3DToKITTIObject(box, v2c)
    x, y, z, dx, dy, dz, phi = [float(v) for v in box]
    c, s = np.cos(phi), np.sin(phi)
    corners = []
    for sx in (-0.5, 0.5):
      for sy in (-0.5, 0.5):
        for sz in (-0.5, 0.5):
          corners.append([x + sx * dx * c - sy * dy * s, y + sx * dx * s + sy * dy * c,
                          z + sz * dz, 1.0])
    uvw = (calib['P2'] @ v2c @ np.asarray(corners).T).T
    if (uvw[:, 2] <= 0).all():
      continue
    uv = uvw[:, :2] / np.maximum(uvw[:, 2:3], 1e-6)
    out.append(dict(class_id=int(cls), score=float(score), location=loc, dimensions=dims,
                    rotation_y=rot_y, bbox=[uv[:, 0].min(), uv[:, 1].min(), uv[:, 0].max(),
                                            uv[:, 1].max()],
                    alpha=rot_y - np.arctan2(loc[0], loc[2])))
  return out


def ExportKITTIDetection(out_dir, source_id, detections, class_names):
  """Writes `<out_dir>/<source_id>.txt` in the 16-column KITTI result format."""
  os.makedirs(out_dir, exist_ok=True)
  path = os.path.join(out_dir, '%s.txt' % source_id)
  with open(path, 'w', encoding='utf-8') as f:
    for d in detections:
      name = class_names[d['class_id']] if d['class_id'] < len(class_names) else 'DontCare'
      h, w, l = d['dimensions']
      f.write('%s -1 -1 %.4f %.2f %.2f %.2f %.2f %.4f %.4f %.4f %.4f %.4f %.4f %.4f %.6f\n' % (
          name, d['alpha'], *d['bbox'], h, w, l, *d['location'], d['rotation_y'], d['score']))
  return path


def main(argv=None):
  ap = argparse.ArgumentParser()
  ap.add_argument('--decoder_path', required=True)
  ap.add_argument('--calib_dir', required=True)
  ap.add_argument('--output_dir', required=True)
  ap.add_argument('--class_names',
                  default='Background,Car,Van,Truck,Pedestrian,Person_sitting,Cyclist,Tram,Misc')
  a = ap.parse_args(argv)
  with open(a.decoder_path, 'rb') as f:
    dump = pickle.load(f)   # noqa: S301  (our own decoder output)
  names = a.class_names.split(',')
  n = 0
  for item in dump:
    d = item[1] if isinstance(item, (tuple, list)) else item
    sid = d['source_id'].decode() if isinstance(d['source_id'], bytes) else str(d['source_id'])
    calib = LoadCalibData(os.path.join(a.calib_dir, sid.strip() + '.txt'))
    ExportKITTIDetection(a.output_dir, sid.strip(), ExtractNpContent(d, calib), names)
    n += 1
  print('wrote %d files to %s' % (n, a.output_dir))
  return 0


if __name__ == '__main__':
  sys.exit(main())
