"""Parsers for the raw KITTI object-detection release (ref
`lingvo/tasks/car/tools/kitti_data.py`): velodyne `.bin` scans, `label_2` text files,
`calib` files, and the camera ↔ velodyne box conversions.

KITTI camera frame: x right, y down, z forward; velodyne frame: x forward, y left, z up.
Labels give box *bottom-centre* in camera coordinates with (h, w, l) and `rotation_y`.
"""

from __future__ import annotations

import math

import numpy as np


def LoadVeloBinFile(filepath):
  """→ dict(xyz [N,3], reflectance [N,1])."""
  scan = np.fromfile(filepath, dtype=np.float32).reshape(-1, 4)
  return {'xyz': scan[:, :3], 'reflectance': scan[:, 3:]}


_LABEL_FIELDS = (('type', str), ('truncated', float), ('occluded', int), ('alpha', float),
                 ('bbox', 4), ('dimensions', 3), ('location', 3), ('rotation_y', float),
                 ('score', float))


def _ParseLabelLine(line):
  parts = line.split()
  if len(parts) not in (15, 16):
    raise ValueError('Expected 15 or 16 fields, got %d: %r' % (len(parts), line))
  obj, i = {}, 0
  for name, kind in _LABEL_FIELDS:
    if name == 'score' and i >= len(parts):
      break
    if isinstance(kind, int):
      obj[name] = [float(x) for x in parts[i:i + kind]]
      i += kind
    else:
      obj[name] = kind(float(parts[i])) if kind is int else kind(parts[i])
      i += 1
  return obj


def _ValidateLabeledObject(obj):
  if not 0.0 <= obj['truncated'] <= 1.0 and obj['truncated'] != -1:
    raise ValueError('truncated out of range: %s' % obj)
  if obj['occluded'] not in (-1, 0, 1, 2, 3):
    raise ValueError('invalid occluded: %s' % obj)
  if not -10.0 <= obj['alpha'] <= 10.0:
    raise ValueError('invalid alpha: %s' % obj)
  x0, y0, x1, y1 = obj['bbox']
  if x1 < x0 or y1 < y0:
    raise ValueError('invalid bbox: %s' % obj)


def LoadLabelFile(filepath):
  """→ list of dicts: type, truncated, occluded, alpha, bbox [xmin, ymin, xmax, ymax],
  dimensions [h, w, l], location [x, y, z] (camera), rotation_y[, score]."""
  objs = []
  with open(filepath, encoding='utf-8') as f:
    for line in f:
      if line.strip():
        o = _ParseLabelLine(line)
        _ValidateLabeledObject(o)
        objs.append(o)
  return objs


def ParseCalibrationDict(raw_calib):
  """{'P2': '…', 'R0_rect': '…', 'Tr_velo_to_cam': '…'} strings → matrices."""
  calib = {}
  for k, v in raw_calib.items():
    vals = np.asarray([float(x) for x in v.split()], np.float64)
    if k.startswith('P'):
      calib[k] = vals.reshape(3, 4)
    elif k == 'R0_rect':
      calib[k] = vals.reshape(3, 3)
    elif k.startswith('Tr_'):
      calib[k] = vals.reshape(3, 4)
  return calib


def LoadCalibrationFile(filepath):
  raw = {}
  with open(filepath, encoding='utf-8') as f:
    for line in f:
      if ':' in line:
        k, v = line.split(':', 1)
        raw[k.strip()] = v.strip()
  return ParseCalibrationDict(raw)


def _Hom(mat, rows=4):
  out = np.eye(4)
  out[:mat.shape[0], :mat.shape[1]] = mat
  return out[:rows] if rows != 4 else out


def VeloToCameraTransformation(calib):
  """4×4: velodyne → rectified camera-0 frame."""
  return _Hom(calib['R0_rect']) @ _Hom(calib['Tr_velo_to_cam'])


def CameraToVeloTransformation(calib):
  return np.linalg.inv(VeloToCameraTransformation(calib))


def VeloToImagePlaneTransformation(calib):
  """3×4: velodyne → image-2 homogeneous pixels."""
  return calib['P2'] @ VeloToCameraTransformation(calib)


def _KITTIObjectHas3DInfo(obj):
  return obj['type'] != 'DontCare' and all(d > 0 for d in obj['dimensions'])


def _KITTIObjectToBBox3D(obj, cam_to_velo_transform):
  """→ [x, y, z, dx(length), dy(width), dz(height), phi] in the velodyne frame."""
  h, w, l = obj['dimensions']
  bottom = np.asarray(obj['location'] + [1.0])
  centre_cam = bottom.copy()
  centre_cam[1] -= h / 2.0                      # camera y points down
  x, y, z, _ = cam_to_velo_transform @ centre_cam
  phi = -obj['rotation_y'] - math.pi / 2.0      # rotation about camera y → about velodyne z
  phi = (phi + math.pi) % (2 * math.pi) - math.pi
  return [float(x), float(y), float(z), float(l), float(w), float(h), float(phi)]


def AnnotateKITTIObjectsWithBBox3D(objects, calib):
  """Adds `has_3d_info` and `bbox3d` (velodyne frame) to every object."""
  cam_to_velo = CameraToVeloTransformation(calib)
  for o in objects:
    o['has_3d_info'] = _KITTIObjectHas3DInfo(o)
    o['bbox3d'] = _KITTIObjectToBBox3D(o, cam_to_velo) if o['has_3d_info'] else [0.0] * 7
  return objects


def BBox3DToKITTIObject(bbox3d, velo_to_cam_transform):
  """Inverse of `_KITTIObjectToBBox3D`: → (location [3] bottom-centre camera frame,
  dimensions [h, w, l], rotation_y)."""
  x, y, z, l, w, h, phi = [float(v) for v in bbox3d]
  cam = velo_to_cam_transform @ np.asarray([x, y, z, 1.0])
  cam[1] += h / 2.0
  rot_y = -phi - math.pi / 2.0
  rot_y = (rot_y + math.pi) % (2 * math.pi) - math.pi
  return [float(cam[0]), float(cam[1]), float(cam[2])], [h, w, l], rot_y
