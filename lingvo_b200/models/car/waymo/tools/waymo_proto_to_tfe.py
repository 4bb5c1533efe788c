"""Waymo Open Dataset `Frame` protos → `tf.Example` (ref
`lingvo/tasks/car/waymo/tools/waymo_proto_to_tfe.py`).

The official reader depends on the `waymo_open_dataset` package + TensorFlow. Neither is
available here, so frames are decoded with the in-repo protobuf wire codec against the
published `dataset.proto` field numbers (the subset this converter needs), range images
(zlib-compressed `MatrixFloat`) are expanded with NumPy, and the output schema is exactly
what `waymo_open_input_generator` reads.
"""

from __future__ import annotations

import zlib

import numpy as np

from lingvo_b200.utils import protowire as pw
from lingvo_b200.utils import tf_example

# --- dataset.proto field numbers -------------------------------------------------------
FRAME = dict(context=1, timestamp_micros=2, pose=3, images=4, lasers=5, laser_labels=6,
             no_label_zones=9)
CONTEXT = dict(name=1, camera_calibrations=2, laser_calibrations=3, stats=4)
STATS = dict(time_of_day=3, location=4, weather=5)
LASER = dict(name=1, ri_return1=2, ri_return2=3)
RANGE_IMAGE = dict(range_image_compressed=2, camera_projection_compressed=3,
                   range_image_pose_compressed=4, range_image_flow_compressed=5)
LASER_CALIB = dict(name=1, beam_inclinations=2, beam_inclination_min=3, beam_inclination_max=4,
                   extrinsic=5)
CAMERA_CALIB = dict(name=1, intrinsic=2, extrinsic=3, width=4, height=5)
IMAGE = dict(name=1, image=2, pose=3)
LABEL = dict(box=1, metadata=2, type=3, id=4, detection_difficulty_level=5,
             tracking_difficulty_level=6, num_lidar_points_in_box=7)
BOX = dict(center_x=1, center_y=2, center_z=3, width=4, length=5, height=6, heading=7)
LABEL_METADATA = dict(speed_x=1, speed_y=2, accel_x=3, accel_y=4)
LASER_NAMES = {1: 'TOP', 2: 'FRONT', 3: 'SIDE_LEFT', 4: 'SIDE_RIGHT', 5: 'REAR'}
CAMERA_NAMES = {1: 'FRONT', 2: 'FRONT_LEFT', 3: 'FRONT_RIGHT', 4: 'SIDE_LEFT', 5: 'SIDE_RIGHT'}


def _One(d, f, default=None):
  v = d.get(f)
  return v[0] if v else default


def _Floats(buf):
  """Packed repeated double/float payload → float32 array."""
  if isinstance(buf, (bytes, bytearray)):
    n = len(buf)
    return (np.frombuffer(buf, '<f8') if n % 8 == 0 and n >= 8 * 16 or n == 8 * 16
            else np.frombuffer(buf, '<f4')).astype(np.float32)
  return np.asarray(buf, np.float32)


def _Transform(buf):
  """`Transform { repeated double transform = 1 }` → 4×4."""
  vals = pw.parse_dict(buf).get(1, [])
  if len(vals) == 1 and isinstance(vals[0], (bytes, bytearray)):
    arr = np.frombuffer(vals[0], '<f8')
  else:
    arr = np.asarray([pw.as_double(v) for v in vals],
                     np.float64)
  return arr.astype(np.float32).reshape(4, 4) if arr.size == 16 else np.eye(4, dtype=np.float32)


def ParseMatrixFloat(compressed):
  """zlib(`MatrixFloat { repeated float data = 1 [packed]; Shape shape = 2 }`) → ndarray."""
  d = pw.parse_dict(zlib.decompress(compressed))
  data = np.frombuffer(_One(d, 1, b''), '<f4')
  shape = pw.parse_dict(_One(d, 2, b'')).get(1, [])
  if len(shape) == 1 and isinstance(shape[0], (bytes, bytearray)):
    shape = pw.parse_packed_varints(shape[0])
  return data.reshape([int(s) for s in shape]) if shape else data


class FrameToTFE:
  """ref :187."""

  def __init__(self, use_range_image_index_as_lidar_feature=None):
    self._use_ri_index = use_range_image_index_as_lidar_feature

  # -- geometry -----------------------------------------------------------------------
  @staticmethod
  def RangeImageToPoints(ri, inclinations, extrinsic):
    """Range image `[H, W, 4]` (range, intensity, elongation, nlz) → points `[n, 6]` =
    xyz (vehicle frame) + the 3 features, for pixels with range > 0."""
    h, w, _ = ri.shape
    az = np.linspace(np.pi, -np.pi, w, endpoint=False, dtype=np.float32)
    az = az - np.arctan2(extrinsic[1, 0], extrinsic[0, 0])
    incl = np.asarray(inclinations, np.float32)[::-1].reshape(h, 1)
    rng = ri[..., 0]
    cos_i = np.cos(incl)
    xyz = np.stack([rng * cos_i * np.cos(az), rng * cos_i * np.sin(az), rng * np.sin(incl)], -1)
    hom = np.concatenate([xyz, np.ones((h, w, 1), np.float32)], -1) @ extrinsic.T
    keep = rng > 0
    return np.concatenate([hom[..., :3][keep], ri[..., 1:4][keep]], -1).astype(np.float32)

  # -- conversion ---------------------------------------------------------------------
  def process(self, frame_bytes):  # pylint: disable=invalid-name
    """Serialized `Frame` → serialized `tf.Example`."""
    fr = pw.parse_dict(frame_bytes)
    ctx = pw.parse_dict(_One(fr, FRAME['context'], b''))
    stats = pw.parse_dict(_One(ctx, CONTEXT['stats'], b''))
    s = lambda d, f: (_One(d, f, b'') or b'')
    feats = {
        'run_segment': [s(ctx, CONTEXT['name'])],
        'run_start_offset': np.asarray([int(_One(fr, FRAME['timestamp_micros'], 0))]),
        'time_of_day': [s(stats, STATS['time_of_day'])],
        'location': [s(stats, STATS['location'])], 'weather': [s(stats, STATS['weather'])],
        'pose': _Transform(_One(fr, FRAME['pose'], b'')).reshape(-1),
    }
    self._AddLasers(fr, ctx, feats)
    self._AddLabels(fr, feats)
    self._AddImages(fr, ctx, feats)
    return tf_example.MakeExample(feats)

  def _LaserCalibrations(self, ctx):
    out = {}
    for buf in ctx.get(CONTEXT['laser_calibrations'], []):
      c = pw.parse_dict(buf)
      name = LASER_NAMES.get(int(_One(c, LASER_CALIB['name'], 0)), 'UNKNOWN')
      incl = _One(c, LASER_CALIB['beam_inclinations'], b'')
      incl = np.frombuffer(incl, '<f8').astype(np.float32) if incl else None
      out[name] = dict(
          inclinations=incl,
          min=pw.as_double(_One(c, LASER_CALIB['beam_inclination_min'], 0)),
          max=pw.as_double(_One(c, LASER_CALIB['beam_inclination_max'], 0)),
          extrinsic=_Transform(_One(c, LASER_CALIB['extrinsic'], b'')))
    return out

  def _AddLasers(self, fr, ctx, feats):
    calibs = self._LaserCalibrations(ctx)
    for buf in fr.get(FRAME['lasers'], []):
      laser = pw.parse_dict(buf)
      name = LASER_NAMES.get(int(_One(laser, LASER['name'], 0)), 'UNKNOWN')
      cal = calibs.get(name, dict(inclinations=None, min=-0.3, max=0.04,
                                  extrinsic=np.eye(4, dtype=np.float32)))
      for ret, field in (('ri1', LASER['ri_return1']), ('ri2', LASER['ri_return2'])):
        ri_msg = pw.parse_dict(_One(laser, field, b''))
        comp = _One(ri_msg, RANGE_IMAGE['range_image_compressed'])
        if not comp:
          feats['laser_%s_%s' % (name, ret)] = np.zeros(0, np.float32)
          continue
        ri = ParseMatrixFloat(comp).astype(np.float32)
        incl = cal['inclinations']
        if incl is None or len(incl) != ri.shape[0]:
          incl = np.linspace(cal['min'], cal['max'], ri.shape[0]).astype(np.float32)
        pts = self.RangeImageToPoints(ri, incl, cal['extrinsic'])
        feats['laser_%s_%s' % (name, ret)] = pts.reshape(-1)
        feats['%s_%s' % (name, ret)] = ri.reshape(-1)
        feats['%s_%s_shape' % (name, ret)] = np.asarray(ri.shape, np.int64)
      feats['%s_extrinsics' % name] = cal['extrinsic'].reshape(-1)
      feats['%s_beam_inclinations' % name] = np.asarray(
          cal['inclinations'] if cal['inclinations'] is not None else [], np.float32)

  def _AddLabels(self, fr, feats):
    labels, ids, det, trk, boxes, npts, meta = [], [], [], [], [], [], []
    dbl = pw.as_double
    for buf in fr.get(FRAME['laser_labels'], []):
      lab = pw.parse_dict(buf)
      box = pw.parse_dict(_One(lab, LABEL['box'], b''))
      g = lambda f: dbl(_One(box, BOX[f], 0))
      # Waymo (length along heading, width) → [x, y, z, dx=length, dy=width, dz=height, phi]
      boxes.append([g('center_x'), g('center_y'), g('center_z'), g('length'), g('width'),
                    g('height'), g('heading')])
      labels.append(int(_One(lab, LABEL['type'], 0)))
      ids.append(_One(lab, LABEL['id'], b'') or b'')
      det.append(int(_One(lab, LABEL['detection_difficulty_level'], 0)))
      trk.append(int(_One(lab, LABEL['tracking_difficulty_level'], 0)))
      n = int(_One(lab, LABEL['num_lidar_points_in_box'], 0))
      npts.append(n)
      md = pw.parse_dict(_One(lab, LABEL['metadata'], b''))
      meta.append([dbl(_One(md, LABEL_METADATA[k], 0)) for k in
                   ('speed_x', 'speed_y', 'accel_x', 'accel_y')])
    # official rule: unset difficulty → LEVEL_2 if ≤ 5 points else LEVEL_1
    single = [d if d else (2 if n <= 5 else 1) for d, n in zip(det, npts)]
    feats.update({
        'labels': np.asarray(labels, np.int64), 'label_ids': ids,
        'detection_difficulties': np.asarray(det, np.int64),
        'single_frame_detection_difficulties': np.asarray(single, np.int64),
        'tracking_difficulties': np.asarray(trk, np.int64),
        'bboxes_3d': np.asarray(boxes, np.float32).reshape(-1),
        'bboxes_3d_num_points': np.asarray(npts, np.int64),
        'label_metadata': np.asarray(meta, np.float32).reshape(-1)})

  def _AddImages(self, fr, ctx, feats):
    calibs = {}
    for buf in ctx.get(CONTEXT['camera_calibrations'], []):
      c = pw.parse_dict(buf)
      name = CAMERA_NAMES.get(int(_One(c, CAMERA_CALIB['name'], 0)), 'UNKNOWN')
      intr = _One(c, CAMERA_CALIB['intrinsic'], b'')
      calibs[name] = dict(
          intrinsics=(np.frombuffer(intr, '<f8').astype(np.float32) if intr else
                      np.zeros(9, np.float32)),
          extrinsics=_Transform(_One(c, CAMERA_CALIB['extrinsic'], b'')),
          width=int(_One(c, CAMERA_CALIB['width'], 0)),
          height=int(_One(c, CAMERA_CALIB['height'], 0)))
    for buf in fr.get(FRAME['images'], []):
      im = pw.parse_dict(buf)
      name = CAMERA_NAMES.get(int(_One(im, IMAGE['name'], 0)), 'UNKNOWN')
      cal = calibs.get(name, dict(intrinsics=np.zeros(9, np.float32),
                                  extrinsics=np.eye(4, dtype=np.float32), width=0, height=0))
      feats['image_%s' % name] = [_One(im, IMAGE['image'], b'') or b'']
      feats['image_%s_shape' % name] = np.asarray([cal['height'], cal['width'], 3], np.int64)
      feats['image_%s_pose' % name] = _Transform(_One(im, IMAGE['pose'], b'')).reshape(-1)
      feats['image_%s_intrinsics' % name] = np.resize(cal['intrinsics'], 9).astype(np.float32)
      feats['image_%s_extrinsics' % name] = cal['extrinsics'].reshape(-1)


class WaymoOpenDatasetConverter:
  """Record-level driver (the reference's Beam `DoFn`; ref :691): TFRecords of `Frame`
  protos in, TFRecords of `tf.Example`s out."""

  def __init__(self, emitter_fn=None):
    self._converter = FrameToTFE()
    self._emit = emitter_fn

  def process(self, item):  # pylint: disable=invalid-name
    ex = self._converter.process(item)
    if self._emit:
      self._emit(ex)
    return [ex]

  def ConvertFile(self, input_path, writer):
    from lingvo_b200 import ops  # pylint: disable=g-import-not-at-top
    y = ops.host().sequential_record_yielder('tfrecord:' + input_path, 1)
    n = 0
    while True:
      rec = y.next()
      if rec is None:
        return n
      writer.write(self._converter.process(rec[0]))
      n += 1
