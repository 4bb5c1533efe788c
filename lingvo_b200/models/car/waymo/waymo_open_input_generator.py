"""Waymo Open Dataset input (ref
`lingvo/tasks/car/waymo/waymo_open_input_generator.py`).

Records are the `tf.Example`s produced by `waymo/tools/waymo_proto_to_tfe.py`:
  laser_<LIDAR>_<ri1|ri2>   flat [n, 3 + 3] = xyz + (intensity, elongation, in-no-label-zone)
  labels / label_ids / bboxes_3d / bboxes_3d_num_points / label_metadata /
  {detection,single_frame_detection,tracking}_difficulties
  pose (4×4), run_segment, run_start_offset, time_of_day, location, weather
  image_<CAMERA>{,_shape,_pose,_intrinsics,_extrinsics,…}
  <LIDAR>_{ri1,ri2}{,_shape}, <LIDAR>_extrinsics, <LIDAR>_beam_inclinations (range images)
"""

from __future__ import annotations

import io

import numpy as np
import torch

from lingvo_b200.core import hyperparams
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.car import base_extractor
from lingvo_b200.models.car import geometry
from lingvo_b200.models.car import input_extractor
from lingvo_b200.models.car import input_preprocessors
from lingvo_b200.models.car.waymo import waymo_metadata

CAMERA_NAMES = ['FRONT', 'FRONT_LEFT', 'FRONT_RIGHT', 'SIDE_LEFT', 'SIDE_RIGHT']
LIDAR_NAMES = ['TOP', 'SIDE_LEFT', 'SIDE_RIGHT', 'FRONT', 'REAR']


def _Id16(b):
  b = b if isinstance(b, bytes) else str(b).encode()
  return np.frombuffer(b.ljust(64, b' ')[:64], np.uint8).copy()


def _PadOrTrim(x, n, fill=0):
  x = np.asarray(x)
  out = np.full((n,) + x.shape[1:], fill, x.dtype)
  k = min(n, len(x))
  out[:k] = x[:k]
  return out


class WaymoFrameMetadataExtractor(input_extractor.FieldsExtractor):
  """Pose and run metadata; can also drop frames by time of day / location / weather
  (ref :48)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('equality_filters', None,
             "[(field, value)] with field in time_of_day/location/weather: keep only matches.")
    return p

  def FeatureMap(self):
    return {'pose': (None, np.float32), 'run_segment': ((), bytes),
            'run_start_offset': ((), np.int64), 'time_of_day': ((), bytes),
            'location': ((), bytes), 'weather': ((), bytes)}

  def _Extract(self, features):
    pose = features['pose'].reshape(4, 4) if features['pose'].size == 16 else np.eye(
        4, dtype=np.float32)
    out = NestedMap(pose=pose.astype(np.float32), run_segment=_Id16(features['run_segment']),
                    run_start_offset=np.int64(features['run_start_offset']))
    for k in ('time_of_day', 'location', 'weather'):
      out[k] = _Id16(features[k])
    self._last_raw = {k: features[k] for k in ('time_of_day', 'location', 'weather')}
    return out

  def Filter(self, outputs):
    p = self.params
    for field, value in (p.equality_filters or []):
      got = self._last_raw[field]
      got = got.decode() if isinstance(got, bytes) else got
      if got != value:
        return input_extractor.BUCKET_UPPER_BOUND
    return 1

  def Shape(self):
    return NestedMap(pose=(4, 4), run_segment=(64,), run_start_offset=(), time_of_day=(64,),
                     location=(64,), weather=(64,))

  def DType(self):
    return NestedMap(pose=np.float32, run_segment=np.uint8, run_start_offset=np.int64,
                     time_of_day=np.uint8, location=np.uint8, weather=np.uint8)


class WaymoImageExtractor(input_extractor.FieldsExtractor):
  """Camera images with intrinsics / extrinsics / pose, one NestedMap per camera
  (ref :196)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('image_output_dtype', np.uint8, 'dtype of the emitted image.')
    p.Define('camera_names', list(CAMERA_NAMES), 'Cameras to read.')
    p.Define('image_shape', [1280, 1920, 3], 'Static image shape (smaller ones are padded).')
    p.Define('decode_image', True, 'Decode JPEG bytes (else emit zeros).')
    return p

  def FeatureMap(self):
    fm = {'pose': (None, np.float32)}
    for c in self.params.camera_names:
      fm['image_%s' % c] = (None, bytes)
      fm['image_%s_shape' % c] = (None, np.int64)
      fm['image_%s_pose' % c] = (None, np.float32)
      fm['image_%s_intrinsics' % c] = (None, np.float32)
      fm['image_%s_extrinsics' % c] = (None, np.float32)
    return fm

  def _Extract(self, features):
    p = self.params
    h, w, _ = p.image_shape
    out = NestedMap()
    for c in p.camera_names:
      img = np.zeros((h, w, 3), p.image_output_dtype)
      raw = features['image_%s' % c]
      if p.decode_image and raw:
        from PIL import Image  # pylint: disable=g-import-not-at-top
        arr = np.asarray(Image.open(io.BytesIO(raw[0])).convert('RGB'))
        img[:min(h, arr.shape[0]), :min(w, arr.shape[1])] = arr[:h, :w].astype(
            p.image_output_dtype)
      mat = lambda k, n: (features[k].reshape(n) if features[k].size == int(np.prod(n))
                          else np.zeros(n, np.float32)).astype(np.float32)
      out[c] = NestedMap(image=img, intrinsics=mat('image_%s_intrinsics' % c, (9,)),
                         extrinsics=mat('image_%s_extrinsics' % c, (4, 4)),
                         pose=mat('image_%s_pose' % c, (4, 4)))
    return out

  def Shape(self):
    p = self.params
    return NestedMap({c: NestedMap(image=tuple(p.image_shape), intrinsics=(9,),
                                   extrinsics=(4, 4), pose=(4, 4)) for c in p.camera_names})

  def DType(self):
    p = self.params
    return NestedMap({c: NestedMap(image=p.image_output_dtype, intrinsics=np.float32,
                                   extrinsics=np.float32, pose=np.float32)
                      for c in p.camera_names})


class WaymoLaserExtractor(input_extractor.LaserExtractor):
  """Merged point cloud of the selected lidars / returns (ref :359). Features:
  intensity, elongation, in-no-label-zone flag."""

  @classmethod
  def Params(cls):
    p = super().Params().Set(max_num_points=None, num_features=3)
    p.Define('lidar_names', list(LIDAR_NAMES), 'Lidars to merge.')
    p.Define('lidar_returns', ['ri1', 'ri2'], 'Returns to merge.')
    return p

  def _Names(self):
    p = self.params
    return ['laser_%s_%s' % (l, r) for l in p.lidar_names for r in p.lidar_returns]

  def FeatureMap(self):
    return {n: (None, np.float32) for n in self._Names()}

  def _Extract(self, features):
    p = self.params
    data = np.concatenate([features[n].reshape(-1, 3 + p.num_features) for n in self._Names()])
    out = self.PadOrTrim(data[:, :3], data[:, 3:])
    if p.max_num_points is None:
      del out['points_padding']
    return out

  def Shape(self):
    s = super().Shape()
    if self.params.max_num_points is None:
      del s['points_padding']
    return s

  def DType(self):
    d = super().DType()
    if self.params.max_num_points is None:
      del d['points_padding']
    return d


class WaymoLaserSceneflowExtractor(WaymoLaserExtractor):
  """Also emits per-point scene flow (`laser_<L>_<r>_flow` = vx, vy, vz, class)
  (ref :418)."""

  def FeatureMap(self):
    fm = super().FeatureMap()
    for n in self._Names():
      fm[n + '_flow'] = (None, np.float32)
    return fm

  def _Extract(self, features):
    p = self.params
    names = self._Names()
    data = np.concatenate([features[n].reshape(-1, 3 + p.num_features) for n in names])
    flow = np.concatenate([features[n + '_flow'].reshape(-1, 4) for n in names])
    n = len(data)
    m = p.max_num_points or n
    out = self.PadOrTrim(data[:, :3], data[:, 3:], rng=np.random.RandomState(0))
    f = np.zeros((m, 4), np.float32)
    f[:min(n, m)] = flow[:m]
    out.points_flow = f[:, :3]
    out.points_flow_class = f[:, 3].astype(np.int32)
    if p.max_num_points is None:
      del out['points_padding']
    return out

  def Shape(self):
    s = super().Shape()
    m = self.params.max_num_points
    s.points_flow, s.points_flow_class = (m, 3), (m,)
    return s

  def DType(self):
    d = super().DType()
    d.points_flow, d.points_flow_class = np.float32, np.int32
    return d


class WaymoLabelExtractor(input_extractor.FieldsExtractor):
  """3-D boxes, classes, difficulties, speed / acceleration (ref :483)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('max_num_objects', 512, 'Objects per frame.')
    p.Define('filter_labels', None, 'Label ids kept in bboxes_3d_mask.')
    return p

  def FeatureMap(self):
    i, f = (None, np.int64), (None, np.float32)
    return {'labels': i, 'label_ids': (None, bytes), 'detection_difficulties': i,
            'single_frame_detection_difficulties': i, 'tracking_difficulties': i,
            'bboxes_3d': f, 'bboxes_3d_num_points': i, 'label_metadata': f}

  def _Extract(self, features):
    p = self.params
    m = p.max_num_objects
    ft = features
    labels = _PadOrTrim(ft['labels'].astype(np.int32), m)
    boxes = ft['bboxes_3d'].reshape(-1, 7).astype(np.float32)
    mask = _PadOrTrim(np.ones(len(boxes), np.float32), m)
    meta = _PadOrTrim(ft['label_metadata'].reshape(-1, 4).astype(np.float32), m)
    unfiltered = mask.copy()
    if p.filter_labels:
      mask = mask * np.isin(labels, np.asarray(p.filter_labels)).astype(np.float32)
    pad_i = lambda k: _PadOrTrim(ft[k].astype(np.int32), m)
    return NestedMap(
        labels=labels, detection_difficulties=pad_i('detection_difficulties'),
        single_frame_detection_difficulties=pad_i('single_frame_detection_difficulties'),
        tracking_difficulties=pad_i('tracking_difficulties'), bboxes_3d=_PadOrTrim(boxes, m),
        bboxes_3d_mask=mask, bboxes_3d_num_points=pad_i('bboxes_3d_num_points'),
        unfiltered_bboxes_3d_mask=unfiltered, speed=meta[:, :2], acceleration=meta[:, 2:])

  def Shape(self):
    m = self.params.max_num_objects
    one = (m,)
    return NestedMap(labels=one, detection_difficulties=one,
                     single_frame_detection_difficulties=one, tracking_difficulties=one,
                     bboxes_3d=(m, 7), bboxes_3d_mask=one, bboxes_3d_num_points=one,
                     unfiltered_bboxes_3d_mask=one, speed=(m, 2), acceleration=(m, 2))

  def DType(self):
    i, f = np.int32, np.float32
    return NestedMap(labels=i, detection_difficulties=i, single_frame_detection_difficulties=i,
                     tracking_difficulties=i, bboxes_3d=f, bboxes_3d_mask=f,
                     bboxes_3d_num_points=i, unfiltered_bboxes_3d_mask=f, speed=f,
                     acceleration=f)


class RangeImageExtractor(input_extractor.FieldsExtractor):
  """Raw range images per lidar (ref :663): `[H, W, 4]` = range, intensity, elongation,
  no-label-zone for each return, plus the xyz of every pixel (`…_xyz`, computed from beam
  inclinations and extrinsics) and a validity mask."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('side_lasers', ['SIDE_LEFT', 'SIDE_RIGHT', 'FRONT', 'REAR'], 'Short-range lidars.')
    p.Define('side_ri_shape', [200, 600, 4], 'Range-image shape of the side lidars.')
    p.Define('top_lasers', ['TOP'], 'Top lidar.')
    p.Define('top_ri_shape', [64, 2650, 4], 'Range-image shape of the top lidar.')
    p.Define('returns', ['ri1', 'ri2'], 'Returns to read.')
    return p

  def _Lasers(self):
    p = self.params
    return [(l, p.top_ri_shape) for l in p.top_lasers] + [(l, p.side_ri_shape)
                                                          for l in p.side_lasers]

  def FeatureMap(self):
    fm = {'pose': (None, np.float32)}
    for laser, _ in self._Lasers():
      fm['%s_beam_inclinations' % laser] = (None, np.float32)
      fm['%s_extrinsics' % laser] = (None, np.float32)
      for r in self.params.returns:
        fm['%s_%s' % (laser, r)] = (None, np.float32)
        fm['%s_%s_shape' % (laser, r)] = (None, np.int64)
    return fm

  @staticmethod
  def PolarToCartesian(range_image, inclinations, extrinsics):
    """range `[H, W]`, per-row inclinations `[H]` → xyz `[H, W, 3]` in the vehicle frame."""
    h, w = range_image.shape
    az = np.linspace(np.pi, -np.pi, w, endpoint=False, dtype=np.float32)
    az = az - np.arctan2(extrinsics[1, 0], extrinsics[0, 0])
    incl = np.asarray(inclinations, np.float32)[::-1].reshape(h, 1)
    cos_i = np.cos(incl)
    xyz = np.stack([range_image * cos_i * np.cos(az), range_image * cos_i * np.sin(az),
                    range_image * np.sin(incl)], -1)
    hom = np.concatenate([xyz, np.ones((h, w, 1), np.float32)], -1) @ extrinsics.T
    return hom[..., :3].astype(np.float32)

  def _Extract(self, features):
    out = NestedMap()
    for laser, shape in self._Lasers():
      h, w, c = shape
      ext = features['%s_extrinsics' % laser]
      ext = ext.reshape(4, 4) if ext.size == 16 else np.eye(4, dtype=np.float32)
      incl = features['%s_beam_inclinations' % laser]
      if incl.size != h:
        incl = np.linspace(-0.3, 0.04, h).astype(np.float32)
      entry = NestedMap()
      for r in self.params.returns:
        raw = features['%s_%s' % (laser, r)]
        ri = np.zeros((h, w, c), np.float32)
        if raw.size:
          shp = features['%s_%s_shape' % (laser, r)]
          src = raw.reshape(tuple(int(s) for s in shp)) if shp.size == 3 else raw.reshape(h, w, c)
          ri[:min(h, src.shape[0]), :min(w, src.shape[1])] = src[:h, :w, :c]
        entry[r] = ri
        entry[r + '_xyz'] = self.PolarToCartesian(ri[..., 0], incl, ext.astype(np.float32))
        entry[r + '_mask'] = (ri[..., 0] > 0).astype(np.float32)
      out[laser] = entry
    return out

  def Shape(self):
    out = NestedMap()
    for laser, shape in self._Lasers():
      h, w, c = shape
      e = NestedMap()
      for r in self.params.returns:
        e[r], e[r + '_xyz'], e[r + '_mask'] = (h, w, c), (h, w, 3), (h, w)
      out[laser] = e
    return out

  def DType(self):
    return self.Shape().Transform(lambda _: np.float32)


class FilterNLZPoints(input_preprocessors.Preprocessor):
  """Removes points inside no-label zones (3rd laser feature == 1) (ref :1003)."""

  def TransformFeatures(self, features):
    las = features.lasers
    if las.get('points_padding') is not None:
      raise ValueError('FilterNLZPoints preprocessor does not support padded lasers.')
    keep = las.points_feature[:, 2] != 1.0
    las.points_xyz, las.points_feature = las.points_xyz[keep], las.points_feature[keep]
    return features


class CellCenterToBestCamera(input_preprocessors.Preprocessor):
  """For every `cell_center_xyz`, finds the camera that sees it (projection inside the
  image, positive depth; first match in `camera_names`) and its pixel (ref :1227). Adds
  `cell_center_camera_id [C]` (−1: none) and `cell_center_pixel [C, 2]`."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('camera_names', list(CAMERA_NAMES), 'Cameras in priority order.')
    p.Define('image_shape', [1280, 1920], '(H, W) of the images.')
    return p

  def TransformFeatures(self, features):
    p = self.params
    pts = features.cell_center_xyz
    n = pts.shape[0]
    cam_id = torch.full((n,), -1, dtype=torch.int32)
    pix = torch.zeros(n, 2)
    hom = torch.cat([pts, torch.ones(n, 1)], 1)
    for ci, name in enumerate(p.camera_names):
      cam = features.images[name]
      k = cam.intrinsics
      f_u, f_v, c_u, c_v = float(k[0]), float(k[1]), float(k[2]), float(k[3])
      cam_pts = hom @ torch.linalg.inv(cam.extrinsics).t()      # vehicle → camera (x forward)
      depth = cam_pts[:, 0]
      u = c_u - f_u * cam_pts[:, 1] / depth.clamp_min(1e-6)
      v = c_v - f_v * cam_pts[:, 2] / depth.clamp_min(1e-6)
      ok = (depth > 0) & (u >= 0) & (u < p.image_shape[1]) & (v >= 0) & (v < p.image_shape[0])
      take = ok & (cam_id < 0)
      cam_id = torch.where(take, torch.full_like(cam_id, ci), cam_id)
      pix = torch.where(take.unsqueeze(1), torch.stack([u, v], 1), pix)
    features.cell_center_camera_id, features.cell_center_pixel = cam_id, pix
    return features

  def TransformShapes(self, shapes):
    n = shapes.cell_center_xyz[0]
    shapes.cell_center_camera_id, shapes.cell_center_pixel = (n,), (n, 2)
    return shapes

  def TransformDTypes(self, dtypes):
    dtypes.cell_center_camera_id, dtypes.cell_center_pixel = np.int32, np.float32
    return dtypes


class RescaleResizeImages(input_preprocessors.Preprocessor):
  """Converts camera images to float in [−1, 1] and resizes them by `resize_ratio`
  (ref :1288)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('camera_names', list(CAMERA_NAMES), 'Cameras to process.')
    p.Define('resize_ratio', 1.0, '1.0: no resizing.')
    p.Define('rescale', True, 'Map [0, 255] → [−1, 1].')
    return p

  def TransformFeatures(self, features):
    p = self.params
    for name in p.camera_names:
      img = features.images[name].image.float()
      if p.rescale:
        img = img / 127.5 - 1.0
      if p.resize_ratio != 1.0:
        chw = img.permute(2, 0, 1).unsqueeze(0)
        chw = torch.nn.functional.interpolate(chw, scale_factor=p.resize_ratio, mode='bilinear',
                                              align_corners=False)
        img = chw[0].permute(1, 2, 0)
      features.images[name].image = img
    return features

  def TransformShapes(self, shapes):
    p = self.params
    for name in p.camera_names:
      h, w, c = shapes.images[name].image
      shapes.images[name].image = (int(h * p.resize_ratio), int(w * p.resize_ratio), c)
    return shapes

  def TransformDTypes(self, dtypes):
    for name in self.params.camera_names:
      dtypes.images[name].image = np.float32
    return dtypes


class WaymoSparseLaser(base_extractor._BaseExtractor):  # pylint: disable=protected-access
  """StarNet-style input for Waymo (ref :1046): frame metadata + labels + merged lasers →
  NLZ filtering → world augmentation (train) → sparse cell sampling → anchors."""

  @classmethod
  def Params(cls):
    ex = hyperparams.Params()
    ex.Define('metadata', WaymoFrameMetadataExtractor.Params(), '')
    ex.Define('labels', WaymoLabelExtractor.Params(), '')
    ex.Define('lasers', WaymoLaserExtractor.Params(), '')
    p = super().Params(ex)
    ip = input_preprocessors
    pre = hyperparams.Params()
    pre.Define('filter_nlz_points', FilterNLZPoints.Params(), '')
    pre.Define('viz_copy', ip.CreateDecoderCopy.Params().Set(
        pad_lasers=ip.PadLaserFeatures.Params().Set(max_num_points=240000)), '')
    pre.Define('select_centers', ip.SparseCenterSelector.Params().Set(num_cell_centers=1024), '')
    pre.Define('gather_features', ip.SparseCellGatherFeatures.Params().Set(
        num_points_per_cell=128, max_distance=2.75), '')
    pre.Define('tile_anchors', ip.TileAnchorBBoxes.Params().Set(
        anchor_box_dimensions=[[4.7, 2.1, 1.7]] * 2, anchor_box_offsets=[[0.0, 0.0, 0.0]] * 2,
        anchor_box_rotations=[0.0, np.pi / 2]), '')
    pre.Define('assign_anchors', ip.AnchorAssignment.Params(), '')
    pre.Define('pad_lasers', ip.PadLaserFeatures.Params().Set(max_num_points=240000), '')
    p.preprocessors = pre
    p.preprocessors_order = ['filter_nlz_points', 'viz_copy', 'select_centers',
                             'gather_features', 'tile_anchors', 'assign_anchors', 'pad_lasers']
    p.file_pattern = ''
    return p

  @property
  def class_names(self):
    return waymo_metadata.WaymoMetadata().ClassNames()



