"""KITTI object-detection input (ref `lingvo/tasks/car/kitti_input_generator.py`).

Records are `tf.Example`s written by `tools/kitti_exporter.py`:
  pointcloud/{xyz,reflectance}, image/{encoded,format,height,width,source_id},
  transform/{velo_to_image_plane [3,4], velo_to_camera [4,4], camera_to_velo [4,4]},
  object/{label, has_3d_info, occlusion, truncation, image/bbox/{xmin,xmax,ymin,ymax},
          velo/bbox/{xyz, dim_xyz, phi}}.

Extractors: `KITTILaserExtractor`, `KITTIImageExtractor`, `KITTILabelExtractor`;
generators: `KITTILaser` (raw points), `KITTISparseLaser` (StarNet cells),
`KITTIGrid` (PointPillars grid).
"""

from __future__ import annotations

import io

import numpy as np

from lingvo_b200.core import hyperparams
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.car import base_extractor
from lingvo_b200.models.car import geometry
from lingvo_b200.models.car import input_extractor
from lingvo_b200.models.car import input_preprocessors
from lingvo_b200.models.car import kitti_metadata

KITTI_CLASS_NAMES = kitti_metadata.KITTIMetadata().ClassNames()


def _NestedMapToParams(nmap):
  p = hyperparams.Params()
  for k, v in nmap.items():
    p.Define(k, v, '')
  return p


def ComputeKITTIDifficulties(box_image_height, occlusion, truncation):
  """3 = easy, 2 = moderate, 1 = hard, 0 = ignored (KITTI protocol; ref :43)."""
  h, o, t = (np.asarray(a, np.float32) for a in (box_image_height, occlusion, truncation))
  easy = ((h >= 40.0) & (o <= 0.0) & (t <= 0.15)).astype(np.int32) * 3
  moderate = ((h >= 25.0) & (o <= 1.0) & (t <= 0.3)).astype(np.int32) * 2
  hard = ((h >= 25.0) & (o <= 2.0) & (t <= 0.5)).astype(np.int32)
  return np.maximum(np.maximum(easy, moderate), hard)


def _PadOrTrim(x, n, fill=0):
  x = np.asarray(x)
  out = np.full((n,) + x.shape[1:], fill, x.dtype)
  k = min(n, len(x))
  out[:k] = x[:k]
  return out


class KITTILaserExtractor(input_extractor.LaserExtractor):
  """ref :62."""

  @classmethod
  def Params(cls):
    return super().Params().Set(max_num_points=None, num_features=1)

  def FeatureMap(self):
    return {'pointcloud/xyz': (None, np.float32), 'pointcloud/reflectance': (None, np.float32)}

  def _Extract(self, features):
    p = self.params
    xyz = features['pointcloud/xyz'].reshape(-1, 3)
    refl = features['pointcloud/reflectance'].reshape(-1, p.num_features)
    return self.PadOrTrim(xyz, refl)


class KITTIImageExtractor(input_extractor.FieldsExtractor):
  """Camera image + calibration (ref :98). The decoded image is resized to a static
  `image_shape` so it can be batched."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('decode_image', True, 'Decode and emit the image.')
    p.Define('image_shape', (375, 1242, 3), 'Static (H, W, 3) of the emitted image.')
    return p

  def FeatureMap(self):
    fm = {
        'image/format': ((), bytes), 'image/height': ((), np.int64), 'image/width': ((), np.int64),
        'image/source_id': ((), bytes),
        'transform/velo_to_image_plane': ((3, 4), np.float32),
        'transform/velo_to_camera': ((4, 4), np.float32),
        'transform/camera_to_velo': ((4, 4), np.float32),
    }
    if self.params.decode_image:
      fm['image/encoded'] = ((), bytes)
    return fm

  def _Extract(self, features):
    p = self.params
    out = NestedMap(
        width=np.int64(features['image/width']), height=np.int64(features['image/height']),
        velo_to_image_plane=features['transform/velo_to_image_plane'],
        velo_to_camera=features['transform/velo_to_camera'],
        camera_to_velo=features['transform/camera_to_velo'])
    if p.decode_image:
      from PIL import Image  # pylint: disable=g-import-not-at-top
      img = Image.open(io.BytesIO(features['image/encoded'])).convert('RGB')
      h, w, _ = p.image_shape
      if img.size != (w, h):
        img = img.resize((w, h))
      out.image = np.asarray(img, np.float32) / 255.0
    return out

  def Shape(self):
    p = self.params
    s = NestedMap(width=(), height=(), velo_to_image_plane=(3, 4), velo_to_camera=(4, 4),
                  camera_to_velo=(4, 4))
    if p.decode_image:
      s.image = tuple(p.image_shape)
    return s

  def DType(self):
    d = NestedMap(width=np.int64, height=np.int64, velo_to_image_plane=np.float32,
                  velo_to_camera=np.float32, camera_to_velo=np.float32)
    if self.params.decode_image:
      d.image = np.float32
    return d


class KITTILabelExtractor(input_extractor.FieldsExtractor):
  """2-D and 3-D boxes, classes and KITTI difficulties, padded to `max_num_objects`
  (ref :220)."""

  KITTI_CLASS_NAMES = KITTI_CLASS_NAMES

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('max_num_objects', 50, 'Objects per example.')
    p.Define('filter_labels', None, 'Label ids kept in bboxes_3d_mask (None: all).')
    return p

  def FeatureMap(self):
    v = lambda dt: (None, dt)
    return {
        'image/source_id': ((), bytes),
        'object/image/bbox/xmin': v(np.float32), 'object/image/bbox/xmax': v(np.float32),
        'object/image/bbox/ymin': v(np.float32), 'object/image/bbox/ymax': v(np.float32),
        'object/label': (None, bytes), 'object/has_3d_info': v(np.int64),
        'object/occlusion': v(np.int64), 'object/truncation': v(np.float32),
        'object/velo/bbox/xyz': v(np.float32), 'object/velo/bbox/dim_xyz': v(np.float32),
        'object/velo/bbox/phi': v(np.float32),
        'transform/velo_to_image_plane': ((3, 4), np.float32),
    }

  def _Extract(self, features):
    import torch  # pylint: disable=g-import-not-at-top
    p = self.params
    m = p.max_num_objects
    f = features
    xmin, xmax = f['object/image/bbox/xmin'], f['object/image/bbox/xmax']
    ymin, ymax = f['object/image/bbox/ymin'], f['object/image/bbox/ymax']
    bboxes = np.stack([ymin, xmin, ymax, xmax], 1).reshape(-1, 4).astype(np.float32)
    n = len(bboxes)
    bboxes_padding = 1.0 - _PadOrTrim(np.ones(n, np.float32), m)
    bboxes_3d = np.concatenate([f['object/velo/bbox/xyz'].reshape(-1, 3),
                                f['object/velo/bbox/dim_xyz'].reshape(-1, 3),
                                f['object/velo/bbox/phi'].reshape(-1, 1)], 1).astype(np.float32)
    cx, cy, dx, dy = bboxes_3d[:, 0], bboxes_3d[:, 1], bboxes_3d[:, 3], bboxes_3d[:, 4]
    bboxes_td = np.stack([cy - dy / 2, cx - dx / 2, cy + dy / 2, cx + dx / 2], -1)
    mask3d = _PadOrTrim(f['object/has_3d_info'].astype(np.float32), m)
    height = _PadOrTrim(ymax - ymin, m) * mask3d
    occlusion = _PadOrTrim(f['object/occlusion'].astype(np.float32), m) * mask3d
    truncation = _PadOrTrim(f['object/truncation'], m) * mask3d
    difficulties = ComputeKITTIDifficulties(height, occlusion, truncation)
    corners = geometry.BBoxCorners(torch.from_numpy(bboxes_3d)).reshape(-1, 3)
    proj = geometry.PointsToImagePlane(
        corners, torch.from_numpy(f['transform/velo_to_image_plane']))[:, :2] if n else \
        torch.zeros(0, 2)
    proj = _PadOrTrim(proj.reshape(-1, 8, 2).numpy(), m)
    texts = [t.decode('utf-8') if isinstance(t, bytes) else t for t in f['object/label']]
    names = self.KITTI_CLASS_NAMES
    labels = _PadOrTrim(np.asarray([names.index(t) if t in names else 0 for t in texts],
                                   np.int32), m)
    filtered = mask3d
    td_mask = mask3d.copy()
    if p.filter_labels is not None:
      ok = np.isin(labels, np.asarray(p.filter_labels)).astype(np.float32)
      bboxes_padding = 1.0 - ok * (1.0 - bboxes_padding)
      filtered = mask3d * ok
      td_mask = td_mask * ok
    return NestedMap(
        source_id=f['image/source_id'], bboxes_count=np.int32(n), bboxes=_PadOrTrim(bboxes, m),
        bboxes_padding=bboxes_padding, bboxes_3d=_PadOrTrim(bboxes_3d, m),
        bboxes_3d_mask=filtered.astype(np.float32),
        unfiltered_bboxes_3d_mask=mask3d.astype(np.float32),
        bboxes3d_proj_to_image_plane=proj.astype(np.float32),
        bboxes_td=_PadOrTrim(bboxes_td.astype(np.float32), m), bboxes_td_mask=td_mask,
        bboxes_3d_num_points=np.zeros(m, np.int32), labels=labels,
        texts=(texts + [''] * m)[:m], box_image_height=height.astype(np.float32),
        occlusion=occlusion.astype(np.float32), truncation=truncation.astype(np.float32),
        difficulties=difficulties.astype(np.int32))

  def Shape(self):
    m = self.params.max_num_objects
    return NestedMap(
        source_id=(), bboxes_count=(), bboxes=(m, 4), bboxes_padding=(m,), bboxes_3d=(m, 7),
        bboxes_3d_mask=(m,), unfiltered_bboxes_3d_mask=(m,),
        bboxes3d_proj_to_image_plane=(m, 8, 2), bboxes_td=(m, 4), bboxes_td_mask=(m,),
        bboxes_3d_num_points=(m,), labels=(m,), texts=(m,), box_image_height=(m,),
        occlusion=(m,), truncation=(m,), difficulties=(m,))

  def DType(self):
    f = np.float32
    return NestedMap(
        source_id=bytes, bboxes_count=np.int32, bboxes=f, bboxes_padding=f, bboxes_3d=f,
        bboxes_3d_mask=f, unfiltered_bboxes_3d_mask=f, bboxes3d_proj_to_image_plane=f,
        bboxes_td=f, bboxes_td_mask=f, bboxes_3d_num_points=np.int32, labels=np.int32,
        texts=bytes, box_image_height=f, occlusion=f, truncation=f, difficulties=np.int32)


class KITTIBase(base_extractor._BaseExtractor):  # pylint: disable=protected-access
  """ref :478."""

  @classmethod
  def Params(cls, *args, **kwargs):
    p = super().Params(*args, **kwargs)
    p.file_datasource = None
    p.file_pattern = ''
    p.cpu_passthrough_keys = ['labels.source_id', 'labels.texts']
    return p

  @property
  def class_names(self):
    return KITTI_CLASS_NAMES

  def ProcessRecord(self, record, source_id=0):
    out = super().ProcessRecord(record, source_id)
    if out is None:
      return None
    feats, bucket = out
    # strings cannot be stacked by the native batcher: ids become fixed-width byte rows,
    # free-text lists are dropped (class ids carry the same information)
    def _Fix(key, v):
      if isinstance(v, (bytes, str)):
        b = v if isinstance(v, bytes) else v.encode()
        return np.frombuffer(b.ljust(16, b' ')[:16], np.uint8).copy()
      if isinstance(v, list):
        return None
      return np.asarray(v) if np.isscalar(v) else v
    feats = feats.TransformWithKey(_Fix).Filter(lambda v: v is not None)
    return feats, bucket


def _Extractors(**kwargs):
  ex = hyperparams.Params()
  ex.Define('labels', KITTILabelExtractor.Params(), '')
  ex.Define('lasers', KITTILaserExtractor.Params(), '')
  ex.Define('images', KITTIImageExtractor.Params().Set(**kwargs), '')
  return ex


class KITTILaser(KITTIBase):
  """Labels + images + the raw (padded) point cloud (ref :500)."""

  @classmethod
  def Params(cls):
    p = super().Params(_Extractors(decode_image=False))
    pre = hyperparams.Params()
    pre.Define('count_points', input_preprocessors.CountNumberOfPointsInBoxes3D.Params(), '')
    pre.Define('viz_copy', input_preprocessors.CreateDecoderCopy.Params(), '')
    pre.Define('pad_lasers', input_preprocessors.PadLaserFeatures.Params().Set(
        max_num_points=72000), '')
    p.preprocessors = pre
    p.preprocessors_order = ['count_points', 'viz_copy', 'pad_lasers']
    return p


class KITTISparseLaser(KITTIBase):
  """StarNet input (ref :536): sampled cell centres with their neighbourhoods, tiled
  anchors and assignments."""

  @classmethod
  def Params(cls):
    p = super().Params(_Extractors(decode_image=False))
    ip = input_preprocessors
    pre = hyperparams.Params()
    pre.Define('count_points', ip.CountNumberOfPointsInBoxes3D.Params(), '')
    pre.Define('viz_copy', ip.CreateDecoderCopy.Params(), '')
    pre.Define('keep_xyz_range', ip.DropLaserPointsOutOfRange.Params().Set(
        keep_x_range=(0.0, 70.4), keep_y_range=(-40.0, 40.0)), '')
    pre.Define('select_centers', ip.SparseCenterSelector.Params().Set(num_cell_centers=256), '')
    pre.Define('gather_features', ip.SparseCellGatherFeatures.Params().Set(
        num_points_per_cell=128, max_distance=3.0), '')
    pre.Define('tile_anchors', ip.SparseCarV1AnchorBoxSettings.Update(
        ip.TileAnchorBBoxes.Params()), '')
    pre.Define('assign_anchors', ip.AnchorAssignment.Params(), '')
    pre.Define('pad_lasers', ip.PadLaserFeatures.Params().Set(max_num_points=72000), '')
    p.preprocessors = pre
    p.preprocessors_order = ['count_points', 'viz_copy', 'keep_xyz_range', 'select_centers',
                             'gather_features', 'tile_anchors', 'assign_anchors', 'pad_lasers']
    return p


class KITTIGrid(KITTIBase):
  """PointPillars input (ref :594): pillars + dense anchors and assignments."""

  @classmethod
  def Params(cls):
    p = super().Params(_Extractors(decode_image=False))
    ip = input_preprocessors
    gs = ip._PointPillarGridSettings   # pylint: disable=protected-access
    pre = hyperparams.Params()
    pre.Define('count_points', ip.CountNumberOfPointsInBoxes3D.Params(), '')
    pre.Define('viz_copy', ip.CreateDecoderCopy.Params(), '')
    pre.Define('keep_xyz_range', ip.DropLaserPointsOutOfRange.Params().Set(
        keep_x_range=gs.GRID_X_RANGE, keep_y_range=gs.GRID_Y_RANGE,
        keep_z_range=gs.GRID_Z_RANGE), '')
    grid = ip.PointsToGrid.Params().Set(num_points_per_cell=100)
    gs.UpdateGridParams(grid)
    pre.Define('points_to_grid', grid, '')
    pre.Define('grid_to_pillars', ip.GridToPillars.Params(), '')
    anchors = ip.GridAnchorCenters.Params()
    gs.UpdateAnchorGridParams(anchors)
    pre.Define('grid_anchor_centers', anchors, '')
    pre.Define('tile_anchors', ip.PointPillarAnchorBoxSettingsCar.Update(
        ip.TileAnchorBBoxes.Params()), '')
    pre.Define('assign_anchors', ip.AnchorAssignment.Params().Set(
        foreground_assignment_threshold=0.6, background_assignment_threshold=0.45), '')
    pre.Define('pad_lasers', ip.PadLaserFeatures.Params().Set(max_num_points=72000), '')
    p.preprocessors = pre
    p.preprocessors_order = ['count_points', 'viz_copy', 'keep_xyz_range', 'points_to_grid',
                             'grid_to_pillars', 'grid_anchor_centers', 'tile_anchors',
                             'assign_anchors', 'pad_lasers']
    return p
