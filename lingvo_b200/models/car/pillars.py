"""PointPillars (ref `lingvo/tasks/car/pillars.py`).

points → pillars (native grid assignment) → per-pillar PointNet (linear+BN+ReLU,
max over points) → scatter to a BEV pseudo-image → 3-block strided conv backbone
with upsample+concat → per-anchor class logits and 7-DOF box residuals.
Loss: focal classification + smooth-L1 localisation (+ direction-invariant sin Δφ).
"""

from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from lingvo_b200 import ops
from lingvo_b200.core import base_layer
from lingvo_b200.core import base_model
from lingvo_b200.core import bn_layers
from lingvo_b200.core import layers
from lingvo_b200.core import py_utils
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.car import detection_3d_lib


class PointsToPillars(base_layer.BaseLayer):
  """Pre-processing: `[N, D]` lasers → padded pillars + grid locations (ref
  `input_preprocessors.py` GridToPillars / native `point_grid_op.cc`)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('grid_x', (-40.0, 40.0, 32), '(min, max, cells).')
    p.Define('grid_y', (-40.0, 40.0, 32), '(min, max, cells).')
    p.Define('max_pillars', 256, 'Max non-empty pillars.')
    p.Define('points_per_pillar', 16, 'Max points per pillar.')
    return p

  def FProp(self, theta, points):
    p = self.params
    pts = points.detach().cpu().numpy().astype(np.float32)
    pp, xy, cnt, used = ops.host().points_to_pillars(
        pts, p.grid_x[0], p.grid_x[1], p.grid_y[0], p.grid_y[1], p.grid_x[2], p.grid_y[2],
        p.max_pillars, p.points_per_pillar)
    return NestedMap(pillar_points=torch.from_numpy(pp), pillar_locations=torch.from_numpy(xy).long(),
                     pillar_count=torch.from_numpy(cnt).long(), num_pillars=used)


class PillarsFeaturizer(base_layer.BaseLayer):
  """Decorates points with offsets to the pillar mean/centre, then PointNet (ref :60-200)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('point_dims', 4, 'Raw point dims (x, y, z, intensity).')
    p.Define('num_features', 64, 'Pillar feature dim.')
    p.Define('grid_x', (-40.0, 40.0, 32), 'Grid.')
    p.Define('grid_y', (-40.0, 40.0, 32), 'Grid.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    idim = p.point_dims + 3 + 2
    self.CreateChild('fc', layers.FCLayer.Params().Set(
        input_dim=idim, output_dim=p.num_features, activation='NONE'))
    self.CreateChild('bn', bn_layers.BatchNormLayer.Params().Set(dim=p.num_features))

  def FProp(self, theta, pillar_points, pillar_locations, pillar_count):
    """[B,P,K,D], [B,P,2], [B,P] → BEV image [B, nx, ny, F]."""
    p = self.params
    b, n_p, k, _ = pillar_points.shape
    valid = (torch.arange(k, device=pillar_points.device).view(1, 1, k) <
             pillar_count.unsqueeze(-1)).float().unsqueeze(-1)
    xyz = pillar_points[..., :3]
    mean = (xyz * valid).sum(2, keepdim=True) / pillar_count.clamp_min(1).view(b, n_p, 1, 1)
    cx = p.grid_x[0] + (pillar_locations[..., 0].float() + 0.5) * (p.grid_x[1] - p.grid_x[0]) / p.grid_x[2]
    cy = p.grid_y[0] + (pillar_locations[..., 1].float() + 0.5) * (p.grid_y[1] - p.grid_y[0]) / p.grid_y[2]
    centre = torch.stack([cx, cy], -1).unsqueeze(2)
    feats = torch.cat([pillar_points, xyz - mean, xyz[..., :2] - centre], -1) * valid
    h = self.fc.FProp(theta.fc, feats)
    h = self.bn.FProp(theta.bn, h, 1.0 - valid)
    h = torch.relu(h) * valid + (valid - 1.0) * 1e9
    pillar_feat = h.max(2).values * (pillar_count > 0).float().unsqueeze(-1)   # [B,P,F]
    image = torch.zeros(b, p.grid_x[2], p.grid_y[2], p.num_features, device=h.device,
                        dtype=pillar_feat.dtype)
    bi = torch.arange(b, device=h.device).view(b, 1).expand(b, n_p)
    occ = pillar_count > 0
    image[bi[occ], pillar_locations[..., 0][occ], pillar_locations[..., 1][occ]] = pillar_feat[occ]
    return image


class PillarsBackbone(base_layer.BaseLayer):
  """Three strided conv blocks, each upsampled back and concatenated (ref :200-330)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('input_dim', 64, 'Input channels.')
    p.Define('block_dims', (64, 128, 256), 'Channels per block.')
    p.Define('block_layers', (2, 2, 2), 'Convs per block (after the strided one).')
    p.Define('upsample_dim', 64, 'Channels of every upsampled branch.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    convs, ups = [], []
    idim = p.input_dim
    for bi, (dim, nl) in enumerate(zip(p.block_dims, p.block_layers)):
      for li in range(nl + 1):
        convs.append(layers.Conv2DLayer.Params().Set(
            name='b%d_c%d' % (bi, li), filter_shape=(3, 3, idim if li == 0 else dim, dim),
            filter_stride=(2, 2) if (li == 0 and bi > 0) else (1, 1), batch_norm=True,
            activation='RELU'))
      idim = dim
      ups.append(layers.Conv2DLayer.Params().Set(
          name='up%d' % bi, filter_shape=(1, 1, dim, p.upsample_dim), filter_stride=(1, 1),
          batch_norm=True, activation='RELU'))
    self.CreateChildren('convs', convs)
    self.CreateChildren('ups', ups)
    self._per_block = [nl + 1 for nl in p.block_layers]

  @property
  def output_dim(self):
    return self.params.upsample_dim * len(self.params.block_dims)

  def FProp(self, theta, image):
    x = image
    ci = 0
    outs = []
    h, w = image.shape[1], image.shape[2]
    for bi, n in enumerate(self._per_block):
      for _ in range(n):
        x = self.convs[ci].FProp(theta.convs[ci], x)
        x = x[0] if isinstance(x, tuple) else x
        ci += 1
      u = self.ups[bi].FProp(theta.ups[bi], x)
      u = u[0] if isinstance(u, tuple) else u
      if u.shape[1] != h:
        u = F.interpolate(u.permute(0, 3, 1, 2), size=(h, w), mode='nearest').permute(0, 2, 3, 1)
      outs.append(u)
    return torch.cat(outs, -1)


class ModelV1(base_model.BaseTask):
  """PointPillars detector (ref `pillars.py` ModelV1 :330-620)."""

  @classmethod
  def Params(cls, num_classes=2):
    p = super().Params()
    p.Define('num_classes', num_classes, 'Classes incl. background (class 0).')
    p.Define('featurizer', PillarsFeaturizer.Params(), 'Pillar featurizer.')
    p.Define('backbone', PillarsBackbone.Params(), 'BEV backbone.')
    p.Define('anchor_box_dimensions', [[3.9, 1.6, 1.56]] * 2, 'Anchor sizes.')
    p.Define('anchor_box_offsets', [[0., 0., -1.0]] * 2, 'Anchor offsets.')
    p.Define('anchor_box_rotations', [0.0, math.pi / 2], 'Anchor headings.')
    p.Define('focal_loss_alpha', 0.25, 'Focal α.')
    p.Define('focal_loss_gamma', 2.0, 'Focal γ.')
    p.Define('localization_loss_weight', 2.0, 'Localisation weight.')
    p.Define('classification_loss_weight', 1.0, 'Classification weight.')
    p.Define('nms_iou_threshold', 0.3, 'NMS IoU.')
    p.Define('nms_score_threshold', 0.05, 'NMS score threshold.')
    p.Define('max_nms_boxes', 32, 'Boxes kept per class.')
    p.name = 'pillars'
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    self._utils = detection_3d_lib.Utils3D()
    self.CreateChild('featurizer', p.featurizer)
    self.CreateChild('backbone', p.backbone.Copy().Set(input_dim=p.featurizer.num_features))
    a = len(p.anchor_box_rotations)
    odim = self.backbone.output_dim
    self.CreateChild('cls_head', layers.Conv2DLayer.Params().Set(
        filter_shape=(1, 1, odim, a * p.num_classes), filter_stride=(1, 1), batch_norm=False,
        activation='NONE', bias=True))
    self.CreateChild('reg_head', layers.Conv2DLayer.Params().Set(
        filter_shape=(1, 1, odim, a * 7), filter_stride=(1, 1), batch_norm=False,
        activation='NONE', bias=True))

  def Anchors(self, device):
    p = self.params
    f = p.featurizer
    centers = self._utils.CreateDenseCoordinates(
        [(f.grid_x[0], f.grid_x[1], f.grid_x[2]), (f.grid_y[0], f.grid_y[1], f.grid_y[2])],
        center_in_cell=True)
    centers = torch.cat([centers, torch.zeros(centers.shape[0], 1)], -1)
    return self._utils.MakeAnchorBoxes(centers, p.anchor_box_dimensions, p.anchor_box_offsets,
                                       p.anchor_box_rotations).to(device)     # [X·Y, A, 7]

  def ComputePredictions(self, theta, batch):
    p = self.params
    img = self.featurizer.FProp(theta.featurizer, batch.pillar_points, batch.pillar_locations,
                                batch.pillar_count)
    feat = self.backbone.FProp(theta.backbone, img)
    cls = self.cls_head.FProp(theta.cls_head, feat)
    reg = self.reg_head.FProp(theta.reg_head, feat)
    cls = cls[0] if isinstance(cls, tuple) else cls
    reg = reg[0] if isinstance(reg, tuple) else reg
    b = cls.shape[0]
    a = len(p.anchor_box_rotations)
    return NestedMap(classification_logits=cls.reshape(b, -1, a, p.num_classes),
                     residuals=reg.reshape(b, -1, a, 7))

  def ComputeLoss(self, theta, predictions, batch):
    p = self.params
    u = self._utils
    logits, res = predictions.classification_logits, predictions.residuals
    b = logits.shape[0]
    dev = logits.device
    anchors = self.Anchors(dev).reshape(-1, 7)
    cls_losses, reg_losses, n_fg = [], [], 0.0
    for i in range(b):
      asg = u.AssignAnchors(anchors, batch.bboxes[i], batch.labels[i], batch.bboxes_mask[i])
      lab = asg['assigned_gt_labels'].to(dev)
      one_hot = F.one_hot(lab, p.num_classes).float()
      cl = u.SigmoidFocalLoss(logits[i].reshape(-1, p.num_classes), one_hot,
                              p.focal_loss_alpha, p.focal_loss_gamma)
      cl = cl[:, 1:].sum(-1) * asg['assigned_cls_mask'].to(dev)
      tgt = u.LocalizationResiduals(anchors, asg['assigned_gt_bbox'].to(dev))
      pred = res[i].reshape(-1, 7)
      # heading: penalise sin(Δφ) so a flipped box costs nothing
      d_rot = torch.sin(pred[:, 6:] - tgt[:, 6:])
      rl = torch.cat([u.ScaledHuberLoss(tgt[:, :6], pred[:, :6], delta=1.0 / 9.0),
                      u.ScaledHuberLoss(torch.zeros_like(d_rot), d_rot, delta=1.0 / 9.0)], -1).sum(-1)
      fg = asg['assigned_reg_mask'].to(dev)
      cls_losses.append(cl.sum())
      reg_losses.append((rl * fg).sum())
      n_fg += float(fg.sum())
    norm = max(n_fg, 1.0)
    cls_loss = torch.stack(cls_losses).sum() / norm
    reg_loss = torch.stack(reg_losses).sum() / norm
    loss = p.classification_loss_weight * cls_loss + p.localization_loss_weight * reg_loss
    return NestedMap(loss=(loss, float(b)), classification_loss=(cls_loss, float(b)),
                     localization_loss=(reg_loss, float(b)),
                     num_foreground=(torch.tensor(n_fg / b), float(b))), NestedMap()

  def Decode(self, batch):
    p = self.params
    with torch.no_grad():
      pred = self.ComputePredictions(self.theta, batch)
      b = pred.residuals.shape[0]
      anchors = self.Anchors(pred.residuals.device).reshape(1, -1, 7).expand(b, -1, 7)
      boxes = self._utils.ResidualsToBBoxes(anchors, pred.residuals.reshape(b, -1, 7))
      scores = torch.sigmoid(pred.classification_logits.reshape(b, -1, p.num_classes))
      idx, mask = self._utils.BatchedNMSIndices(
          boxes, scores, p.nms_iou_threshold, p.nms_score_threshold, p.max_nms_boxes)
    return NestedMap(per_class_predicted_bboxes=boxes, per_class_scores=scores,
                     per_class_indices=idx, per_class_valid_mask=mask)


# --------------------------------------------------------------------------------------
# Reference-named building blocks (ref `pillars.py`: PointsToGridFeaturizer :33, Builder
# :92, LossNormType :325, DynamicVoxelizationFeaturizer :647). They compose the same model
# from `builder_lib` recipes and consume the `KITTIGrid` / dynamic-voxel input formats.
# --------------------------------------------------------------------------------------
import enum  # pylint: disable=g-import-not-at-top,wrong-import-position

from lingvo_b200.models.car import builder_lib  # pylint: disable=wrong-import-position
from lingvo_b200.models.car import car_layers  # pylint: disable=wrong-import-position
from lingvo_b200.models.car import point_detector  # pylint: disable=wrong-import-position


class LossNormType(enum.Enum):
  NO_NORM = 0
  NORM_BY_NUM_POSITIVES = 1


class PointsToGridFeaturizer(base_layer.BaseLayer):
  """Pillars from the `GridToPillars` preprocessor → BEV feature image (ref :33): augment
  every point with its offset to the pillar's point mean and to the pillar centre, run the
  per-point `featurizer`, max over the pillar's points, scatter the pillar vectors to
  `[B, gx, gy, C]`."""

  @classmethod
  def Params(cls, num_laser_features=1, num_output_features=64):
    p = super().Params()
    b = builder_lib.ModelBuilderBase(builder_lib.ModelBuilderBase.Params())
    idims = 3 + num_laser_features + 3 + 3
    p.Define('num_laser_features', num_laser_features, 'Laser features per point.')
    p.Define('featurizer', b._FC('pillar_fc', idims, num_output_features),   # pylint: disable=protected-access
             'Per-point featurizer.')
    p.Define('num_output_features', num_output_features, 'Pillar feature dim.')
    p.Define('grid_size', (432, 496, 1), '(gx, gy, gz).')
    return p

  def __init__(self, params):
    super().__init__(params)
    self.CreateChild('featurizer', self.params.featurizer)

  def FProp(self, theta, input_batch):
    p = self.params
    pts = input_batch.pillar_points                         # [B, N, K, 3+F]
    cnt = input_batch.point_count.to(pts.dtype)             # [B, N]
    loc = input_batch.point_locations.long()                # [B, N, 3]
    b, n, k, _ = pts.shape
    valid = (torch.arange(k, device=pts.device).view(1, 1, k) < cnt.unsqueeze(-1)).to(pts.dtype)
    mean = (pts[..., :3] * valid.unsqueeze(-1)).sum(2, keepdim=True) / cnt.clamp_min(1.0).view(
        b, n, 1, 1)
    centers = input_batch.get('pillar_centers')
    if centers is None:
      centers = mean.squeeze(2)
    feats = torch.cat([pts, pts[..., :3] - mean, pts[..., :3] - centers.unsqueeze(2)], -1)
    feats = self.featurizer.FProp(theta.featurizer, feats * valid.unsqueeze(-1))
    neg = torch.finfo(feats.dtype).min
    pooled = feats.masked_fill(valid.unsqueeze(-1) < 0.5, neg).max(2).values
    pooled = torch.where(cnt.unsqueeze(-1) > 0, pooled, torch.zeros_like(pooled))
    gx, gy, _ = p.grid_size
    flat = loc[..., 0] * gy + loc[..., 1]
    image = torch.zeros(b, gx * gy, pooled.shape[-1], device=pts.device, dtype=pooled.dtype)
    image.scatter_(1, flat.unsqueeze(-1).expand(-1, -1, pooled.shape[-1]),
                   pooled * (cnt.unsqueeze(-1) > 0))
    return image.reshape(b, gx, gy, -1)


class DynamicVoxelizationFeaturizer(base_layer.BaseLayer):
  """Raw padded points → BEV feature image through dynamic voxelisation (no per-pillar
  point budget) (ref :647)."""

  @classmethod
  def Params(cls, num_laser_features=1, num_output_features=64):
    p = super().Params()
    b = builder_lib.ModelBuilderBase(builder_lib.ModelBuilderBase.Params())
    dv = car_layers.DynamicVoxelization.Params().Set(num_laser_features=num_laser_features)
    enc = car_layers.PointEncoder.Params().Set(name='enc').Instantiate().NumEncodingFeatures(
        num_laser_features)
    dv.featurizer = b._FC('point_fc', enc, num_output_features)   # pylint: disable=protected-access
    p.Define('dynamic_voxelization', dv, 'Voxelisation + encoding + pooling.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self.CreateChild('dynamic_voxelization', self.params.dynamic_voxelization)

  def FProp(self, theta, input_batch):
    las = input_batch.lasers
    return self.dynamic_voxelization.FProp(theta.dynamic_voxelization, las.points_xyz,
                                           las.points_feature, las.points_padding)


def SparseToDense(grid_shape, locations, feats):
  """Scatters pillar features back onto the dense grid (ref :36): `locations [B, P, 3]`
  integer (x, y, z) cells, `feats [B, P, F]` → `[B, nx, ny, nz·F]`. Pillars sharing a cell
  are summed (as `tf.scatter_nd` does); padded pillars should carry zero features."""
  nx, ny, nz = grid_shape
  b, p, f = feats.shape
  assert tuple(locations.shape) == (b, p, 3)
  loc = locations.long()
  flat = ((torch.arange(b, device=feats.device).view(b, 1) * nx + loc[..., 0]) * ny +
          loc[..., 1]) * nz + loc[..., 2]
  grid = feats.new_zeros(b * nx * ny * nz, f)
  grid.index_add_(0, flat.reshape(-1), feats.reshape(-1, f))
  return grid.reshape(b, nx, ny, nz * f)


class Builder(builder_lib.ModelBuilderBase):
  """PointPillars layer recipes (ref :148)."""

  def Featurizer(self, name, idims, odims):
    return self._FC(name, idims, odims)

  def _Block(self, name, stride, repeats, idims, odims):
    layers_ = [self._Conv('c0', (3, 3, idims, odims), (stride, stride))]
    layers_ += [self._Conv('c%d' % (i + 1), (3, 3, odims, odims)) for i in range(repeats)]
    return self._Seq(name, *layers_)

  def _FetchBlock(self, name, stride, repeats, idims, odims, activation=None):
    """Strided 3×3 conv + `repeats` 3×3 convs whose output is a fetch point `<name>.final`
    (ref :169)."""
    act = activation or 'RELU'
    return self._Seq(
        name,
        self._Conv('c3x3', (3, 3, idims, odims), (stride, stride), activation_fn=act),
        self._Rep('rep', repeats, self._Conv('c3x3', (3, 3, odims, odims), activation_fn=act)),
        self._Fetch('final'))

  def _TopDown(self, name, strides=(2, 2, 2), channel_multiplier=1, activation=None,
               idims=None, repeats=(3, 5, 5)):
    """The three strided blocks of [PointPillars §2.2] (ref :180)."""
    if len(strides) != 3:
      raise ValueError('`strides` expected to be list/tuple of len 3.')
    m = channel_multiplier
    return self._Seq(
        name,
        self._FetchBlock('b0', strides[0], repeats[0], idims or m * 64, m * 64, activation),
        self._FetchBlock('b1', strides[1], repeats[1], m * 64, m * 128, activation),
        self._FetchBlock('b2', strides[2], repeats[2], m * 128, m * 256, activation))

  def _Upsample(self, name, stride, idims, odims, activation=None):
    """Transposed conv (kernel = stride) + BN + activation (ref :195)."""
    act = activation or 'RELU'
    return builder_lib._Conv2D.Params().Set(   # pylint: disable=protected-access
        name=name, filter_shape=(stride, stride, idims, odims), stride=(stride, stride),
        transpose=True, use_bn=True, activation=act)

  def Contract(self, down_strides=(2, 2, 2), channel_multiplier=1, activation=None,
               idims=None, repeats=(3, 5, 5)):
    """Contracting half (ref :208): runs the blocks once and returns
    (b2 output, b1 output, b0 output) — the finer maps are *fetched*, not recomputed."""
    return self._Branch(
        'branch',
        self._TopDown('topdown', strides=down_strides, channel_multiplier=channel_multiplier,
                      activation=activation, idims=idims, repeats=repeats),
        ['b1.final', 'b0.final'])

  def Expand(self, odims, channel_multiplier=1, activation=None):
    """Expanding half (ref :218): every scale is upsampled to the b0 resolution (×4, ×2, ×1)
    and the three maps are concatenated → `3·odims` channels."""
    m = channel_multiplier
    return self._Concat(
        'concat',
        self._Seq('b2', self._ArgIdx('idx', [0]), self._Upsample('ups', 4, m * 256, odims,
                                                                  activation)),
        self._Seq('b1', self._ArgIdx('idx', [1]), self._Upsample('ups', 2, m * 128, odims,
                                                                  activation)),
        self._Seq('b0', self._ArgIdx('idx', [2]), self._Upsample('ups', 1, m * 64, odims,
                                                                  activation)))

  def Backbone(self, idims, dims=(64, 128, 256), repeats=(3, 5, 5), up_dims=128,
               first_stride=2, channel_multiplier=None, activation=None):
    """Contract → Expand (ref :238): `idims` input channels → `3·up_dims` channels at
    1/`first_stride` of the input resolution. Every block runs exactly once."""
    if channel_multiplier is None:
      assert tuple(dims) == (dims[0], 2 * dims[0], 4 * dims[0]) and dims[0] % 64 == 0, dims
      channel_multiplier = dims[0] // 64
    return self._Seq(
        'backbone',
        self.Contract((first_stride, 2, 2), channel_multiplier, activation, idims=idims,
                      repeats=tuple(repeats)),
        self.Expand(up_dims, channel_multiplier, activation))

  def MLPFeaturizer(self, name, dims, use_bn=True, activation_fn='RELU'):
    """Per-point MLP over `.features` of a points NestedMap → the feature tensor (ref :261)."""
    return self._Seq(
        name,
        self._FeaturesMLP('feat', dims, use_bn=use_bn, activation_fn=activation_fn),
        self._GetValue('get_features', 'features'))

  def ScalePillarsFeaturizer(self, name, input_dims, output_dims):
    """The wider swish featurizer of the scaled-up pillars models (ref :268)."""
    del name
    return self.MLPFeaturizer('feat', [input_dims, 256, 256, 256, output_dims],
                              activation_fn='SWISH')

  def Detector(self, name, idims, odims, conv_init_method=None, bias_params_init=None):
    del conv_init_method
    del bias_params_init
    return self._ConvPlain(name, (3, 3, idims, odims))


class ModelV2(point_detector.PointDetectorBase):
  """PointPillars on the extractor-based input (`KITTIGrid`): `PointsToGridFeaturizer` →
  `Builder.Backbone` → conv heads; anchors / assignments come from the preprocessors, NMS
  and metrics from `PointDetectorBase` and the output decoder (ref :330 `ModelV1`)."""

  NUM_OUTPUT_CHANNELS = 128

  @classmethod
  def Params(cls, grid_size_z=1, num_anchors=2, num_classes=2, num_laser_features=1):
    p = super().Params(num_classes)
    b = Builder(Builder.Params())
    c = cls.NUM_OUTPUT_CHANNELS
    p.Define('grid_size_z', grid_size_z, 'Grid size along z.')
    p.Define('num_anchors', num_anchors, 'Anchors per cell.')
    p.Define('num_laser_features', num_laser_features, 'Laser features per point.')
    p.Define('input_featurizer', PointsToGridFeaturizer.Params(num_laser_features, 64),
             'Points → BEV image.')
    p.Define('backbone', b.Backbone(64 * grid_size_z, up_dims=c), 'BEV backbone.')
    p.Define('class_detector', b.Detector('class', 3 * c, num_anchors * num_classes),
             'Classification head.')
    p.Define('regression_detector', b.Detector('reg', 3 * c, num_anchors * 7), 'Box head.')
    p.Define('focal_loss_alpha', 0.25, 'Focal α.')
    p.Define('focal_loss_gamma', 2.0, 'Focal γ.')
    p.Define('localization_loss_weight', 2.0, 'Localisation weight.')
    p.Define('classification_loss_weight', 1.0, 'Classification weight.')
    p.Define('loss_norm_type', LossNormType.NORM_BY_NUM_POSITIVES, 'Normalisation.')
    p.Define('huber_loss_delta', 1.0 / 9.0, 'Huber δ.')
    p.name = 'pillars_v2'
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    for n in ('input_featurizer', 'backbone', 'class_detector', 'regression_detector'):
      self.CreateChild(n, p.Get(n))

  def ComputePredictions(self, theta, input_batch):
    p = self.params
    img = self.input_featurizer.FProp(theta.input_featurizer, input_batch)
    feat = self.backbone.FProp(theta.backbone, img)
    cls = self.class_detector.FProp(theta.class_detector, feat)
    reg = self.regression_detector.FProp(theta.regression_detector, feat)
    b = cls.shape[0]
    return NestedMap(classification_logits=cls.reshape(b, -1, p.num_anchors, p.num_classes),
                     residuals=reg.reshape(b, -1, p.num_anchors, 7))

  def ComputeLoss(self, theta, predictions, input_batch):
    p = self.params
    u = self._utils_3d
    res, logits = predictions.residuals, predictions.classification_logits
    b = res.shape[0]
    gt = input_batch.anchor_localization_residuals.reshape(res.shape).to(res.dtype)
    labels = input_batch.assigned_gt_labels.reshape(res.shape[:-1]).long().clamp(
        max=p.num_classes - 1)
    cls_w = input_batch.assigned_cls_mask.reshape(res.shape[:-1]).to(res.dtype)
    reg_w = input_batch.assigned_reg_mask.reshape(res.shape[:-1]).to(res.dtype)
    focal = u.SigmoidFocalLoss(logits.float(), F.one_hot(labels, p.num_classes).float(),
                               p.focal_loss_alpha, p.focal_loss_gamma)
    cls_loss = focal[..., 1:].sum(-1) * cls_w
    d_rot = torch.sin(res[..., 6:] - gt[..., 6:])
    reg = torch.cat([u.ScaledHuberLoss(gt[..., :6], res[..., :6], delta=p.huber_loss_delta),
                     u.ScaledHuberLoss(torch.zeros_like(d_rot), d_rot, delta=p.huber_loss_delta)],
                    -1).sum(-1) * reg_w
    norm = reg_w.sum().clamp_min(1.0) if p.loss_norm_type == LossNormType.NORM_BY_NUM_POSITIVES \
        else torch.tensor(float(b), device=res.device)
    cls_total, reg_total = cls_loss.sum() / norm, reg.sum() / norm
    loss = p.classification_loss_weight * cls_total + p.localization_loss_weight * reg_total
    bs = float(b)
    return NestedMap(loss=(loss, bs), **{'loss/classification': (cls_total, bs),
                                         'loss/localization': (reg_total, bs)}), NestedMap()
