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
      rl = torch.cat([u.ScaledHuberLoss(tgt[:, :6], pred[:, :6]),
                      u.ScaledHuberLoss(torch.zeros_like(d_rot), d_rot)], -1).sum(-1)
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
