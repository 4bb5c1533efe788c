"""Anchor-free pillars / CenterPoint-style detector (ref
`lingvo/tasks/car/pillars_anchor_free.py`).

Every BEV cell predicts, for the object whose box contains the cell centre:
class logits, optional centerness, centre offset (Δxyz), log-dimensions and heading as
`angle_bin_num` bins + per-bin residual. Targets come from the `PointAssignment`
preprocessor run on the grid anchor centres. Location / dimension losses are pluggable
(`HuberLoss`, or `LaplaceKL` which also predicts a scale = aleatoric uncertainty).
Decoding: plain NMS, CenterNet max-pool heat-map NMS, or none.
"""

from __future__ import annotations

import enum
import math

import torch
import torch.nn.functional as F

from lingvo_b200.core import base_layer
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.car import detection_decoder
from lingvo_b200.models.car import pillars
from lingvo_b200.models.car import point_detector


class NMSDecoderType(enum.Enum):
  NMS_DECODER = 0
  HEATMAP_NMS_DECODER = 1
  NO_NMS_DECODER = 2


class ClassLossFN(enum.Enum):
  SIGMOID_LOSS = 0
  FOCAL_SIGMOID_LOSS = 1


def HeatMapNMS(heat_map_scores, kernel_size):
  """Keeps local maxima of `[B, gx, gy, C]`, zeroing everything else (ref :42)."""
  kh, kw = (kernel_size[1], kernel_size[2]) if len(kernel_size) == 4 else tuple(kernel_size)
  hm = heat_map_scores.permute(0, 3, 1, 2)
  pooled = F.max_pool2d(hm, (kh, kw), stride=1, padding=(kh // 2, kw // 2))
  return (hm * (hm == pooled).to(hm.dtype)).permute(0, 2, 3, 1)


class _LossInterface(base_layer.BaseLayer):
  """A regression loss that owns the parameterisation of its predictions (ref :68)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('num_params_per_prediction', 1, 'Raw outputs per regressed scalar.')
    return p

  def MeanPrediction(self, theta, prediction_tensors):
    """[..., n · num_params] raw → [..., n] point estimates."""
    raise NotImplementedError()

  def FProp(self, theta, prediction_tensors, labels):
    raise NotImplementedError()


class HuberLoss(_LossInterface):
  """Smooth-L1 on the raw prediction (ref :98)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('delta', 1.0 / (3.0 ** 2), 'Transition point.')
    p.num_params_per_prediction = 1
    return p

  def MeanPrediction(self, theta, prediction_tensors):
    return prediction_tensors

  def FProp(self, theta, prediction_tensors, labels, transform_fn=None):
    pred = self.MeanPrediction(theta, prediction_tensors)
    if transform_fn is not None:
      pred, labels = transform_fn(pred), transform_fn(labels)
    d = (pred - labels).abs()
    delta = self.params.delta
    return torch.where(d < delta, 0.5 * d * d / delta, d - 0.5 * delta)


class LaplaceKL(_LossInterface):
  """KL(Laplace(label, label_scale) ‖ Laplace(μ, b)) with predicted (μ, log b) — a
  regression loss with learned per-output uncertainty (ref :130)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('label_scale', 0.1, 'Scale of the (narrow) label distribution.')
    p.Define('min_scale', 1e-3, 'Lower bound of the predicted scale.')
    p.num_params_per_prediction = 2
    return p

  def _SplitPredictionParams(self, prediction_tensors):
    n = prediction_tensors.shape[-1] // 2
    return prediction_tensors[..., :n], prediction_tensors[..., n:]

  def MeanPrediction(self, theta, prediction_tensors):
    return self._SplitPredictionParams(prediction_tensors)[0]

  def Scale(self, theta, prediction_tensors):
    return F.softplus(self._SplitPredictionParams(prediction_tensors)[1]) + self.params.min_scale

  def FProp(self, theta, prediction_tensors, labels, transform_fn=None):
    mu = self.MeanPrediction(theta, prediction_tensors)
    b2 = self.Scale(theta, prediction_tensors)
    if transform_fn is not None:
      mu, labels = transform_fn(mu), transform_fn(labels)
    b1 = self.params.label_scale
    d = (mu - labels).abs()
    return torch.log(b2 / b1) + (b1 * torch.exp(-d / b1) + d) / b2 - 1.0


class AnchorFreePillarsBase(point_detector.PointDetectorBase):
  """ref :182."""

  NUM_OUTPUT_CHANNELS = 128

  @classmethod
  def Params(cls, grid_size_z=1, num_classes=2, num_laser_features=1, angle_bin_num=12):
    p = super().Params(num_classes)
    b = pillars.Builder(pillars.Builder.Params())
    c = cls.NUM_OUTPUT_CHANNELS
    p.Define('grid_size_z', grid_size_z, 'Grid size along z.')
    p.Define('num_laser_features', num_laser_features, 'Laser features per point.')
    p.Define('angle_bin_num', angle_bin_num, 'Heading bins.')
    p.Define('input_featurizer', pillars.PointsToGridFeaturizer.Params(num_laser_features, 64),
             'Points → BEV image.')
    p.Define('backbone', b.Backbone(64 * grid_size_z, up_dims=c, first_stride=1),
             'BEV backbone (output at the input resolution).')
    p.Define('location_loss', HuberLoss.Params(), 'Loss of the centre offsets.')
    p.Define('dimensions_loss', HuberLoss.Params(), 'Loss of the log-dimensions.')
    p.Define('class_detector', None, 'Built in __init__ unless given.')
    p.Define('centerness_detector', None, 'Built in __init__ unless given.')
    p.Define('regression_detector', None, 'Built in __init__ unless given.')
    p.Define('classification_loss_fn', ClassLossFN.FOCAL_SIGMOID_LOSS, 'Classification loss.')
    p.Define('focal_loss_alpha', 0.25, 'Focal α.')
    p.Define('focal_loss_gamma', 2.0, 'Focal γ.')
    p.Define('location_loss_weight', 1.0, 'Weight of the centre-offset loss.')
    p.Define('dimension_loss_weight', 1.0, 'Weight of the size loss.')
    p.Define('centerness_loss_weight', 0.0, 'Weight of the centerness loss (0: no head).')
    p.Define('rot_cls_loss_weight', 1.0, 'Weight of the heading-bin classification loss.')
    p.Define('rot_reg_loss_weight', 1.0, 'Weight of the in-bin heading residual loss.')
    p.Define('corner_loss_weight', 0.0, 'Weight of the corner loss.')
    p.Define('classification_loss_weight', 1.0, 'Weight of the classification loss.')
    p.Define('loss_norm_type', pillars.LossNormType.NORM_BY_NUM_POSITIVES, 'Normalisation.')
    p.Define('nms_decoder_type', NMSDecoderType.NMS_DECODER, 'Decode strategy.')
    p.Define('heatmap_nms_kernel_size', [1, 3, 3, 1], 'Max-pool kernel of heat-map NMS.')
    p.Define('heatmap_nms_score_threshold', 0.1, 'Minimum peak score.')
    p.name = 'anchor_free_pillars'
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    b = pillars.Builder(pillars.Builder.Params())
    idims = 3 * self.NUM_OUTPUT_CHANNELS
    self.CreateChild('location_loss', p.location_loss)
    self.CreateChild('dimensions_loss', p.dimensions_loss)
    n_reg = (3 * p.location_loss.num_params_per_prediction +
             3 * p.dimensions_loss.num_params_per_prediction + 2 * p.angle_bin_num)
    self._num_residual_dims = n_reg
    self.CreateChild('input_featurizer', p.input_featurizer)
    self.CreateChild('backbone', p.backbone)
    self.CreateChild('class_detector', p.class_detector or b.Detector(
        'class', idims, p.grid_size_z * p.num_classes))
    self.CreateChild('regression_detector', p.regression_detector or b.Detector(
        'reg', idims, p.grid_size_z * n_reg))
    if p.centerness_loss_weight > 0:
      self.CreateChild('centerness_detector', p.centerness_detector or b.Detector(
          'centerness', idims, p.grid_size_z * p.num_classes))

  # -- forward ----------------------------------------------------------------------
  def ComputePredictions(self, theta, input_batch):
    p = self.params
    img = self.input_featurizer.FProp(theta.input_featurizer, input_batch)
    feat = self.backbone.FProp(theta.backbone, img)
    b, gx, gy, _ = feat.shape
    n = gx * gy * p.grid_size_z
    out = NestedMap(
        classification_logits=self.class_detector.FProp(theta.class_detector, feat).reshape(
            b, n, p.num_classes),
        residuals=self.regression_detector.FProp(theta.regression_detector, feat).reshape(
            b, n, self._num_residual_dims),
        points=input_batch.anchor_centers.reshape(b, n, 3), grid_shape=(gx, gy, p.grid_size_z))
    if p.centerness_loss_weight > 0:
      out.centerness_logits = self.centerness_detector.FProp(
          theta.centerness_detector, feat).reshape(b, n, p.num_classes)
    return out

  def _SplitResiduals(self, theta, residuals):
    """raw `[..., R]` → (loc_raw, dim_raw, bin_logits [..., A], bin_res [..., A])."""
    p = self.params
    nl = 3 * p.location_loss.num_params_per_prediction
    nd = 3 * p.dimensions_loss.num_params_per_prediction
    a = p.angle_bin_num
    return (residuals[..., :nl], residuals[..., nl:nl + nd],
            residuals[..., nl + nd:nl + nd + a], residuals[..., nl + nd + a:])

  def _AngleToBin(self, phi):
    """heading → (bin index, residual normalised to [−1, 1])."""
    a = self.params.angle_bin_num
    width = 2 * math.pi / a
    shifted = torch.remainder(phi + math.pi, 2 * math.pi)
    idx = torch.clamp((shifted / width).long(), max=a - 1)
    centre = (idx.to(phi.dtype) + 0.5) * width
    return idx, (shifted - centre) / (width / 2)

  def _BinToAngle(self, bin_logits, bin_res):
    a = self.params.angle_bin_num
    width = 2 * math.pi / a
    idx = bin_logits.argmax(-1)
    res = bin_res.gather(-1, idx.unsqueeze(-1)).squeeze(-1)
    return (idx.to(res.dtype) + 0.5) * width + res * (width / 2) - math.pi

  # -- targets & loss -----------------------------------------------------------------
  def GenerateTarget(self, predictions, input_batch):
    """Flattens the per-cell assignments next to the predictions (ref :548)."""
    p = self.params
    b, n, _ = predictions.residuals.shape
    reg_w = input_batch.assigned_reg_mask.reshape(b, n, -1).sum(-1)
    ret = NestedMap(
        points=predictions.points, residuals=predictions.residuals,
        classification_logits=predictions.classification_logits,
        class_weights=input_batch.assigned_cls_mask.reshape(b, n),
        assigned_gt_labels=input_batch.assigned_gt_labels.reshape(b, n),
        reg_weights=reg_w, assigned_gt_bboxes=input_batch.assigned_gt_bbox.reshape(b, n, 7),
        target_predictions=input_batch.target_predictions.reshape(b, n, 7))
    if p.centerness_loss_weight > 0:
      ret.assigned_gt_centerness = input_batch.assigned_gt_center_ness.reshape(b, n)
      ret.centerness_logits = predictions.centerness_logits
    return ret

  def _ComputeClassificationLoss(self, logits, labels, class_weights):
    p = self.params
    one_hot = F.one_hot(labels.long().clamp(max=p.num_classes - 1), p.num_classes).to(
        logits.dtype)
    if p.classification_loss_fn == ClassLossFN.FOCAL_SIGMOID_LOSS:
      loss = self._utils_3d.SigmoidFocalLoss(logits.float(), one_hot, p.focal_loss_alpha,
                                             p.focal_loss_gamma)
    else:
      loss = F.binary_cross_entropy_with_logits(logits.float(), one_hot, reduction='none')
    return loss[..., 1:].sum(-1) * class_weights

  def _ComputeRegressionLoss(self, theta, tgt):
    """→ dict of per-cell losses (already multiplied by reg_weights)."""
    p = self.params
    loc_raw, dim_raw, bin_logits, bin_res = self._SplitResiduals(theta, tgt.residuals)
    w = tgt.reg_weights
    target = tgt.target_predictions                        # Δxyz, log dims, φ
    loc = self.location_loss.FProp(theta.location_loss, loc_raw, target[..., :3]).sum(-1) * w
    dim = self.dimensions_loss.FProp(theta.dimensions_loss, dim_raw, target[..., 3:6]).sum(-1) * w
    idx, res = self._AngleToBin(target[..., 6])
    rot_cls = F.cross_entropy(bin_logits.reshape(-1, p.angle_bin_num).float(), idx.reshape(-1),
                              reduction='none').reshape(idx.shape) * w
    pred_res = bin_res.gather(-1, idx.unsqueeze(-1)).squeeze(-1)
    rot_reg = self._utils_3d.ScaledHuberLoss(res, pred_res, delta=1.0 / 9.0) * w
    return dict(location=loc, dimension=dim, rot_cls=rot_cls, rot_reg=rot_reg)

  def ComputeLoss(self, theta, predictions, input_batch):
    p = self.params
    tgt = self.GenerateTarget(predictions, input_batch)
    b = tgt.residuals.shape[0]
    cls = self._ComputeClassificationLoss(tgt.classification_logits, tgt.assigned_gt_labels,
                                          tgt.class_weights)
    reg = self._ComputeRegressionLoss(theta, tgt)
    norm = tgt.reg_weights.sum().clamp_min(1.0) \
        if p.loss_norm_type == pillars.LossNormType.NORM_BY_NUM_POSITIVES else \
        torch.tensor(float(b), device=cls.device)
    bs = float(b)
    terms = {'classification': (cls.sum() / norm, p.classification_loss_weight),
             'location': (reg['location'].sum() / norm, p.location_loss_weight),
             'dimension': (reg['dimension'].sum() / norm, p.dimension_loss_weight),
             'rot_cls': (reg['rot_cls'].sum() / norm, p.rot_cls_loss_weight),
             'rot_reg': (reg['rot_reg'].sum() / norm, p.rot_reg_loss_weight)}
    if p.centerness_loss_weight > 0:
      fg = (tgt.reg_weights > 0).to(cls.dtype)
      lab_cls = (tgt.assigned_gt_labels.long() - 0).clamp(0, p.num_classes - 1)
      logit = tgt.centerness_logits.gather(-1, lab_cls.unsqueeze(-1)).squeeze(-1)
      cn = F.binary_cross_entropy_with_logits(logit.float(), tgt.assigned_gt_centerness,
                                              reduction='none') * fg
      terms['centerness'] = (cn.sum() / norm, p.centerness_loss_weight)
    if p.corner_loss_weight > 0:
      boxes = self._BBoxesAndLogits(input_batch, predictions).predicted_bboxes
      corner = (self._utils_3d.CornerLoss(tgt.assigned_gt_bboxes, boxes) *
                tgt.reg_weights).sum() / norm
      terms['corner'] = (corner, p.corner_loss_weight)
    loss = sum(v * w for v, w in terms.values())
    metrics = NestedMap(loss=(loss, bs), num_positives=(tgt.reg_weights.sum() / bs, bs))
    for k, (v, _) in terms.items():
      metrics['loss/' + k] = (v, bs)
    return metrics, NestedMap(residuals=tgt.residuals)

  # -- decoding ---------------------------------------------------------------------
  def _BBoxesAndLogits(self, input_batch, predictions):
    theta = self.theta
    loc_raw, dim_raw, bin_logits, bin_res = self._SplitResiduals(theta, predictions.residuals)
    centre = predictions.points + self.location_loss.MeanPrediction(theta.location_loss, loc_raw)
    dims = torch.exp(self.dimensions_loss.MeanPrediction(theta.dimensions_loss, dim_raw).clamp(
        max=6.0))
    phi = self._BinToAngle(bin_logits, bin_res)
    boxes = torch.cat([centre, dims, phi.unsqueeze(-1)], -1)
    logits = predictions.classification_logits
    if 'centerness_logits' in predictions:
      # score = class prob × centerness: fold in as a logit adjustment
      logits = torch.logit((torch.sigmoid(logits.float()) *
                            torch.sigmoid(predictions.centerness_logits.float())).clamp(
                                1e-6, 1 - 1e-6))
    return NestedMap(predicted_bboxes=boxes, classification_logits=logits)

  def _NoPostProcessDecoder(self, predicted_bboxes, classification_scores, max_boxes):
    """Top-`max_boxes` per class by score, no suppression (ref :819)."""
    b, n, c = classification_scores.shape
    k = min(max_boxes, n)
    scores, idx = classification_scores.transpose(1, 2).topk(k, -1)          # [B, C, K]
    boxes = predicted_bboxes.gather(
        1, idx.reshape(b, c * k, 1).expand(-1, -1, 7)).reshape(b, c, k, 7)
    return idx, boxes, scores, torch.ones_like(scores)

  def _DecodeImpl(self, input_batch):
    p = self.params
    if p.nms_decoder_type == NMSDecoderType.NMS_DECODER:
      return super()._DecodeImpl(input_batch)
    predictions = self.ComputePredictions(self.theta, input_batch)
    bl = self._BBoxesAndLogits(input_batch, predictions)
    scores = torch.sigmoid(bl.classification_logits.float())
    b = scores.shape[0]
    if p.nms_decoder_type == NMSDecoderType.HEATMAP_NMS_DECODER:
      gx, gy, gz = predictions.grid_shape
      hm = scores.reshape(b, gx, gy, gz * p.num_classes)[..., :p.num_classes]
      _, boxes, sc, mask = detection_decoder.DecodeWithMaxPoolNMS(
          bl.predicted_bboxes, scores, hm, tuple(p.heatmap_nms_kernel_size[1:3]),
          p.max_nms_boxes, p.heatmap_nms_score_threshold)
    else:
      _, boxes, sc, mask = self._NoPostProcessDecoder(bl.predicted_bboxes, scores,
                                                      p.max_nms_boxes)
    sc = sc * mask
    viz = torch.where(sc >= p.visualization_classification_threshold, sc, torch.zeros_like(sc))
    return NestedMap(per_class_predicted_bboxes=boxes, per_class_predicted_bbox_scores=sc,
                     per_class_valid_mask=mask, visualization_weights=viz)


class ModelV1(AnchorFreePillarsBase):
  """The default configuration (kept as a separate registered name like the reference's
  params files expect)."""
