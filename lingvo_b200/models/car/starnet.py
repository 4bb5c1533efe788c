"""StarNet: sparse, targeted point-cloud detection (ref `lingvo/tasks/car/starnet.py`,
arXiv 1908.11069).

The input pipeline (`KITTISparseLaser` / `WaymoSparseLaser`) samples cell centres, gathers
each centre's neighbourhood and tiles anchors at the centres. The model featurises every
cell independently with a point network (padded MLP-max, or GIN blocks), then predicts per
anchor a 7-DOF residual and class logits. `ModelV2` adds a self-attention stage so that
cells exchange context before prediction.
"""

from __future__ import annotations

import enum
import math

import torch
import torch.nn.functional as F

from lingvo_b200.core import py_utils
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.car import builder_lib
from lingvo_b200.models.car import point_detector


class Builder(builder_lib.ModelBuilderBase):
  """Layer recipes of StarNet (ref :46)."""

  def _FeaturesFC(self, name, idims, odims, use_bn=True, activation_fn='RELU'):
    return self._SeqOnFeatures(name, self._FC('fc', idims, odims, use_bn, activation_fn))

  def _FeaturesMLP(self, name, dims, use_bn=True):
    return self._SeqOnFeatures(name, self._MLP('mlp', dims, use_bn))

  def _PaddedMean(self, name):
    def Fn(inp):
      w = (1.0 - inp.padding).unsqueeze(-1)
      return (inp.features * w).sum(-2) / w.sum(-2).clamp_min(1.0)
    return self._Fn(name, Fn)

  def MLPMaxFeaturizer(self, dims):
    return self._Seq('feat', self._GetValue('get_value', 'features'), self._MLP('mlp', dims),
                     self._Max('max'))

  def PaddedMLPMaxFeaturizer(self, idims, dims, use_bn=True):
    return self._Seq('feat', self._FeaturesFC('input_fc', idims, dims[0], use_bn=False),
                     self._FeaturesMLP('mlp', dims, use_bn=use_bn), self._PaddedMax('max'))

  def FC(self, name, idims, odims, use_bn=True, activation_fn='RELU'):
    return self._FC(name, idims, odims, use_bn, activation_fn)

  def Linear(self, name, idims, odims, params_init=None):
    p = self._Linear(name, idims, odims)
    if params_init is not None:
      p.params_init = params_init
    return p

  def Bias(self, name, dims, params_init=None):
    p = self._Bias(name, dims)
    if params_init is not None:
      p.params_init = params_init
    return p

  def LinearWithBias(self, name, idims, odims, linear_params_init=None, bias_params_init=None):
    return self._Seq(name, self.Linear('linear', idims, odims, linear_params_init),
                     self.Bias('bias', odims, bias_params_init))

  def Atten(self, name, depth, dims, hdims, heads, odims, keep_prob=1.0,
            linear_params_init=None, bias_params_init=None):
    return self._Seq(name, self._SelfAttenStack('attens', depth, dims, hdims, heads, keep_prob),
                     self.LinearWithBias('proj', dims, odims, linear_params_init,
                                         bias_params_init))

  def _GINBlock(self, name, idims, odims):
    """One graph-isomorphism block on a cell: every point gets MLP([own ‖ max over the
    cell]); returns the updated points tensor."""
    def Combine(inp, agg):
      out = inp.copy()
      out.features = torch.cat([inp.features, agg.unsqueeze(-2).expand_as(inp.features)], -1)
      return out
    return self._Seq(name, self._Par('par', self._Identity('id'), self._PaddedMax('agg')),
                     self._Fn('combine', Combine), self._FeaturesFC('fc', 2 * idims, odims))

  def _GIN(self, name, mlp_dims, combine_method='concat'):
    """Stacked GIN blocks; the cell descriptor concatenates (max ‖ mean) read-outs of every
    block."""
    assert combine_method == 'concat'
    n = len(mlp_dims) - 1
    blocks = [self._GINBlock('gin%d' % i, mlp_dims[i], mlp_dims[i + 1]) for i in range(n)]
    readout = lambda tag: self._Concat('readout_' + tag, self._PaddedMax('max'),
                                       self._PaddedMean('mean'))
    # run blocks sequentially, reading out after each
    def Chain(k):
      if k == n:
        return None
      return blocks[k]
    stages = []
    cur = []
    for k in range(n):
      cur = cur + [blocks[k]]
      stages.append(self._Seq('stage%d' % k, *[b.Copy() for b in cur], readout(str(k))))
    del Chain
    return self._Concat(name, *stages)

  def GINOutputDim(self, mlp_dims):
    return 2 * sum(mlp_dims[1:])

  def GINFeaturizer(self, name, fc_dims, mlp_dims, num_laser_features=1):
    """Drops the absolute cell-centre coordinates, then FC + GIN (ref :106)."""
    idims = 3 + num_laser_features
    return self._Seq(
        name, self._SeqOnFeatures('drop_cell_center_xyz', self._Fn('drop', lambda t: t[..., 3:])),
        self._FeaturesFC('fc0', idims, fc_dims), self._GIN('gin', [fc_dims] + list(mlp_dims)))

  def GINFeaturizerV2(self, name, fc_dims, mlp_dims, num_laser_features=1, fc_use_bn=True):
    idims = 3 + num_laser_features
    return self._Seq(name, self._FeaturesFC('fc0', idims, fc_dims, use_bn=fc_use_bn),
                     self._GIN('gin', [fc_dims] + list(mlp_dims)))

  def ZerosCellFeaturizer(self, name, dims):
    return self._Fn(name, lambda inp: torch.zeros(inp.points.shape[:-2] + (dims,),
                                                  device=inp.points.device))


class LossNormType(enum.Enum):
  NO_NORM = 0
  NORM_BY_NUM_POSITIVES = 1


class ModelBase(point_detector.PointDetectorBase):
  """Losses and box decoding shared by the StarNet variants (ref :161)."""

  @classmethod
  def Params(cls, num_classes=2, num_anchor_bboxes_offsets=25, num_anchor_bboxes_rotations=4,
             num_anchor_bboxes_dimensions=1):
    p = super().Params(num_classes)
    p.Define('num_anchor_bboxes_per_center',
             num_anchor_bboxes_offsets * num_anchor_bboxes_rotations *
             num_anchor_bboxes_dimensions, 'Anchors tiled at every centre.')
    p.Define('focal_loss_alpha', 0.25, 'Focal-loss α.')
    p.Define('focal_loss_gamma', 2.0, 'Focal-loss γ.')
    p.Define('huber_loss_delta', 1.0 / (3.0 ** 2), 'Huber δ of the localisation loss.')
    p.Define('loss_weight_localization', 2.0, 'Weight of the localisation loss.')
    p.Define('loss_weight_classification', 1.0, 'Weight of the classification loss.')
    p.Define('loss_norm_type', LossNormType.NORM_BY_NUM_POSITIVES, 'Loss normalisation.')
    p.Define('squash_rotation_predictions', False, 'φ residual = π·tanh(raw).')
    p.Define('corner_loss_weight', 0.0, 'Weight of the 8-corner loss.')
    p.Define('per_class_loss_weight', None, 'Per-class weights of the classification loss.')
    p.Define('location_loss_weight', 1.0, 'Weight of the x/y/z residual terms.')
    p.Define('dimension_loss_weight', 1.0, 'Weight of the size residual terms.')
    p.Define('rotation_loss_weight', 1.0, 'Weight of the heading residual term.')
    p.Define('direction_classifier_weight', 0.0, 'Kept for parity.')
    p.Define('direction_aware_rot_loss', False, 'Kept for parity.')
    return p

  def ComputeLoss(self, theta, predictions, input_batch):
    p = self.params
    u = self._utils_3d
    res, logits = predictions.residuals, predictions.classification_logits
    b = res.shape[0]
    gt_res = input_batch.anchor_localization_residuals.to(res.dtype)
    labels = input_batch.assigned_gt_labels.long()
    cls_w = input_batch.assigned_cls_mask.to(res.dtype)
    reg_w = input_batch.assigned_reg_mask.to(res.dtype)
    if 'cell_center_padding' in input_batch:
      live = (1.0 - input_batch.cell_center_padding).to(res.dtype).unsqueeze(-1)
      cls_w, reg_w = cls_w * live, reg_w * live
    one_hot = F.one_hot(labels, p.num_classes).to(res.dtype)
    focal = u.SigmoidFocalLoss(logits.float(), one_hot, p.focal_loss_alpha, p.focal_loss_gamma)
    if p.per_class_loss_weight is not None:
      focal = focal * torch.tensor(p.per_class_loss_weight, device=focal.device)
    cls_loss = (focal[..., 1:].sum(-1) * cls_w)
    # heading: sin(Δφ) so that a box flipped by π costs nothing
    d_rot = torch.sin(res[..., 6:] - gt_res[..., 6:])
    loc = u.ScaledHuberLoss(gt_res[..., :3], res[..., :3], delta=p.huber_loss_delta).sum(-1)
    dim = u.ScaledHuberLoss(gt_res[..., 3:6], res[..., 3:6], delta=p.huber_loss_delta).sum(-1)
    rot = u.ScaledHuberLoss(torch.zeros_like(d_rot), d_rot, delta=p.huber_loss_delta).sum(-1)
    reg_loss = (p.location_loss_weight * loc + p.dimension_loss_weight * dim +
                p.rotation_loss_weight * rot) * reg_w
    if p.loss_norm_type == LossNormType.NORM_BY_NUM_POSITIVES:
      norm = reg_w.sum().clamp_min(1.0)
    else:
      norm = torch.tensor(float(b), device=res.device)
    cls_total, reg_total = cls_loss.sum() / norm, reg_loss.sum() / norm
    loss = p.loss_weight_classification * cls_total + p.loss_weight_localization * reg_total
    bs = float(b)
    metrics = NestedMap(
        loss=(loss, bs), **{'loss/localization': (reg_total, bs),
                            'loss/classification': (cls_total, bs)},
        num_positives=(reg_w.sum() / bs, bs))
    if p.corner_loss_weight > 0:
      pred_boxes = u.ResidualsToBBoxes(input_batch.anchor_bboxes, res)
      corner = (u.CornerLoss(input_batch.assigned_gt_bbox, pred_boxes) * reg_w).sum() / norm
      loss = loss + p.corner_loss_weight * corner
      metrics['loss/corner'] = (corner, bs)
      metrics.loss = (loss, bs)
    with torch.no_grad():
      pred_boxes = u.ResidualsToBBoxes(input_batch.anchor_bboxes, res)
      metrics.update(self._BBoxDimensionErrors(input_batch.assigned_gt_bbox, pred_boxes, reg_w))
    per_example = NestedMap(residuals=res, classification_logits=logits)
    return metrics, per_example

  def _Squash(self, residuals):
    if self.params.squash_rotation_predictions:
      residuals = torch.cat([residuals[..., :6], math.pi * torch.tanh(residuals[..., 6:])], -1)
    return residuals

  def _PointInput(self, input_batch):
    """Cell neighbourhoods as a points tensor: features = [centre ‖ centred xyz ‖ laser]."""
    centre = input_batch.cell_center_xyz.unsqueeze(2)
    centred = input_batch.cell_points_xyz - centre
    feat = torch.cat([centre.expand_as(centred), centred, input_batch.cell_feature], -1)
    return NestedMap(points=centred, features=feat, padding=input_batch.cell_points_padding)


def _FocalBiasInit(num_classes, prior=0.01):
  """Bias initialisation that starts every class at `prior` probability."""
  del num_classes
  return py_utils.WeightInit.Constant(-math.log((1.0 - prior) / prior))


class ModelV1(ModelBase):
  """Independent cells (ref :516)."""

  @classmethod
  def Params(cls, num_classes=2, num_anchor_bboxes_offsets=25, num_anchor_bboxes_rotations=4,
             num_anchor_bboxes_dimensions=1, num_laser_features=1):
    p = super().Params(num_classes, num_anchor_bboxes_offsets, num_anchor_bboxes_rotations,
                       num_anchor_bboxes_dimensions)
    b = Builder(Builder.Params())
    dims = [64, 128, 256, 512]
    p.Define('cell_featurizer',
             b.PaddedMLPMaxFeaturizer(3 + 3 + num_laser_features, dims),
             'Points tensor → [B, C, cell_feature_dims].')
    p.Define('cell_feature_dims', dims[-1], 'Output dim of the featurizer.')
    p.Define('anchor_projected_feature_dims', 512, 'Kept for parity.')
    p.name = 'starnet_v1'
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    b = Builder(Builder.Params())
    a = p.num_anchor_bboxes_per_center
    self.CreateChild('cell_featurizer', p.cell_featurizer)
    self.CreateChild('localization_regressor', b.LinearWithBias(
        'localization_regressor', p.cell_feature_dims, a * 7))
    self.CreateChild('classifier', b.LinearWithBias(
        'classifier', p.cell_feature_dims, a * p.num_classes,
        bias_params_init=_FocalBiasInit(p.num_classes)))

  def ComputePredictions(self, theta, input_batch):
    p = self.params
    feat = self.cell_featurizer.FProp(theta.cell_featurizer, self._PointInput(input_batch))
    b, c = feat.shape[:2]
    a = p.num_anchor_bboxes_per_center
    res = self.localization_regressor.FProp(theta.localization_regressor, feat).reshape(
        b, c, a, 7)
    logits = self.classifier.FProp(theta.classifier, feat).reshape(b, c, a, p.num_classes)
    return NestedMap(residuals=self._Squash(res), classification_logits=logits)


class ModelV2(ModelBase):
  """Cells → (optional) self-attention across cells → per-anchor heads that also see an
  embedding of the anchor geometry (ref :652)."""

  @classmethod
  def Params(cls, num_classes=2, num_anchor_bboxes_offsets=25, num_anchor_bboxes_rotations=4,
             num_anchor_bboxes_dimensions=1, num_laser_features=1):
    p = super().Params(num_classes, num_anchor_bboxes_offsets, num_anchor_bboxes_rotations,
                       num_anchor_bboxes_dimensions)
    b = Builder(Builder.Params())
    gin = [128, 128, 256]
    p.Define('cell_featurizer', b.GINFeaturizerV2('feat', 64, gin, 3 + num_laser_features),
             'Points tensor → [B, C, cell_feature_dims].')
    p.Define('cell_feature_dims', b.GINOutputDim([64] + gin), 'Featurizer output dim.')
    p.Define('anchor_projected_feature_dims', 128, 'Dim of the anchor-geometry embedding.')
    p.Define('num_attention_layers', 0, 'Self-attention layers across cells (0: none).')
    p.Define('attention_heads', 4, 'Attention heads.')
    p.Define('head_hidden_dims', 256, 'Hidden dim of the prediction heads.')
    p.Define('oracle_location', False, 'Debug: use ground-truth location residuals.')
    p.Define('oracle_dimension', False, 'Debug: use ground-truth size residuals.')
    p.Define('oracle_rotation', False, 'Debug: use ground-truth heading residuals.')
    p.name = 'starnet_v2'
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    b = Builder(Builder.Params())
    d = p.cell_feature_dims
    self.CreateChild('cell_featurizer', p.cell_featurizer)
    if p.num_attention_layers:
      self.CreateChild('cell_attention', b._SelfAttenStack(   # pylint: disable=protected-access
          'cell_attention', p.num_attention_layers, d, 2 * d, p.attention_heads, 1.0))
    self.CreateChild('anchor_projection', b.FC('anchor_projection', 7,
                                               p.anchor_projected_feature_dims, use_bn=False))
    h = d + p.anchor_projected_feature_dims
    self.CreateChild('localization_regressor', b._Seq(   # pylint: disable=protected-access
        'localization_regressor', b.FC('fc', h, p.head_hidden_dims, use_bn=False),
        b.LinearWithBias('out', p.head_hidden_dims, 7)))
    self.CreateChild('classifier', b._Seq(   # pylint: disable=protected-access
        'classifier', b.FC('fc', h, p.head_hidden_dims, use_bn=False),
        b.LinearWithBias('out', p.head_hidden_dims, p.num_classes,
                         bias_params_init=_FocalBiasInit(p.num_classes))))

  def _CellFeaturizer(self, theta, input_batch):
    feat = self.cell_featurizer.FProp(theta.cell_featurizer, self._PointInput(input_batch))
    if self.params.num_attention_layers:
      feat = self.cell_attention.FProp(theta.cell_attention, feat)
    return feat

  def ComputePredictions(self, theta, input_batch):
    p = self.params
    feat = self._CellFeaturizer(theta, input_batch)                         # [B, C, D]
    anchors = input_batch.anchor_bboxes                                      # [B, C, A, 7]
    rel = torch.cat([anchors[..., :3] - input_batch.cell_center_xyz.unsqueeze(2),
                     anchors[..., 3:]], -1)
    emb = self.anchor_projection.FProp(theta.anchor_projection, rel)
    joint = torch.cat([feat.unsqueeze(2).expand(-1, -1, anchors.shape[2], -1), emb], -1)
    res = self._Squash(self.localization_regressor.FProp(theta.localization_regressor, joint))
    logits = self.classifier.FProp(theta.classifier, joint)
    gt = input_batch.get('anchor_localization_residuals')
    if gt is not None:
      parts = [gt[..., :3] if p.oracle_location else res[..., :3],
               gt[..., 3:6] if p.oracle_dimension else res[..., 3:6],
               gt[..., 6:] if p.oracle_rotation else res[..., 6:]]
      res = torch.cat(parts, -1)
    return NestedMap(residuals=res, classification_logits=logits)
