"""DeepFusion: lidar-camera feature fusion for 3-D detection (ref
`lingvo/tasks/car/deep_fusion.py`, arXiv 2203.08195).

Pillar features query the camera feature maps at the pixels their points project to
(`SinglePointAligner` / `MultiPointsAligner`) and fuse them either by concatenation or
with *LearnableAlign* cross attention (`DeepFusionAligner`): the lidar feature is the
query, the camera features of the pillar's points are keys / values.
`MultiModalFeaturizer` wraps this into a drop-in replacement for the pillars input
featurizer.
"""

from __future__ import annotations

import torch
import torch.nn.functional as F

from lingvo_b200.core import base_layer
from lingvo_b200.models.car import pillars


class ImageFeatureExtractorBuilder(pillars.Builder):
  """Small conv tower producing 1/4-resolution camera features (ref :31)."""

  def ImageFeatureExtractor(self, name, out_channels=192):
    return self._Seq(name, self._Conv('c0', (3, 3, 3, 32), (2, 2)),
                     self._Conv('c1', (3, 3, 32, 64), (1, 1)),
                     self._Conv('c2', (3, 3, 64, 128), (2, 2)),
                     self._Conv('c3', (3, 3, 128, out_channels), (1, 1)))


class LearnableAlignBuilder(pillars.Builder):
  """Projections of the LearnableAlign attention block (ref :40)."""

  def __init__(self, lidar_channels=64, image_channels=192, qkv_channels=128):
    super().__init__(pillars.Builder.Params())
    self.lidar_channels, self.image_channels, self.qkv_channels = (
        lidar_channels, image_channels, qkv_channels)

  def Fusion(self, name):
    """[lidar ‖ attended image] → lidar_channels."""
    return self._FC(name, self.lidar_channels + self.qkv_channels, self.lidar_channels)

  def LidarEmbedding(self, name):
    return self._Linear(name, self.lidar_channels, self.qkv_channels)

  def ImageEmbedding(self, name):
    return self._Linear(name, self.image_channels, self.qkv_channels)

  def Dropout(self, name, keep_prob=0.7):
    return self._Dropout(name, keep_prob)

  def FC(self, name):
    return self._FC(name, self.qkv_channels, self.qkv_channels, use_bn=False)


class SinglePointAligner(base_layer.BaseLayer):
  """Bilinear sample of the camera features at ONE projected point per pillar (its centre)
  (ref :175). `image_features [B,H,W,C]`; `points_projected [B,N,(K,)2]` in full-resolution
  pixels; `feat_ratio` = (feature h / image h, feature w / image w)."""

  def _Sample(self, image_features, feat_ratio, uv):
    """uv `[B, M, 2]` (u = x, v = y pixels) → `[B, M, C]` (zeros outside the image)."""
    b, h, w, c = image_features.shape
    x = uv[..., 0] * feat_ratio[1]
    y = uv[..., 1] * feat_ratio[0]
    gx = 2.0 * (x + 0.5) / w - 1.0
    gy = 2.0 * (y + 0.5) / h - 1.0
    grid = torch.stack([gx, gy], -1).unsqueeze(2)                       # [B, M, 1, 2]
    out = F.grid_sample(image_features.permute(0, 3, 1, 2), grid.to(image_features.dtype),
                        mode='bilinear', padding_mode='zeros', align_corners=False)
    return out.squeeze(-1).transpose(1, 2)

  def FProp(self, theta, image_features, feat_ratio, points_projected):
    uv = points_projected if points_projected.dim() == 3 else points_projected[:, :, 0]
    return self._Sample(image_features, feat_ratio, uv)


class MultiPointsAligner(SinglePointAligner):
  """Samples the camera features at ALL K points of a pillar and averages the real ones
  (ref :233)."""

  def FProp(self, theta, image_features, feat_ratio, points_projected, points_mask=None):
    b, n, k, _ = points_projected.shape
    feats = self._Sample(image_features, feat_ratio, points_projected.reshape(b, n * k, 2))
    feats = feats.reshape(b, n, k, -1)
    if points_mask is None:
      return feats.mean(2)
    m = points_mask.unsqueeze(-1).to(feats.dtype)
    return (feats * m).sum(2) / m.sum(2).clamp_min(1.0)


class DeepFusionAligner(SinglePointAligner):
  """LearnableAlign (ref :287): softmax(q_lidar · k_imgᵀ) v_img over the K points of each
  pillar, then FC([lidar ‖ attended])."""

  @classmethod
  def Params(cls, lidar_channels=64, image_channels=192, qkv_channels=128):
    p = super().Params()
    b = LearnableAlignBuilder(lidar_channels, image_channels, qkv_channels)
    p.Define('q_embedding', b.LidarEmbedding('q_embedding'), 'Query projection.')
    p.Define('k_embedding', b.ImageEmbedding('k_embedding'), 'Key projection.')
    p.Define('v_embedding', b.ImageEmbedding('v_embedding'), 'Value projection.')
    p.Define('attn_dropout', b.Dropout('attn_dropout'), 'Attention dropout.')
    p.Define('fc', b.FC('fc'), 'Post-attention FC.')
    p.Define('fusion', b.Fusion('fusion'), 'Fusion FC.')
    p.Define('qkv_channels', qkv_channels, 'Attention dim.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    for n in ('q_embedding', 'k_embedding', 'v_embedding', 'attn_dropout', 'fc', 'fusion'):
      self.CreateChild(n, p.Get(n))

  def FProp(self, theta, image_features, feat_ratio, points_projected, lidar_features,
            points_mask=None):
    p = self.params
    b, n, k, _ = points_projected.shape
    img = self._Sample(image_features, feat_ratio, points_projected.reshape(b, n * k, 2))
    img = img.reshape(b, n, k, -1)
    q = self.q_embedding.FProp(theta.q_embedding, lidar_features)                 # [B,N,D]
    kk = self.k_embedding.FProp(theta.k_embedding, img)                           # [B,N,K,D]
    vv = self.v_embedding.FProp(theta.v_embedding, img)
    logits = (q.unsqueeze(2) * kk).sum(-1) / (p.qkv_channels ** 0.5)              # [B,N,K]
    if points_mask is not None:
      logits = logits.masked_fill(points_mask < 0.5, -1e9)
    att = self.attn_dropout.FProp(theta.attn_dropout, torch.softmax(logits.float(), -1).to(
        vv.dtype))
    ctx = self.fc.FProp(theta.fc, (att.unsqueeze(-1) * vv).sum(2))
    if points_mask is not None:
      ctx = ctx * (points_mask.sum(-1, keepdim=True) > 0).to(ctx.dtype)
    return self.fusion.FProp(theta.fusion, torch.cat([lidar_features, ctx], -1))


class MultiModalFeaturizer(base_layer.BaseLayer):
  """Pillars featurizer with camera fusion (ref :76). Expects, besides the pillar inputs,
  `input_batch.images.image [B, H, W, 3]` and `input_batch.pillar_points_projected
  [B, N, K, 2]` (pixels of every pillar point, from the calibration)."""

  @classmethod
  def Params(cls, num_laser_features=1, num_output_features=64, image_channels=192,
             aligner='deep_fusion'):
    p = super().Params()
    p.Define('lidar_featurizer', pillars.PointsToGridFeaturizer.Params(
        num_laser_features, num_output_features), 'Lidar pillar featurizer.')
    p.Define('image_feature_extractor', ImageFeatureExtractorBuilder(
        pillars.Builder.Params()).ImageFeatureExtractor('image_fe', image_channels),
             'Camera tower.')
    if aligner == 'deep_fusion':
      ap = DeepFusionAligner.Params(num_output_features, image_channels)
    elif aligner == 'multi':
      ap = MultiPointsAligner.Params()
    else:
      ap = SinglePointAligner.Params()
    p.Define('aligner', ap, 'Lidar ↔ camera alignment / fusion.')
    p.Define('num_output_features', num_output_features, 'Pillar feature dim.')
    p.Define('image_channels', image_channels, 'Camera feature dim.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    self.CreateChild('lidar_featurizer', p.lidar_featurizer)
    self.CreateChild('image_feature_extractor', p.image_feature_extractor)
    self.CreateChild('aligner', p.aligner)
    if not isinstance(self.aligner, DeepFusionAligner):
      from lingvo_b200.models.car import builder_lib  # pylint: disable=g-import-not-at-top
      b = builder_lib.ModelBuilderBase(builder_lib.ModelBuilderBase.Params())
      self.CreateChild('concat_fc', b._FC(   # pylint: disable=protected-access
          'concat_fc', p.num_output_features + p.image_channels, p.num_output_features))

  def FProp(self, theta, input_batch):
    p = self.params
    lf = self.lidar_featurizer
    image = input_batch.images.image
    img_feat = self.image_feature_extractor.FProp(theta.image_feature_extractor, image)
    ratio = (img_feat.shape[1] / image.shape[1], img_feat.shape[2] / image.shape[2])
    # pillar vectors before scattering: reuse the lidar featurizer on a 1-cell "grid" trick
    bev = lf.FProp(theta.lidar_featurizer, input_batch)                  # [B, gx, gy, C]
    b, gx, gy, c = bev.shape
    loc = input_batch.point_locations.long()
    flat = loc[..., 0] * gy + loc[..., 1]
    pillar_vec = bev.reshape(b, gx * gy, c).gather(1, flat.unsqueeze(-1).expand(-1, -1, c))
    k = input_batch.pillar_points.shape[2]
    mask = (torch.arange(k, device=bev.device).view(1, 1, k) <
            input_batch.point_count.unsqueeze(-1)).to(bev.dtype)
    proj = input_batch.pillar_points_projected
    if isinstance(self.aligner, DeepFusionAligner):
      fused = self.aligner.FProp(theta.aligner, img_feat, ratio, proj, pillar_vec, mask)
    else:
      if isinstance(self.aligner, MultiPointsAligner):
        img_vec = self.aligner.FProp(theta.aligner, img_feat, ratio, proj, mask)
      else:
        img_vec = self.aligner.FProp(theta.aligner, img_feat, ratio, proj)
      fused = self.concat_fc.FProp(theta.concat_fc, torch.cat([pillar_vec, img_vec], -1))
    live = (input_batch.point_count > 0).unsqueeze(-1).to(fused.dtype)
    out = torch.zeros(b, gx * gy, c, device=bev.device, dtype=fused.dtype)
    out.scatter_(1, flat.unsqueeze(-1).expand(-1, -1, c), fused * live)
    return out.reshape(b, gx, gy, c)
