"""PointNet / PointNet++ builders (ref `lingvo/tasks/car/pointnet.py`).

All nets consume a points tensor `NestedMap(points [B,P,3], features [B,P,F],
padding [B,P])`; classifiers return `[B, output_dim]`, segmentation nets a points tensor
whose `features` are `[B, P, output_dim]`.
"""

from __future__ import annotations

import torch

from lingvo_b200.core import py_utils
from lingvo_b200.models.car import builder_lib
from lingvo_b200.models.car import car_layers


class PointNet(builder_lib.ModelBuilderBase):
  """ref :28."""

  def _ConcatWithBranch(self, name, tile, *subs):
    """[per-point features ‖ global feature from `subs` (tiled over points if `tile`)]."""
    def Merge(x, g):
      if tile:
        g = g.unsqueeze(-2).expand(x.shape[:-1] + g.shape[-1:])
      return torch.cat([x, g], -1)
    return self._Seq(name, self._Par('par', self._Identity('id'), self._Seq('branch', *subs)),
                     self._Fn('merge', Merge))

  def _ConcatWithOnehot(self, name):
    """Appends a one-hot object-category vector (`inp.cat [B]`, 16 ShapeNet classes) to the
    global feature."""
    def Fn(inp):
      oh = torch.nn.functional.one_hot(inp.cat.long(), 16).to(inp.features.dtype)
      return torch.cat([inp.features, oh], -1)
    return self._Fn(name, Fn)

  def _TNet(self, name, idims):
    """Input transform: predicts a `idims × idims` matrix (initialised to identity) and
    applies it to every point feature."""
    d = idims
    def ToMatrix(v):
      eye = torch.eye(d, device=v.device, dtype=v.dtype).reshape(1, d * d)
      return (v + eye).reshape(-1, d, d)
    predict = self._Seq(
        'predict', self._MLP('mlp0', [d, 64, 128, 1024]), self._Max('max'),
        self._MLP('mlp1', [1024, 512, 256]),
        self._Linear('linear', 256, d * d).Set(
            params_init=py_utils.WeightInit.Constant(0.0)),
        self._Fn('to_matrix', ToMatrix))
    return self._Seq(name, self._Par('par', self._Identity('id'), predict),
                     self._Fn('apply', torch.matmul))

  def Classifier(self, name='pointnet_classifier', input_dims=3, feature_dims=256,
                 keep_prob=0.7):
    """Per-point MLP → padded max → MLP (ref :91)."""
    p = self._Seq(
        name,
        self._SeqToKey('point_features', 'features', self._MakeInputFeatureFromPoints('in'),
                       self._MLP('mlp0', [input_dims, 64, 64, 64, 128, 1024])),
        self._PaddedMax('max'),
        self._MLP('mlp1', [1024, 512, feature_dims]), self._Dropout('dropout', keep_prob))
    p.Define('output_dim', feature_dims, 'Final output dimension.')
    return p

  def Segmentation(self, name='pointnet_segmentation', input_dims=3):
    """Per-point features concatenated with the global feature (ref :120)."""
    local = self._MLP('mlp0', [input_dims, 64, 64])
    glob = self._Seq('global', self._MLP('mlp1', [64, 64, 128, 1024]), self._Max('max'))
    feats = self._Seq('feats', self._MakeInputFeatureFromPoints('in'), local,
                      self._ConcatWithBranch('concat', True, glob),
                      self._MLP('mlp2', [1088, 512, 256, 128, 128]))
    p = self._SeqToKey(name, 'features', feats)
    p.Define('output_dim', 128, 'Final output dimension.')
    return p

  def SegmentationShapeNet(self, name='pointnet_shapenet', keep_prob=0.8, input_dims=3):
    """ShapeNet part-segmentation variant with input T-Net (ref :146)."""
    local = self._Seq('local', self._TNet('tnet', input_dims),
                      self._MLP('mlp0', [input_dims, 64, 128, 128]))
    glob = self._Seq('global', self._MLP('mlp1', [128, 512, 2048]), self._Max('max'))
    feats = self._Seq('feats', self._GetValue('pts', 'points'), local,
                      self._ConcatWithBranch('concat', True, glob),
                      self._MLP('mlp2', [2176, 256, 256, 128]),
                      self._Dropout('dropout', keep_prob))
    p = self._SeqToKey(name, 'features', feats)
    p.Define('output_dim', 128, 'Final output dimension.')
    return p


class PointNetPP(builder_lib.ModelBuilderBase):
  """ref :173."""

  def _SetAbstractionWithMLPMax(self, name, feature_extraction_sub, num_samples, ball_radius,
                                group_size):
    """Sample + group, featurise each group, max over the group → smaller points tensor."""
    sg = car_layers.SamplingAndGroupingLayer.Params().Set(
        name='sample_group', num_samples=num_samples, ball_radius=ball_radius,
        group_size=group_size)
    def Pool(grouped, query):
      from lingvo_b200.core.nested_map import NestedMap  # pylint: disable=g-import-not-at-top
      neg = torch.finfo(grouped.features.dtype).min
      f = grouped.features.masked_fill(grouped.padding.unsqueeze(-1) > 0.5, neg).max(-2).values
      f = torch.where(f == neg, torch.zeros_like(f), f)
      return NestedMap(points=query.points, features=f, padding=query.padding)
    featurise = self._Seq(
        'featurise', self._Par('par', self._Seq('grouped', self._Arg('g', 0),
                                                self._SeqOnFeatures('mlp', feature_extraction_sub)),
                               self._Arg('q', 1)))
    return self._Seq(name, sg, featurise, self._Fn('pool', Pool))

  def _ModelNet40Featurizer(self, input_dims):
    return self._Seq(
        'modelnet40',
        self._SetAbstractionWithMLPMax('sa0', self._MLP('mlp', [input_dims + 3, 64, 64, 128]),
                                       512, 0.2, 32),
        self._SetAbstractionWithMLPMax('sa1', self._MLP('mlp', [128 + 3, 128, 128, 256]),
                                       128, 0.4, 64),
        self._SeqToKey('feat', 'features', self._MakeInputFeatureFromPoints('in'),
                       self._MLP('mlp', [256 + 3, 256, 512, 1024])),
        self._PaddedMax('max'))

  def Classifier(self, name='pointnetpp_classifier', input_dims=3, feature_dims=256,
                 keep_prob=0.6, num_points=1024):
    del num_points
    p = self._Seq(name, self._ModelNet40Featurizer(input_dims),
                  self._FC('fc0', 1024, 512), self._Dropout('d0', keep_prob),
                  self._FC('fc1', 512, feature_dims), self._Dropout('d1', keep_prob))
    p.Define('output_dim', feature_dims, 'Final output dimension.')
    return p
