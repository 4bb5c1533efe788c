"""Layers shared by car models (ref `lingvo/tasks/car/car_layers.py`)."""

from __future__ import annotations

import torch

from lingvo_b200.core import base_layer
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.car import car_lib


class SamplingAndGroupingLayer(base_layer.BaseLayer):
  """PointNet++ set abstraction front end (ref :26): farthest-point sample `num_samples`
  centres, group `group_size` neighbours within `ball_radius`, express the neighbours
  relative to their centre. Input / output are points tensors; the output is
  `(grouped NestedMap(points [B,S,K,3], features [B,S,K,F'], padding [B,S,K]),
  query NestedMap(points [B,S,3], padding [B,S]))`."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('num_samples', 0, 'Centres sampled (S).')
    p.Define('ball_radius', 0, 'Neighbourhood radius.')
    p.Define('group_size', 0, 'Neighbours per centre (K).')
    p.Define('sample_neighbors_uniformly', True, 'Random neighbours within the ball.')
    return p

  def FProp(self, theta, input_data):
    p = self.params
    pts, feats, pad = input_data.points, input_data.features, input_data.padding
    b = pts.shape[0]
    idx, _ = car_lib.FarthestPointSampler(pts, pad, p.num_samples)
    centres = pts.gather(1, idx.unsqueeze(-1).expand(-1, -1, 3))
    centre_pad = pad.gather(1, idx)
    nidx, npad = car_lib.NeighborhoodIndices(
        pts, centres, p.group_size, pad > 0.5, p.ball_radius,
        sample_neighbors_uniformly=p.sample_neighbors_uniformly)
    flat = nidx.reshape(b, -1)
    g_pts = pts.gather(1, flat.unsqueeze(-1).expand(-1, -1, 3)).reshape(
        b, p.num_samples, p.group_size, 3)
    g_feat = feats.gather(1, flat.unsqueeze(-1).expand(-1, -1, feats.shape[-1])).reshape(
        b, p.num_samples, p.group_size, -1)
    g_pts = g_pts - centres.unsqueeze(2)                      # centre-relative coordinates
    grouped = NestedMap(points=g_pts, features=torch.cat([g_pts, g_feat], -1),
                        padding=torch.maximum(npad, centre_pad.unsqueeze(-1)))
    return grouped, NestedMap(points=centres, padding=centre_pad)


class PointEncoder(base_layer.BaseLayer):
  """Per-point input features for dynamic voxelisation (ref :114): any of xyz, xyz
  relative to the voxel centroid / centre, laser features, centroid, voxel centre,
  covariance."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('include_xyz', False, 'Raw coordinates.')
    p.Define('include_xyz_norm_by_centroid', True, 'xyz − voxel centroid.')
    p.Define('include_xyz_norm_by_center', True, 'xyz − voxel centre.')
    p.Define('include_features', True, 'Laser features.')
    p.Define('include_centroid', True, 'Voxel centroid.')
    p.Define('include_centers', True, 'Voxel centre.')
    p.Define('include_covariance', False, 'Voxel covariance (9).')
    return p

  def NumEncodingFeatures(self, num_laser_features):
    p = self.params
    return (3 * (p.include_xyz + p.include_xyz_norm_by_centroid + p.include_xyz_norm_by_center +
                 p.include_centroid + p.include_centers) +
            num_laser_features * p.include_features + 9 * p.include_covariance)

  def FProp(self, unused_theta, points_xyz, dynamic_voxels, dynamic_voxel_statistics,
            points_feature):
    p = self.params
    st = dynamic_voxel_statistics
    parts = []
    if p.include_xyz:
      parts.append(points_xyz)
    if p.include_xyz_norm_by_centroid:
      parts.append(st.centered_xyz)
    if p.include_xyz_norm_by_center:
      parts.append(points_xyz - dynamic_voxels.centers)
    if p.include_features:
      parts.append(points_feature)
    if p.include_centroid:
      parts.append(st.centroids)
    if p.include_centers:
      parts.append(dynamic_voxels.centers)
    if p.include_covariance:
      parts.append(st.covariance)
    return torch.cat(parts, -1) * (1.0 - dynamic_voxels.padding).unsqueeze(-1)


class DynamicVoxelization(base_layer.BaseLayer):
  """Points → dense pillar/voxel feature grid without a fixed points-per-voxel budget
  (ref :189): encode every point (`PointEncoder`), featurise it with `featurizer`
  (per-point MLP), max-pool per voxel with a scatter-reduce, and lay the voxels out as
  `[B, gx, gy, gz · C]`."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('grid_size', (40, 40, 1), '(gx, gy, gz).')
    p.Define('grid_range_x', (0, 40), 'x range.')
    p.Define('grid_range_y', (-40, 40), 'y range.')
    p.Define('grid_range_z', (-3, 3), 'z range.')
    p.Define('num_laser_features', 1, 'Laser features per point.')
    p.Define('point_encoder', PointEncoder.Params(), 'Point encoding.')
    p.Define('featurizer', None, 'Per-point layer `[B,P,F_in] → [B,P,C]` (None: identity).')
    p.Define('pooling_method', 'max', 'max | mean.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    self.CreateChild('point_encoder', p.point_encoder)
    if p.featurizer is not None:
      self.CreateChild('featurizer', p.featurizer)

  def _VoxelizeAndEncodePoints(self, theta, points_xyz, points_feature, points_padding):
    p = self.params
    dv = car_lib.DynamicVoxelization(points_xyz, points_padding, p.grid_size, p.grid_range_x,
                                     p.grid_range_y, p.grid_range_z)
    stats = car_lib.DynamicVoxelStatistics(points_xyz, dv)
    enc = self.point_encoder.FProp(theta.point_encoder, points_xyz, dv, stats, points_feature)
    return dv, enc

  def _ComputeVoxelFeatures(self, dynamic_voxels, featurized_points):
    p = self.params
    dv = dynamic_voxels
    pooled = car_lib._BatchedUnsortedSegmentFn(   # pylint: disable=protected-access
        featurized_points, dv.indices, dv.num_voxels, p.pooling_method, dv.padding)
    if p.pooling_method == 'max':
      # voxel 0 also receives every out-of-range point: they were zeroed, keep max ≥ real
      pass
    b, c = pooled.shape[0], pooled.shape[-1]
    gx, gy, gz = p.grid_size
    return pooled.reshape(b, gx, gy, gz * c)

  def FProp(self, theta, points_xyz, points_feature, points_padding):
    p = self.params
    dv, enc = self._VoxelizeAndEncodePoints(theta, points_xyz, points_feature, points_padding)
    if p.featurizer is not None:
      enc = self.featurizer.FProp(theta.featurizer, enc)
    enc = enc * (1.0 - dv.padding).unsqueeze(-1)
    return self._ComputeVoxelFeatures(dv, enc)
