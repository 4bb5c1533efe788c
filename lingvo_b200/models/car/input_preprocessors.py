"""Per-example input preprocessors for 3-D detection (ref
`lingvo/tasks/car/input_preprocessors.py`).

A `Preprocessor` maps the NestedMap of ONE example's features (CPU torch tensors, no
batch dim) to a new NestedMap: `TransformFeatures`, plus `TransformShapes` /
`TransformDTypes` describing the effect on static shapes / dtypes. Conventions:

  lasers.points_xyz [P,3]  lasers.points_feature [P,F]  lasers.points_padding [P]
  labels.labels [L]  labels.bboxes_3d [L,7]  labels.bboxes_3d_mask [L]
  (+ labels.difficulties, labels.bboxes_3d_num_points, labels.unfiltered_bboxes_3d_mask)

Randomised preprocessors draw from a per-layer `torch.Generator` seeded from
`p.random_seed` (None → nondeterministic), so the augmentation stream is reproducible and
independent of other RNG consumers. Every augmentation records what it did in
`features.world_*` so the `Inverse*` preprocessors can undo it on predictions.
"""

from __future__ import annotations

import math

import numpy as np
import torch

from lingvo_b200.core import base_layer
from lingvo_b200.core import hyperparams
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.car import car_lib
from lingvo_b200.models.car import detection_3d_lib
from lingvo_b200.models.car import geometry
from lingvo_b200.models.car import ops as car_ops


def _ConsistentShuffle(tensors, gen):
  """Same random permutation of dim 0 for every tensor (ref :36)."""
  n = tensors[0].shape[0]
  perm = torch.randperm(n, generator=gen)
  return tuple(t[perm] for t in tensors)


def _GetApplyPointMaskFn(points_mask):
  """→ fn gathering the kept points of any per-point tensor (ref :47)."""
  idx = torch.nonzero(points_mask, as_tuple=False).squeeze(1)
  return lambda t: t.index_select(0, idx)


class Preprocessor(base_layer.BaseLayer):
  """ref :58."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.name = cls.__name__
    return p

  def __init__(self, params):
    super().__init__(params)
    self._gen = None

  def _Gen(self):
    if self._gen is None:
      self._gen = torch.Generator()
      seed = self.params.random_seed
      if seed is None:
        self._gen.seed()
      else:
        self._gen.manual_seed(int(seed))
    return self._gen

  def _Uniform(self, lo, hi, shape=()):
    return torch.empty(shape).uniform_(float(lo), float(hi), generator=self._Gen())

  def FProp(self, theta, features):
    del theta
    return self.TransformFeatures(features)

  def TransformFeatures(self, features):
    raise NotImplementedError()

  def TransformBatchedFeatures(self, features):
    """Applies `TransformFeatures` to every example of a batched NestedMap."""
    n = next(iter(features.Flatten())).shape[0]
    outs = [self.TransformFeatures(features.Transform(lambda t, i=i: t[i])) for i in range(n)]
    return outs[0].Pack([torch.stack(vs) for vs in zip(*[o.Flatten() for o in outs])])

  def TransformShapes(self, shapes):
    """Default: shapes unchanged."""
    return shapes

  def TransformDTypes(self, dtypes):
    return dtypes


class EntryPreprocessor(Preprocessor):
  """Runs a sub-preprocessor on every entry of `features[input_field]` (a NestedMap of
  NestedMaps, e.g. one per camera or per frame) (ref :150)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('prefixes', ['pseudo_ri'], 'Entries of features to process.')
    p.Define('subprocessors', [], 'Preprocessors applied (in order) to each entry.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self.CreateChildren('subprocessors', list(self.params.subprocessors))

  def _Apply(self, entry, method):
    for sp in self.subprocessors:
      entry = getattr(sp, method)(entry)
    return entry

  def TransformFeatures(self, features):
    for prefix in self.params.prefixes:
      features[prefix] = self._Apply(features[prefix], 'TransformFeatures')
    return features

  def TransformShapes(self, shapes):
    for prefix in self.params.prefixes:
      shapes[prefix] = self._Apply(shapes[prefix], 'TransformShapes')
    return shapes

  def TransformDTypes(self, dtypes):
    for prefix in self.params.prefixes:
      dtypes[prefix] = self._Apply(dtypes[prefix], 'TransformDTypes')
    return dtypes


class CreateDecoderCopy(Preprocessor):
  """Saves un-augmented copies of lasers / images / labels under `decoder_copy` (padded to
  fixed sizes) for decode-time visualisation and metrics (ref :243)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('keys', ['lasers', 'labels', 'images'], 'Top-level keys to copy.')
    p.Define('parent_key', 'decoder_copy', 'Where the copies go.')
    p.Define('pad_lasers', PadLaserFeatures.Params(), 'Padding applied to the laser copy.')
    return p

  def __init__(self, params):
    super().__init__(params)
    if self.params.pad_lasers is not None:
      self.CreateChild('pad_lasers', self.params.pad_lasers)

  def TransformFeatures(self, features):
    p = self.params
    copy = NestedMap()
    for k in p.keys:
      if k in features:
        copy[k] = features[k].DeepCopy().Transform(
            lambda t: t.clone() if isinstance(t, torch.Tensor) else t)
    if p.pad_lasers is not None and 'lasers' in copy:
      copy = self.pad_lasers.TransformFeatures(copy)
    features[p.parent_key] = copy
    return features

  def TransformShapes(self, shapes):
    p = self.params
    copy = NestedMap({k: shapes[k].DeepCopy() for k in p.keys if k in shapes})
    if p.pad_lasers is not None and 'lasers' in copy:
      copy = self.pad_lasers.TransformShapes(copy)
    shapes[p.parent_key] = copy
    return shapes

  def TransformDTypes(self, dtypes):
    p = self.params
    dtypes[p.parent_key] = NestedMap({k: dtypes[k].DeepCopy() for k in p.keys if k in dtypes})
    return dtypes


class FilterByKey(Preprocessor):
  """Keeps only the listed top-level / dotted keys (ref :316)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('keep_key_prefixes', [''], 'Key prefixes to keep.')
    return p

  def _Filter(self, nmap):
    prefixes = self.params.keep_key_prefixes
    return nmap.FilterKeyVal(lambda k, _: any(k.startswith(pre) for pre in prefixes))

  def TransformFeatures(self, features):
    return self._Filter(features)

  def TransformShapes(self, shapes):
    return self._Filter(shapes)

  def TransformDTypes(self, dtypes):
    return self._Filter(dtypes)


class FilterGroundTruthByNumPoints(Preprocessor):
  """Masks out boxes with fewer than `min_num_points` laser points (needs
  `labels.bboxes_3d_num_points`, see CountNumberOfPointsInBoxes3D) (ref :352)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('min_num_points', 1, 'Boxes with fewer points are turned off.')
    return p

  def TransformFeatures(self, features):
    lab = features.labels
    keep = (lab.bboxes_3d_num_points >= self.params.min_num_points).to(lab.bboxes_3d_mask.dtype)
    lab.bboxes_3d_mask = lab.bboxes_3d_mask * keep
    return features


class FilterGroundTruthByDifficulty(Preprocessor):
  """Masks out boxes whose difficulty is not in `difficulty_threshold`'s allowed set:
  keeps `background_id` boxes and those with `difficulties >= threshold` (ref :400)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('background_id', 0, 'Label id of background (always kept).')
    p.Define('difficulty_threshold', 1, 'Keep boxes at least this easy.')
    return p

  def TransformFeatures(self, features):
    p = self.params
    lab = features.labels
    keep = (lab.difficulties >= p.difficulty_threshold) | (lab.labels == p.background_id)
    lab.bboxes_3d_mask = lab.bboxes_3d_mask * keep.to(lab.bboxes_3d_mask.dtype)
    if 'labels' in lab:
      lab.labels = torch.where(keep, lab.labels, torch.full_like(lab.labels, p.background_id))
    return features


class CountNumberOfPointsInBoxes3D(Preprocessor):
  """Adds `labels.bboxes_3d_num_points [L]` (ref :443)."""

  def TransformFeatures(self, features):
    las, lab = features.lasers, features.labels
    inside = geometry.IsWithinBBox3D(las.points_xyz, lab.bboxes_3d)          # [P, L]
    if 'points_padding' in las:
      inside = inside & (las.points_padding < 0.5).unsqueeze(1)
    lab.bboxes_3d_num_points = inside.sum(0).to(torch.int32)
    return features

  def TransformShapes(self, shapes):
    shapes.labels.bboxes_3d_num_points = tuple(shapes.labels.bboxes_3d[:1])
    return shapes

  def TransformDTypes(self, dtypes):
    dtypes.labels.bboxes_3d_num_points = np.int32
    return dtypes


class AddPerPointLabels(Preprocessor):
  """Per-point class / box labels (ref :484): `lasers.points_label [P]`,
  `lasers.points_bbox_id [P]` (−1: none), `lasers.points_bbox_3d [P,7]`; with
  `per_dimension_adjustment` boxes are inflated before the inside test."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('per_dimension_adjustment', None, '[dx, dy, dz] added to box extents.')
    p.Define('minimum_bbox_size', None, '[dx, dy, dz] lower bound on box extents.')
    return p

  def TransformFeatures(self, features):
    p = self.params
    las, lab = features.lasers, features.labels
    boxes = lab.bboxes_3d.clone()
    if p.per_dimension_adjustment:
      boxes[:, 3:6] += torch.tensor(p.per_dimension_adjustment)
    if p.minimum_bbox_size:
      boxes[:, 3:6] = torch.maximum(boxes[:, 3:6], torch.tensor(p.minimum_bbox_size))
    inside = geometry.IsWithinBBox3D(las.points_xyz, boxes) & (lab.bboxes_3d_mask > 0).unsqueeze(0)
    any_in = inside.any(1)
    bid = torch.where(any_in, inside.float().argmax(1), torch.full((inside.shape[0],), -1))
    safe = bid.clamp_min(0)
    las.points_bbox_id = bid.to(torch.int32)
    las.points_label = torch.where(any_in, lab.labels[safe], torch.zeros_like(lab.labels[safe]))
    las.points_bbox_3d = torch.where(any_in.unsqueeze(1), lab.bboxes_3d[safe],
                                     torch.zeros_like(lab.bboxes_3d[safe]))
    return features

  def TransformShapes(self, shapes):
    n = shapes.lasers.points_xyz[0]
    shapes.lasers.points_label, shapes.lasers.points_bbox_id = (n,), (n,)
    shapes.lasers.points_bbox_3d = (n, 7)
    return shapes

  def TransformDTypes(self, dtypes):
    dtypes.lasers.points_label = dtypes.labels.labels
    dtypes.lasers.points_bbox_id = np.int32
    dtypes.lasers.points_bbox_3d = np.float32
    return dtypes


class PointsToGrid(Preprocessor):
  """Bins points into a dense `[gx, gy, gz]` grid with up to `num_points_per_cell` points
  per cell (ref :583): adds `grid_centers`, `grid_num_points`, `laser_grid`."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('num_points_per_cell', 100, 'Points kept per cell.')
    p.Define('grid_size', (40, 40, 1), '(gx, gy, gz).')
    p.Define('grid_range_x', (-80, 80), 'x range.')
    p.Define('grid_range_y', (-80, 80), 'y range.')
    p.Define('grid_range_z', (-2, 4), 'z range.')
    p.Define('normalize_td_labels', True, 'Kept for parity (top-down labels).')
    return p

  def _Centers(self):
    p = self.params
    rng = [p.grid_range_x, p.grid_range_y, p.grid_range_z]
    return detection_3d_lib.Utils3D().CreateDenseCoordinates(
        [(r[0], r[1], g) for r, g in zip(rng, p.grid_size)], center_in_cell=True).reshape(
            tuple(p.grid_size) + (3,))

  def TransformFeatures(self, features):
    p = self.params
    las = features.lasers
    pts = torch.cat([las.points_xyz, las.points_feature], -1)
    pad = las.get('points_padding')
    if pad is None:
      pad = torch.zeros(pts.shape[0])
    dv = car_lib.DynamicVoxelization(las.points_xyz.unsqueeze(0), pad.unsqueeze(0),
                                     p.grid_size, p.grid_range_x, p.grid_range_y,
                                     p.grid_range_z)
    ok = dv.padding[0] < 0.5
    cell = dv.indices[0][ok]
    data = pts[ok]
    n_cells = dv.num_voxels
    order = torch.argsort(cell, stable=True)
    cell, data = cell[order], data[order]
    counts = torch.bincount(cell, minlength=n_cells)
    start = torch.cumsum(counts, 0) - counts
    slot = torch.arange(cell.numel()) - start[cell]
    keep = slot < p.num_points_per_cell
    grid = torch.zeros(n_cells, p.num_points_per_cell, pts.shape[1])
    grid[cell[keep], slot[keep]] = data[keep]
    features.laser_grid = grid.reshape(tuple(p.grid_size) + grid.shape[1:])
    features.grid_num_points = counts.clamp_max(p.num_points_per_cell).reshape(
        tuple(p.grid_size)).to(torch.int32)
    features.grid_centers = self._Centers()
    return features

  def TransformShapes(self, shapes):
    p = self.params
    f = 3 + shapes.lasers.points_feature[-1]
    shapes.grid_centers = tuple(p.grid_size) + (3,)
    shapes.grid_num_points = tuple(p.grid_size)
    shapes.laser_grid = tuple(p.grid_size) + (p.num_points_per_cell, f)
    return shapes

  def TransformDTypes(self, dtypes):
    dtypes.grid_centers, dtypes.laser_grid = np.float32, np.float32
    dtypes.grid_num_points = np.int32
    return dtypes


class _PointPillarGridSettings:
  """PointPillars grid defaults (KITTI): 0.16 m pillars over x∈[0,69.12], y∈[−39.68,39.68]
  (ref :702)."""
  GRID_X, GRID_Y, GRID_Z = 432, 496, 1
  GRID_X_RANGE = (0.0, 69.12)
  GRID_Y_RANGE = (-39.68, 39.68)
  GRID_Z_RANGE = (-3.0, 1.0)

  @classmethod
  def UpdateGridParams(cls, grid_params):
    grid_params.grid_size = (cls.GRID_X, cls.GRID_Y, cls.GRID_Z)
    grid_params.grid_range_x = cls.GRID_X_RANGE
    grid_params.grid_range_y = cls.GRID_Y_RANGE
    grid_params.grid_range_z = cls.GRID_Z_RANGE

  @classmethod
  def UpdateAnchorGridParams(cls, anchor_params, output_stride=2):
    anchor_params.grid_size = (cls.GRID_X // output_stride, cls.GRID_Y // output_stride,
                               cls.GRID_Z)
    anchor_params.grid_range_x = cls.GRID_X_RANGE
    anchor_params.grid_range_y = cls.GRID_Y_RANGE
    anchor_params.grid_range_z = (-1.0, -1.0)


def MakeGridSettings(grid_x_range, grid_y_range, grid_z_range, grid_x, grid_y, grid_z):
  """Class with custom pillar grid settings (ref :754)."""
  class GridSettings(_PointPillarGridSettings):
    GRID_X_RANGE, GRID_Y_RANGE, GRID_Z_RANGE = grid_x_range, grid_y_range, grid_z_range
    GRID_X, GRID_Y, GRID_Z = grid_x, grid_y, grid_z
  return GridSettings


class GridToPillars(Preprocessor):
  """Keeps up to `num_pillars` non-empty cells of the grid as pillars (ref :772): adds
  `point_count [N]`, `point_locations [N,3]` (grid coords), `pillar_points [N,K,F]`."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('num_points_per_pillar', 100, 'Points per pillar.')
    p.Define('num_pillars', 12000, 'Pillars kept.')
    p.Define('drop_laser_grid', True, 'Remove laser_grid afterwards.')
    p.Define('shuffle', True, 'Pick the kept pillars at random (else in grid order).')
    return p

  def TransformFeatures(self, features):
    p = self.params
    grid = features.laser_grid
    gshape = grid.shape[:3]
    flat = grid.reshape((-1,) + tuple(grid.shape[3:]))
    counts = features.grid_num_points.reshape(-1)
    occupied = torch.nonzero(counts > 0, as_tuple=False).squeeze(1)
    if p.shuffle and occupied.numel() > 0:
      occupied = occupied[torch.randperm(occupied.numel(), generator=self._Gen())]
    occupied = occupied[:p.num_pillars]
    n = occupied.numel()
    k = p.num_points_per_pillar
    pillar_points = torch.zeros(p.num_pillars, k, flat.shape[-1])
    pillar_points[:n] = flat[occupied][:, :k]
    point_count = torch.zeros(p.num_pillars, dtype=torch.int32)
    point_count[:n] = counts[occupied].clamp_max(k).to(torch.int32)
    loc = torch.zeros(p.num_pillars, 3, dtype=torch.int32)
    gx, gy, gz = gshape
    loc[:n] = torch.stack([occupied // (gy * gz), (occupied // gz) % gy, occupied % gz], 1).to(
        torch.int32)
    features.pillar_points, features.point_count, features.point_locations = (
        pillar_points, point_count, loc)
    if p.drop_laser_grid:
      del features['laser_grid']
    return features

  def TransformShapes(self, shapes):
    p = self.params
    f = shapes.laser_grid[-1]
    shapes.pillar_points = (p.num_pillars, p.num_points_per_pillar, f)
    shapes.point_count, shapes.point_locations = (p.num_pillars,), (p.num_pillars, 3)
    if p.drop_laser_grid:
      del shapes['laser_grid']
    return shapes

  def TransformDTypes(self, dtypes):
    dtypes.pillar_points = np.float32
    dtypes.point_count, dtypes.point_locations = np.int32, np.int32
    if self.params.drop_laser_grid:
      del dtypes['laser_grid']
    return dtypes


class GridAnchorCenters(Preprocessor):
  """`anchor_centers [gx·gy·gz, 3]` at cell centres (ref :927)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('grid_size', (20, 20, 1), '(gx, gy, gz).')
    p.Define('grid_range_x', (-25, 25), 'x range.')
    p.Define('grid_range_y', (-25, 25), 'y range.')
    p.Define('grid_range_z', (0, 0), 'z range.')
    return p

  def TransformFeatures(self, features):
    p = self.params
    features.anchor_centers = detection_3d_lib.Utils3D().CreateDenseCoordinates(
        [(p.grid_range_x[0], p.grid_range_x[1], p.grid_size[0]),
         (p.grid_range_y[0], p.grid_range_y[1], p.grid_size[1]),
         (p.grid_range_z[0], p.grid_range_z[1], p.grid_size[2])], center_in_cell=True)
    return features

  def TransformShapes(self, shapes):
    g = self.params.grid_size
    shapes.anchor_centers = (g[0] * g[1] * g[2], 3)
    return shapes

  def TransformDTypes(self, dtypes):
    dtypes.anchor_centers = np.float32
    return dtypes


class SparseCenterSelector(Preprocessor):
  """Picks `num_cell_centers` points as cell / anchor centres by farthest-point or uniform
  sampling (ref :980): adds `anchor_centers`, `cell_center_xyz`."""

  _SAMPLING_METHODS = ['farthest_point', 'random_uniform']

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('num_cell_centers', 256, 'Centres to select.')
    p.Define('features_preparation_layers', [], 'Preprocessors run on a copy first.')
    p.Define('sampling_method', 'farthest_point', 'farthest_point | random_uniform.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    if p.sampling_method not in self._SAMPLING_METHODS:
      raise ValueError('Param `sampling_method` must be one of {}.'.format(self._SAMPLING_METHODS))
    self.CreateChildren('features_preparation_layers', list(p.features_preparation_layers))

  def _Sample(self, xyz, pad, num_seeded=0):
    p = self.params
    if p.sampling_method == 'farthest_point':
      idx, _ = car_lib.FarthestPointSampler(
          xyz.unsqueeze(0), pad.unsqueeze(0), p.num_cell_centers,
          num_seeded_points=int(num_seeded),
          random_seed=int(torch.randint(0, 2 ** 31 - 1, (1,), generator=self._Gen())))
      return idx[0]
    noise = torch.rand(xyz.shape[0], generator=self._Gen()).masked_fill(pad > 0.5, -1.0)
    return noise.topk(min(p.num_cell_centers, xyz.shape[0])).indices

  def TransformFeatures(self, features):
    prepared = features.DeepCopy()
    for layer in self.features_preparation_layers:
      prepared = layer.TransformFeatures(prepared)
    las = prepared.lasers
    pad = las.get('points_padding')
    if pad is None:
      pad = torch.zeros(las.points_xyz.shape[0])
    idx = self._Sample(las.points_xyz, pad, las.get('num_seeded_points', 0))
    centers = las.points_xyz[idx]
    features.cell_center_xyz = centers
    features.anchor_centers = centers.clone()
    return features

  def TransformShapes(self, shapes):
    n = self.params.num_cell_centers
    shapes.anchor_centers, shapes.cell_center_xyz = (n, 3), (n, 3)
    return shapes

  def TransformDTypes(self, dtypes):
    dtypes.anchor_centers, dtypes.cell_center_xyz = np.float32, np.float32
    return dtypes


class SparseCellGatherFeatures(Preprocessor):
  """Gathers `num_points_per_cell` neighbours of every cell centre (ref :1148): adds
  `cell_points_xyz [C,K,3]`, `cell_feature [C,K,F]`, `cell_points_padding [C,K]`."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('num_points_per_cell', 128, 'Points per cell.')
    p.Define('max_distance', 3.0, 'Neighbourhood radius (metres).')
    p.Define('sample_neighbors_uniformly', False, 'Random neighbours within the radius.')
    return p

  def TransformFeatures(self, features):
    p = self.params
    las = features.lasers
    pad = las.get('points_padding')
    idx, ipad = car_lib.NeighborhoodIndices(
        las.points_xyz.unsqueeze(0), features.cell_center_xyz.unsqueeze(0),
        p.num_points_per_cell, None if pad is None else (pad > 0.5).unsqueeze(0),
        p.max_distance, p.sample_neighbors_uniformly)
    idx = idx[0]
    features.cell_points_xyz = las.points_xyz[idx]
    features.cell_feature = las.points_feature[idx]
    features.cell_points_padding = ipad[0]
    return features

  def TransformShapes(self, shapes):
    c, k = shapes.cell_center_xyz[0], self.params.num_points_per_cell
    shapes.cell_points_xyz = (c, k, 3)
    shapes.cell_feature = (c, k, shapes.lasers.points_feature[-1])
    shapes.cell_points_padding = (c, k)
    return shapes

  def TransformDTypes(self, dtypes):
    dtypes.cell_points_xyz = dtypes.cell_feature = dtypes.cell_points_padding = np.float32
    return dtypes


class SparseCellCentersTopK(Preprocessor):
  """Keeps the `num_cell_centers` cells with the most real points; every `cell_*` and
  `anchor_centers` tensor is re-ordered consistently (ref :1249)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('num_cell_centers', 512, 'Cells kept.')
    p.Define('sorting_function', None, 'fn(features) → score per cell (default: #points).')
    return p

  def TransformFeatures(self, features):
    p = self.params
    score = p.sorting_function(features) if p.sorting_function else \
        (1.0 - features.cell_points_padding).sum(1)
    idx = score.topk(min(p.num_cell_centers, score.numel())).indices
    for k in list(features.keys()):
      if k.startswith('cell_') or k == 'anchor_centers':
        features[k] = features[k][idx]
    return features

  def TransformShapes(self, shapes):
    n = self.params.num_cell_centers
    for k in list(shapes.keys()):
      if k.startswith('cell_') or k == 'anchor_centers':
        shapes[k] = (n,) + tuple(shapes[k][1:])
    return shapes


class _AnchorBoxSettings:
  """Anchor templates: per class one box size and center-z, tiled at `ROTATIONS`
  (ref :1375)."""
  # (dx, dy, dz) per class and centre heights
  DIMENSION_PRIORS = [(3.9, 1.6, 1.56)]
  ROTATIONS = [0, math.pi / 2]
  CENTER_X_OFFSETS = [0.0]
  CENTER_Y_OFFSETS = [0.0]
  CENTER_Z_OFFSETS = [-1.0]

  @classmethod
  def NumAnchors(cls):
    return (len(cls.DIMENSION_PRIORS) * len(cls.ROTATIONS) * len(cls.CENTER_X_OFFSETS) *
            len(cls.CENTER_Y_OFFSETS) * len(cls.CENTER_Z_OFFSETS))

  @classmethod
  def GenerateAnchorSettings(cls):
    """→ array [A, 7] of (ox, oy, oz, dx, dy, dz, rot)."""
    out = []
    for dims in cls.DIMENSION_PRIORS:
      for rot in cls.ROTATIONS:
        for ox in cls.CENTER_X_OFFSETS:
          for oy in cls.CENTER_Y_OFFSETS:
            for oz in cls.CENTER_Z_OFFSETS:
              out.append((ox, oy, oz) + tuple(dims) + (rot,))
    return np.asarray(out, np.float32)

  @classmethod
  def Update(cls, params):
    s = cls.GenerateAnchorSettings()
    params.anchor_box_dimensions = s[:, 3:6].tolist()
    params.anchor_box_offsets = s[:, 0:3].tolist()
    params.anchor_box_rotations = s[:, 6].tolist()
    return params


def MakeAnchorBoxSettings(dimension_priors, rotations, center_x_offsets, center_y_offsets,
                          center_z_offsets):
  """Class with custom anchor templates (ref :1443)."""
  class CustomAnchorBoxSettings(_AnchorBoxSettings):
    DIMENSION_PRIORS, ROTATIONS = dimension_priors, rotations
    CENTER_X_OFFSETS, CENTER_Y_OFFSETS, CENTER_Z_OFFSETS = (
        center_x_offsets, center_y_offsets, center_z_offsets)
  return CustomAnchorBoxSettings


class SparseCarV1AnchorBoxSettings(_AnchorBoxSettings):
  """StarNet KITTI car anchors (ref :1471)."""
  DIMENSION_PRIORS = [(1.6, 3.9, 1.56)]
  ROTATIONS = [0, math.pi / 2, 3 * math.pi / 4]
  CENTER_X_OFFSETS = [-1.5, 1.5]
  CENTER_Y_OFFSETS = [-1.5, 1.5]
  CENTER_Z_OFFSETS = [0.0]


class PointPillarAnchorBoxSettingsCar(_AnchorBoxSettings):
  DIMENSION_PRIORS = [(1.6, 3.9, 1.56)]
  ROTATIONS = [0, math.pi / 2]
  CENTER_Z_OFFSETS = [-1.0]


class PointPillarAnchorBoxSettingsPed(PointPillarAnchorBoxSettingsCar):
  DIMENSION_PRIORS = [(0.6, 0.8, 1.73)]
  CENTER_Z_OFFSETS = [-0.6]


class PointPillarAnchorBoxSettingsCyc(PointPillarAnchorBoxSettingsCar):
  DIMENSION_PRIORS = [(0.6, 1.76, 1.73)]
  CENTER_Z_OFFSETS = [-0.6]


class PointPillarAnchorBoxSettingsPedCyc(PointPillarAnchorBoxSettingsCar):
  DIMENSION_PRIORS = [(0.6, 0.8, 1.73), (0.6, 1.76, 1.73)]
  CENTER_Z_OFFSETS = [-0.6]


class TileAnchorBBoxes(Preprocessor):
  """`anchor_centers [..., 3]` × A templates → `anchor_bboxes [..., A, 7]` (ref :1321)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('anchor_box_dimensions', [], '[A, 3] sizes.')
    p.Define('anchor_box_offsets', [], '[A, 3] centre offsets.')
    p.Define('anchor_box_rotations', [], '[A] headings.')
    return p

  def TransformFeatures(self, features):
    p = self.params
    c = features.anchor_centers
    flat = c.reshape(-1, 3)
    boxes = detection_3d_lib.Utils3D().MakeAnchorBoxes(
        flat, p.anchor_box_dimensions, p.anchor_box_offsets, p.anchor_box_rotations)
    features.anchor_bboxes = boxes.reshape(tuple(c.shape[:-1]) + boxes.shape[1:])
    return features

  def TransformShapes(self, shapes):
    a = len(self.params.anchor_box_dimensions)
    shapes.anchor_bboxes = tuple(shapes.anchor_centers[:-1]) + (a, 7)
    return shapes

  def TransformDTypes(self, dtypes):
    dtypes.anchor_bboxes = np.float32
    return dtypes


class AnchorAssignment(Preprocessor):
  """IoU-based anchor ↔ ground-truth assignment + regression targets (ref :1510)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('foreground_assignment_threshold', 0.5, 'IoU ≥ this → foreground.')
    p.Define('background_assignment_threshold', 0.35, 'IoU ≤ this → background.')
    return p

  def TransformFeatures(self, features):
    p = self.params
    u = detection_3d_lib.Utils3D()
    ab = features.anchor_bboxes
    flat = ab.reshape(-1, 7)
    lab = features.labels
    a = u.AssignAnchors(flat, lab.bboxes_3d, lab.labels, lab.bboxes_3d_mask,
                        p.foreground_assignment_threshold, p.background_assignment_threshold)
    base = tuple(ab.shape[:-1])
    residuals = u.LocalizationResiduals(flat, a['assigned_gt_bbox'])
    features.anchor_localization_residuals = residuals.reshape(base + (7,))
    features.assigned_gt_idx = a['assigned_gt_idx'].reshape(base).to(torch.int32)
    features.assigned_gt_bbox = a['assigned_gt_bbox'].reshape(base + (7,))
    features.assigned_gt_labels = a['assigned_gt_labels'].reshape(base)
    iou = car_ops.pairwise_iou3d(flat, lab.bboxes_3d) if lab.bboxes_3d.shape[0] else \
        torch.zeros(flat.shape[0], 0)
    sim = iou.gather(1, a['assigned_gt_idx'].clamp_min(0).unsqueeze(1)).squeeze(1) \
        if iou.shape[1] else torch.zeros(flat.shape[0])
    features.assigned_gt_similarity_score = torch.where(
        a['assigned_gt_idx'] >= 0, sim, torch.zeros_like(sim)).reshape(base)
    features.assigned_cls_mask = a['assigned_cls_mask'].reshape(base)
    features.assigned_reg_mask = a['assigned_reg_mask'].reshape(base)
    return features

  def TransformShapes(self, shapes):
    base = tuple(shapes.anchor_bboxes[:-1])
    shapes.anchor_localization_residuals = shapes.assigned_gt_bbox = base + (7,)
    for k in ('assigned_gt_idx', 'assigned_gt_labels', 'assigned_gt_similarity_score',
              'assigned_cls_mask', 'assigned_reg_mask'):
      shapes[k] = base
    return shapes

  def TransformDTypes(self, dtypes):
    for k in ('anchor_localization_residuals', 'assigned_gt_bbox',
              'assigned_gt_similarity_score', 'assigned_cls_mask', 'assigned_reg_mask'):
      dtypes[k] = np.float32
    dtypes.assigned_gt_idx = np.int32
    dtypes.assigned_gt_labels = dtypes.labels.labels
    return dtypes


def _ApplyPointMask(las, keep, mode):
  """Drops points (`mode='remove'`) or flags them padded (`mode='pad'`)."""
  if mode == 'remove':
    fn = _GetApplyPointMaskFn(keep)
    n = las.points_xyz.shape[0]
    for k, v in list(las.items()):
      if isinstance(v, torch.Tensor) and v.dim() >= 1 and v.shape[0] == n:
        las[k] = fn(v)
  else:
    pad = las.get('points_padding')
    base = pad if pad is not None else torch.zeros(las.points_xyz.shape[0])
    las.points_padding = torch.maximum(base, (~keep).float())
  return las


class DropLaserPointsOutOfRange(Preprocessor):
  """Drops (or pads) points outside `keep_*_range` (ref :1615)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    inf = float('inf')
    p.Define('keep_x_range', (-inf, inf), 'x range kept.')
    p.Define('keep_y_range', (-inf, inf), 'y range kept.')
    p.Define('keep_z_range', (-inf, inf), 'z range kept.')
    return p

  def TransformFeatures(self, features):
    p = self.params
    las = features.lasers
    xyz = las.points_xyz
    keep = torch.ones(xyz.shape[0], dtype=torch.bool)
    for d, (lo, hi) in enumerate((p.keep_x_range, p.keep_y_range, p.keep_z_range)):
      keep &= (xyz[:, d] >= lo) & (xyz[:, d] <= hi)
    _ApplyPointMask(las, keep, 'pad' if 'points_padding' in las else 'remove')
    return features


class KITTIDropPointsOutOfFrustum(Preprocessor):
  """Keeps points that project into the camera image (needs `images.velo_to_image_plane
  [3,4]`, `images.width`, `images.height`) (ref :1696)."""

  def TransformFeatures(self, features):
    img = features.images
    uvz = geometry.PointsToImagePlane(features.lasers.points_xyz, img.velo_to_image_plane)
    w, h = float(img.width), float(img.height)
    keep = (uvz[:, 2] >= 0) & (uvz[:, 0] >= 0) & (uvz[:, 0] <= w) & (uvz[:, 1] >= 0) & \
        (uvz[:, 1] <= h)
    _ApplyPointMask(features.lasers, keep,
                    'pad' if 'points_padding' in features.lasers else 'remove')
    return features


class RandomWorldRotationAboutZAxis(Preprocessor):
  """Rotates points and boxes by U(−max_rotation, max_rotation) about z; records
  `world_rot_z` (ref :1754)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('max_rotation', None, 'Max |angle| in radians.')
    p.Define('include_world_rot_z', True, 'Record the angle in features.world_rot_z.')
    return p

  def __init__(self, params):
    super().__init__(params)
    if self.params.max_rotation is None:
      raise ValueError('max_rotation needs to be specified, instead of None.')

  def TransformFeatures(self, features):
    p = self.params
    rot = self._Uniform(-p.max_rotation, p.max_rotation)
    m = geometry.BatchMakeRotationMatrix(rot)
    features.lasers.points_xyz = features.lasers.points_xyz @ m.t()
    if 'labels' in features:
      b = features.labels.bboxes_3d
      features.labels.bboxes_3d = torch.cat(
          [b[:, :3] @ m.t(), b[:, 3:6], geometry.WrapAngleRad(b[:, 6:7] + rot)], -1)
    if p.include_world_rot_z:
      features.world_rot_z = rot.reshape(())
    return features

  def TransformShapes(self, shapes):
    if self.params.include_world_rot_z:
      shapes.world_rot_z = ()
    return shapes

  def TransformDTypes(self, dtypes):
    if self.params.include_world_rot_z:
      dtypes.world_rot_z = np.float32
    return dtypes


class DropPointsOutOfFrustum(Preprocessor):
  """Keeps points within given inclination (θ) and azimuth (φ) ranges (ref :1857)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('keep_theta_range', (0.0, math.pi), 'Inclination range kept.')
    p.Define('keep_phi_range', (0.0, 2 * math.pi), 'Azimuth range kept (radians in [0, 2π)).')
    return p

  def TransformFeatures(self, features):
    p = self.params
    sph = geometry.SphericalCoordinatesTransform(features.lasers.points_xyz)
    theta, phi = sph[:, 1], torch.remainder(sph[:, 2], 2 * math.pi)
    keep = ((theta >= p.keep_theta_range[0]) & (theta <= p.keep_theta_range[1]) &
            (phi >= p.keep_phi_range[0]) & (phi <= p.keep_phi_range[1]))
    _ApplyPointMask(features.lasers, keep,
                    'pad' if 'points_padding' in features.lasers else 'remove')
    return features


class DropBoxesOutOfRange(Preprocessor):
  """Masks boxes whose centre (or any corner) lies outside the range (ref :1956)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    inf = float('inf')
    p.Define('keep_x_range', (-inf, inf), 'x range kept.')
    p.Define('keep_y_range', (-inf, inf), 'y range kept.')
    p.Define('keep_z_range', (-inf, inf), 'z range kept.')
    p.Define('use_corners', False, 'Require every corner (not just the centre) inside.')
    return p

  def TransformFeatures(self, features):
    p = self.params
    lab = features.labels
    pts = geometry.BBoxCorners(lab.bboxes_3d) if p.use_corners else lab.bboxes_3d[:, None, :3]
    keep = torch.ones(lab.bboxes_3d.shape[0], dtype=torch.bool)
    for d, (lo, hi) in enumerate((p.keep_x_range, p.keep_y_range, p.keep_z_range)):
      keep &= ((pts[..., d] >= lo) & (pts[..., d] <= hi)).all(-1)
    lab.bboxes_3d_mask = lab.bboxes_3d_mask * keep.to(lab.bboxes_3d_mask.dtype)
    return features


class PadLaserFeatures(Preprocessor):
  """Pads / randomly trims the point cloud to `max_num_points`; adds `points_padding`
  (ref :2023)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('max_num_points', 128500, 'Static number of points.')
    return p

  def TransformFeatures(self, features):
    p = self.params
    las = features.lasers
    n = las.points_xyz.shape[0]
    if 'points_padding' in las:                   # real points first
      order = torch.argsort(las.points_padding, stable=True)
      n_real = int((las.points_padding < 0.5).sum())
    else:
      order, n_real = torch.arange(n), n
    if n_real > p.max_num_points:
      pick = torch.randperm(n_real, generator=self._Gen())[:p.max_num_points]
      order = order[pick.sort().values]
      n_real = p.max_num_points
    else:
      order = order[:n_real]
    m = p.max_num_points
    for k, v in list(las.items()):
      if isinstance(v, torch.Tensor) and v.dim() >= 1 and v.shape[0] == n and k != 'points_padding':
        out = torch.zeros((m,) + tuple(v.shape[1:]), dtype=v.dtype)
        out[:n_real] = v[order]
        las[k] = out
    pad = torch.ones(m)
    pad[:n_real] = 0.0
    las.points_padding = pad
    return features

  def TransformShapes(self, shapes):
    m = self.params.max_num_points
    las = shapes.lasers
    for k in list(las.keys()):
      if las[k] is not None and k != 'points_padding' and len(las[k]) >= 1:
        las[k] = (m,) + tuple(las[k][1:])
    las.points_padding = (m,)
    return shapes

  def TransformDTypes(self, dtypes):
    dtypes.lasers.points_padding = np.float32
    return dtypes


class WorldScaling(Preprocessor):
  """Scales points and boxes by U(scaling); records `world_scaling` (ref :2088)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('scaling', None, '(min, max) scale.')
    return p

  def __init__(self, params):
    super().__init__(params)
    s = self.params.scaling
    if s is None or len(s) != 2 or s[0] > s[1]:
      raise ValueError('scaling needs to be a (min, max) pair.')

  def TransformFeatures(self, features):
    s = self._Uniform(*self.params.scaling)
    features.lasers.points_xyz = features.lasers.points_xyz * s
    if 'labels' in features:
      b = features.labels.bboxes_3d
      features.labels.bboxes_3d = torch.cat([b[:, :6] * s, b[:, 6:]], -1)
    features.world_scaling = s.reshape(())
    return features

  def TransformShapes(self, shapes):
    shapes.world_scaling = ()
    return shapes

  def TransformDTypes(self, dtypes):
    dtypes.world_scaling = np.float32
    return dtypes


class RandomDropLaserPoints(Preprocessor):
  """Keeps each point with probability `keep_prob` (ref :2156)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('keep_prob', 0.95, 'Keep probability.')
    return p

  def TransformFeatures(self, features):
    las = features.lasers
    keep = torch.rand(las.points_xyz.shape[0], generator=self._Gen()) < self.params.keep_prob
    _ApplyPointMask(las, keep, 'pad' if 'points_padding' in las else 'remove')
    return features


class RandomFlipY(Preprocessor):
  """Mirrors the scene across the x axis (y → −y, φ → −φ) with `flip_probability`; records
  `world_flip_y` (ref :2204)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('flip_probability', 0.5, 'Probability of flipping.')
    return p

  def TransformFeatures(self, features):
    flip = bool(torch.rand((), generator=self._Gen()) < self.params.flip_probability)
    if flip:
      xyz = features.lasers.points_xyz
      features.lasers.points_xyz = xyz * torch.tensor([1.0, -1.0, 1.0])
      if 'labels' in features:
        b = features.labels.bboxes_3d.clone()
        b[:, 1] = -b[:, 1]
        b[:, 6] = geometry.WrapAngleRad(-b[:, 6])
        features.labels.bboxes_3d = b
    features.world_flip_y = torch.tensor(1.0 if flip else 0.0)
    return features

  def TransformShapes(self, shapes):
    shapes.world_flip_y = ()
    return shapes

  def TransformDTypes(self, dtypes):
    dtypes.world_flip_y = np.float32
    return dtypes


class GlobalTranslateNoise(Preprocessor):
  """Adds N(0, noise_std) to all points and box centres; records `world_translate`
  (ref :2278)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('noise_std', [0.2, 0.2, 0.2], 'Std-dev per axis.')
    return p

  def TransformFeatures(self, features):
    t = torch.randn(3, generator=self._Gen()) * torch.tensor(self.params.noise_std)
    features.lasers.points_xyz = features.lasers.points_xyz + t
    if 'labels' in features:
      b = features.labels.bboxes_3d
      features.labels.bboxes_3d = torch.cat([b[:, :3] + t, b[:, 3:]], -1)
    features.world_translate = t
    return features

  def TransformShapes(self, shapes):
    shapes.world_translate = (3,)
    return shapes

  def TransformDTypes(self, dtypes):
    dtypes.world_translate = np.float32
    return dtypes


class RandomBBoxTransform(Preprocessor):
  """Per-object augmentation (ref :2361): every real box (and the points inside it) gets
  its own random rotation about its centre, translation noise and scaling; a move is
  rejected (retried up to `max_attempts`) when the new box would collide with another."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('max_rotation', None, 'Max |rotation| about the box centre (radians).')
    p.Define('noise_std', None, '[sx, sy, sz] translation noise.')
    p.Define('max_scaling', None, 'Scale drawn from U(1−m, 1+m) per axis ([mx,my,mz]).')
    p.Define('max_shearing', None, 'Kept for parity (unused).')
    p.Define('max_attempts', 5, 'Retries per box before leaving it unchanged.')
    return p

  def TransformFeatures(self, features):
    p = self.params
    las, lab = features.lasers, features.labels
    boxes = lab.bboxes_3d.clone()
    xyz = las.points_xyz.clone()
    real = torch.nonzero(lab.bboxes_3d_mask > 0, as_tuple=False).squeeze(1).tolist()
    inside = geometry.IsWithinBBox3D(xyz, boxes)
    for i in real:
      for _ in range(p.max_attempts):
        rot = float(self._Uniform(-p.max_rotation, p.max_rotation)) if p.max_rotation else 0.0
        shift = torch.randn(3, generator=self._Gen()) * torch.tensor(p.noise_std) \
            if p.noise_std else torch.zeros(3)
        scale = (1.0 + (torch.rand(3, generator=self._Gen()) * 2 - 1) *
                 torch.tensor(p.max_scaling)) if p.max_scaling else torch.ones(3)
        cand = boxes[i].clone()
        cand[:3] += shift
        cand[3:6] *= scale
        cand[6] = geometry.WrapAngleRad(cand[6] + rot)
        others = [j for j in real if j != i]
        if others and float(car_ops.pairwise_iou3d(cand[None], boxes[others]).max()) > 0:
          continue
        pts = inside[:, i]
        if pts.any():
          local = car_lib.LocalTransform(xyz[pts], boxes[i].expand(int(pts.sum()), 7)) * scale
          c, s = math.cos(float(cand[6])), math.sin(float(cand[6]))
          xyz[pts] = torch.stack([local[:, 0] * c - local[:, 1] * s + cand[0],
                                  local[:, 0] * s + local[:, 1] * c + cand[1],
                                  local[:, 2] + cand[2]], -1)
        boxes[i] = cand
        break
    las.points_xyz, lab.bboxes_3d = xyz, boxes
    return features


class GroundTruthAugmentor(Preprocessor):
  """Pastes objects (boxes + their points) sampled from a database into the scene
  (ref :2708). The database is a list of dicts {bbox_3d [7], label, points_xyz [n,3],
  points_feature [n,F]} built offline; candidates overlapping existing boxes are skipped
  and points of the scene inside a pasted box are removed."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('groundtruth_database', None, 'List of object dicts or a path to a .npy dump.')
    p.Define('num_db_objects', None, 'Use only the first N database objects.')
    p.Define('max_num_points_per_bbox', 2048, 'Points kept per pasted object.')
    p.Define('filter_min_points', 0, 'Ignore database objects with fewer points.')
    p.Define('filter_max_points', None, 'Ignore database objects with more points.')
    p.Define('difficulty_sampling_probability', None, 'Kept for parity.')
    p.Define('class_sampling_probability', None, 'Per-class acceptance probability.')
    p.Define('filter_min_difficulty', 0, 'Kept for parity.')
    p.Define('max_augmented_bboxes', 15, 'Objects pasted per scene.')
    p.Define('label_filter', [], 'Only paste these labels (empty: all).')
    p.Define('batch_mode', False, 'Kept for parity.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    db = p.groundtruth_database
    if isinstance(db, str):
      db = list(np.load(db, allow_pickle=True))
    db = list(db or [])
    if p.num_db_objects:
      db = db[:p.num_db_objects]
    def _Ok(o):
      n = len(o['points_xyz'])
      if n < p.filter_min_points or (p.filter_max_points and n > p.filter_max_points):
        return False
      return not p.label_filter or int(o['label']) in p.label_filter
    self._db = [o for o in db if _Ok(o)]

  def TransformFeatures(self, features):
    p = self.params
    if not self._db:
      return features
    las, lab = features.lasers, features.labels
    boxes, mask, labels = lab.bboxes_3d.clone(), lab.bboxes_3d_mask.clone(), lab.labels.clone()
    free = torch.nonzero(mask <= 0, as_tuple=False).squeeze(1).tolist()
    new_xyz, new_feat = [], []
    order = torch.randperm(len(self._db), generator=self._Gen()).tolist()
    pasted = 0
    for oi in order:
      if pasted >= p.max_augmented_bboxes or not free:
        break
      o = self._db[oi]
      if p.class_sampling_probability is not None:
        prob = p.class_sampling_probability[int(o['label'])]
        if float(torch.rand((), generator=self._Gen())) > prob:
          continue
      cand = torch.as_tensor(np.asarray(o['bbox_3d'], np.float32))
      live = boxes[mask > 0]
      if live.shape[0] and float(car_ops.pairwise_iou3d(cand[None], live).max()) > 0:
        continue
      slot = free.pop(0)
      boxes[slot], mask[slot] = cand, 1.0
      labels[slot] = int(o['label'])
      pts = torch.as_tensor(np.asarray(o['points_xyz'], np.float32))[:p.max_num_points_per_bbox]
      ft = torch.as_tensor(np.asarray(o['points_feature'], np.float32)).reshape(
          len(o['points_xyz']), -1)[:p.max_num_points_per_bbox]
      new_xyz.append(pts)
      new_feat.append(ft)
      pasted += 1
    if pasted:
      added = boxes[[i for i in range(len(mask)) if mask[i] > 0 and lab.bboxes_3d_mask[i] <= 0]]
      keep = ~geometry.IsWithinBBox3D(las.points_xyz, added).any(1)
      if 'points_padding' in las:
        keep = keep & (las.points_padding < 0.5)
      las.points_xyz = torch.cat([las.points_xyz[keep]] + new_xyz)
      las.points_feature = torch.cat([las.points_feature[keep]] + new_feat)
      if 'points_padding' in las:
        las.points_padding = torch.zeros(las.points_xyz.shape[0])
      lab.bboxes_3d, lab.bboxes_3d_mask, lab.labels = boxes, mask, labels
    return features


class FrustumDropout(Preprocessor):
  """Drops (or noises) the points within a random cone around a random real point: all
  points whose (θ, φ) lie within `theta_width / phi_width` of it, optionally only beyond
  `distance`, each dropped with prob 1 − keep_prob (ref :3093)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('theta_width', 0.03, 'Inclination half-width (radians).')
    p.Define('phi_width', 0.0, 'Azimuth half-width (radians).')
    p.Define('distance', 0.0, 'Only points farther than this from the sensor.')
    p.Define('keep_prob', 0.0, 'Probability of keeping a point in the frustum.')
    p.Define('drop_type', 'union', "'union': θ OR φ within width; 'intersection': both.")
    return p

  def __init__(self, params):
    super().__init__(params)
    if self.params.drop_type not in ('union', 'intersection'):
      raise ValueError('drop_type must be union or intersection.')

  def TransformFeatures(self, features):
    p = self.params
    las = features.lasers
    real = (las.points_padding < 0.5) if 'points_padding' in las else torch.ones(
        las.points_xyz.shape[0], dtype=torch.bool)
    cand = torch.nonzero(real, as_tuple=False).squeeze(1)
    if cand.numel() == 0:
      return features
    seed = cand[int(torch.randint(0, cand.numel(), (1,), generator=self._Gen()))]
    sph = geometry.SphericalCoordinatesTransform(las.points_xyz)
    d_theta = (sph[:, 1] - sph[seed, 1]).abs()
    d_phi = geometry.WrapAngleRad(sph[:, 2] - sph[seed, 2]).abs()
    in_t, in_p = d_theta < p.theta_width, d_phi < p.phi_width
    sel = (in_t | in_p) if p.drop_type == 'union' else (in_t & in_p)
    sel &= sph[:, 0] > p.distance
    drop = sel & (torch.rand(sel.shape[0], generator=self._Gen()) >= p.keep_prob)
    _ApplyPointMask(las, ~drop, 'pad' if 'points_padding' in las else 'remove')
    return features


class RepeatPreprocessor(Preprocessor):
  """Applies a sub-preprocessor `repeat_count` times (ref :3247)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('repeat_count', 1, 'Repetitions.')
    p.Define('subprocessor', None, 'Preprocessor params.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self.CreateChild('subprocessor', self.params.subprocessor)

  def TransformFeatures(self, features):
    for _ in range(self.params.repeat_count):
      features = self.subprocessor.TransformFeatures(features)
    return features

  def TransformShapes(self, shapes):
    return self.subprocessor.TransformShapes(shapes)

  def TransformDTypes(self, dtypes):
    return self.subprocessor.TransformDTypes(dtypes)


class RandomApplyPreprocessor(Preprocessor):
  """Applies the sub-preprocessor with probability `prob` (shapes must not change);
  records the coin in `features.<name>_applied` when `record_choice` (ref :3298)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('prob', 1.0, 'Probability of applying.')
    p.Define('subprocessor', None, 'Preprocessor params.')
    p.Define('choice_key', None, 'If set, store 1/0 under this features key.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self.CreateChild('subprocessor', self.params.subprocessor)

  def TransformFeatures(self, features):
    p = self.params
    apply = bool(torch.rand((), generator=self._Gen()) <= p.prob)
    if apply:
      features = self.subprocessor.TransformFeatures(features)
    if p.choice_key:
      features[p.choice_key] = torch.tensor(1.0 if apply else 0.0)
    return features

  def TransformShapes(self, shapes):
    before = shapes.DeepCopy()
    after = self.subprocessor.TransformShapes(shapes)
    if sorted(after.FlattenItems()) != sorted(before.FlattenItems()):
      raise ValueError('RandomApplyPreprocessor: the sub-preprocessor must not change shapes.')
    if self.params.choice_key:
      after[self.params.choice_key] = ()
    return after

  def TransformDTypes(self, dtypes):
    dtypes = self.subprocessor.TransformDTypes(dtypes)
    if self.params.choice_key:
      dtypes[self.params.choice_key] = np.float32
    return dtypes


class ConstantPreprocessor(Preprocessor):
  """Adds constant features (ref :3398)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('constants', {}, 'dotted key → python / numpy constant.')
    return p

  def TransformFeatures(self, features):
    for k, v in self.params.constants.items():
      features.Set(k, torch.as_tensor(np.asarray(v)))
    return features

  def TransformShapes(self, shapes):
    for k, v in self.params.constants.items():
      shapes.Set(k, tuple(np.asarray(v).shape))
    return shapes

  def TransformDTypes(self, dtypes):
    for k, v in self.params.constants.items():
      dtypes.Set(k, np.asarray(v).dtype.type)
    return dtypes


class IdentityPreprocessor(Preprocessor):
  """Does nothing (ref :3427)."""

  def TransformFeatures(self, features):
    return features


class RandomChoicePreprocessor(Preprocessor):
  """Applies exactly one of several sub-preprocessors, chosen with the given weights
  (constants or schedules) (ref :3445). All choices must produce identical shapes."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('subprocessors', [], 'List of (preprocessor params, weight or schedule params).')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    if not p.subprocessors:
      raise ValueError('No subprocessors were specified.')
    subs, self._weights = [], []
    for sp, w in p.subprocessors:
      subs.append(sp)
      self._weights.append(w.Instantiate() if hasattr(w, 'Instantiate') else float(w))
    self.CreateChildren('subprocessors', subs)

  def _Probabilities(self):
    w = torch.tensor([float(x.Value()) if hasattr(x, 'Value') else x for x in self._weights])
    return w / w.sum()

  def TransformFeatures(self, features):
    i = int(torch.multinomial(self._Probabilities(), 1, generator=self._Gen()))
    return self.subprocessors[i].TransformFeatures(features)

  def TransformShapes(self, shapes):
    outs = [sp.TransformShapes(shapes.DeepCopy()) for sp in self.subprocessors]
    ref = sorted(outs[0].FlattenItems())
    if any(sorted(o.FlattenItems()) != ref for o in outs[1:]):
      raise ValueError('Shapes not compatible across the choices.')
    return outs[0]

  def TransformDTypes(self, dtypes):
    return self.subprocessors[0].TransformDTypes(dtypes)


class Sequence(Preprocessor):
  """Runs a list of preprocessors in order (ref :3527)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('preprocessors', [], 'Preprocessor params.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self.CreateChildren('preprocessors', list(self.params.preprocessors))

  def TransformFeatures(self, features):
    for pre in self.preprocessors:
      features = pre.TransformFeatures(features)
    return features

  def TransformShapes(self, shapes):
    for pre in self.preprocessors:
      shapes = pre.TransformShapes(shapes)
    return shapes

  def TransformDTypes(self, dtypes):
    for pre in self.preprocessors:
      dtypes = pre.TransformDTypes(dtypes)
    return dtypes


class SparseSampler(Preprocessor):
  """Fused centre selection + neighbour gathering through the native sampling op
  (ref :3559): adds `anchor_centers`, `cell_center_xyz`, `cell_center_padding`,
  `cell_points_xyz`, `cell_feature`, `cell_points_padding`."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('center_selector', 'farthest', 'farthest | uniform.')
    p.Define('neighbor_sampler', 'uniform', 'uniform | closest.')
    p.Define('num_centers', 16, 'Centres.')
    p.Define('features_preparation_layers', [], 'Preprocessors run on a copy first.')
    p.Define('keep_z_range', (-float('inf'), float('inf')), 'Only centres with z in range.')
    p.Define('num_neighbors', 64, 'Points per cell.')
    p.Define('max_distance', 1.0, 'Neighbourhood radius.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self.CreateChildren('features_preparation_layers',
                        list(self.params.features_preparation_layers))

  def TransformFeatures(self, features):
    p = self.params
    prepared = features.DeepCopy()
    for layer in self.features_preparation_layers:
      prepared = layer.TransformFeatures(prepared)
    las = prepared.lasers
    pad = las.get('points_padding')
    pad = pad if pad is not None else torch.zeros(las.points_xyz.shape[0])
    seed = int(torch.randint(0, 2 ** 31 - 1, (1,), generator=self._Gen()))
    # One native call: centre selection restricted to keep_z_range + ball query.
    c, cpad, idx, ipad = car_ops.sample_points(
        las.points_xyz.unsqueeze(0), pad.unsqueeze(0), p.num_centers, p.num_neighbors,
        p.max_distance, p.center_selector, seed, neighbor_sampler=p.neighbor_sampler,
        center_z_min=p.keep_z_range[0], center_z_max=p.keep_z_range[1])
    centers = las.points_xyz[c[0]]
    features.cell_center_xyz = centers
    features.anchor_centers = centers.clone()
    features.cell_center_padding = cpad[0]
    features.cell_points_xyz = las.points_xyz[idx[0]]
    features.cell_feature = las.points_feature[idx[0]]
    features.cell_points_padding = ipad[0]
    return features

  def TransformShapes(self, shapes):
    p = self.params
    c, k = p.num_centers, p.num_neighbors
    shapes.anchor_centers = shapes.cell_center_xyz = (c, 3)
    shapes.cell_center_padding = (c,)
    shapes.cell_points_xyz = (c, k, 3)
    shapes.cell_feature = (c, k, shapes.lasers.points_feature[-1])
    shapes.cell_points_padding = (c, k)
    return shapes

  def TransformDTypes(self, dtypes):
    for k in ('anchor_centers', 'cell_center_xyz', 'cell_center_padding', 'cell_points_xyz',
              'cell_feature', 'cell_points_padding'):
      dtypes[k] = np.float32
    return dtypes


class PointAssignment(Preprocessor):
  """Anchor-free assignment (ref :3700): each `anchor_centers` point inside a real box is
  foreground for it; targets are the box in the point's frame (Δxyz, log dims, sin/cos-free
  Δφ), plus FCOS centerness."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('extra_box_size', (0.0, 0.0, 0.0), 'Inflation of boxes for the inside test.')
    p.Define('centerness_range', (0.0, 1.0), 'Range of the centerness label.')
    p.Define('num_classes', 1, 'Number of foreground classes.')
    return p

  def TransformFeatures(self, features):
    p = self.params
    ctr = features.anchor_centers
    base = tuple(ctr.shape[:-1])
    pts = ctr.reshape(-1, 3)
    lab = features.labels
    boxes = lab.bboxes_3d.clone()
    boxes[:, 3:6] += torch.tensor(p.extra_box_size)
    inside = geometry.IsWithinBBox3D(pts, boxes) & (lab.bboxes_3d_mask > 0).unsqueeze(0)
    # a point inside several boxes goes to the one whose centre is nearest
    d = car_lib.SquaredDistanceMatrix(pts.unsqueeze(0), lab.bboxes_3d[:, :3].unsqueeze(0))[0]
    d = d.masked_fill(~inside, float('inf'))
    fg = inside.any(1)
    idx = torch.where(fg, d.argmin(1), torch.full((pts.shape[0],), -1))
    safe = idx.clamp_min(0)
    gt = torch.where(fg.unsqueeze(1), lab.bboxes_3d[safe], torch.zeros(pts.shape[0], 7))
    residual = torch.cat([gt[:, :3] - pts, torch.log(gt[:, 3:6].clamp_min(1e-6)), gt[:, 6:7]], -1)
    residual = torch.where(fg.unsqueeze(1), residual, torch.zeros_like(residual))
    features.target_predictions = residual.reshape(base + (7,))
    features.assigned_gt_idx = idx.reshape(base).to(torch.int32)
    features.assigned_gt_bbox = gt.reshape(base + (7,))
    features.assigned_gt_labels = torch.where(fg, lab.labels[safe],
                                              torch.zeros_like(lab.labels[safe])).reshape(base)
    cn = car_lib.GenerateCenternessLabel(pts, torch.where(
        fg.unsqueeze(1), gt, torch.ones(pts.shape[0], 7)), p.centerness_range)
    features.assigned_gt_center_ness = torch.where(fg, cn, torch.zeros_like(cn)).reshape(base)
    features.assigned_cls_mask = torch.ones(base)
    reg = torch.zeros(pts.shape[0], p.num_classes)
    cls = (features.assigned_gt_labels.reshape(-1) - 1).clamp(0, p.num_classes - 1).long()
    reg[torch.arange(pts.shape[0]), cls] = fg.float()
    features.assigned_reg_mask = reg.reshape(base + (p.num_classes,))
    return features

  def TransformShapes(self, shapes):
    base = tuple(shapes.anchor_centers[:-1])
    shapes.target_predictions = shapes.assigned_gt_bbox = base + (7,)
    for k in ('assigned_gt_idx', 'assigned_gt_labels', 'assigned_gt_center_ness',
              'assigned_cls_mask'):
      shapes[k] = base
    shapes.assigned_reg_mask = base + (self.params.num_classes,)
    return shapes

  def TransformDTypes(self, dtypes):
    for k in ('target_predictions', 'assigned_gt_bbox', 'assigned_gt_center_ness',
              'assigned_cls_mask', 'assigned_reg_mask'):
      dtypes[k] = np.float32
    dtypes.assigned_gt_idx = np.int32
    dtypes.assigned_gt_labels = dtypes.labels.labels
    return dtypes


class FrustumNoise(Preprocessor):
  """Like `FrustumDropout`, but perturbs the selected points' range by U(±noise) along
  their ray instead of dropping them (ref :3865)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('theta_width', 0.03, 'Inclination half-width.')
    p.Define('phi_width', 0.0, 'Azimuth half-width.')
    p.Define('distance', 0.0, 'Only points farther than this.')
    p.Define('noise_magnitude', 1.0, 'Max range perturbation (metres).')
    p.Define('drop_type', 'union', 'union | intersection.')
    return p

  def TransformFeatures(self, features):
    p = self.params
    las = features.lasers
    n = las.points_xyz.shape[0]
    seed = int(torch.randint(0, n, (1,), generator=self._Gen()))
    sph = geometry.SphericalCoordinatesTransform(las.points_xyz)
    in_t = (sph[:, 1] - sph[seed, 1]).abs() < p.theta_width
    in_p = geometry.WrapAngleRad(sph[:, 2] - sph[seed, 2]).abs() < p.phi_width
    sel = ((in_t | in_p) if p.drop_type == 'union' else (in_t & in_p)) & (sph[:, 0] > p.distance)
    noise = (torch.rand(n, generator=self._Gen()) * 2 - 1) * p.noise_magnitude
    scale = torch.where(sel, (sph[:, 0] + noise).clamp_min(0.0) / sph[:, 0].clamp_min(1e-6),
                        torch.ones(n))
    las.points_xyz = las.points_xyz * scale.unsqueeze(1)
    return features


class PerPillarPointCloudCenters(Preprocessor):
  """Adds `pillar_centers [N, 3]`: the metric centre of each pillar from its grid location
  (ref :3998)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('grid_size', (432, 496, 1), '(gx, gy, gz).')
    p.Define('grid_range_x', (0.0, 69.12), 'x range.')
    p.Define('grid_range_y', (-39.68, 39.68), 'y range.')
    p.Define('grid_range_z', (-3.0, 1.0), 'z range.')
    return p

  def TransformFeatures(self, features):
    p = self.params
    lo = torch.tensor([p.grid_range_x[0], p.grid_range_y[0], p.grid_range_z[0]])
    hi = torch.tensor([p.grid_range_x[1], p.grid_range_y[1], p.grid_range_z[1]])
    size = (hi - lo) / torch.tensor(p.grid_size, dtype=torch.float32)
    features.pillar_centers = (features.point_locations.float() + 0.5) * size + lo
    return features

  def TransformShapes(self, shapes):
    shapes.pillar_centers = tuple(shapes.point_locations)
    return shapes

  def TransformDTypes(self, dtypes):
    dtypes.pillar_centers = np.float32
    return dtypes


class CopyFeatures(Preprocessor):
  """Copies `features[src]` to `features[dst]` for each (src, dst) (ref :4050)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('source_target_pairs', [], '[(src key, dst key)].')
    return p

  def _Copy(self, nmap, clone):
    for src, dst in self.params.source_target_pairs:
      v = nmap.GetItem(src)
      nmap.Set(dst, v.clone() if clone and isinstance(v, torch.Tensor) else v)
    return nmap

  def TransformFeatures(self, features):
    return self._Copy(features, True)

  def TransformShapes(self, shapes):
    return self._Copy(shapes, False)

  def TransformDTypes(self, dtypes):
    return self._Copy(dtypes, False)


# ---- inverses: undo world augmentations on predicted boxes at decode time ------------
class InverseRandomApplyPreprocessor(Preprocessor):
  """Runs the (inverse) sub-preprocessor iff `features[choice_key]` says the forward one
  was applied (ref :4077)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('subprocessor', None, 'Inverse preprocessor params.')
    p.Define('choice_key', None, 'Key written by RandomApplyPreprocessor.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self.CreateChild('subprocessor', self.params.subprocessor)

  def TransformFeatures(self, features):
    if float(features[self.params.choice_key]) > 0.5:
      features = self.subprocessor.TransformFeatures(features)
    return features


class _InverseBase(Preprocessor):
  """Operates on `features[bbox_key] [..., 7]` (e.g. predicted boxes)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('bbox_key', 'predicted_bboxes', 'Key of the boxes to transform.')
    return p


class InverseWorldScaling(_InverseBase):
  """ref :4148."""

  def TransformFeatures(self, features):
    k = self.params.bbox_key
    b = features[k]
    features[k] = torch.cat([b[..., :6] / features.world_scaling, b[..., 6:]], -1)
    return features


class InverseGlobalTranslateNoise(_InverseBase):
  """ref :4195."""

  def TransformFeatures(self, features):
    k = self.params.bbox_key
    b = features[k]
    features[k] = torch.cat([b[..., :3] - features.world_translate, b[..., 3:]], -1)
    return features


class InverseRandomFlipY(_InverseBase):
  """ref :4243."""

  def TransformFeatures(self, features):
    if float(features.world_flip_y) > 0.5:
      k = self.params.bbox_key
      b = features[k].clone()
      b[..., 1] = -b[..., 1]
      b[..., 6] = geometry.WrapAngleRad(-b[..., 6])
      features[k] = b
    return features


class InverseRandomWorldRotationAboutZAxis(_InverseBase):
  """ref :4301."""

  def TransformFeatures(self, features):
    k = self.params.bbox_key
    b = features[k]
    rot = -features.world_rot_z
    m = geometry.BatchMakeRotationMatrix(rot)
    features[k] = torch.cat([b[..., :3] @ m.t(), b[..., 3:6],
                             geometry.WrapAngleRad(b[..., 6:7] + rot)], -1)
    return features



