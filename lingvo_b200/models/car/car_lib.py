"""Tensor routines for point-cloud layers and models (ref `lingvo/tasks/car/car_lib.py`).

Neighbourhood search, farthest-point sampling, point pooling, dynamic voxelisation and
label helpers, all as batched torch code (device-resident, no host loops except the
inherently sequential FPS loop, which stays on the device).
"""

from __future__ import annotations

import torch

from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.car import geometry


def SquaredDistanceMatrix(pa, pb, mem_optimized=False):
  """pa `[N, P1, D]`, pb `[N, P2, D]` → squared distances `[N, P1, P2]` (ref :26).
  `mem_optimized` uses |a|² + |b|² − 2a·b (one GEMM, no `[N,P1,P2,D]` temporary)."""
  if mem_optimized:
    a2 = pa.square().sum(-1, keepdim=True)
    b2 = pb.square().sum(-1).unsqueeze(-2)
    return (a2 + b2 - 2.0 * torch.matmul(pa, pb.transpose(-1, -2))).clamp_min(0.0)
  return (pa.unsqueeze(-2) - pb.unsqueeze(-3)).square().sum(-1)


def NeighborSquaredDistanceMatrix(points, neighbor_points):
  """points `[N, P, 3]`, neighbour_points `[N, P, K, 3]` → `[N, P, K]` (ref :67)."""
  return (points.unsqueeze(2) - neighbor_points).square().sum(-1)


def KnnIndices(points, query_points, k, valid_num=None, max_distance=None):
  """k nearest neighbours of each query among the first `valid_num` points (ref :89)."""
  padding = None
  if valid_num is not None:
    p1 = points.shape[1]
    padding = torch.arange(p1, device=points.device).unsqueeze(0) >= valid_num.unsqueeze(-1)
  return NeighborhoodIndices(points, query_points, k, padding, max_distance)


def NeighborhoodIndices(points, query_points, k, points_padding=None, max_distance=None,
                        sample_neighbors_uniformly=False):
  """→ (indices `[N, P2, k]`, padding `[N, P2, k]`) (ref :139). Real (unpadded) results are
  distinct real points; slots that cannot be filled (too few points, or beyond
  `max_distance`) repeat the closest point and are marked padded. With
  `sample_neighbors_uniformly`, the k are drawn at random among the in-range points."""
  n, p1, _ = points.shape
  d = SquaredDistanceMatrix(query_points, points, mem_optimized=True)      # [N, P2, P1]
  big = torch.finfo(d.dtype).max / 4
  if points_padding is not None:
    d = d.masked_fill(points_padding.bool().unsqueeze(1), big)
  in_range = d < big
  if max_distance is not None:
    in_range = in_range & (d <= max_distance * max_distance)
  if sample_neighbors_uniformly:
    key = torch.rand_like(d)
    key = torch.where(in_range, key, 2.0 + d / d.max().clamp_min(1.0))
  else:
    key = torch.where(in_range, d, big + d)        # out-of-range sorted after in-range
  kk = min(k, p1)
  _, idx = torch.topk(key, kk, dim=-1, largest=False)
  ok = in_range.gather(-1, idx)
  if kk < k:
    idx = torch.cat([idx, idx[..., :1].expand(-1, -1, k - kk)], -1)
    ok = torch.cat([ok, torch.zeros_like(ok[..., :1]).expand(-1, -1, k - kk)], -1)
  closest = d.argmin(-1, keepdim=True)
  idx = torch.where(ok, idx, closest.expand_as(idx))
  return idx, (~ok).to(points.dtype)


def FarthestPointSampler(points, padding, num_sampled_points, precomputed_squared_distance=None,
                         num_seeded_points=0, random_seed=None):
  """Farthest-first traversal (ref :244) → (sampled_idx `[N, S]`, closest_idx `[N, P1]`).
  The first point is random among the real ones (or the seeds are taken first);
  `closest_idx[n, p]` is the position in `sampled_idx` of the sample nearest to point p."""
  n, p1, _ = points.shape
  dev = points.device
  real = padding < 0.5
  gen = None
  if random_seed is not None:
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(random_seed))
  big = torch.finfo(points.dtype).max / 4
  min_d = torch.full((n, p1), big, device=dev, dtype=points.dtype)
  min_d = min_d.masked_fill(~real, -1.0)                    # padded points are never picked
  closest = torch.zeros(n, p1, dtype=torch.long, device=dev)
  sampled = torch.zeros(n, num_sampled_points, dtype=torch.long, device=dev)
  rows = torch.arange(n, device=dev)
  for s in range(num_sampled_points):
    if s < num_seeded_points:
      pick = torch.full((n,), s, dtype=torch.long, device=dev)
    elif s == 0:
      noise = torch.rand(n, p1, device=dev, generator=gen).masked_fill(~real, -1.0)
      pick = noise.argmax(-1)
    else:
      pick = min_d.argmax(-1)
    sampled[:, s] = pick
    if precomputed_squared_distance is not None:
      d = precomputed_squared_distance[rows, pick]
    else:
      d = (points - points[rows, pick].unsqueeze(1)).square().sum(-1)
    closer = (d < min_d) & real
    closest = torch.where(closer, torch.full_like(closest, s), closest)
    min_d = torch.where(closer, d, min_d)
    min_d[rows, pick] = torch.where(real[rows, pick], torch.zeros_like(d[:, 0]),
                                    min_d[rows, pick])
  return sampled, closest


def _SegmentReduce(values, segment_ids, num_segments, method):
  """values `[N, P, C]`, segment_ids `[N, P]` → `[N, num_segments, C]`."""
  n, p, c = values.shape
  idx = segment_ids.unsqueeze(-1).expand(n, p, c)
  if method == 'mean':
    out = torch.zeros(n, num_segments, c, device=values.device, dtype=values.dtype)
    out.scatter_add_(1, idx, values)
    cnt = torch.zeros(n, num_segments, 1, device=values.device, dtype=values.dtype)
    cnt.scatter_add_(1, segment_ids.unsqueeze(-1), torch.ones_like(values[..., :1]))
    return out / cnt.clamp_min(1.0)
  reduce = 'amax' if method == 'max' else 'amin'
  fill = torch.finfo(values.dtype).min if method == 'max' else torch.finfo(values.dtype).max
  out = torch.full((n, num_segments, c), fill, device=values.device, dtype=values.dtype)
  out.scatter_reduce_(1, idx, values, reduce=reduce, include_self=True)
  return torch.where(out == fill, torch.zeros_like(out), out)


def SegmentPool3D(points, point_features, pooling_idx, closest_idx, pooling_method='max'):
  """Pools features of the points assigned (by `closest_idx`) to each kept point
  (ref :525) → (pooled_points `[N, P2, 3]`, pooled_features `[N, P2, C]`)."""
  assert pooling_method in ('min', 'max', 'mean')
  p2 = pooling_idx.shape[1]
  pooled_points = points.gather(1, pooling_idx.unsqueeze(-1).expand(-1, -1, points.shape[-1]))
  return pooled_points, _SegmentReduce(point_features, closest_idx, p2, pooling_method)


def MaxPool3D(points, point_features, pooling_idx, closest_idx):
  """ref :457."""
  return SegmentPool3D(points, point_features, pooling_idx, closest_idx, 'max')


def WhereBroadcast(conditional, true_result, false_result):
  """`torch.where` with `conditional` broadcast over trailing dims (ref :590)."""
  c = conditional
  while c.dim() < true_result.dim():
    c = c.unsqueeze(-1)
  return torch.where(c.bool(), true_result, false_result)


def RavelIndex(coords, dims):
  """`[..., len(dims)]` integer coordinates → flat C-order index (ref :827)."""
  mult = []
  m = 1
  for d in reversed(list(dims)):
    mult.append(m)
    m *= int(d)
  mult = torch.tensor(list(reversed(mult)), device=coords.device, dtype=coords.dtype)
  return (coords * mult).sum(-1)


def DynamicVoxelization(points_xyz, points_padding, grid_size, grid_range_x, grid_range_y,
                        grid_range_z):
  """Point → voxel maps without materialising voxel tensors (ref :730):
  `coords [B,P,3]`, `centers [B,P,3]`, `indices [B,P]` (always valid, 0 when out of
  range), `padding [B,P]`, `num_voxels`."""
  lo = torch.tensor([grid_range_x[0], grid_range_y[0], grid_range_z[0]],
                    device=points_xyz.device, dtype=points_xyz.dtype)
  hi = torch.tensor([grid_range_x[1], grid_range_y[1], grid_range_z[1]],
                    device=points_xyz.device, dtype=points_xyz.dtype)
  gs = torch.tensor(list(grid_size), device=points_xyz.device, dtype=points_xyz.dtype)
  size = (hi - lo) / gs
  coords = torch.floor((points_xyz - lo) / size).long()
  inside = ((coords >= 0) & (coords < gs.long())).all(-1)
  padding = ((points_padding > 0.5) | ~inside).to(points_xyz.dtype)
  centers = (coords.to(points_xyz.dtype) + 0.5) * size + lo
  indices = RavelIndex(coords, grid_size)
  indices = torch.where(padding > 0.5, torch.zeros_like(indices), indices)
  n_vox = 1
  for g in grid_size:
    n_vox *= int(g)
  return NestedMap(coords=coords, centers=centers, indices=indices, padding=padding,
                   num_voxels=n_vox)


def _BatchedUnsortedSegmentFn(batched_data, batched_segment_ids, num_segments, method='sum',
                              batched_padding=None):
  """Per-batch-row segment reduction `[B, P, C]` → `[B, num_segments, C]` (ref :848)."""
  if batched_padding is not None:
    w = (1.0 - batched_padding).unsqueeze(-1)
    if method in ('sum', 'mean'):
      batched_data = batched_data * w
  b, p, c = batched_data.shape
  idx = batched_segment_ids.unsqueeze(-1).expand(b, p, c)
  if method == 'sum':
    out = torch.zeros(b, num_segments, c, device=batched_data.device, dtype=batched_data.dtype)
    return out.scatter_add_(1, idx, batched_data)
  if method == 'mean':
    out = torch.zeros(b, num_segments, c, device=batched_data.device, dtype=batched_data.dtype)
    out.scatter_add_(1, idx, batched_data)
    ones = torch.ones_like(batched_data[..., :1]) if batched_padding is None else w
    cnt = torch.zeros(b, num_segments, 1, device=batched_data.device, dtype=batched_data.dtype)
    cnt.scatter_add_(1, batched_segment_ids.unsqueeze(-1), ones)
    return out / cnt.clamp_min(1.0)
  return _SegmentReduce(batched_data, batched_segment_ids, num_segments, method)


def DynamicVoxelStatistics(points_xyz, dynamic_voxels):
  """Per-point voxel statistics (ref :626): `centroids [B,P,3]` (mean xyz of the point's
  voxel), `covariance [B,P,9]`, `centered_xyz [B,P,3]`, `num_points [B,P,1]`."""
  dv = dynamic_voxels
  w = (1.0 - dv.padding).unsqueeze(-1)
  ones = w
  seg = dv.indices
  nv = dv.num_voxels
  cnt = _BatchedUnsortedSegmentFn(ones, seg, nv, 'sum')
  mean = _BatchedUnsortedSegmentFn(points_xyz * w, seg, nv, 'sum') / cnt.clamp_min(1.0)
  gather = lambda t: t.gather(1, seg.unsqueeze(-1).expand(-1, -1, t.shape[-1]))
  centroids = gather(mean)
  centered = (points_xyz - centroids) * w
  outer = (centered.unsqueeze(-1) * centered.unsqueeze(-2)).flatten(-2)
  cov = _BatchedUnsortedSegmentFn(outer, seg, nv, 'sum') / (cnt - 1.0).clamp_min(1.0)
  return NestedMap(centroids=centroids * w, covariance=gather(cov) * w, centered_xyz=centered,
                   num_points=gather(cnt) * w)


def LocalTransform(points, bboxes_3d):
  """Points `[..., 3]` into the frame of their assigned boxes `[..., 7]`: translate to the
  box centre, rotate by −phi (ref :948)."""
  rel = points - bboxes_3d[..., :3]
  c, s = torch.cos(bboxes_3d[..., 6]), torch.sin(bboxes_3d[..., 6])
  return torch.stack([rel[..., 0] * c + rel[..., 1] * s, -rel[..., 0] * s + rel[..., 1] * c,
                      rel[..., 2]], -1)


def GenerateCenternessLabel(points, assigned_gt_bboxes, centerness_range, ignore_z=False,
                            epsilon=1e-6):
  """FCOS centerness of each point inside its assigned box, rescaled into
  `centerness_range = (lo, hi)` (ref :984): sqrt-free product of min/max ratios along each
  local axis, raised to 1/#axes."""
  local = LocalTransform(points, assigned_gt_bboxes)
  half = assigned_gt_bboxes[..., 3:6] / 2
  dims = 2 if ignore_z else 3
  near = (half[..., :dims] - local[..., :dims].abs()).clamp_min(0.0)
  far = half[..., :dims] + local[..., :dims].abs()
  ratio = (near / far.clamp_min(epsilon)).clamp(0.0, 1.0)
  centerness = ratio.prod(-1).clamp_min(0.0) ** (1.0 / dims)
  lo, hi = centerness_range
  return lo + centerness * (hi - lo)


def ComputeFeatureRatio(images, image_features):
  """(height ratio, width ratio) between feature maps `[..., h, w, C]` and images
  `[..., H, W, 3]` (ref :1075)."""
  return (image_features.shape[-3] / images.shape[-3], image_features.shape[-2] / images.shape[-2])


def StackCameraImages(images, camera_names=None):
  """NestedMap camera → {image, …} → stacked `[B, num_cameras, H, W, 3]` (ref :1094)."""
  names = camera_names or sorted(images.keys())
  return torch.stack([images[n].image for n in names], 1)



