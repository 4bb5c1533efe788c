"""Python front-end of the native 3-D detection ops (ref `lingvo/tasks/car/ops/__init__.py`;
C++ in `lingvo_b200/ops/csrc_host/car_ops.cpp`). Inputs may be torch tensors or arrays;
outputs are torch tensors on the CPU (these ops sit in the input pipeline and in decode
post-processing)."""

from __future__ import annotations

import numpy as np
import torch

from lingvo_b200 import ops as _ops


def _F(x):
  x = x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
  return np.ascontiguousarray(x, np.float32)


def _I(x):
  x = x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
  return np.ascontiguousarray(x, np.int32)


def pairwise_iou3d(boxes_a, boxes_b):  # pylint: disable=invalid-name
  """`[N, 7]`, `[M, 7]` → IoU `[N, M]` (exact rotated-box intersection)."""
  a, b = _F(boxes_a).reshape(-1, 7), _F(boxes_b).reshape(-1, 7)
  if a.shape[0] == 0 or b.shape[0] == 0:
    return torch.zeros(a.shape[0], b.shape[0])
  return torch.from_numpy(_ops.host().pairwise_iou_3d(a, b))


def non_max_suppression_3d(bboxes, scores, nms_iou_threshold, score_threshold,  # pylint: disable=invalid-name
                           max_boxes_per_class):
  """Per-class greedy NMS: bboxes `[N, 7]`, scores `[N, C]`, per-class thresholds →
  (indices `[C, max_boxes]` padded with −1, mask `[C, max_boxes]`)."""
  s = _F(scores)
  if s.ndim == 1:
    s = s[:, None]
  c = s.shape[1]
  as_list = lambda v: [float(v)] * c if np.isscalar(v) else [float(x) for x in v]
  idx = _ops.host().nms_3d(_F(bboxes).reshape(-1, 7), s, as_list(nms_iou_threshold),
                           as_list(score_threshold), int(max_boxes_per_class))
  idx = torch.from_numpy(np.asarray(idx)).long()
  return idx, (idx >= 0).float()


def average_precision3d(iou_threshold, groundtruth_bbox, groundtruth_imageid,  # pylint: disable=invalid-name
                        groundtruth_ignore, prediction_bbox, prediction_imageid,
                        prediction_ignore, prediction_score, num_recall_points=1,
                        algorithm='KITTI'):
  """→ (AP scalar, precision_recall `[num_recall_points, 2]`, score_and_hit `[M, 2]`)."""
  ap, pr, sh = _ops.host().average_precision_3d(
      float(iou_threshold), _F(groundtruth_bbox).reshape(-1, 7), _I(groundtruth_imageid),
      _I(groundtruth_ignore), _F(prediction_bbox).reshape(-1, 7), _I(prediction_imageid),
      _I(prediction_ignore), _F(prediction_score), int(num_recall_points), algorithm)
  return float(ap), torch.from_numpy(np.asarray(pr)), torch.from_numpy(np.asarray(sh))


def point_to_grid(points, x_range, y_range, grid_size, max_pillars, points_per_pillar):  # pylint: disable=invalid-name
  """Pillar bucketing of one scene: → (pillar_points `[P, K, D]`, pillar_xy `[P, 2]`,
  pillar_count `[P]`, num_occupied)."""
  pts = _F(points)
  out = _ops.host().points_to_pillars(pts, float(x_range[0]), float(x_range[1]),
                                      float(y_range[0]), float(y_range[1]), int(grid_size[0]),
                                      int(grid_size[1]), int(max_pillars), int(points_per_pillar))
  pp, xy, cnt, n = out
  return (torch.from_numpy(np.asarray(pp)), torch.from_numpy(np.asarray(xy)).long(),
          torch.from_numpy(np.asarray(cnt)).long(), int(n))


def sample_points(points, points_padding, num_centers, num_neighbors, max_distance=None,  # pylint: disable=invalid-name
                  center_selector='farthest', random_seed=-1):
  """Centre selection (farthest-point or uniform) + neighbour gathering for one batch of
  scenes `[B, P, 3]` → (center `[B, C]`, center_padding, indices `[B, C, K]`,
  indices_padding)."""
  from lingvo_b200.models.car import car_lib  # pylint: disable=g-import-not-at-top
  pts = torch.as_tensor(_F(points))
  pad = torch.as_tensor(_F(points_padding))
  b, p, _ = pts.shape
  if center_selector == 'farthest':
    centers = []
    for i in range(b):
      real = np.nonzero(pad[i].numpy() < 0.5)[0]
      k = min(num_centers, len(real))
      if k == 0:
        centers.append(np.zeros(num_centers, np.int64))
        continue
      loc = np.asarray(_ops.host().farthest_point_sample(pts[i, real].numpy(), k))
      idx = real[loc]
      centers.append(np.concatenate([idx, np.repeat(idx[:1], num_centers - k)]))
    center = torch.from_numpy(np.stack(centers)).long()
  else:
    g = torch.Generator()
    if random_seed >= 0:
      g.manual_seed(int(random_seed))
    noise = torch.rand(b, p, generator=g).masked_fill(pad > 0.5, -1.0)
    center = noise.topk(num_centers, -1).indices
  n_real = (pad < 0.5).sum(1, keepdim=True)
  center_padding = (torch.arange(num_centers).unsqueeze(0) >= n_real).float()
  q = pts.gather(1, center.unsqueeze(-1).expand(-1, -1, 3))
  idx, idx_pad = car_lib.NeighborhoodIndices(pts, q, num_neighbors, pad > 0.5, max_distance)
  return center, center_padding, idx, idx_pad
