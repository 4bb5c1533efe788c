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
                  center_selector='farthest', random_seed=-1, num_seeded_points=0,
                  neighbor_sampler='closest', neighbor_algorithm='auto',
                  center_z_min=None, center_z_max=None):
  """Centre selection (farthest-point or uniform) + neighbourhood gathering for a batch of
  scenes `[B, P, ≥3]` → (center `[B, M]`, center_padding, indices `[B, M, K]`,
  indices_padding); ref `car_ops.cc:189-255` / `ps_utils.cc`. Runs in the native host
  library, scenes in parallel.

  The first `num_seeded_points` points seed the farthest-point criterion but are never
  returned; only points with z in [center_z_min, center_z_max] can be centres;
  `neighbor_sampler` 'closest' keeps the K nearest points within `max_distance`, 'uniform'
  a uniform sample of them; `neighbor_algorithm='hash'` forces the grid-hash ball query.
  """
  big = 3.4e38
  pts = np.ascontiguousarray(_F(points), np.float32)
  pad = np.ascontiguousarray(_F(points_padding), np.float32)
  c, cp, idx, ip = _ops.host().sample_points(
      pts, pad, int(num_seeded_points), center_selector, neighbor_sampler, neighbor_algorithm,
      int(num_centers), -big if center_z_min is None else float(center_z_min),
      big if center_z_max is None else float(center_z_max), int(num_neighbors),
      big if max_distance is None else float(max_distance), int(random_seed))
  return (torch.from_numpy(np.asarray(c)).long(), torch.from_numpy(np.asarray(cp)),
          torch.from_numpy(np.asarray(idx)).long(), torch.from_numpy(np.asarray(ip)))



def average_precision2d(iou_threshold, groundtruth_bbox, groundtruth_imageid,  # pylint: disable=invalid-name
                        groundtruth_ignore, prediction_bbox, prediction_imageid,
                        prediction_ignore, prediction_score, num_recall_points=1,
                        algorithm='VOC'):
  """Image-plane AP over `(ymin, xmin, ymax, xmax)` boxes with the same matching protocol as
  `average_precision3d` (ref `image_metrics.cc` `AveragePrecision<Box2D>`)."""
  ap, pr, sh = _ops.host().average_precision_2d(
      float(iou_threshold), _F(groundtruth_bbox).reshape(-1, 4), _I(groundtruth_imageid),
      _I(groundtruth_ignore), _F(prediction_bbox).reshape(-1, 4), _I(prediction_imageid),
      _I(prediction_ignore), _F(prediction_score), int(num_recall_points), algorithm)
  return float(ap), torch.from_numpy(np.asarray(pr)), torch.from_numpy(np.asarray(sh))
