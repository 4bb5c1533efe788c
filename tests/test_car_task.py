"""car task: native geometry ops + a tiny PointPillars train/decode (CPU)."""

import math

import numpy as np
import torch

from lingvo_b200 import model_registry
from lingvo_b200 import ops
from lingvo_b200.models.car import detection_3d_lib
import lingvo_b200.models.car.params.kitti  # noqa: F401


def test_rotated_iou_and_nms():
  h = ops.host()
  a = np.array([[0, 0, 0, 2, 2, 2, 0], [0, 0, 0, 2, 2, 2, math.pi / 4], [1, 0, 0, 2, 2, 2, 0],
                [10, 10, 0, 2, 2, 2, 0], [0, 0, 5, 2, 2, 2, 0]], np.float32)
  iou = h.pairwise_iou_3d(a, a)
  assert abs(iou[0, 1] - (8 * (math.sqrt(2) - 1)) / (8 - 8 * (math.sqrt(2) - 1))) < 1e-4
  assert abs(iou[0, 2] - 1.0 / 3.0) < 1e-5        # half overlap: 4 / (8 + 8 - 4)
  assert iou[0, 3] == 0 and iou[0, 4] == 0
  scores = np.array([[0.9], [0.8], [0.7], [0.6], [0.01]], np.float32)
  keep = h.nms_3d(a, scores, [0.3], [0.05], 4)
  assert keep[0].tolist() == [0, 3, -1, -1]       # 1 and 2 overlap box 0; 4 is below threshold


def test_pillars_and_fps():
  h = ops.host()
  pts = np.array([[0.1, 0.1, 0, 1], [0.2, 0.3, 0, 1], [3.5, 3.5, 0, 1], [9, 9, 0, 1]], np.float32)
  pp, xy, cnt, used = h.points_to_pillars(pts, 0, 4, 0, 4, 4, 4, 8, 3)
  assert used == 2 and cnt[:2].tolist() == [2, 1]
  assert xy[0].tolist() == [0, 0] and xy[1].tolist() == [3, 3]
  np.testing.assert_allclose(pp[0, 1], pts[1])
  idx = h.farthest_point_sample(np.array([[0, 0, 0], [1, 0, 0], [10, 0, 0], [5, 0, 0]], np.float32), 3)
  assert idx == [0, 2, 3]


def test_residual_roundtrip():
  u = detection_3d_lib.Utils3D()
  anchors = torch.tensor([[0., 0, 0, 4, 2, 1.5, 0.0], [5, 5, 0, 4, 2, 1.5, 1.57]])
  gt = torch.tensor([[0.5, -0.3, 0.1, 4.2, 1.9, 1.6, 0.2], [5.5, 4.0, -0.2, 3.8, 2.1, 1.4, 1.3]])
  res = u.LocalizationResiduals(anchors, gt)
  torch.testing.assert_close(u.ResidualsToBBoxes(anchors, res), gt, atol=1e-5, rtol=1e-5)


def test_point_pillars_tiny_trains_and_decodes():
  cfg = model_registry.GetParams('car.kitti.PointPillarsCarTiny', 'Train')
  model = cfg.Instantiate()
  task = model.tasks[0]
  losses = []
  for _ in range(12):
    m, _ = task.TrainStep()
    losses.append(float(m['loss'][0]))
  assert np.isfinite(losses).all()
  assert min(losses[-3:]) < losses[0]
  out = task.Decode(task.input.GetPreprocessedInputBatch())
  assert out.per_class_indices.shape[1] == 2 and out.per_class_predicted_bboxes.shape[-1] == 7
