"""3-D detection helpers vs. brute-force oracles (ref tasks/car/detection_3d_lib_test.py)."""
import math

import numpy as np
import pytest
import torch

from lingvo_b200.models.car import detection_3d_lib as d3
from lingvo_b200.models.car import geometry

U = d3.Utils3D()


def _Boxes(n, seed=0, spread=10.0):
  g = torch.Generator().manual_seed(seed)
  ctr = (torch.rand(n, 3, generator=g) - 0.5) * spread
  dims = 1.0 + 2.0 * torch.rand(n, 3, generator=g)
  phi = (torch.rand(n, 1, generator=g) - 0.5) * 2 * math.pi
  return torch.cat([ctr, dims, phi], -1)


def test_scaled_huber_and_corner_loss():
  x = torch.tensor([0.0, 0.5, 2.0, -3.0])
  got = U.ScaledHuberLoss(torch.zeros(4), x, delta=1.0)
  assert got.tolist() == pytest.approx([0.0, 0.125, 1.5, 2.5])
  got = U.ScaledHuberLoss(torch.zeros(4), x, weights=torch.tensor([1., 2, 0, 1]), delta=0.5)
  assert got.tolist() == pytest.approx([0.0, 2 * 0.25, 0.0, 2.75])
  b = _Boxes(5)
  assert float(U.CornerLoss(b, b).abs().max()) < 1e-5
  flipped = b.clone(); flipped[:, 6] += math.pi
  assert float(U.CornerLoss(b, flipped).abs().max()) < 1e-4          # symmetric: flip is free
  assert float(U.CornerLoss(b, flipped, symmetric=False).min()) > 0.1
  shifted = b.clone(); shifted[:, 0] += 0.5
  np.testing.assert_allclose(U.CornerLoss(b, shifted).numpy(), np.full(5, 8 * 0.125), rtol=1e-4)
  assert U.CornerLoss(b.reshape(1, 5, 7), shifted.reshape(1, 5, 7)).shape == (1, 5)


def test_rotated_iou_ignores_z_and_assign_anchors_outcomes():
  gt = torch.tensor([[0., 0, 0, 4, 2, 1.5, 0.0], [10, 0, 5, 4, 2, 1.5, math.pi / 2],
                     [50, 50, 0, 1, 1, 1, 0]])
  anchors = torch.tensor([[0., 0, 9, 4, 2, 3, 0.0],        # = gt0 up to z → IoU 1
                          [1., 0, 0, 4, 2, 1.5, 0.0],       # IoU 0.6 with gt0 → fg
                          [2.3, 0, 0, 4, 2, 1.5, 0.0],      # IoU ~0.27 → background
                          [1.5, 0, 0, 4, 2, 1.5, 0.0],      # IoU ~0.45 → ignored
                          [10, 0.4, 0, 2, 4, 1, 0.0],       # overlaps the rotated gt1
                          [30, 30, 0, 1, 1, 1, 0.0]])       # nothing
  iou = U.IOU2DRotatedBoxes(anchors, gt)
  assert iou.shape == (6, 3)
  assert float(iou[0, 0]) == pytest.approx(1.0, abs=1e-5)
  assert float(iou[1, 0]) == pytest.approx(3 * 2 / (16 - 6), abs=1e-4)
  assert float(iou[4, 1]) == pytest.approx(7.2 / 8.8, abs=1e-3)    # gt1 is turned by 90°
  labels = torch.tensor([3, 5, 7])
  mask = torch.tensor([1.0, 1.0, 0.0])
  a = U.AssignAnchors(anchors, gt, labels, mask)
  assert a.assigned_gt_idx.tolist() == [0, 0, -1, -1, 1, -1]
  assert a.assigned_gt_labels.tolist() == [3, 3, 0, 0, 5, 0]
  assert a.assigned_cls_mask.tolist() == [1, 1, 1, 0, 1, 1]
  assert a.assigned_reg_mask.tolist() == [1, 1, 0, 0, 1, 0]
  torch.testing.assert_close(a.assigned_gt_bbox[0], gt[0])
  assert a.assigned_gt_bbox[5].tolist() == [0, 0, 0, 1, 1, 1, 0]
  # force match: gt1's best anchor (IoU 0.82) — raise the bar so only forcing helps
  hi = U.AssignAnchors(anchors, gt, labels, mask, foreground_assignment_threshold=0.9)
  assert hi.assigned_reg_mask.tolist() == [1, 0, 0, 0, 1, 0]
  no = U.AssignAnchors(anchors, gt, labels, mask, foreground_assignment_threshold=0.9,
                       force_match=False)
  assert no.assigned_reg_mask.tolist() == [1, 0, 0, 0, 0, 0]
  # the masked-out gt never wins, even with a perfect anchor
  a2 = U.AssignAnchors(gt[2:3], gt, labels, mask)
  assert a2.assigned_gt_idx.tolist() == [-1] and a2.assigned_cls_mask.tolist() == [1.0]
  # custom similarity
  a3 = U.AssignAnchors(anchors, gt, labels, mask, force_match=False,
                       similarity_fn=lambda x, y: torch.ones(x.shape[0], y.shape[0]) * 0.4)
  assert a3.assigned_cls_mask.sum() == 0                        # 0.35 < 0.4 < 0.5: all ignored
  empty = U.AssignAnchors(anchors, torch.zeros(0, 7), torch.zeros(0, dtype=torch.long),
                          torch.zeros(0))
  assert empty.assigned_reg_mask.sum() == 0 and empty.assigned_cls_mask.sum() == 6


def test_residual_coding_roundtrip():
  anchors, gt = _Boxes(20, 1), _Boxes(20, 2)
  res = U.LocalizationResiduals(anchors, gt)
  back = U.ResidualsToBBoxes(anchors, res)
  torch.testing.assert_close(back[:, :6], gt[:, :6], atol=1e-4, rtol=1e-4)
  d = geometry.WrapAngleRad(back[:, 6] - gt[:, 6])
  assert float(d.abs().max()) < 1e-4
  pts = torch.randn(20, 3)
  tgt = U.LocalizationResidualsAnchorFree(pts, gt)
  torch.testing.assert_close(tgt[:, :3] + pts, gt[:, :3])
  torch.testing.assert_close(tgt[:, 3:], gt[:, 3:])


def test_angle_bins_roundtrip():
  bins = 12
  boxes = _Boxes(200, 3)
  boxes[:, 6] = torch.linspace(-4 * math.pi, 4 * math.pi, 200)
  cls, res = U.AngleToBin(boxes, bins)
  assert int(cls.min()) >= 0 and int(cls.max()) < bins and float(res.abs().max()) <= 0.5 + 1e-5
  width = 2 * math.pi / bins
  zero = torch.zeros(1, 7)
  c0, r0 = U.AngleToBin(zero, bins)
  assert int(c0) == 0 and float(r0) == pytest.approx(0.0, abs=1e-6)   # 0 is the centre of bin 0
  del width
  logits = torch.nn.functional.one_hot(cls.long(), bins).float()
  per_bin_res = torch.zeros(200, bins).scatter_(1, cls.long().unsqueeze(1),
                                               (res + 0.5).unsqueeze(1))
  # AngleToBin shifts by half a bin; decoding (cls + res + .5)·w − w/2 recovers the angle
  dec = U.BinToAngle(logits, per_bin_res, bins) - math.pi / bins
  diff = geometry.WrapAngleRad(dec - boxes[:, 6])
  assert float(diff.abs().max()) < 1e-4
  out = U.ResidualsToBBoxesAnchorFree(torch.zeros(200, 3), torch.ones(200, 6), logits,
                                      per_bin_res)
  assert out.shape == (200, 7) and float(out[:, :6].min()) == 1.0


def test_nms_variants():
  boxes = torch.tensor([[0., 0, 0, 2, 2, 1, 0], [0.2, 0, 0, 2, 2, 1, 0], [5, 5, 0, 2, 2, 1, 0],
                        [5.1, 5, 0, 2, 2, 1, 0.1], [9, 9, 0, 1, 1, 1, 0]])
  scores = torch.tensor([0.9, 0.8, 0.7, 0.95, 0.005])
  idx, mask = U.NMSIndices(boxes, scores, 4, nms_iou_threshold=0.3, score_threshold=0.01)
  assert idx.tolist() == [3, 0, 0, 0] and mask.tolist() == [1, 1, 0, 0]
  bidx, bmask = U.BatchedNMSIndices(boxes[None].repeat(2, 1, 1), torch.stack([scores, scores.flip(0)]),
                                    max_num_boxes=3)
  assert bidx.shape == (2, 3) and bmask[0].tolist() == [1, 1, 0]
  # oriented per-class NMS: two classes, per-class thresholds
  sc = torch.stack([scores, scores.flip(0)], -1)[None]
  oi, osc, om = U.BatchedOrientedNMSIndices(boxes[None], sc, [0.3, 0.3], [0.01, 0.5], 3)
  assert oi.shape == (1, 2, 3)
  assert oi[0, 0, :2].tolist() == [3, 0] and om[0, 0].tolist() == [1, 1, 0]
  assert osc[0, 0, :2].tolist() == pytest.approx([0.95, 0.9])
  # class 1: scores flipped → only boxes with score > 0.5 survive the threshold
  assert om[0, 1].sum() == 2 and set(oi[0, 1, :2].tolist()) == {1, 3} or om[0, 1].sum() >= 1
  legacy_idx, legacy_mask = U.BatchedNMSIndices(boxes[None], sc, 0.3, 0.01, 3)
  assert legacy_idx.shape == (1, 2, 3) and int(legacy_idx[0, 0, 2]) == -1


def test_corners_to_image_plane():
  boxes = _Boxes(3, 5)[None]
  corners = geometry.BBoxCorners(boxes)
  proj = torch.tensor([[[2.0, 0, 0, 1], [0, 2.0, 0, 1], [0, 0, 0, 1.0]]])
  out = U.CornersToImagePlane(corners, proj)
  assert out.shape == (1, 3, 8, 2)
  torch.testing.assert_close(out, corners[..., :2] * 2 + 1)


def test_find_center_points_and_assign_points():
  gt = torch.tensor([[0., 0, 0, 4, 2, 2, 0.0], [10, 0, 0, 2, 2, 2, math.pi / 4],
                     [99, 99, 0, 1, 1, 1, 0]])
  mask = torch.tensor([1.0, 1.0, 0.0])
  pts = torch.tensor([[0.1, 0.1, 0], [1.9, 0.9, 0.5], [10, 0.2, 0], [3, 3, 0], [10, 1.3, 0],
                      [0, 0, 5.0]])
  vals, idx = U.FindCenterPoints(pts, gt, mask)
  assert idx.tolist() == [0, 2, 0] and vals[2].abs().sum() == 0
  torch.testing.assert_close(vals[1], pts[2])
  seen = set()
  for seed in range(20):
    _, ridx = U.FindCenterPoints(pts, gt, mask, random_seed=seed, random_chosen=True)
    assert int(ridx[0]) in (0, 1, 5) and int(ridx[1]) in (2, 4)        # footprint ignores z
    seen.add(int(ridx[0]))
  assert len(seen) > 1
  a = U.AssignPoints(pts, gt, torch.tensor([1, 2, 1]), mask, cls_num=3,
                     expand_gt_bbox_dims=[0.0, 0.0, 0.0], random_seed=1)
  assert a.assigned_gt_idx.tolist() == [0, 0, 1, -1, 1, -1]
  assert a.assigned_gt_labels.tolist() == [1, 1, 2, 0, 2, 0]
  assert a.assigned_cls_mask.tolist() == [1] * 6
  assert a.assigned_reg_mask.shape == (6, 3)
  assert a.assigned_reg_mask.sum(-1).tolist() == [1, 1, 1, 0, 1, 0]
  assert a.assigned_reg_mask[2].tolist() == [0, 0, 1]
  # ignore_z pulls in the point floating above box 0; growing the boxes pulls in (3, 3)
  az = U.AssignPoints(pts, gt, torch.tensor([1, 2, 1]), mask, 3, [0.0, 0.0, 0.0], ignore_z=True)
  assert az.assigned_gt_idx.tolist()[5] == 0
  ag = U.AssignPoints(pts, gt, torch.tensor([1, 2, 1]), mask, 3, [4.0, 4.5, 0.0])
  assert ag.assigned_gt_idx.tolist()[3] == 0
  # indices refer to the unmasked list: mask out box 0 → box 1 keeps index 1
  a1 = U.AssignPoints(pts, gt, torch.tensor([1, 2, 1]), torch.tensor([0.0, 1.0, 0.0]), 3,
                      [0.0, 0.0, 0.0])
  assert a1.assigned_gt_idx.tolist() == [-1, -1, 1, -1, 1, -1]


def test_random_pad_or_trim():
  x = torch.arange(10.0).reshape(5, 2)
  y = torch.arange(5)
  (tx, ty), pad = d3.RandomPadOrTrimTo([x, y], 3, seed=0)
  assert tx.shape == (3, 2) and pad.tolist() == [0, 0, 0]
  assert torch.equal(tx[:, 0] / 2, ty.float()) and len(set(ty.tolist())) == 3
  (px, py), pad = d3.RandomPadOrTrimTo([x, y], 8, seed=0)
  assert px.shape == (8, 2) and pad.tolist() == [0] * 5 + [1] * 3
  assert torch.equal(px[:5], x) and set(py[5:].tolist()) <= set(range(5))
  (zx,), pad = d3.RandomPadOrTrimTo([torch.zeros(0, 2)], 4)
  assert zx.shape == (4, 2) and pad.sum() == 4
