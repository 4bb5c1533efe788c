"""3-D detection utilities (ref `lingvo/tasks/car/detection_3d_lib.py`, `geometry.py`).

7-DOF boxes `(x, y, z, dx, dy, dz, phi)`. Anchor grids, SECOND/PointPillars
residual encoding (ref :700-800), IoU-based anchor assignment (ref :300-560),
focal + smooth-L1 losses, oriented NMS decode (native `_H.nms_3d`).
"""

from __future__ import annotations

import math

import numpy as np
import torch

from lingvo_b200 import ops


class Utils3D:

  def CreateDenseCoordinates(self, ranges, center_in_cell=False):
    """ranges: [(min, max, num)] per dim → [prod(num), ndims] grid of centres."""
    axes = []
    for lo, hi, n in ranges:
      if center_in_cell:
        step = (hi - lo) / n
        axes.append(torch.linspace(lo + step / 2, hi - step / 2, n))
      else:
        axes.append(torch.linspace(lo, hi, n))
    grid = torch.meshgrid(*axes, indexing='ij')
    return torch.stack([g.reshape(-1) for g in grid], -1)

  def MakeAnchorBoxes(self, anchor_centers, anchor_box_dimensions, anchor_box_offsets,
                      anchor_box_rotations):
    """centers [N,3] × A templates → [N, A, 7]."""
    dims = torch.as_tensor(anchor_box_dimensions, dtype=torch.float32)     # [A,3]
    offs = torch.as_tensor(anchor_box_offsets, dtype=torch.float32)        # [A,3]
    rots = torch.as_tensor(anchor_box_rotations, dtype=torch.float32)      # [A]
    n, a = anchor_centers.shape[0], dims.shape[0]
    ctr = anchor_centers.unsqueeze(1) + offs.unsqueeze(0)
    return torch.cat([ctr, dims.unsqueeze(0).expand(n, a, 3),
                      rots.view(1, a, 1).expand(n, a, 1)], -1)

  def LocalizationResiduals(self, anchor_bboxes, assigned_gt_bboxes):
    """SECOND encoding: Δxy by the anchor diagonal, Δz by height, log size ratios, Δφ."""
    xa, ya, za, dxa, dya, dza, pa = anchor_bboxes.unbind(-1)
    xg, yg, zg, dxg, dyg, dzg, pg = assigned_gt_bboxes.unbind(-1)
    diag = torch.sqrt(dxa ** 2 + dya ** 2)
    eps = 1e-8
    return torch.stack([
        (xg - xa) / diag, (yg - ya) / diag, (zg - za) / dza,
        torch.log(dxg.clamp_min(eps) / dxa), torch.log(dyg.clamp_min(eps) / dya),
        torch.log(dzg.clamp_min(eps) / dza), pg - pa], -1)

  def ResidualsToBBoxes(self, anchor_bboxes, residuals, min_angle_rad=-math.pi,
                        max_angle_rad=math.pi):
    xa, ya, za, dxa, dya, dza, pa = anchor_bboxes.unbind(-1)
    rx, ry, rz, rdx, rdy, rdz, rp = residuals.unbind(-1)
    diag = torch.sqrt(dxa ** 2 + dya ** 2)
    phi = pa + rp
    span = max_angle_rad - min_angle_rad
    phi = torch.remainder(phi - min_angle_rad, span) + min_angle_rad
    return torch.stack([rx * diag + xa, ry * diag + ya, rz * dza + za,
                        torch.exp(rdx) * dxa, torch.exp(rdy) * dya, torch.exp(rdz) * dza, phi], -1)

  def AssignAnchors(self, anchor_bboxes, gt_bboxes, gt_bboxes_labels, gt_bboxes_mask,
                    foreground_assignment_threshold=0.5, background_assignment_threshold=0.35):
    """Per-example assignment (numpy/native IoU): returns dict of tensors over anchors:
    assigned_gt_idx, assigned_gt_bbox, assigned_gt_labels, assigned_cls_mask (1 = use in
    the classification loss), assigned_reg_mask (1 = foreground)."""
    a = anchor_bboxes.detach().cpu().numpy().astype(np.float32)
    g = gt_bboxes.detach().cpu().numpy().astype(np.float32)
    mask = gt_bboxes_mask.detach().cpu().numpy() > 0
    n = a.shape[0]
    iou = ops.host().pairwise_iou_3d(a, g) if g.shape[0] else np.zeros((n, 0), np.float32)
    iou[:, ~mask] = -1.0
    best = iou.argmax(1) if g.shape[0] else np.zeros(n, np.int64)
    best_iou = iou.max(1) if g.shape[0] else np.full(n, -1.0, np.float32)
    fg = best_iou >= foreground_assignment_threshold
    bg = best_iou <= background_assignment_threshold
    # force-match: every real gt box owns its best anchor
    if g.shape[0]:
      for j in np.nonzero(mask)[0]:
        i = int(iou[:, j].argmax())
        if iou[i, j] > 0:
          fg[i], bg[i], best[i] = True, False, j
    idx = torch.from_numpy(np.where(fg, best, -1).astype(np.int64))
    labels = torch.where(idx >= 0, gt_bboxes_labels.cpu()[idx.clamp_min(0)],
                         torch.zeros_like(idx))
    bbox = torch.where((idx >= 0).unsqueeze(-1), gt_bboxes.cpu()[idx.clamp_min(0)],
                       anchor_bboxes.cpu())
    return dict(assigned_gt_idx=idx, assigned_gt_bbox=bbox, assigned_gt_labels=labels,
                assigned_cls_mask=torch.from_numpy((fg | bg).astype(np.float32)),
                assigned_reg_mask=torch.from_numpy(fg.astype(np.float32)))

  def SigmoidFocalLoss(self, logits, one_hot, alpha=0.25, gamma=2.0):
    p = torch.sigmoid(logits)
    ce = torch.nn.functional.binary_cross_entropy_with_logits(logits, one_hot, reduction='none')
    pt = p * one_hot + (1 - p) * (1 - one_hot)
    w = alpha * one_hot + (1 - alpha) * (1 - one_hot)
    return w * (1 - pt) ** gamma * ce

  def ScaledHuberLoss(self, labels, predictions, delta=1.0 / 9.0):
    d = (predictions - labels).abs()
    return torch.where(d < delta, 0.5 * d * d / delta, d - 0.5 * delta)

  def CornerLoss(self, gt_bboxes, predicted_bboxes):
    """Smooth-L1 between the 8 corners (min over the heading flip)."""
    def corners(b):
      x, y, z, dx, dy, dz, phi = b.unbind(-1)
      c, s = torch.cos(phi), torch.sin(phi)
      out = []
      for sx in (-0.5, 0.5):
        for sy in (-0.5, 0.5):
          for sz in (-0.5, 0.5):
            out.append(torch.stack([x + sx * dx * c - sy * dy * s,
                                    y + sx * dx * s + sy * dy * c, z + sz * dz], -1))
      return torch.stack(out, -2)
    flipped = gt_bboxes.clone()
    flipped[..., 6] = flipped[..., 6] + math.pi
    cp = corners(predicted_bboxes)
    l1 = self.ScaledHuberLoss(corners(gt_bboxes), cp, 1.0).sum((-1, -2))
    l2 = self.ScaledHuberLoss(corners(flipped), cp, 1.0).sum((-1, -2))
    return torch.minimum(l1, l2)

  def BatchedNMSIndices(self, bboxes, scores, nms_iou_threshold=0.3, score_threshold=0.01,
                        max_num_boxes=None):
    """bboxes [B,N,7], scores [B,N,C] → indices [B,C,K] (−1 padded) + mask."""
    b, n, _ = bboxes.shape
    c = scores.shape[-1]
    k = max_num_boxes or n
    thr = nms_iou_threshold if isinstance(nms_iou_threshold, (list, tuple)) else [nms_iou_threshold] * c
    sthr = score_threshold if isinstance(score_threshold, (list, tuple)) else [score_threshold] * c
    out = []
    for i in range(b):
      out.append(ops.host().nms_3d(bboxes[i].detach().cpu().numpy().astype(np.float32),
                                   scores[i].detach().cpu().numpy().astype(np.float32),
                                   list(thr), list(sthr), k))
    idx = torch.from_numpy(np.stack(out)).long()
    return idx, (idx >= 0).float()
