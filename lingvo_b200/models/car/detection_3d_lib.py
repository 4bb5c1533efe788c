"""3-D detection utilities (ref `lingvo/tasks/car/detection_3d_lib.py`).

7-DOF boxes `(x, y, z, dx, dy, dz, phi)`. Anchor grids and SECOND / PointPillars residual
coding (ref :453-615), similarity-based anchor assignment with force matching (ref :262),
point → box assignment and centre-point search for anchor-free heads (ref :1024-1286),
angle-bin coding (ref :817-920), Huber / corner losses (ref :57-142), axis-aligned and
oriented NMS (ref :617-778; the oriented kernel and rotated IoU are the native `_H` ops).

The assignment helpers run per example inside input preprocessing (host tensors); the
losses / decoders are device code.
"""

from __future__ import annotations

import math

import numpy as np
import torch

from lingvo_b200 import ops
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.car import geometry


def _Np(x, dtype=np.float32):
  return (x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)).astype(dtype)


class Utils3D:
  """Helper routines for 3D detection problems."""

  # ------------------------------------------------------------------------------ losses --
  def ScaledHuberLoss(self, labels, predictions, weights=1.0, delta=1.0):
    """(1/δ)·½x² for |x| ≤ δ, |x| − ½δ beyond; x = labels − predictions (ref :57).
    `delta` moves the quadratic bowl without changing the linear tails."""
    d = (predictions - labels).abs()
    loss = torch.where(d <= delta, 0.5 * d * d / delta, d - 0.5 * delta)
    return loss * weights

  def SigmoidFocalLoss(self, logits, one_hot, alpha=0.25, gamma=2.0):
    p = torch.sigmoid(logits)
    ce = torch.nn.functional.binary_cross_entropy_with_logits(logits, one_hot, reduction='none')
    pt = p * one_hot + (1 - p) * (1 - one_hot)
    w = alpha * one_hot + (1 - alpha) * (1 - one_hot)
    return w * (1 - pt) ** gamma * ce

  def CornerLoss(self, gt_bboxes, predicted_bboxes, symmetric=True):
    """Huber loss between the 8 corners of predicted and ground-truth boxes `[..., 7]` →
    `[...]` (ref :93; Frustum-PointNets). `symmetric`: the smaller of the losses against
    the ground truth and the ground truth turned by π (heading-flip invariant)."""
    shape = gt_bboxes.shape[:-1]
    gt = gt_bboxes.reshape(-1, 7)
    pred = predicted_bboxes.reshape(-1, 7)
    pred_corners = geometry.BBoxCorners(pred)
    loss = self.ScaledHuberLoss(geometry.BBoxCorners(gt), pred_corners).sum((-2, -1))
    if symmetric:
      rot = gt.new_tensor([0., 0., 0., 0., 0., 0., math.pi])
      flipped = self.ScaledHuberLoss(geometry.BBoxCorners(gt + rot), pred_corners).sum((-2, -1))
      loss = torch.minimum(loss, flipped)
    return loss.reshape(shape)

  # ----------------------------------------------------------------------------- anchors --
  def CreateDenseCoordinates(self, ranges, center_in_cell=False):
    """ranges: [(min, max, num)] per dim → [prod(num), ndims] grid (ref :144); with
    `center_in_cell` the points sit in the middle of `num` equal cells."""
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
    """centers `[..., 3]` × A templates → `[..., A, 7]` (ref :185)."""
    dims = torch.as_tensor(anchor_box_dimensions, dtype=torch.float32)     # [A,3]
    offs = torch.as_tensor(anchor_box_offsets, dtype=torch.float32)        # [A,3]
    rots = torch.as_tensor(anchor_box_rotations, dtype=torch.float32)      # [A]
    lead = anchor_centers.shape[:-1]
    a = dims.shape[0]
    ctr = anchor_centers.unsqueeze(-2) + offs
    return torch.cat([ctr, dims.expand(*lead, a, 3),
                      rots.view(a, 1).expand(*lead, a, 1)], -1)

  def IOU2DRotatedBoxes(self, bboxes_u, bboxes_v):
    """Rotated bird's-eye-view IoU of every pair `[U, V]`; z is ignored (ref :234)."""
    def Flat(b):
      b = _Np(b)[:, :7].copy()
      b[:, 2], b[:, 5] = 0.0, 1.0
      return b
    u, v = Flat(bboxes_u), Flat(bboxes_v)
    if not u.shape[0] or not v.shape[0]:
      return torch.zeros(u.shape[0], v.shape[0])
    return torch.from_numpy(np.asarray(ops.host().pairwise_iou_3d(u, v)))

  def AssignAnchors(self, anchor_bboxes, gt_bboxes, gt_bboxes_labels, gt_bboxes_mask,
                    foreground_assignment_threshold=0.5, background_assignment_threshold=0.35,
                    background_class_id=0, force_match=True, similarity_fn=None):
    """SSD-style assignment of every anchor `[A, 7]` to its most similar ground-truth box
    `[G, 7]` (ref :262): score ≥ fg threshold → foreground; ≤ bg threshold → background;
    in between → ignored (cls mask 0). `force_match`: an anchor that is a ground-truth box's
    best match (score > 0) is foreground regardless of the threshold.

    Returns NestedMap(assigned_gt_idx [A] (−1: none), assigned_gt_bbox [A, 7],
    assigned_gt_similarity_score [A], assigned_gt_labels [A], assigned_cls_mask [A],
    assigned_reg_mask [A])."""
    similarity_fn = similarity_fn or self.IOU2DRotatedBoxes
    anchor_bboxes = anchor_bboxes.detach().cpu().float()
    gt_bboxes = gt_bboxes.detach().cpu().float()
    gt_mask = gt_bboxes_mask.detach().cpu().float()
    gt_labels = gt_bboxes_labels.detach().cpu()
    a, g = anchor_bboxes.shape[0], gt_bboxes.shape[0]
    assert anchor_bboxes.shape[1] == 7 and gt_bboxes.shape[1] == 7
    if g == 0:
      score = torch.zeros(a, 1)
      gt_bboxes = torch.zeros(1, 7)
      gt_mask = torch.zeros(1)
      gt_labels = torch.full((1,), background_class_id, dtype=gt_labels.dtype)
      g = 1
    else:
      score = similarity_fn(anchor_bboxes, gt_bboxes).float()
    assert tuple(score.shape) == (a, g)
    max_score, max_idx = score.max(1)
    forced = torch.zeros(a, dtype=torch.bool)
    if force_match:
      gt_best = score.max(0, keepdim=True).values
      matches = ((score == gt_best) & (score == max_score.unsqueeze(1)) & (score > 0) &
                 (gt_mask > 0).unsqueeze(0))
      forced = matches.any(1)
      max_idx = torch.where(forced, matches.int().argmax(1), max_idx)
    max_score = torch.where(gt_mask[max_idx] == 1, max_score, torch.zeros_like(max_score))
    bg = max_score <= background_assignment_threshold
    fg = max_score >= foreground_assignment_threshold
    if force_match:
      bg &= ~forced
      fg |= forced
    dummy = gt_bboxes.new_tensor([[0, 0, 0, 1, 1, 1, 0]])
    boxes = torch.cat([gt_bboxes, dummy], 0)
    labels = torch.cat([gt_labels, gt_labels.new_tensor([background_class_id])], 0)
    gather = torch.where(fg, max_idx, torch.full_like(max_idx, g))
    return NestedMap(
        assigned_gt_idx=torch.where(gather == g, torch.full_like(gather, -1), gather).int(),
        assigned_gt_bbox=boxes[gather],
        assigned_gt_similarity_score=max_score,
        assigned_gt_labels=labels[gather],
        assigned_cls_mask=(bg | fg).float(),
        assigned_reg_mask=fg.float())

  def LocalizationResiduals(self, anchor_bboxes, assigned_gt_bboxes):
    """SECOND coding (ref :453): Δxy over the anchor's ground diagonal, Δz over its height,
    log size ratios, Δφ."""
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
    """Inverse of `LocalizationResiduals`, heading wrapped into [min, max) (ref :540)."""
    xa, ya, za, dxa, dya, dza, pa = anchor_bboxes.unbind(-1)
    rx, ry, rz, rdx, rdy, rdz, rp = residuals.unbind(-1)
    diag = torch.sqrt(dxa ** 2 + dya ** 2)
    phi = geometry.WrapAngleRad(pa + rp, min_angle_rad, max_angle_rad)
    return torch.stack([rx * diag + xa, ry * diag + ya, rz * dza + za,
                        torch.exp(rdx) * dxa, torch.exp(rdy) * dya, torch.exp(rdz) * dza, phi], -1)

  # --------------------------------------------------------------------------------- NMS --
  def NMSIndices(self, bboxes, scores, max_output_size, nms_iou_threshold=0.3,
                 score_threshold=0.01):
    """Axis-aligned bird's-eye-view NMS of `[N, 7]` boxes with scores `[N]` (ref :617): the
    heading is ignored (boxes are their xy extents). → (indices `[max_output_size]`, 0 where
    padded; mask with 1 for real picks)."""
    assert bboxes.shape[-1] == 7
    b = bboxes.detach()
    x1, x2 = b[:, 0] - b[:, 3] / 2, b[:, 0] + b[:, 3] / 2
    y1, y2 = b[:, 1] - b[:, 4] / 2, b[:, 1] + b[:, 4] / 2
    area = (x2 - x1) * (y2 - y1)
    s = scores.detach()
    order = torch.argsort(s, descending=True, stable=True)
    order = order[s[order] > score_threshold]
    keep = []
    while order.numel() and len(keep) < max_output_size:
      i = order[0]
      keep.append(int(i))
      rest = order[1:]
      iw = (torch.minimum(x2[i], x2[rest]) - torch.maximum(x1[i], x1[rest])).clamp_min(0)
      ih = (torch.minimum(y2[i], y2[rest]) - torch.maximum(y1[i], y1[rest])).clamp_min(0)
      inter = iw * ih
      iou = inter / (area[i] + area[rest] - inter).clamp_min(1e-12)
      order = rest[iou <= nms_iou_threshold]
    idx = torch.zeros(max_output_size, dtype=torch.int32, device=bboxes.device)
    mask = torch.zeros(max_output_size, device=bboxes.device)
    if keep:
      idx[:len(keep)] = torch.tensor(keep, dtype=torch.int32, device=bboxes.device)
      mask[:len(keep)] = 1.0
    return idx, mask

  def BatchedNMSIndices(self, bboxes, scores, nms_iou_threshold=0.3, score_threshold=0.01,
                        max_num_boxes=None):
    """`NMSIndices` over a batch: bboxes `[B, N, 7]`, scores `[B, N]` → indices / mask
    `[B, K]` (ref :672).

    With per-class scores `[B, N, C]` this is the oriented per-class NMS instead (indices
    `[B, C, K]`, −1 padded) — the form the decoders of this repository use."""
    if scores.dim() == 3:
      idx, _, mask = self.BatchedOrientedNMSIndices(
          bboxes, scores, nms_iou_threshold, score_threshold, max_num_boxes or bboxes.shape[1])
      return torch.where(mask > 0, idx, torch.full_like(idx, -1)), mask
    k = max_num_boxes or bboxes.shape[1]
    outs = [self.NMSIndices(bboxes[i], scores[i], k, nms_iou_threshold, score_threshold)
            for i in range(bboxes.shape[0])]
    return torch.stack([o[0] for o in outs]), torch.stack([o[1] for o in outs])

  def BatchedOrientedNMSIndices(self, bboxes, scores, nms_iou_threshold, score_threshold,
                                max_boxes_per_class):
    """Per-class rotated-box NMS (ref :719): bboxes `[B, N, 7]`, scores `[B, N, C]`;
    thresholds are floats or per-class lists. → (bbox_indices, bbox_scores, valid_mask), each
    `[B, C, max_boxes_per_class]`, picks in descending score order."""
    b, n, _ = bboxes.shape
    assert tuple(scores.shape[:2]) == (b, n)
    c = scores.shape[-1]
    def PerClass(v):
      return [float(x) for x in v] if isinstance(v, (list, tuple)) else [float(v)] * c
    iou_thr, score_thr = PerClass(nms_iou_threshold), PerClass(score_threshold)
    k = int(max_boxes_per_class)
    idx_np = np.stack([
        ops.host().nms_3d(_Np(bboxes[i]), _Np(scores[i]), iou_thr, score_thr, k)
        for i in range(b)])
    idx = torch.from_numpy(idx_np).long().to(bboxes.device)                 # −1 padded
    mask = (idx >= 0).float()
    safe = idx.clamp_min(0)
    picked = scores.detach().transpose(1, 2).gather(2, safe) * mask
    return safe.int() * mask.int(), picked, mask

  # -------------------------------------------------------------------------- projections --
  def CornersToImagePlane(self, corners, velo_to_image_plane):
    """corners `[B, N, 8, 3]`, projection `[B, 3, 4]` → pixel corners `[B, N, 8, 2]`
    (ref :780)."""
    b, n = corners.shape[:2]
    out = [geometry.PointsToImagePlane(corners[i].reshape(-1, 3), velo_to_image_plane[i])[:, :2]
           for i in range(b)]
    return torch.stack(out).reshape(b, n, 8, 2)

  # ----------------------------------------------------------------------------- angle bins --
  def AngleToBin(self, assigned_gt_bboxes, angle_bins, min_angle_val=0,
                 max_angle_val=2 * math.pi):
    """Heading → (bin id `[...]` int, residual `[...]` in bin widths, within ±0.5) (ref :817).
    Bins are centred on k·width: an angle of 0 sits in the middle of bin 0."""
    assert assigned_gt_bboxes.shape[-1] == 7
    width = (max_angle_val - min_angle_val) / float(angle_bins)
    angle = geometry.WrapAngleRad(assigned_gt_bboxes[..., -1], min_angle_val, max_angle_val)
    angle = geometry.WrapAngleRad(angle + width / 2, min_angle_val, max_angle_val)
    angle = angle - min_angle_val
    class_id = torch.floor(angle / width).to(torch.int32)
    residual = (angle - (class_id.float() + 0.5) * width) / width
    return class_id, residual

  def BinToAngle(self, predicted_angle_cls, predicted_angle_res, angle_bins, min_angle_val=0,
                 max_angle_val=2 * math.pi):
    """(bin scores `[..., bins]`, per-bin residuals `[..., bins]`) → angle `[...]`
    (ref :867)."""
    assert predicted_angle_cls.shape[-1] == angle_bins
    assert predicted_angle_res.shape == predicted_angle_cls.shape
    width = (max_angle_val - min_angle_val) / float(angle_bins)
    cls = predicted_angle_cls.argmax(-1)
    res = predicted_angle_res.gather(-1, cls.unsqueeze(-1)).squeeze(-1)
    return (cls.to(res.dtype) + res) * width + min_angle_val

  # ----------------------------------------------------------------------------- anchor-free --
  def ResidualsToBBoxesAnchorFree(self, points, residuals, angle_cls, angle_res):
    """points `[..., 3]` + (Δxyz, dims) `[..., 6]` + angle bins → boxes `[..., 7]` (ref :921)."""
    assert points.shape[-1] == 3 and residuals.shape[-1] == 6
    phi = self.BinToAngle(angle_cls, angle_res, angle_cls.shape[-1])
    return torch.cat([residuals[..., :3] + points, residuals[..., 3:], phi.unsqueeze(-1)], -1)

  def LocalizationResidualsAnchorFree(self, points, assigned_gt_bboxes):
    """Targets of an anchor-free head: (box centre − point, box dims, box heading) (ref :975)."""
    assert points.shape[-1] == 3 and assigned_gt_bboxes.shape[-1] == 7
    return torch.cat([assigned_gt_bboxes[..., :3] - points, assigned_gt_bboxes[..., 3:]], -1)

  def FindCenterPoints(self, points, gt_bboxes, gt_bboxes_mask, random_seed=None,
                       random_chosen=False):
    """For every real ground-truth box the point `[N, 3]` closest (L1) to its centre — or,
    with `random_chosen`, a random point inside its footprint when there is one (ref :1024).
    → (coordinates `[M, 3]`, indices `[M]`); rows of masked-out boxes are 0."""
    n, m = points.shape[0], gt_bboxes.shape[0]
    assert points.shape[1] == 3 and gt_bboxes.shape[1] == 7
    real = gt_bboxes_mask > 0
    boxes = gt_bboxes[real]
    dist = (points[:, None, :] - boxes[None, :, :3]).abs().sum(-1)               # [N, R]
    if random_chosen and boxes.shape[0]:
      footprint = geometry.BBoxCorners(boxes)[:, 0:4, 0:2]                       # [R, 4, 2]
      inside = geometry.IsWithinBBox(points[None, :, :2].expand(boxes.shape[0], n, 2),
                                     footprint).t().float()                    # [N, R]
      gen = None
      if random_seed is not None:
        gen = torch.Generator().manual_seed(int(random_seed))
      prob = torch.rand(n, boxes.shape[0], generator=gen).to(points.device)
      has_points = inside.max(0, keepdim=True).values
      dist = dist * (1 - has_points) + (1 - inside * prob) * has_points
    values = points.new_zeros(m, 3)
    indices = torch.zeros(m, dtype=torch.int32, device=points.device)
    if boxes.shape[0] and n:
      pick = dist.argmin(0)
      values[real] = points[pick]
      indices[real] = pick.int()
    return values, indices

  def AssignPoints(self, points, gt_bboxes, gt_labels, gt_bboxes_mask, cls_num,
                   expand_gt_bbox_dims, random_seed=None, background_class_id=0,
                   ignore_z=False):
    """Assigns every point `[N, 3]` to a ground-truth box it lies in (boxes grown by
    `expand_gt_bbox_dims` to catch context points; ties broken at random) (ref :1108).

    Returns NestedMap(assigned_gt_idx [N] (index into the *unmasked* box list, −1: none),
    assigned_gt_bbox [N, 7], assigned_gt_labels [N], assigned_cls_mask [N] (all 1: points
    outside every box are background), assigned_reg_mask [N, cls_num] (1 at the assigned
    box's class for interior points))."""
    assert points.shape[1] == 3 and gt_bboxes.shape[1] == 7
    n = points.shape[0]
    real = gt_bboxes_mask > 0
    valid_idx = torch.cat([torch.nonzero(real)[:, 0], real.new_tensor([-1], dtype=torch.long)])
    boxes, labels = gt_bboxes[real], gt_labels[real]
    grown = boxes.clone()
    grown[:, 3:6] = grown[:, 3:6] + torch.as_tensor(expand_gt_bbox_dims, dtype=boxes.dtype,
                                                    device=boxes.device).reshape(1, 3)
    r = boxes.shape[0]
    if r == 0:
      inside = points.new_zeros(n, 1)
    elif ignore_z:
      footprint = geometry.BBoxCorners(grown)[:, 0:4, 0:2]
      inside = geometry.IsWithinBBox(points[None, :, :2].expand(r, n, 2), footprint).t().float()
    else:
      inside = geometry.IsWithinBBox3D(points, grown).float()
    gen = None
    if random_seed is not None:
      gen = torch.Generator().manual_seed(int(random_seed))
    prob = 0.1 + 0.9 * torch.rand(inside.shape, generator=gen).to(points.device)
    scored = inside * prob
    best = scored.argmax(-1)
    is_in = (scored.max(-1).values > 0)
    dummy = boxes.new_tensor([[0, 0, 0, 1, 1, 1, 0]])
    all_boxes = torch.cat([boxes, dummy], 0)
    all_labels = torch.cat([labels, labels.new_tensor([background_class_id])], 0)
    assigned = torch.where(is_in, best, torch.full_like(best, r))
    assigned_labels = all_labels[assigned]
    reg_mask = is_in.float().unsqueeze(-1) * torch.nn.functional.one_hot(
        assigned_labels.long(), cls_num).float()
    return NestedMap(
        assigned_gt_idx=valid_idx[assigned].int(),
        assigned_gt_bbox=all_boxes[assigned],
        assigned_gt_labels=assigned_labels,
        assigned_cls_mask=torch.ones(n, device=points.device),
        assigned_reg_mask=reg_mask)


def RandomPadOrTrimTo(tensor_list, num_points_out, seed=None):
  """Brings every tensor's leading dim to `num_points_out` (ref :1288): a random subset when
  there are more points; random *duplicates* of real points when there are fewer (zeros if
  there is nothing to duplicate). → (tensors, padding `[num_points_out]`, 1 at duplicates)."""
  actual = int(tensor_list[0].shape[0])
  gen = None
  if seed is not None:
    gen = torch.Generator().manual_seed(int(seed))
  dev = tensor_list[0].device
  padding = (torch.arange(num_points_out, device=dev) >= actual).float()
  if actual > num_points_out:
    idx = torch.randperm(actual, generator=gen)[:num_points_out].to(dev)
    return [t[idx] for t in tensor_list], padding
  if actual == 0:
    return [t.new_zeros((num_points_out,) + tuple(t.shape[1:])) for t in tensor_list], padding
  extra = torch.randint(0, actual, (num_points_out - actual,), generator=gen).to(dev)
  return [torch.cat([t, t[extra]], 0) for t in tensor_list], padding
