"""Turning detector outputs into final boxes (ref `lingvo/tasks/car/detection_decoder.py`).

  DecodeWithNMS          rotated-IoU NMS per class (native op) or class-agnostic
  HeatMapNMS             peak picking on a centre heat map (max-pool NMS, on device)
  DecodeWithMaxPoolNMS   CenterNet-style decode: heat-map peaks index the box regression
"""

from __future__ import annotations

import torch
import torch.nn.functional as F

from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.car import ops as car_ops


def _PerClass(v, c):
  return [float(v)] * c if not isinstance(v, (list, tuple)) else [float(x) for x in v]


def _MultiClassOrientedDecodeWithNMS(predicted_bboxes, classification_scores,
                                     nms_iou_threshold, score_threshold,
                                     max_boxes_per_class=None):
  """Independent oriented NMS per class (ref :73)."""
  b, n, _ = predicted_bboxes.shape
  c = classification_scores.shape[-1]
  k = max_boxes_per_class or n
  dev = predicted_bboxes.device
  idx = torch.zeros(b, c, k, dtype=torch.long)
  mask = torch.zeros(b, c, k)
  for i in range(b):
    ii, mm = car_ops.non_max_suppression_3d(
        predicted_bboxes[i], classification_scores[i], _PerClass(nms_iou_threshold, c),
        _PerClass(score_threshold, c), k)
    idx[i], mask[i] = ii.clamp_min(0), mm
  idx, mask = idx.to(dev), mask.to(dev)
  flat = idx.reshape(b, c * k)
  boxes = predicted_bboxes.gather(1, flat.unsqueeze(-1).expand(-1, -1, 7)).reshape(b, c, k, 7)
  scores = classification_scores.transpose(1, 2).gather(2, idx)          # [B, C, K]
  return idx, boxes * mask.unsqueeze(-1), scores * mask, mask


def _SingleClassDecodeWithNMS(predicted_bboxes, classification_scores, nms_iou_threshold,
                              score_threshold, max_boxes_per_class=None):
  """One class-agnostic NMS on the per-box max score; every class then reads its own score
  at the surviving boxes (ref :133)."""
  b, n, _ = predicted_bboxes.shape
  c = classification_scores.shape[-1]
  k = max_boxes_per_class or n
  dev = predicted_bboxes.device
  best = classification_scores.max(-1).values
  idx = torch.zeros(b, k, dtype=torch.long)
  mask = torch.zeros(b, k)
  thr = nms_iou_threshold if not isinstance(nms_iou_threshold, (list, tuple)) else \
      nms_iou_threshold[0]
  sthr = score_threshold if not isinstance(score_threshold, (list, tuple)) else \
      min(score_threshold)
  for i in range(b):
    ii, mm = car_ops.non_max_suppression_3d(predicted_bboxes[i], best[i].unsqueeze(-1),
                                            [thr], [sthr], k)
    idx[i], mask[i] = ii[0].clamp_min(0), mm[0]
  idx, mask = idx.to(dev), mask.to(dev)
  boxes = predicted_bboxes.gather(1, idx.unsqueeze(-1).expand(-1, -1, 7))
  scores = classification_scores.gather(1, idx.unsqueeze(-1).expand(-1, -1, c))
  boxes = boxes.unsqueeze(1).expand(b, c, k, 7) * mask.view(b, 1, k, 1)
  scores = scores.transpose(1, 2) * mask.unsqueeze(1)
  return idx, boxes, scores, mask.unsqueeze(1).expand(b, c, k).contiguous()


def DecodeWithNMS(predicted_bboxes, classification_scores, nms_iou_threshold, score_threshold,
                  max_boxes_per_class=None, use_oriented_per_class_nms=False):
  """→ (indices, boxes `[B,C,K,7]`, scores `[B,C,K]`, valid mask `[B,C,K]`) (ref :22)."""
  fn = (_MultiClassOrientedDecodeWithNMS if use_oriented_per_class_nms
        else _SingleClassDecodeWithNMS)
  return fn(predicted_bboxes, classification_scores, nms_iou_threshold, score_threshold,
            max_boxes_per_class)


def HeatMapNMS(heat_map_scores, kernel_size, max_num_objects, score_threshold):
  """Local maxima of `heat_map_scores [B, gx, gy, C]` (ref :222) → NestedMap(
  top_k_indices `[B, C, K, 2]` (gx, gy), top_k_scores `[B, C, K]`, peak_heat_map
  `[B, gx, gy, C]`)."""
  b, gx, gy, c = heat_map_scores.shape
  kh, kw = (kernel_size[1], kernel_size[2]) if len(kernel_size) == 4 else tuple(kernel_size)
  hm = heat_map_scores.permute(0, 3, 1, 2)
  pooled = F.max_pool2d(hm, (kh, kw), stride=1, padding=(kh // 2, kw // 2))
  peaks = hm * (hm == pooled).to(hm.dtype)
  scores, flat = peaks.reshape(b, c, -1).topk(min(max_num_objects, gx * gy), -1)
  idx = torch.stack([flat // gy, flat % gy], -1)
  keep = (scores > score_threshold).to(hm.dtype)
  peak_map = torch.zeros(b, c, gx * gy, device=hm.device, dtype=hm.dtype)
  peak_map.scatter_(2, flat, keep)
  return NestedMap(top_k_indices=idx, top_k_scores=scores * keep,
                   peak_heat_map=peak_map.reshape(b, c, gx, gy).permute(0, 2, 3, 1))


def DecodeWithMaxPoolNMS(predicted_bboxes, classification_scores, heatmap, kernel_size=(3, 3),
                         max_boxes_per_class=64, score_threshold=0.1):
  """CenterNet-style decode (ref :299): `heatmap [B,gx,gy,C]` peaks select rows of
  `predicted_bboxes [B, gx·gy, 7]`. → (indices `[B,C,K]`, boxes, scores, mask)."""
  del classification_scores
  b, gx, gy, c = heatmap.shape
  nms = HeatMapNMS(heatmap, kernel_size, max_boxes_per_class, score_threshold)
  flat = nms.top_k_indices[..., 0] * gy + nms.top_k_indices[..., 1]       # [B, C, K]
  k = flat.shape[-1]
  boxes = predicted_bboxes.gather(1, flat.reshape(b, c * k, 1).expand(-1, -1, 7)).reshape(
      b, c, k, 7)
  mask = (nms.top_k_scores > 0).to(boxes.dtype)
  return flat, boxes * mask.unsqueeze(-1), nms.top_k_scores, mask
