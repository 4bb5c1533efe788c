"""KITTI output decoder (ref `lingvo/tasks/car/kitti_decoder.py`): filters detections to
the camera frustum, projects them to 2-D image boxes (needed for the min-height rule and
the official text format), feeds the KITTI AP metric."""

from __future__ import annotations

import numpy as np
import torch

from lingvo_b200.core import metrics as metrics_lib
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.car import base_decoder
from lingvo_b200.models.car import breakdown_metric
from lingvo_b200.models.car import detection_3d_metrics
from lingvo_b200.models.car import geometry
from lingvo_b200.models.car import kitti_ap_metric
from lingvo_b200.models.car import kitti_metadata


class KITTIDecoder(base_decoder.BaseDecoder):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('filter_predictions_outside_frustum', False,
             'Drop detections whose projection misses the camera image.')
    p.Define('truncation_threshold', 0.0,
             'Drop detections with more than this fraction outside the image.')
    p.ap_metric = kitti_ap_metric.KITTIAPMetrics.Params(kitti_metadata.KITTIMetadata())
    return p

  def CreateDecoderMetrics(self):
    p = self.params
    ap = p.ap_metric.Copy()
    m = {'num_samples_in_batch': metrics_lib.AverageMetric(),
         'kitti_AP_v2': ap.Instantiate()}
    if p.draw_visualizations:
      m['top_down_visualization'] = detection_3d_metrics.TopDownVisualizationMetric()
      m['camera_visualization'] = detection_3d_metrics.CameraVisualization()
    return m

  def _BBox2DImage(self, bbox_corners_image, width, height):
    """Corner projections `[..., 8, 2]` → clipped 2-D boxes `[..., 4]` = ymin, xmin, ymax,
    xmax and the fraction of the un-clipped box lying inside the image."""
    x, y = bbox_corners_image[..., 0], bbox_corners_image[..., 1]
    xmin, xmax, ymin, ymax = x.min(-1).values, x.max(-1).values, y.min(-1).values, y.max(-1).values
    w = width.view(-1, *([1] * (xmin.dim() - 1))).to(x.dtype)
    h = height.view(-1, *([1] * (xmin.dim() - 1))).to(x.dtype)
    cx0, cx1 = xmin.clamp_min(0), torch.minimum(xmax, w)
    cy0, cy1 = ymin.clamp_min(0), torch.minimum(ymax, h)
    full = ((xmax - xmin) * (ymax - ymin)).clamp_min(1e-6)
    inside = ((cx1 - cx0).clamp_min(0) * (cy1 - cy0).clamp_min(0)) / full
    return torch.stack([cy0, cx0, cy1, cx1], -1), inside

  def ProcessOutputs(self, input_batch, model_outputs):
    """model_outputs: per_class_predicted_bboxes `[B,C,K,7]`, …_bbox_scores, …_valid_mask
    (ref :152)."""
    p = self.params
    boxes = model_outputs.per_class_predicted_bboxes
    scores = model_outputs.per_class_predicted_bbox_scores
    mask = model_outputs.per_class_valid_mask
    b, c, k, _ = boxes.shape
    img = input_batch.images
    corners = geometry.BBoxCorners(boxes)                                  # [B,C,K,8,3]
    proj = torch.stack([
        geometry.PointsToImagePlane(corners[i].reshape(-1, 3), img.velo_to_image_plane[i])
        for i in range(b)]).reshape(b, c, k, 8, 3)
    in_front = (proj[..., 2] > 0).all(-1)
    bbox2d, inside = self._BBox2DImage(proj[..., :2], img.width, img.height)
    heights = (bbox2d[..., 2] - bbox2d[..., 0]).clamp_min(0)
    if p.filter_predictions_outside_frustum:
      keep = in_front & (inside > p.truncation_threshold)
      scores = scores * keep.to(scores.dtype)
      mask = mask * keep.to(mask.dtype)
    lab = input_batch.decoder_copy.labels if 'decoder_copy' in input_batch else input_batch.labels
    out = NestedMap(
        per_class_predicted_bboxes=boxes, per_class_predicted_bbox_scores=scores,
        per_class_valid_mask=mask, per_class_predicted_bboxes_2d=bbox2d,
        per_class_predicted_bbox_heights=heights * mask,
        per_class_predicted_bbox_corners_image=proj[..., :2],
        source_ids=lab.source_id, gt_bboxes_3d=lab.bboxes_3d,
        gt_bboxes_3d_mask=lab.get('unfiltered_bboxes_3d_mask', lab.bboxes_3d_mask),
        gt_labels=lab.labels, gt_difficulties=lab.difficulties,
        gt_bboxes_3d_num_points=lab.bboxes_3d_num_points)
    if p.draw_visualizations and 'decoder_copy' in input_batch:
      pts, pad = self._SampleLaserForVisualization(
          input_batch.decoder_copy.lasers.points_xyz,
          input_batch.decoder_copy.lasers.points_padding)
      out.points_sampled, out.points_sampled_padding = pts, pad
    return out

  def PostProcessDecodeOut(self, dec_out_dict, dec_metrics_dict):
    """Updates metrics from one decoded batch (numpy dict) (ref :250)."""
    d = dec_out_dict
    to_np = lambda x: x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
    boxes, scores = to_np(d['per_class_predicted_bboxes']), to_np(d['per_class_predicted_bbox_scores'])
    heights = to_np(d['per_class_predicted_bbox_heights'])
    b, c = scores.shape[:2]
    n_cls = kitti_metadata.KITTIMetadata().NumClasses()
    dec_metrics_dict['num_samples_in_batch'].Update(b)
    for i in range(b):
      gm = to_np(d['gt_bboxes_3d_mask'])[i] > 0
      det_scores = np.zeros((n_cls, scores.shape[2]), np.float32)
      det_boxes = np.zeros((n_cls, scores.shape[2], 7), np.float32)
      det_heights = np.zeros((n_cls, scores.shape[2]), np.float32)
      det_scores[:c], det_boxes[:c], det_heights[:c] = scores[i], boxes[i], heights[i]
      sid = to_np(d['source_ids'])[i]
      sid = bytes(sid.tolist()).decode().strip() if sid.dtype == np.uint8 else str(sid)
      dec_metrics_dict['kitti_AP_v2'].Update(sid, NestedMap(
          groundtruth_labels=to_np(d['gt_labels'])[i][gm],
          groundtruth_bboxes=to_np(d['gt_bboxes_3d'])[i][gm],
          groundtruth_difficulties=to_np(d['gt_difficulties'])[i][gm],
          groundtruth_num_points=to_np(d['gt_bboxes_3d_num_points'])[i][gm],
          detection_scores=det_scores, detection_boxes=det_boxes,
          detection_heights_in_pixels=det_heights))
    return []



