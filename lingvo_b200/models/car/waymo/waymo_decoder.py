"""Waymo Open Dataset output decoder (ref `lingvo/tasks/car/waymo/waymo_decoder.py`)."""

from __future__ import annotations

import numpy as np
import torch

from lingvo_b200.core import metrics as metrics_lib
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.car import base_decoder
from lingvo_b200.models.car import detection_3d_metrics
from lingvo_b200.models.car.waymo import waymo_ap_metric
from lingvo_b200.models.car.waymo import waymo_metadata


class WaymoOpenDatasetDecoder(base_decoder.BaseDecoder):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('extra_ap_metrics', {}, 'name → extra AP metric params (e.g. BEV).')
    p.Define('save_residuals', False, 'Keep regression residuals in the decode output.')
    p.ap_metric = waymo_ap_metric.WaymoAPMetrics.Params(waymo_metadata.WaymoMetadata())
    return p

  def CreateDecoderMetrics(self):
    p = self.params
    m = {'num_samples_in_batch': metrics_lib.AverageMetric(),
         'waymo_metrics': p.ap_metric.Copy().Instantiate()}
    for name, mp in p.extra_ap_metrics.items():
      m[name] = mp.Copy().Instantiate()
    if p.draw_visualizations:
      m['top_down_visualization'] = detection_3d_metrics.TopDownVisualizationMetric()
      m['world_viewer'] = detection_3d_metrics.WorldViewer()
    return m

  def ProcessOutputs(self, input_batch, model_outputs):
    lab = input_batch.decoder_copy.labels if 'decoder_copy' in input_batch else input_batch.labels
    md = input_batch.get('metadata')
    out = NestedMap(
        per_class_predicted_bboxes=model_outputs.per_class_predicted_bboxes,
        per_class_predicted_bbox_scores=model_outputs.per_class_predicted_bbox_scores,
        per_class_valid_mask=model_outputs.per_class_valid_mask,
        gt_bboxes_3d=lab.bboxes_3d,
        gt_bboxes_3d_mask=lab.get('unfiltered_bboxes_3d_mask', lab.bboxes_3d_mask),
        gt_labels=lab.labels,
        gt_difficulties=lab.get('single_frame_detection_difficulties',
                                lab.get('detection_difficulties')),
        gt_bboxes_3d_num_points=lab.bboxes_3d_num_points, gt_speed=lab.get('speed'))
    if md is not None:
      out.run_segment, out.run_start_offset, out.pose = (
          md.run_segment, md.run_start_offset, md.pose)
    return out

  @staticmethod
  def _SourceId(d, i):
    to_np = lambda x: x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
    if 'run_segment' in d:
      seg = to_np(d['run_segment'])[i]
      seg = bytes(seg.tolist()).decode().strip() if seg.dtype == np.uint8 else str(seg)
      return '%s_%d' % (seg, int(to_np(d['run_start_offset'])[i]))
    return str(i)

  def PostProcessDecodeOut(self, dec_out_dict, dec_metrics_dict):
    d = dec_out_dict
    to_np = lambda x: x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
    boxes, scores = to_np(d['per_class_predicted_bboxes']), to_np(d['per_class_predicted_bbox_scores'])
    b, c, k = scores.shape
    n_cls = waymo_metadata.WaymoMetadata().NumClasses()
    dec_metrics_dict['num_samples_in_batch'].Update(b)
    ap_keys = [n for n in dec_metrics_dict if hasattr(dec_metrics_dict[n], '_GetData')]
    for i in range(b):
      gm = to_np(d['gt_bboxes_3d_mask'])[i] > 0
      det_scores = np.zeros((n_cls, k), np.float32)
      det_boxes = np.zeros((n_cls, k, 7), np.float32)
      det_scores[:c], det_boxes[:c] = scores[i], boxes[i]
      res = NestedMap(
          groundtruth_labels=to_np(d['gt_labels'])[i][gm],
          groundtruth_bboxes=to_np(d['gt_bboxes_3d'])[i][gm],
          groundtruth_difficulties=to_np(d['gt_difficulties'])[i][gm],
          groundtruth_num_points=to_np(d['gt_bboxes_3d_num_points'])[i][gm],
          detection_scores=det_scores, detection_boxes=det_boxes)
      if d.get('gt_speed') is not None:
        res.groundtruth_speed = to_np(d['gt_speed'])[i][gm]
      for n in ap_keys:
        dec_metrics_dict[n].Update(self._SourceId(d, i), res)
    return []
