"""Base task for point-cloud detectors (ref `lingvo/tasks/car/point_detector.py`).

Subclasses implement `ComputePredictions` (→ `residuals`, `classification_logits`, …) and
`ComputeLoss`; this class turns predictions into boxes, runs NMS, hands the result to the
dataset's output decoder, and exposes the inference entry point.
"""

from __future__ import annotations

import torch

from lingvo_b200.core import base_model
from lingvo_b200.core import py_utils
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.car import detection_3d_lib
from lingvo_b200.models.car import detection_decoder
from lingvo_b200.models.car import kitti_decoder


class PointDetectorBase(base_model.BaseTask):
  """ref :27."""

  @classmethod
  def Params(cls, num_classes=2):
    p = super().Params()
    p.Define('num_classes', num_classes, 'Classes including background (index 0).')
    p.Define('max_nms_boxes', 1024, 'Boxes kept per class after NMS.')
    p.Define('nms_iou_threshold', 0.3, 'NMS IoU threshold (scalar or per class).')
    p.Define('nms_score_threshold', 0.01, 'Score threshold before NMS (scalar or per class).')
    p.Define('visualization_classification_threshold', 0.25, 'Score threshold for drawing.')
    p.Define('output_decoder', kitti_decoder.KITTIDecoder.Params(), 'Dataset decoder.')
    p.Define('use_oriented_per_class_nms', False, 'Oriented NMS per class.')
    p.Define('inference_batch_size', None, 'Static batch size of the inference graph.')
    p.Define('decode_include_residuals', False, 'Add residual debug tensors to the decode output.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self._utils_3d = detection_3d_lib.Utils3D()
    self.CreateChild('output_decoder', self.params.output_decoder)

  def CreateDecoderMetrics(self):
    return self.output_decoder.CreateDecoderMetrics()

  def _BBoxesAndLogits(self, input_batch, predictions):
    """Default for anchor-based heads: decode residuals against the anchors."""
    p = self.params
    b = predictions.residuals.shape[0]
    boxes = self._utils_3d.ResidualsToBBoxes(input_batch.anchor_bboxes, predictions.residuals)
    return NestedMap(predicted_bboxes=boxes.reshape(b, -1, 7),
                     classification_logits=predictions.classification_logits.reshape(
                         b, -1, p.num_classes))

  def _BBoxDimensionErrors(self, gt_bboxes, pred_bboxes, regression_weights, epsilon=1e-6):
    """Masked mean absolute errors of each box dimension (ref :126)."""
    w = regression_weights.reshape(regression_weights.shape + (1,))
    denom = regression_weights.sum().clamp_min(epsilon)
    err = ((gt_bboxes - pred_bboxes).abs() * w).reshape(-1, 7).sum(0) / denom
    phi = (torch.remainder(gt_bboxes[..., 6] - pred_bboxes[..., 6] + torch.pi / 2, torch.pi) -
           torch.pi / 2).abs()
    names = ('x', 'y', 'z', 'dx', 'dy', 'dz')
    out = {'error/' + n: (err[i], denom) for i, n in enumerate(names)}
    out['error/phi'] = ((phi * regression_weights).sum() / denom, denom)
    return out

  def Inference(self):
    """{'default': fn(input_batch NestedMap) → per-class boxes / scores / mask}."""
    def _Default(input_batch):
      with torch.no_grad():
        return self._DecodeImpl(input_batch).Filter(lambda v: isinstance(v, torch.Tensor))
    return {'default': _Default}

  def _DecodeImpl(self, input_batch):
    p = self.params
    predictions = self.ComputePredictions(self.theta, input_batch)
    bl = self._BBoxesAndLogits(input_batch, predictions)
    boxes = bl.predicted_bboxes
    b, n, _ = boxes.shape
    logits = bl.classification_logits.reshape(b, n, p.num_classes)
    scores = torch.sigmoid(logits.float())
    idx, cls_boxes, cls_scores, valid = detection_decoder.DecodeWithNMS(
        boxes.float(), scores, nms_iou_threshold=p.nms_iou_threshold,
        score_threshold=p.nms_score_threshold, max_boxes_per_class=p.max_nms_boxes,
        use_oriented_per_class_nms=p.use_oriented_per_class_nms)
    cls_scores = cls_scores * valid
    viz = torch.where(cls_scores >= p.visualization_classification_threshold, cls_scores,
                      torch.zeros_like(cls_scores))
    out = NestedMap(per_class_predicted_bboxes=cls_boxes,
                    per_class_predicted_bbox_scores=cls_scores, per_class_valid_mask=valid,
                    visualization_weights=viz)
    if p.decode_include_residuals and 'residuals' in predictions:
      def _Gather(t):
        t = t.reshape(b, n, -1)
        if idx.dim() == 3:
          flat = idx.reshape(b, -1)
          g = t.gather(1, flat.unsqueeze(-1).expand(-1, -1, t.shape[-1]))
          return g.reshape(b, idx.shape[1], idx.shape[2], -1)
        g = t.gather(1, idx.unsqueeze(-1).expand(-1, -1, t.shape[-1]))
        return g.unsqueeze(1).expand(-1, p.num_classes, -1, -1)
      out.per_class_residuals = _Gather(predictions.residuals)
      out.per_class_logits = _Gather(predictions.classification_logits)
      if 'anchor_localization_residuals' in input_batch:
        out.per_class_gt_residuals = _Gather(input_batch.anchor_localization_residuals)
        out.per_class_gt_labels = _Gather(input_batch.assigned_gt_labels.float())
        out.per_class_anchor_boxes = _Gather(input_batch.anchor_bboxes)
    return out

  def Decode(self, input_batch):
    with torch.no_grad():
      out = self._DecodeImpl(input_batch)
      out.update(self.output_decoder.ProcessOutputs(input_batch, out))
      out.global_step = torch.tensor(py_utils.GetGlobalStep())
    return out

  def PostProcessDecodeOut(self, dec_out_dict, dec_metrics_dict):
    return self.output_decoder.PostProcessDecodeOut(dec_out_dict, dec_metrics_dict)

  def DecodeFinalize(self, decode_finalize_args):
    self.output_decoder.DecodeFinalize(decode_finalize_args)
