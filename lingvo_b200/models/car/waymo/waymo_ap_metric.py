"""Waymo Open Dataset flavour of the AP metric (ref
`lingvo/tasks/car/waymo/waymo_ap_metric.py`).

The reference defers to the `waymo_open_dataset` metrics library. That library is not
available here, so the official definitions are implemented on the native AP op:
  * AP = area under the interpolated PR curve (VOC style, 101-point reporting);
  * LEVEL_2 evaluates every box, LEVEL_1 only boxes labelled LEVEL_1 (LEVEL_2 boxes are
    ignore-first-match);
  * APH weights each true positive by heading accuracy 1 − min(|Δθ|, 2π − |Δθ|) / π;
  * `box_type='2d'` evaluates bird's-eye-view IoU;
  * breakdowns by range ([0,30), [30,50), [50,∞) m) via `WaymoBreakdownMetric`.
"""

from __future__ import annotations

import numpy as np

from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.car import ap_metric
from lingvo_b200.models.car import breakdown_metric
from lingvo_b200.models.car import ops as car_ops

RANGE_EDGES = (0.0, 30.0, 50.0, np.inf)


def BuildWaymoMetricConfig(metadata, box_type, waymo_breakdown_metrics):
  """Plain-dict equivalent of the WOD metrics `Config` proto (ref :38)."""
  names = metadata.ClassNames()
  return dict(
      num_desired_score_cutoffs=metadata.NumberOfPrecisionRecallPoints() - 1,
      box_type={'2d': 'TYPE_2D', '3d': 'TYPE_3D'}[box_type],
      iou_thresholds=[metadata.IoUThresholds().get(n, 0.0) for n in names],
      breakdown_generator_ids=['ONE_SHARD'] + list(waymo_breakdown_metrics),
      difficulties=['LEVEL_1', 'LEVEL_2'], matcher_type='TYPE_HUNGARIAN')


class WaymoAPMetrics(ap_metric.APMetrics):
  """ref :73."""

  @classmethod
  def Params(cls, metadata):
    p = super().Params(metadata)
    p.Define('waymo_breakdown_metrics', [], "Any of 'RANGE', 'VELOCITY'.")
    p.ap_algorithm = 'VOC'
    return p

  def __init__(self, params):
    super().__init__(params)
    self._waymo_metric_config = BuildWaymoMetricConfig(
        self.metadata, self.params.box_type, self.params.waymo_breakdown_metrics)
    self._aph = {}
    self._waymo_breakdowns = {}
    for name in self.params.waymo_breakdown_metrics:
      self._waymo_breakdowns[name] = WaymoBreakdownMetric(
          WaymoBreakdownMetric.Params().Set(metadata=self.metadata, breakdown_list=[name]))

  def _GetData(self, classid, difficulty=None, distance=None, num_points=None, rotation=None,
               range_bin=None):
    data = super()._GetData(classid, None, distance, num_points, rotation)
    if data is None:
      return None
    g = self._LoadBoundingBoxes('groundtruth', classid, distance, num_points, rotation)
    ignore = np.zeros(len(g), np.int32)
    if difficulty == 'LEVEL_1':
      ignore = (g.difficulties != self.metadata.DifficultyLevels()['LEVEL_1']).astype(np.int32)
    if range_bin is not None:
      lo, hi = RANGE_EDGES[range_bin], RANGE_EDGES[range_bin + 1]
      d = np.linalg.norm(g.boxes[:, :2], axis=1)
      ignore = np.where((d >= lo) & (d < hi), ignore, 1).astype(np.int32)
      pd_d = np.linalg.norm(data.pd.bbox[:, :2], axis=1)
      data.pd.ignore = (~((pd_d >= lo) & (pd_d < hi))).astype(np.int32)
    data.gt.ignore = ignore
    return data

  def _HeadingWeightedAP(self, data):
    """APH: re-run the matching and weight hits by heading accuracy."""
    gt_b, pd_b = data.gt.bbox, data.pd.bbox
    if self.params.box_type == '2d':
      gt_b, pd_b = self._Flatten2D(gt_b), self._Flatten2D(pd_b)
    _, _, sh = car_ops.average_precision3d(
        data.iou_threshold, gt_b, data.gt.imgid, data.gt.ignore, pd_b, data.pd.imgid,
        data.pd.ignore, data.pd.score, num_recall_points=1, algorithm='VOC')
    hit = sh[:, 1].numpy() > 0
    n_gt = int((data.gt.ignore == 0).sum())
    if not n_gt or not hit.any():
      return 0.0
    # heading accuracy of each hit against its best-IoU ground truth in the same image
    iou = car_ops.pairwise_iou3d(pd_b[hit], gt_b).numpy()
    same = data.pd.imgid[hit][:, None] == data.gt.imgid[None, :]
    j = np.where(same, iou, -1.0).argmax(1)
    dth = np.abs(pd_b[hit][:, 6] - gt_b[j][:, 6]) % (2 * np.pi)
    acc = 1.0 - np.minimum(dth, 2 * np.pi - dth) / np.pi
    order = np.argsort(-data.pd.score, kind='stable')
    w = np.zeros(len(order), np.float32)
    w[np.nonzero(hit)[0]] = acc
    counted = (hit | (data.pd.ignore == 0))[order]
    tp_w = np.cumsum(w[order])
    tp = np.cumsum(hit[order].astype(np.float32))
    fp = np.cumsum((~hit[order] & counted).astype(np.float32))
    prec = tp_w / np.maximum(tp + fp, 1e-9)
    rec = tp / n_gt
    prec = np.maximum.accumulate(prec[::-1])[::-1]
    return float(np.sum(np.diff(np.concatenate([[0.0], rec])) * prec))

  def _ComputeFinalMetrics(self, classids=None, difficulty=None, distance=None,
                           num_points=None, rotation=None, range_bin=None):
    classids = classids or self.metadata.EvalClassIndices()
    pts = self.metadata.NumberOfPrecisionRecallPoints()
    aps = np.full(len(classids), np.nan, np.float32)
    prs = np.zeros((len(classids), pts, 2), np.float32)
    any_data = False
    for i, c in enumerate(classids):
      data = self._GetData(c, difficulty, distance, num_points, rotation, range_bin)
      if data is None:
        continue
      any_data = True
      aps[i], prs[i] = self._BuildMetric(data, c)
      if distance is None and num_points is None and rotation is None:
        key = (c, difficulty, range_bin)
        self._aph[key] = self._HeadingWeightedAP(data)
    return (aps, prs) if any_data else None

  def _EvaluateIfNecessary(self):
    if self._is_eval_complete:
      return
    super()._EvaluateIfNecessary()
    for m in self._waymo_breakdowns.values():
      m.ComputeMetrics(self._ComputeFinalMetrics)

  @property
  def value(self):
    """Mean LEVEL_2 (= all boxes) AP over the evaluated classes."""
    self._EvaluateIfNecessary()
    aps = self._AveragePrecisionByDifficulty()
    v = aps.get('LEVEL_2', aps['default'])
    return float(np.nanmean(v)) if np.any(~np.isnan(v)) else 0.0

  def Scalars(self, name):
    out = super().Scalars(name)
    names = self.metadata.ClassNames()
    for (c, level, rb), aph in sorted(self._aph.items(), key=str):
      tag = '%s/APH_%s_%s' % (name, names[c].lower(), level or 'default')
      if rb is not None:
        tag += '_range%d' % rb
      out[tag] = aph
    for m in self._waymo_breakdowns.values():
      out.update(m.Scalars(name))
    return out


class WaymoBreakdownMetric(breakdown_metric.BreakdownMetric):
  """AP per range bucket (ref :360)."""

  SELECTOR = 'range_bin'

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('breakdown_list', ['RANGE'], 'Breakdowns to compute.')
    return p

  def NumBinsOfHistogram(self):
    return len(RANGE_EDGES) - 1

  def BinLabels(self):
    return ['%g-%gm' % (RANGE_EDGES[i], RANGE_EDGES[i + 1]) for i in range(len(RANGE_EDGES) - 1)]

  def Discretize(self, bboxes):
    d = np.linalg.norm(np.asarray(bboxes, np.float32).reshape(-1, 7)[:, :2], axis=1)
    return np.digitize(d, RANGE_EDGES[1:-1])

  def AccumulateHistogram(self, result):
    self._Accumulate(self.Discretize(result.bboxes), result.labels)
