"""Average-precision metric for 3-D detection (ref `lingvo/tasks/car/ap_metric.py`).

`APMetrics.Update(str_id, result)` is called once per evaluated scene with ground truth
and per-class detections; `value` is the mean AP over the evaluated classes, `Summary`
emits per-class / per-difficulty scalars and the breakdown plots. Boxes are kept in
columnar numpy storage (`Boxes3D`) and sliced with boolean masks; the AP itself is the
native `AveragePrecision3D` op.
"""

from __future__ import annotations

import numpy as np

from lingvo_b200.core import hyperparams
from lingvo_b200.core import metrics as metrics_lib
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.car import breakdown_metric
from lingvo_b200.models.car import ops as car_ops

_COLUMNS = (('imgids', np.int32, ()), ('scores', np.float32, ()), ('boxes', np.float32, (7,)),
            ('difficulties', np.int32, ()), ('distances', np.int32, ()),
            ('num_points', np.int32, ()), ('rotations', np.int32, ()),
            ('heights_in_pixels', np.float32, ()), ('speeds', np.float32, (2,)))


class Boxes3D:
  """Growable column store of boxes and their attributes (ref :28)."""

  def __init__(self, capacity=64):
    self._n = 0
    self._cols = {name: np.zeros((capacity,) + shape, dt) for name, dt, shape in _COLUMNS}

  def __len__(self):
    return self._n

  def _Reserve(self, extra):
    cap = len(self._cols['imgids'])
    if self._n + extra <= cap:
      return
    new_cap = max(2 * cap, self._n + extra)
    for name, arr in self._cols.items():
      grown = np.zeros((new_cap,) + arr.shape[1:], arr.dtype)
      grown[:self._n] = arr[:self._n]
      self._cols[name] = grown

  def Add(self, img_id, score, box, difficulty, distance, num_points, rotation,
          height_in_pixels, speed):
    self.Extend(imgids=[img_id], scores=[score], boxes=[box], difficulties=[difficulty],
                distances=[distance], num_points=[num_points], rotations=[rotation],
                heights_in_pixels=[height_in_pixels],
                speeds=[np.asarray(speed, np.float32).reshape(-1)[:2]])

  def Extend(self, **columns):
    k = len(columns['imgids'])
    self._Reserve(k)
    for name, dt, shape in _COLUMNS:
      v = np.asarray(columns[name], dt).reshape((k,) + shape)
      self._cols[name][self._n:self._n + k] = v
    self._n += k

  def Select(self, mask):
    out = Boxes3D(max(int(mask.sum()), 1))
    out._n = int(mask.sum())   # pylint: disable=protected-access
    for name in self._cols:
      out._cols[name][:out._n] = self._cols[name][:self._n][mask]   # pylint: disable=protected-access
    return out

  def __getattr__(self, name):
    cols = self.__dict__.get('_cols', {})
    if name in cols:
      return cols[name][:self._n]
    raise AttributeError(name)


class APMetrics(metrics_lib.BaseMetric):
  """ref :128."""

  @classmethod
  def Params(cls, metadata):
    p = hyperparams.InstantiableParams(cls)
    p.Define('metadata', metadata, 'EvaluationMetadata of the dataset.')
    p.Define('breakdown_metrics', [], "Any of 'distance', 'num_points', 'rotation'.")
    p.Define('metric_weights', None,
             '{difficulty: weights over EvalClassIndices} for the scalar `value`.')
    p.Define('box_type', '3d', "'3d' (volume IoU) or '2d' (bird's-eye view).")
    p.Define('ap_algorithm', 'KITTI', "'KITTI' (N-point interpolation) or 'VOC'.")
    return p

  def __init__(self, params):
    self.params = params.Copy()
    p = self.params
    self.metadata = p.metadata
    assert p.box_type in ('2d', '3d')
    self._groundtruth, self._prediction = {}, {}
    self._str_to_imgid = {}
    self._iou_thresholds = self.metadata.IoUThresholds()
    bp = breakdown_metric.ByDifficulty.Params().Set(metadata=self.metadata)
    self._breakdown_metrics = {'difficulty': breakdown_metric.ByDifficulty(bp)}
    for name in p.breakdown_metrics:
      cls = breakdown_metric.ByName(name)
      self._breakdown_metrics[name] = cls(cls.Params().Set(metadata=self.metadata))
    self._is_eval_complete = False

  # -- bookkeeping ------------------------------------------------------------------
  def _GetImageId(self, str_id):
    return self._str_to_imgid.setdefault(str_id, len(self._str_to_imgid))

  def _Boxes(self, store, classid):
    if classid not in store:
      store[classid] = Boxes3D()
    return store[classid]

  def Update(self, str_id, result):
    """`result`: groundtruth_{labels [G], bboxes [G,7], difficulties [G], num_points [G],
    speed [G,2]?} and detection_{scores [C,N], boxes [C,N,7], heights_in_pixels [C,N]?}
    (ref :356)."""
    md = self.metadata
    img = self._GetImageId(str_id)
    labels = np.asarray(result.groundtruth_labels, np.int32).reshape(-1)
    g = len(labels)
    bboxes = np.asarray(result.groundtruth_bboxes, np.float32).reshape(g, 7)
    diff = np.asarray(result.groundtruth_difficulties, np.int32).reshape(-1)
    npts = np.asarray(result.get('groundtruth_num_points', np.zeros(g)), np.int32).reshape(-1)
    speed = np.asarray(result.get('groundtruth_speed', np.zeros((g, 2))), np.float32).reshape(g, 2)
    gt_view = NestedMap(bboxes=bboxes, num_points=npts, difficulties=diff, labels=labels)
    for m in self._breakdown_metrics.values():
      m.AccumulateHistogram(gt_view)
      m.AccumulateCumulative(gt_view)
    zeros = np.zeros(g, np.int32)
    bins = {k: (self._breakdown_metrics[k].Discretize(v) if k in self._breakdown_metrics
                else zeros)
            for k, v in (('distance', bboxes), ('num_points', npts), ('rotation', bboxes))}
    for c in np.unique(labels):
      if not 0 < c < md.NumClasses():
        continue
      sel = labels == c
      self._Boxes(self._groundtruth, int(c)).Extend(
          imgids=np.full(sel.sum(), img), scores=np.ones(sel.sum()), boxes=bboxes[sel],
          difficulties=diff[sel], distances=bins['distance'][sel],
          num_points=bins['num_points'][sel], rotations=bins['rotation'][sel],
          heights_in_pixels=np.full(sel.sum(), -1.0), speeds=speed[sel])
    scores = np.asarray(result.detection_scores, np.float32)
    assert scores.shape[0] == md.NumClasses(), '%s vs. %s' % (scores.shape[0], md.NumClasses())
    det_boxes = np.asarray(result.detection_boxes, np.float32)
    heights = np.asarray(result.get('detection_heights_in_pixels', np.zeros_like(scores)),
                         np.float32)
    for c in range(1, md.NumClasses()):
      keep = scores[c] > 0
      k = int(keep.sum())
      if not k:
        continue
      b = det_boxes[c][keep]
      zk = np.zeros(k, np.int32)
      self._Boxes(self._prediction, c).Extend(
          imgids=np.full(k, img), scores=scores[c][keep], boxes=b, difficulties=zk,
          distances=(self._breakdown_metrics['distance'].Discretize(b)
                     if 'distance' in self._breakdown_metrics else zk),
          num_points=zk,
          rotations=(self._breakdown_metrics['rotation'].Discretize(b)
                     if 'rotation' in self._breakdown_metrics else zk),
          heights_in_pixels=heights[c][keep], speeds=np.zeros((k, 2)))
    self._is_eval_complete = False

  # -- data selection ---------------------------------------------------------------
  def _LoadBoundingBoxes(self, box_type, class_id, distance=None, num_points=None,
                         rotation=None):
    store = self._groundtruth if box_type == 'groundtruth' else self._prediction
    boxes = store.get(class_id)
    if boxes is None or not len(boxes):
      return None
    mask = np.ones(len(boxes), bool)
    if distance is not None:
      mask &= boxes.distances == distance
    if num_points is not None:
      mask &= boxes.num_points == num_points
    if rotation is not None:
      mask &= boxes.rotations == rotation
    if mask.all():
      return boxes
    return boxes.Select(mask) if mask.any() else None

  def _GetData(self, classid, difficulty=None, distance=None, num_points=None, rotation=None):
    """Ground truth + predictions of one class (restricted to one breakdown bin); ground
    truth easier/harder than `difficulty` is flagged ignore rather than removed."""
    g = self._LoadBoundingBoxes('groundtruth', classid, distance, num_points, rotation)
    # predictions carry no point counts: never filtered by num_points
    p = self._LoadBoundingBoxes('prediction', classid, distance, None, rotation)
    if g is None or p is None:
      return None
    gt_ignore = np.zeros(len(g), np.int32)
    if difficulty is not None:
      level = self.metadata.DifficultyLevels()[difficulty]
      gt_ignore = (g.difficulties != level).astype(np.int32)
    name = self.metadata.ClassNames()[classid]
    return NestedMap(
        iou_threshold=self._iou_thresholds[name],
        gt=NestedMap(imgid=g.imgids, bbox=g.boxes, ignore=gt_ignore),
        pd=NestedMap(imgid=p.imgids, bbox=p.boxes, score=p.scores,
                     ignore=np.zeros(len(p), np.int32)))

  def _Flatten2D(self, boxes):
    """Bird's-eye-view evaluation: collapse z so volume IoU equals area IoU."""
    b = np.array(boxes, np.float32, copy=True)
    b[:, 2], b[:, 5] = 0.0, 1.0
    return b

  def _BuildMetric(self, feed_data, classid):
    """→ (ap, precision_recall [pts, 2]) via the native op."""
    del classid
    p = self.params
    gt_b, pd_b = feed_data.gt.bbox, feed_data.pd.bbox
    if p.box_type == '2d':
      gt_b, pd_b = self._Flatten2D(gt_b), self._Flatten2D(pd_b)
    ap, pr, _ = car_ops.average_precision3d(
        feed_data.iou_threshold, gt_b, feed_data.gt.imgid, feed_data.gt.ignore, pd_b,
        feed_data.pd.imgid, feed_data.pd.ignore, feed_data.pd.score,
        num_recall_points=self.metadata.NumberOfPrecisionRecallPoints(),
        algorithm=p.ap_algorithm)
    return ap, pr.numpy()

  def _ComputeFinalMetrics(self, classids=None, difficulty=None, distance=None,
                           num_points=None, rotation=None):
    """(ap [n_eval], pr [n_eval, pts, 2]) or None when the selection is empty."""
    classids = classids or self.metadata.EvalClassIndices()
    pts = self.metadata.NumberOfPrecisionRecallPoints()
    aps = np.full(len(classids), np.nan, np.float32)
    prs = np.zeros((len(classids), pts, 2), np.float32)
    any_data = False
    for i, c in enumerate(classids):
      data = self._GetData(c, difficulty, distance, num_points, rotation)
      if data is None:
        continue
      any_data = True
      aps[i], prs[i] = self._BuildMetric(data, c)
    return (aps, prs) if any_data else None

  def _EvaluateIfNecessary(self):
    if self._is_eval_complete:
      return
    for m in self._breakdown_metrics.values():
      m.ComputeMetrics(self._ComputeFinalMetrics)
    self._is_eval_complete = True

  # -- BaseMetric API ---------------------------------------------------------------
  def _AveragePrecisionByDifficulty(self):
    self._EvaluateIfNecessary()
    m = self._breakdown_metrics['difficulty']
    return dict(zip(m.BinLabels(), m._average_precisions))   # pylint: disable=protected-access

  @property
  def value(self):
    """(Weighted) mean AP over the evaluated classes at the default difficulty."""
    aps = self._AveragePrecisionByDifficulty()
    w = self.params.metric_weights
    if w:
      total = 0.0
      for level, weights in w.items():
        total += float(np.nansum(aps[level] * np.asarray(weights)))
      return total
    v = aps['default']
    return float(np.nanmean(v)) if np.any(~np.isnan(v)) else 0.0

  def Scalars(self, name):
    """All scalar summaries {tag: value}."""
    self._EvaluateIfNecessary()
    out = {name + '/mAP': self.value}
    for m in self._breakdown_metrics.values():
      out.update(m.Scalars(name))
    return out

  def Summary(self, name):
    return self.Scalars(name)
