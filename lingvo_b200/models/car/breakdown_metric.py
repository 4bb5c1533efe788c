"""Precision/recall broken down by a property of the ground truth (ref
`lingvo/tasks/car/breakdown_metric.py`): difficulty, distance from the sensor, number of
laser points in the box, heading.

A `BreakdownMetric` bins ground-truth boxes (`Discretize`), keeps histograms of what the
eval set contains, and at the end asks the owning `APMetrics` — through
`compute_metrics_fn(**bin_selector)` — for the AP / PR curve restricted to each bin.
Results live in `_average_precisions[bin][class]`, `_precision_recall[bin][class]`.
"""

from __future__ import annotations

import numpy as np

from lingvo_b200.core import hyperparams
from lingvo_b200.core import plot


def _FindRecallAtGivenPrecision(precision_recall, precision_level):
  """`[n, m, 2]` (precision, recall) curves → best recall with precision ≥ level, per
  class (0 when never reached)."""
  assert precision_recall.ndim == 3 and precision_recall.shape[-1] == 2
  assert 0.0 < precision_level < 1.0
  ok = precision_recall[..., 0] >= precision_level
  return np.where(ok, precision_recall[..., 1], 0.0).max(1).astype(np.float32)


def _FindMaximumRecall(precision_recall):
  """Largest recall reached with non-zero precision, per class."""
  assert precision_recall.ndim == 3 and precision_recall.shape[-1] == 2
  ok = precision_recall[..., 0] > 0.0
  return np.where(ok, precision_recall[..., 1], 0.0).max(1).astype(np.float32)


class BreakdownMetric:
  """Base class (ref :94)."""

  @classmethod
  def Params(cls):
    p = hyperparams.InstantiableParams(cls)
    p.Define('metadata', None, 'EvaluationMetadata of the dataset.')
    return p

  SELECTOR = None          # keyword of APMetrics._GetData that selects one bin

  def __init__(self, p):
    self.params = p
    self._meta = p.metadata
    n_bins, n_cls = self.NumBinsOfHistogram(), p.metadata.NumClasses()
    n_eval = len(p.metadata.EvalClassIndices())
    pr_pts = p.metadata.NumberOfPrecisionRecallPoints()
    self._histogram = np.zeros((n_bins, n_cls), np.int64)
    self._values = np.zeros((n_bins, 1), np.float32)
    self._average_precisions = np.full((n_bins, n_eval), np.nan, np.float32)
    self._precision_recall = np.zeros((n_bins, n_eval, pr_pts, 2), np.float32)
    self._recall_at_precision = {
        lvl: np.zeros((n_bins, n_eval), np.float32) for lvl in p.metadata.RecallAtPrecision()}
    self._max_recall = np.zeros((n_bins, n_eval), np.float32)

  def NumBinsOfHistogram(self):
    raise NotImplementedError()

  def Discretize(self, values):
    raise NotImplementedError()

  def AccumulateHistogram(self, result):
    """Counts this example's ground-truth boxes per (bin, class)."""
    raise NotImplementedError()

  def AccumulateCumulative(self, result):
    """Optional running statistics (e.g. a CDF); default none."""

  def _Accumulate(self, bins, labels):
    np.add.at(self._histogram, (np.asarray(bins, np.int64), np.asarray(labels, np.int64)), 1)

  def _BinSelector(self, b):
    return {self.SELECTOR: b}

  def ComputeMetrics(self, compute_metrics_fn):
    """Fills the per-bin AP tables; bins without ground truth stay NaN."""
    for b in range(self.NumBinsOfHistogram()):
      m = compute_metrics_fn(**self._BinSelector(b))
      if m is None:
        continue
      ap, pr = m
      self._average_precisions[b] = ap
      self._precision_recall[b] = pr
      self._max_recall[b] = _FindMaximumRecall(pr)
      for lvl in self._recall_at_precision:
        self._recall_at_precision[lvl][b] = _FindRecallAtGivenPrecision(pr, lvl)

  def BinLabels(self):
    return [str(b) for b in range(self.NumBinsOfHistogram())]

  def Scalars(self, name):
    """{summary tag: value} for every (bin, evaluated class) with data."""
    out = {}
    names = self._meta.ClassNames()
    for ci, c in enumerate(self._meta.EvalClassIndices()):
      for b, lab in enumerate(self.BinLabels()):
        ap = self._average_precisions[b, ci]
        if not np.isnan(ap):
          out['%s/%s/%s_%s' % (name, type(self).__name__, names[c], lab)] = float(ap)
    return out

  def GenerateSummaries(self, name):
    """AP-per-bin plots as image summaries `[(tag, HxWx3 uint8)]` plus the scalars."""
    names = self._meta.ClassNames()
    images = []
    for ci, c in enumerate(self._meta.EvalClassIndices()):
      ys = self._average_precisions[:, ci]
      if np.all(np.isnan(ys)):
        continue
      def _Setter(fig, axes, ys=ys, c=c):
        xs = np.arange(len(ys))
        axes.bar(xs, np.nan_to_num(ys))
        axes.set_xticks(xs)
        axes.set_xticklabels(self.BinLabels(), rotation=45, fontsize=6)
        axes.set_ylim(0, 1)
        axes.set_title('%s AP by %s' % (names[c], type(self).__name__))
      tag = '%s/%s/%s' % (name, type(self).__name__, names[c])
      png = plot.Custom(tag, (5, 3), _Setter)       # None when matplotlib is not installed
      if png is not None:
        images.append((tag, png))
    return self.Scalars(name), images


class ByDifficulty(BreakdownMetric):
  """One bin per difficulty level; the default AP table (ref :653)."""

  SELECTOR = 'difficulty'

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('ap_key', 'ap', 'Kept for parity.')
    return p

  def _Levels(self):
    # ordered by level id so bin index ↔ level is stable
    return [k for k, _ in sorted(self._meta.DifficultyLevels().items(), key=lambda kv: kv[1])]

  def NumBinsOfHistogram(self):
    return len(self._meta.DifficultyLevels()) + 1        # + "all"

  def BinLabels(self):
    return self._Levels() + ['default']

  def Discretize(self, difficulties):
    ids = sorted(self._meta.DifficultyLevels().values())
    return np.searchsorted(ids, np.asarray(difficulties)).clip(0, len(ids) - 1)

  def AccumulateHistogram(self, result):
    self._Accumulate(self.Discretize(result.difficulties), result.labels)

  def _BinSelector(self, b):
    labels = self.BinLabels()
    return {'difficulty': None if labels[b] == 'default' else labels[b]}


class ByDistance(BreakdownMetric):
  """Bins of `DistanceBinWidth` metres of ground-plane distance (ref :241)."""

  SELECTOR = 'distance'

  def NumBinsOfHistogram(self):
    m = self._meta
    return int(np.ceil(m.MaximumDistance() / m.DistanceBinWidth()))

  def BinLabels(self):
    w = self._meta.DistanceBinWidth()
    return ['%g-%gm' % (b * w, (b + 1) * w) for b in range(self.NumBinsOfHistogram())]

  @classmethod
  def _CalculateEuclideanDistanceFromOrigin(cls, bboxes):
    return np.linalg.norm(np.asarray(bboxes, np.float32).reshape(-1, 7)[:, :3], axis=1)

  def Discretize(self, bboxes):
    d = self._CalculateEuclideanDistanceFromOrigin(bboxes)
    return np.minimum((d / self._meta.DistanceBinWidth()).astype(np.int64),
                      self.NumBinsOfHistogram() - 1)

  def AccumulateHistogram(self, result):
    self._Accumulate(self.Discretize(result.bboxes), result.labels)


class ByNumPoints(BreakdownMetric):
  """Log-spaced bins of the number of laser points inside the box (ref :357)."""

  SELECTOR = 'num_points'

  def NumBinsOfHistogram(self):
    return self._meta.NumberOfPointsBins()

  def _LogSpacedBinEdgesofPoints(self):
    return np.logspace(0, np.log10(self._meta.MaximumNumberOfPoints()),
                       self.NumBinsOfHistogram() + 1)

  def BinLabels(self):
    e = self._LogSpacedBinEdgesofPoints()
    return ['%d-%d' % (e[i], e[i + 1]) for i in range(self.NumBinsOfHistogram())]

  def Discretize(self, num_points):
    e = self._LogSpacedBinEdgesofPoints()
    b = np.digitize(np.asarray(num_points, np.float64), e[1:-1])
    return b.clip(0, self.NumBinsOfHistogram() - 1)

  def AccumulateHistogram(self, result):
    self._Accumulate(self.Discretize(result.num_points), result.labels)

  def AccumulateCumulative(self, result):
    n = np.asarray(result.num_points)
    self._values[:, 0] += np.bincount(self.Discretize(n), weights=n,
                                      minlength=self.NumBinsOfHistogram())


class ByRotation(BreakdownMetric):
  """Linear bins of |heading| folded into [0, π) (ref :523)."""

  SELECTOR = 'rotation'

  def NumBinsOfHistogram(self):
    return self._meta.NumberOfRotationBins()

  def BinLabels(self):
    w = self._meta.MaximumRotation() / self.NumBinsOfHistogram()
    return ['%.0f°' % np.degrees((b + 0.5) * w) for b in range(self.NumBinsOfHistogram())]

  def _CalculateRotation(self, bboxes):
    phi = np.asarray(bboxes, np.float32).reshape(-1, 7)[:, 6]
    return np.mod(phi, self._meta.MaximumRotation())

  def Discretize(self, bboxes):
    w = self._meta.MaximumRotation() / self.NumBinsOfHistogram()
    return np.minimum((self._CalculateRotation(bboxes) / w).astype(np.int64),
                      self.NumBinsOfHistogram() - 1)

  def AccumulateHistogram(self, result):
    self._Accumulate(self.Discretize(result.bboxes), result.labels)


_BY_NAME = {'difficulty': ByDifficulty, 'distance': ByDistance, 'num_points': ByNumPoints,
            'rotation': ByRotation}


def ByName(breakdown_metric_name):
  """'distance' → ByDistance … (ref :227)."""
  if breakdown_metric_name not in _BY_NAME:
    raise ValueError('Invalid breakdown name: %s, valid names are %s' % (
        breakdown_metric_name, sorted(_BY_NAME)))
  return _BY_NAME[breakdown_metric_name]
