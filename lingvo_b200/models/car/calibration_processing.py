"""Calibration of detection scores (ref `lingvo/tasks/car/calibration_processing.py`).

Expected calibration error and reliability curves from (score, hit) pairs — the
`score_and_hit` output of the AP op.
"""

from __future__ import annotations

import numpy as np

from lingvo_b200.core import plot


def ExpectedCalibrationError(confidence, empirical_accuracy, num_examples, min_confidence=None):
  """Σ_bins (n_b / N) · |accuracy_b − confidence_b| over bins with confidence ≥
  `min_confidence` (ref :22)."""
  confidence, empirical_accuracy = np.asarray(confidence, np.float64), np.asarray(
      empirical_accuracy, np.float64)
  num_examples = np.asarray(num_examples, np.float64)
  assert confidence.shape == empirical_accuracy.shape == num_examples.shape
  keep = num_examples > 0
  if min_confidence is not None:
    keep &= confidence >= min_confidence
  n = num_examples[keep].sum()
  if n == 0:
    return 0.0
  return float((num_examples[keep] / n * np.abs(empirical_accuracy[keep] - confidence[keep])).sum())


def CalibrationCurve(scores, hits, num_bins):
  """Equal-width score bins → (mean score, hit rate, count) per bin (ref :59)."""
  scores, hits = np.asarray(scores, np.float64), np.asarray(hits, np.float64)
  edges = np.linspace(0.0, 1.0, num_bins + 1)
  b = np.digitize(scores, edges[1:-1])
  count = np.bincount(b, minlength=num_bins).astype(np.float64)
  sum_s = np.bincount(b, weights=scores, minlength=num_bins)
  sum_h = np.bincount(b, weights=hits, minlength=num_bins)
  safe = np.maximum(count, 1.0)
  return sum_s / safe, sum_h / safe, count


class CalibrationCalculator:
  """Per-class reliability curves + ECE from an evaluated `APMetrics` (ref :111)."""

  def __init__(self, metadata):
    self._metadata = metadata
    self._num_bins = metadata.NumberOfCalibrationBins()
    self._results = {}

  def Calculate(self, metrics):
    """`metrics`: an APMetrics whose Update calls are done."""
    from lingvo_b200.models.car import ops as car_ops  # pylint: disable=g-import-not-at-top
    names = self._metadata.ClassNames()
    for c in self._metadata.EvalClassIndices():
      data = metrics._GetData(c)   # pylint: disable=protected-access
      if data is None:
        continue
      _, _, sh = car_ops.average_precision3d(
          data.iou_threshold, data.gt.bbox, data.gt.imgid, data.gt.ignore, data.pd.bbox,
          data.pd.imgid, data.pd.ignore, data.pd.score, num_recall_points=1)
      sh = sh.numpy()
      conf, acc, n = CalibrationCurve(sh[:, 0], sh[:, 1], self._num_bins)
      self._results[names[c]] = dict(mean_predicted_accuracies=conf,
                                     empirical_accuracies=acc, num_examples=n,
                                     ece=ExpectedCalibrationError(conf, acc, n))
    return self._results

  def Summary(self, name):
    """→ ({tag: ece}, [(tag, image)])."""
    scalars, images = {}, []
    for cls_name, r in self._results.items():
      scalars['%s/calibration_ece_%s' % (name, cls_name)] = r['ece']
      def _CalibrationSetter(fig, axes, r=r, cls_name=cls_name):
        axes.plot([0, 1], [0, 1], 'k--', linewidth=0.5)
        m = r['num_examples'] > 0
        axes.plot(r['mean_predicted_accuracies'][m], r['empirical_accuracies'][m], 'o-')
        axes.set_xlabel('score')
        axes.set_ylabel('hit rate')
        axes.set_title('%s (ECE %.3f)' % (cls_name, r['ece']))
      try:
        images.append(('%s/calibration_curve_%s' % (name, cls_name),
                       plot.Image(_CalibrationSetter, figsize=(4, 4))))
      except Exception:  # pylint: disable=broad-except
        pass
    return scalars, images
