"""Dataset facts the 3-D detection metrics need (ref
`lingvo/tasks/car/evaluation_metadata.py`).

Subclasses fill a `SPEC` dict; the accessor methods keep the reference's names so the
metric code is dataset-agnostic.
"""

from __future__ import annotations

import math


class EvaluationMetadata:

  SPEC = {}

  def __init__(self, name):
    self.name = name

  def _Get(self, key):
    if key not in self.SPEC:
      raise NotImplementedError('%s.%s' % (type(self).__name__, key))
    return self.SPEC[key]

  def ClassNames(self):
    return list(self._Get('class_names'))

  def LabelMap(self):
    return dict(enumerate(self.ClassNames()))

  def NumClasses(self):
    return len(self.ClassNames())

  def DifficultyLevels(self):
    return dict(self._Get('difficulty_levels'))

  def IoUThresholds(self):
    return dict(self._Get('iou_thresholds'))

  def EvalClassIndices(self):
    names = self.ClassNames()
    order = self.SPEC.get('eval_classes') or sorted(self.IoUThresholds())
    return [names.index(n) for n in order]

  def IgnoreClassIndices(self):
    names = self.ClassNames()
    return {names.index(k): [names.index(v) for v in vs]
            for k, vs in self._Get('ignore_neighbors').items()}

  def NumberOfPrecisionRecallPoints(self):
    return self._Get('pr_points')

  def MaximumDistance(self):
    return self._Get('max_distance')

  def DistanceBinWidth(self):
    return self._Get('distance_bin_width')

  def MaximumNumberOfPoints(self):
    return self._Get('max_num_points')

  def NumberOfPointsBins(self):
    return self._Get('num_points_bins')

  def MaximumRotation(self):
    return math.pi

  def NumberOfRotationBins(self):
    return self._Get('rotation_bins')

  def NumberOfCalibrationBins(self):
    return self.SPEC.get('calibration_bins', 15)

  def MinHeight2D(self):
    return dict(self._Get('min_height_2d'))

  def RecallAtPrecision(self):
    return [0.50, 0.95]
