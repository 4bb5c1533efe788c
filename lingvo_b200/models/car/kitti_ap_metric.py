"""KITTI flavour of the AP metric (ref `lingvo/tasks/car/kitti_ap_metric.py`).

KITTI-specific rules on top of `APMetrics`:
  * difficulty d evaluates boxes at least as easy as d — harder ones are *ignored on first
    match* (not false negatives, and a detection matching them is not a false positive);
  * detections whose 2-D height is below the level's minimum never count as FP;
  * neighbour classes (Van for Car, Person_sitting for Pedestrian) are ignore-first-match,
    `DontCare` regions ignore every match.
"""

from __future__ import annotations

import numpy as np

from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.car import ap_metric


class KITTIAPMetrics(ap_metric.APMetrics):

  def _GetData(self, classid, difficulty=None, distance=None, num_points=None, rotation=None):
    md = self.metadata
    g = self._LoadBoundingBoxes('groundtruth', classid, distance, num_points, rotation)
    p = self._LoadBoundingBoxes('prediction', classid, distance, None, rotation)
    if g is None or p is None:
      return None
    min_height = md.MinHeight2D()[difficulty] if difficulty else 0
    pd_ignore = (p.heights_in_pixels < min_height).astype(np.int32)
    gt_img, gt_box = [g.imgids], [g.boxes]
    if difficulty:
      gt_ign = [(g.difficulties < md.DifficultyLevels()[difficulty]).astype(np.int32)]
    else:
      gt_ign = [np.zeros(len(g), np.int32)]
    def _AddIgnored(class_id, code):
      extra = self._LoadBoundingBoxes('groundtruth', class_id)
      if extra is not None:
        gt_img.append(extra.imgids)
        gt_box.append(extra.boxes)
        gt_ign.append(np.full(len(extra), code, np.int32))
    for neighbour in md.IgnoreClassIndices().get(classid, []):
      _AddIgnored(neighbour, 1)
    if 'DontCare' in md.ClassNames():
      _AddIgnored(md.ClassNames().index('DontCare'), 2)
    return NestedMap(
        iou_threshold=self._iou_thresholds[md.ClassNames()[classid]],
        gt=NestedMap(imgid=np.concatenate(gt_img), bbox=np.concatenate(gt_box),
                     ignore=np.concatenate(gt_ign)),
        pd=NestedMap(imgid=p.imgids, bbox=p.boxes, score=p.scores, ignore=pd_ignore))

  def Scalars(self, name):
    """Adds the conventional `AP_<class>_<difficulty>` tags."""
    out = super().Scalars(name)
    names = self.metadata.ClassNames()
    for level, aps in self._AveragePrecisionByDifficulty().items():
      for ci, c in enumerate(self.metadata.EvalClassIndices()):
        if not np.isnan(aps[ci]):
          out['%s/AP_%s_%s' % (name, names[c].lower(), level)] = float(aps[ci])
    return out
