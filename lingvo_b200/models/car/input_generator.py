"""Synthetic 3-D detection input (stand-in for the KITTI/Waymo pipelines of
`lingvo/tasks/car/{kitti,waymo}_input_generator.py`, which need the datasets):
random scenes with box-shaped point clusters, pre-pillarised with the native op."""

from __future__ import annotations

import numpy as np
import torch

from lingvo_b200.core import base_input_generator
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.car import pillars


class SyntheticPillarsInput(base_input_generator.BaseInputGenerator):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('pillars', pillars.PointsToPillars.Params(), 'Pillar grid.')
    p.Define('max_boxes', 4, 'GT boxes per scene (padded).')
    p.Define('points_per_box', 60, 'Lidar returns per object.')
    p.Define('clutter_points', 200, 'Background returns.')
    p.Define('seed', 0, 'RNG seed.')
    p.batch_size = 4
    return p

  def __init__(self, params):
    super().__init__(params)
    self.CreateChild('pillars', self.params.pillars)
    self._rng = np.random.RandomState(self.params.seed)

  def _Scene(self):
    p = self.params
    gx, gy = p.pillars.grid_x, p.pillars.grid_y
    n = self._rng.randint(1, p.max_boxes + 1)
    boxes = np.zeros((p.max_boxes, 7), np.float32)
    pts = [np.concatenate([
        self._rng.uniform([gx[0], gy[0], -2.0], [gx[1], gy[1], -1.6], (p.clutter_points, 3)),
        self._rng.uniform(0, 0.2, (p.clutter_points, 1))], 1)]
    for i in range(n):
      c = self._rng.uniform([gx[0] * 0.8, gy[0] * 0.8, -1.0], [gx[1] * 0.8, gy[1] * 0.8, -0.8])
      phi = self._rng.choice([0.0, np.pi / 2]) + self._rng.uniform(-0.1, 0.1)
      d = np.array([3.9, 1.6, 1.56]) * self._rng.uniform(0.9, 1.1, 3)
      boxes[i] = [*c, *d, phi]
      local = self._rng.uniform(-0.5, 0.5, (p.points_per_box, 3)) * d
      rot = np.array([[np.cos(phi), -np.sin(phi)], [np.sin(phi), np.cos(phi)]])
      xy = local[:, :2] @ rot.T + c[:2]
      pts.append(np.concatenate([xy, local[:, 2:] + c[2], np.full((p.points_per_box, 1), 0.9)], 1))
    mask = np.zeros(p.max_boxes, np.float32)
    mask[:n] = 1
    labels = np.zeros(p.max_boxes, np.int64)
    labels[:n] = 1
    return np.concatenate(pts).astype(np.float32), boxes, labels, mask

  def _InputBatch(self):
    p = self.params
    outs = []
    for _ in range(p.batch_size):
      pts, boxes, labels, mask = self._Scene()
      pil = self.pillars.FProp(None, torch.from_numpy(pts))
      outs.append(NestedMap(pillar_points=pil.pillar_points, pillar_locations=pil.pillar_locations,
                            pillar_count=pil.pillar_count, bboxes=torch.from_numpy(boxes),
                            labels=torch.from_numpy(labels), bboxes_mask=torch.from_numpy(mask)))
    return outs[0].Pack([torch.stack(x) for x in zip(*[o.Flatten() for o in outs])])
