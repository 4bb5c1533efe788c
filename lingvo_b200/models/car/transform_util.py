"""2-D/3-D transform helpers for top-down rendering (ref
`lingvo/tasks/car/transform_util.py`)."""

from __future__ import annotations

import math

import numpy as np


class Box2D:
  """An oriented rectangle in the ground plane (ref :22): centre, width (along heading),
  height (across), heading in radians."""

  def __init__(self, x, y, width, height, angle):
    self._center = np.array([x, y], np.float64)
    self._width, self._height, self._angle = float(width), float(height), float(angle)
    self._corners = self._ComputeCorners()

  def _ComputeCorners(self):
    c, s = math.cos(self._angle), math.sin(self._angle)
    ax = np.array([c, s]) * self._width / 2
    ay = np.array([-s, c]) * self._height / 2
    ctr = self._center
    return np.stack([ctr - ax - ay, ctr + ax - ay, ctr + ax + ay, ctr - ax + ay])

  @property
  def center(self):
    return self._center

  @property
  def corners(self):
    return self._corners

  @property
  def width(self):
    return self._width

  @property
  def height(self):
    return self._height

  @property
  def angle(self):
    return self._angle

  def Extrema(self):
    """(min_x, max_x, min_y, max_y) of the corners."""
    c = self._corners
    return c[:, 0].min(), c[:, 0].max(), c[:, 1].min(), c[:, 1].max()

  def Apply(self, transform):
    """New Box2D under a 4×4 transform (scale taken from the transform's x/y axes)."""
    x, y, _ = TransformPoint(transform, self._center[0], self._center[1], 0.0)
    sx = math.hypot(transform[0, 0], transform[1, 0])
    sy = math.hypot(transform[0, 1], transform[1, 1])
    return Box2D(x, y, self._width * sx, self._height * sy,
                 TransformHeading(transform, self._angle))


def TransformHeading(transform, heading):
  """Heading of the unit vector (cos h, sin h) after applying `transform` (ref :119)."""
  x, y = math.cos(heading), math.sin(heading)
  x0, y0, _ = TransformPoint(transform, 0.0, 0.0, 0.0)
  x1, y1, _ = TransformPoint(transform, x, y, 0.0)
  return math.atan2(y1 - y0, x1 - x0)


def TransformPoint(transform, x, y, z):
  """Applies a 4×4 homogeneous transform to one point (ref :153)."""
  v = np.asarray(transform, np.float64) @ np.array([x, y, z, 1.0])
  return v[0], v[1], v[2]


def CopyTransform(transform):
  return np.array(transform, np.float64, copy=True)


def MakeCarToImageTransform(pixels_per_meter, image_ref_x, image_ref_y, flip_axes=True):
  """Car frame (x forward, y left) → top-down image pixels: scale by `pixels_per_meter`,
  optionally swap/flip so that forward is up, then translate the car origin to
  (image_ref_x, image_ref_y) (ref :164)."""
  ppm = float(pixels_per_meter)
  if flip_axes:
    # image x = −car y, image y = −car x
    m = np.array([[0.0, -ppm, 0.0, image_ref_x],
                  [-ppm, 0.0, 0.0, image_ref_y],
                  [0.0, 0.0, 1.0, 0.0],
                  [0.0, 0.0, 0.0, 1.0]])
  else:
    m = np.array([[ppm, 0.0, 0.0, image_ref_x],
                  [0.0, ppm, 0.0, image_ref_y],
                  [0.0, 0.0, 1.0, 0.0],
                  [0.0, 0.0, 0.0, 1.0]])
  return m
