"""Geometric routines on points and boxes (ref `lingvo/tasks/car/geometry.py`).

Conventions: 3-D boxes are `[x, y, z, dx, dy, dz, phi]` (centre, extents, heading around
+z); 2-D image boxes are `[ymin, xmin, ymax, xmax]`; transforms are 4×4 homogeneous
matrices acting on column vectors. Everything is batched torch code that runs on the
device of its inputs.
"""

from __future__ import annotations

import math

import torch


def _BroadcastMatmul(x, y):
  """`x [..., N, K] @ y [..., K, M]` with the batch dims of `y` broadcast to `x`'s."""
  while y.dim() < x.dim():
    y = y.unsqueeze(0)
  return torch.matmul(x, y)


def _MakeRotationMatrix(yaw, roll, pitch):
  """R = Rz(yaw) · Ry(pitch) · Rx(roll) for scalar angles → [3, 3] (ref :51)."""
  yaw, roll, pitch = (torch.as_tensor(a, dtype=torch.float32) for a in (yaw, roll, pitch))
  cy, sy, cr, sr, cp, sp = (torch.cos(yaw), torch.sin(yaw), torch.cos(roll), torch.sin(roll),
                            torch.cos(pitch), torch.sin(pitch))
  o, z = torch.ones(()), torch.zeros(())
  rz = torch.stack([cy, -sy, z, sy, cy, z, z, z, o]).reshape(3, 3)
  ry = torch.stack([cp, z, sp, z, o, z, -sp, z, cp]).reshape(3, 3)
  rx = torch.stack([o, z, z, z, cr, -sr, z, sr, cr]).reshape(3, 3)
  return rz @ ry @ rx


def BatchMakeRotationMatrix(yaw, clockwise=False):
  """yaw `[...]` → z-rotation matrices `[..., 3, 3]` (ref :87)."""
  c, s = torch.cos(yaw), torch.sin(yaw)
  if clockwise:
    s = -s
  o, z = torch.ones_like(c), torch.zeros_like(c)
  return torch.stack([c, -s, z, s, c, z, z, z, o], -1).reshape(yaw.shape + (3, 3))


def CoordinateTransform(points, pose):
  """Rotates `points [..., 3]` by (yaw, roll, pitch) of `pose [6] = x,y,z,yaw,roll,pitch`
  and translates by (x, y, z) (ref :118)."""
  rot = _MakeRotationMatrix(pose[3], pose[4], pose[5]).to(points)
  return torch.matmul(points, rot.t()) + pose[:3].to(points)


def TransformPoints(points, transforms):
  """points `[..., N, 3]`, transforms `[..., 4, 4]` → `[..., N, 3]` (ref :154)."""
  ones = torch.ones_like(points[..., :1])
  hom = torch.cat([points, ones], -1)
  out = _BroadcastMatmul(hom, transforms.transpose(-1, -2).to(points))
  return out[..., :3]


def WrapAngleRad(angles_rad, min_val=-math.pi, max_val=math.pi):
  """Wraps angles into [min_val, max_val) (ref :176)."""
  span = max_val - min_val
  return torch.remainder(angles_rad - min_val, span) + min_val


def TransformBBoxes3D(bboxes_3d, transforms):
  """Applies rigid transforms to 7-DOF boxes: centres move, headings gain the transform's
  yaw, extents stay (ref :182)."""
  centers = TransformPoints(bboxes_3d[..., :3], transforms)
  t = transforms.to(bboxes_3d)
  yaw = torch.atan2(t[..., 1, 0], t[..., 0, 0]).unsqueeze(-1)
  phi = WrapAngleRad(bboxes_3d[..., 6] + yaw)
  return torch.cat([centers, bboxes_3d[..., 3:6], phi.unsqueeze(-1)], -1)


def XYWHToBBoxes(xywh):
  """centre-x, centre-y, w, h → ymin, xmin, ymax, xmax (ref :212)."""
  x, y, w, h = xywh.unbind(-1)
  return torch.stack([y - h / 2, x - w / 2, y + h / 2, x + w / 2], -1)


def BBoxesToXYWH(bboxes):
  """ymin, xmin, ymax, xmax → centre-x, centre-y, w, h (ref :256)."""
  ymin, xmin, ymax, xmax = bboxes.unbind(-1)
  return torch.stack([(xmin + xmax) / 2, (ymin + ymax) / 2, xmax - xmin, ymax - ymin], -1)


def PointsToImagePlane(points, velo_to_image_plane):
  """Projects lidar points `[N, 3]` with a `[3, 4]` projection → `[N, 3]` =
  (u, v, depth) (ref :226)."""
  hom = torch.cat([points, torch.ones_like(points[:, :1])], -1)
  proj = hom @ velo_to_image_plane.to(points).t()
  depth = proj[:, 2:3]
  return torch.cat([proj[:, :2] / depth, depth], -1)


def BBoxesCentroid(bboxes):
  """Centre (x, y) of ymin/xmin/ymax/xmax boxes (ref :270)."""
  return BBoxesToXYWH(bboxes)[..., :2]


def ReorderIndicesByPhi(anchor, bboxes):
  """Permutation ordering boxes by the signed angle of their centroid relative to the
  direction of `anchor (x0, y0)`, counter-clockwise side first (ref :282)."""
  n = bboxes.shape[0]
  if n == 0:
    return torch.zeros(0, dtype=torch.long, device=bboxes.device)
  c = BBoxesCentroid(bboxes)
  a = anchor.to(c)
  norm = a.norm() * c.norm(dim=1)
  cosine = torch.where(norm > 0, (c @ a) / norm.clamp_min(1e-30), torch.zeros_like(norm))
  cross_z = a[0] * c[:, 1] - a[1] * c[:, 0]
  score = torch.where(cross_z > 0, -1 - cosine, 1 + cosine)
  return torch.argsort(score, descending=True, stable=True)


def _SmoothL1Norm(a):
  return torch.where(a.abs() < 1, 0.5 * a * a, a.abs() - 0.5)


def DistanceBetweenCentroidsAndBBoxesFastAndFurious(centroids, bboxes, masks):
  """Smooth-L1 of size-normalised centre offsets and log size ratios between predicted
  x/y/w/h and ground-truth boxes ("Fast and Furious", Luo et al. 2018) (ref :346)."""
  x, y, w, h = centroids.unbind(-1)
  xg, yg, wg, hg = BBoxesToXYWH(bboxes).unbind(-1)
  pos = lambda t: t.clamp_min(1e-8)
  terms = (masks * (x - xg) / pos(wg), masks * (y - yg) / pos(hg),
           masks * torch.log(pos(w) / pos(wg)), masks * torch.log(pos(h) / pos(hg)))
  return sum(_SmoothL1Norm(t) for t in terms)


def DistanceBetweenCentroids(u, v, masks):
  """Masked smooth-L1 distance between x/y/w/h vectors (ref :379)."""
  return masks * _SmoothL1Norm(u - v).sum(-1)


def _IsOnLeftHandSideOrOn(point, v1, v2):
  """`point [..., 2]` is on or to the left of the directed edge v1→v2."""
  d = v2 - v1
  rel = point - v1
  return (d[..., 0] * rel[..., 1] - d[..., 1] * rel[..., 0]) >= 0


def _IsCounterClockwiseDirection(v1, v2, v3):
  return ((v2[..., 0] - v1[..., 0]) * (v3[..., 1] - v2[..., 1]) -
          (v2[..., 1] - v1[..., 1]) * (v3[..., 0] - v2[..., 0])) > 0


def _BBoxArea(bbox):
  """Shoelace area of `[..., 4, 2]` corner loops."""
  x, y = bbox[..., 0], bbox[..., 1]
  return 0.5 * (x * y.roll(-1, -1) - y * x.roll(-1, -1)).sum(-1).abs()


def IsWithinBBox(points, bbox):
  """points `[..., N, 2]` inside (or on) the convex quadrilateral `bbox [..., 4, 2]` given
  in either winding order → bool `[..., N]` (ref :474)."""
  v = [bbox[..., i, :].unsqueeze(-2) for i in range(4)]
  ccw = _IsCounterClockwiseDirection(bbox[..., 0, :], bbox[..., 1, :], bbox[..., 2, :])
  inside_ccw = torch.ones(points.shape[:-1], dtype=torch.bool, device=points.device)
  inside_cw = inside_ccw.clone()
  for i in range(4):
    a, b = v[i], v[(i + 1) % 4]
    inside_ccw &= _IsOnLeftHandSideOrOn(points, a, b)
    inside_cw &= _IsOnLeftHandSideOrOn(points, b, a)
  return torch.where(ccw.unsqueeze(-1), inside_ccw, inside_cw)


def BBoxCorners2D(bboxes):
  """`[..., 5] = x, y, dx, dy, phi` → corners `[..., 4, 2]` (ref :525)."""
  x, y, dx, dy, phi = bboxes.unbind(-1)
  c, s = torch.cos(phi), torch.sin(phi)
  sx = torch.tensor([0.5, -0.5, -0.5, 0.5], device=bboxes.device, dtype=bboxes.dtype)
  sy = torch.tensor([0.5, 0.5, -0.5, -0.5], device=bboxes.device, dtype=bboxes.dtype)
  lx, ly = dx.unsqueeze(-1) * sx, dy.unsqueeze(-1) * sy
  cx = x.unsqueeze(-1) + lx * c.unsqueeze(-1) - ly * s.unsqueeze(-1)
  cy = y.unsqueeze(-1) + lx * s.unsqueeze(-1) + ly * c.unsqueeze(-1)
  return torch.stack([cx, cy], -1)


def BBoxCorners(bboxes):
  """7-DOF boxes `[..., 7]` → corners `[..., 8, 3]`: the 4 top corners then the 4 bottom
  ones, each loop in the `BBoxCorners2D` order (ref :567)."""
  flat = BBoxCorners2D(torch.cat([bboxes[..., 0:2], bboxes[..., 3:5], bboxes[..., 6:7]], -1))
  z, dz = bboxes[..., 2], bboxes[..., 5]
  top = torch.cat([flat, (z + dz / 2).unsqueeze(-1).unsqueeze(-1).expand(flat.shape[:-1] + (1,))], -1)
  bot = torch.cat([flat, (z - dz / 2).unsqueeze(-1).unsqueeze(-1).expand(flat.shape[:-1] + (1,))], -1)
  return torch.cat([top, bot], -2)


def IsWithinBBox3D(points_3d, bboxes_3d):
  """points `[N, 3]`, boxes `[M, 7]` → bool `[N, M]` (ref :621). Works in each box's local
  frame: |R(-phi)(p − c)| ≤ extent / 2."""
  rel = points_3d[:, None, :] - bboxes_3d[None, :, :3]
  c, s = torch.cos(bboxes_3d[:, 6]), torch.sin(bboxes_3d[:, 6])
  lx = rel[..., 0] * c + rel[..., 1] * s
  ly = -rel[..., 0] * s + rel[..., 1] * c
  half = bboxes_3d[None, :, 3:6] / 2
  return ((lx.abs() <= half[..., 0]) & (ly.abs() <= half[..., 1]) &
          (rel[..., 2].abs() <= half[..., 2]))


def SphericalCoordinatesTransform(points_xyz):
  """xyz → (distance, inclination θ from +z, azimuth φ) (ref :685)."""
  dist = points_xyz.norm(dim=-1)
  theta = torch.acos((points_xyz[..., 2] / dist.clamp_min(1e-7)).clamp(-1, 1))
  phi = torch.atan2(points_xyz[..., 1], points_xyz[..., 0])
  return torch.stack([dist, theta, phi], -1)


def TargetTransforms(original_transforms, target_transform):
  """`[F, 4, 4]` frame→world poses and one target pose → frame→target transforms
  (ref :708)."""
  inv = torch.linalg.inv(target_transform.to(original_transforms))
  return torch.matmul(inv.unsqueeze(0), original_transforms)
