"""Drawing helpers for detection summaries (ref `lingvo/tasks/car/summary.py`).

Top-down (bird's-eye) rasters of laser points and oriented boxes, boxes on camera images,
trajectory comparisons. Images are `uint8 [H, W, 3]` numpy arrays drawn with PIL.
"""

from __future__ import annotations

import math
import re

import numpy as np
from PIL import Image
from PIL import ImageColor
from PIL import ImageDraw

from lingvo_b200.models.car import transform_util

TOP_DOWN_SIZE = 1024
PIXELS_PER_METER = 32.0 / 3.0


def _PILColorList():
  """A stable list of distinguishable PIL colour names (ref :41)."""
  names = sorted(n for n in ImageColor.colormap
                 if not re.search(r'gr[ae]y|white|black|snow|ivory|linen', n))
  return names


PIL_COLOR_LIST = _PILColorList()


def ExtractRunIds(run_segments):
  """`'<run_id>_<segment>…'` byte strings → run-id strings (ref :50)."""
  out = []
  for s in run_segments:
    s = s.decode('utf-8') if isinstance(s, bytes) else str(s)
    out.append(s.split('_')[0] if s else '')
  return out


def CameraImageSummary(frontal_images, run_segment_strings, figsize=(6, 4)):
  """Camera frames annotated with their run ids → list of uint8 images (ref :67)."""
  del figsize
  out = []
  for img, run in zip(np.asarray(frontal_images), ExtractRunIds(run_segment_strings)):
    pil = Image.fromarray(_ToUint8(img))
    ImageDraw.Draw(pil).text((4, 4), run, fill='yellow')
    out.append(np.asarray(pil))
  return out


def _ToUint8(img):
  img = np.asarray(img)
  if img.dtype != np.uint8:
    img = (np.clip(img, 0.0, 1.0) * 255).astype(np.uint8) if img.max() <= 1.0 + 1e-6 else \
        np.clip(img, 0, 255).astype(np.uint8)
  if img.ndim == 2:
    img = np.stack([img] * 3, -1)
  return img


def _CarToImageTransform():
  return transform_util.MakeCarToImageTransform(
      pixels_per_meter=PIXELS_PER_METER, image_ref_x=TOP_DOWN_SIZE / 2,
      image_ref_y=TOP_DOWN_SIZE * 3 / 4, flip_axes=True)


def DrawTopDown(lasers):
  """Rasterises one frame's laser points `[P, ≥3]` (car frame) into a top-down image with
  height-coded intensity (ref :114)."""
  pts = np.asarray(lasers, np.float32).reshape(-1, np.asarray(lasers).shape[-1])
  t = _CarToImageTransform()
  hom = np.concatenate([pts[:, :3], np.ones((len(pts), 1), np.float32)], 1)
  uv = hom @ t.T
  u, v = uv[:, 0].astype(np.int64), uv[:, 1].astype(np.int64)
  ok = (u >= 0) & (u < TOP_DOWN_SIZE) & (v >= 0) & (v < TOP_DOWN_SIZE)
  img = np.zeros((TOP_DOWN_SIZE, TOP_DOWN_SIZE, 3), np.uint8)
  z = np.clip((pts[:, 2] + 2.0) / 5.0, 0.0, 1.0)
  shade = (80 + 175 * z).astype(np.uint8)
  img[v[ok], u[ok]] = np.stack([shade[ok]] * 3, -1)
  return img


def MakeRectangle(l, w, theta, offset=(0, 0)):
  """Corners of an l×w rectangle rotated by theta around its centre at `offset` (ref :157)."""
  c, s = math.cos(theta), math.sin(theta)
  corners = [(l / 2, w / 2), (l / 2, -w / 2), (-l / 2, -w / 2), (-l / 2, w / 2)]
  return [(c * x - s * y + offset[0], s * x + c * y + offset[1]) for x, y in corners]


def DrawHeadingTriangle(draw, x, y, heading, color, scale=25):
  """Small triangle pointing along `heading` (ref :166)."""
  tip = (x + scale * math.cos(heading), y + scale * math.sin(heading))
  left = (x + scale / 3 * math.cos(heading + 2.5), y + scale / 3 * math.sin(heading + 2.5))
  right = (x + scale / 3 * math.cos(heading - 2.5), y + scale / 3 * math.sin(heading - 2.5))
  draw.polygon([tip, left, right], outline=color)


def DrawCircle(draw, x, y, fill, outline, circle_size=5):
  draw.ellipse([x - circle_size, y - circle_size, x + circle_size, y + circle_size],
               fill=fill, outline=outline)


def DrawBoundingBoxOnImage(image, box, color='red', thickness=4, display_str='',
                           text_loc='TOP', fill=None):
  """Draws one oriented box `(cx, cy, w, h, heading)` in pixel units on a PIL image
  (ref :185)."""
  draw = ImageDraw.Draw(image)
  cx, cy, w, h, heading = box
  pts = MakeRectangle(w, h, heading, (cx, cy))
  if fill is not None:
    draw.polygon(pts, fill=fill)
  draw.line(pts + [pts[0]], width=thickness, fill=color)
  DrawHeadingTriangle(draw, cx, cy, heading, color, scale=max(w, h) / 2)
  if display_str:
    ys = [p[1] for p in pts]
    ty = min(ys) - 12 if text_loc == 'TOP' else max(ys) + 2
    draw.text((min(p[0] for p in pts), ty), display_str, fill=color)
  return image


def VisualizeBoxes(image, boxes, classes, scores, class_id_to_name, min_score_thresh=0.25,
                   line_thickness=4, groundtruth_box_visualization_color='black',
                   skip_scores=False, skip_labels=False, text_loc='TOP'):
  """Draws all boxes with score ≥ threshold (scores None ⇒ ground truth colour) on a
  uint8 image (ref :242)."""
  pil = Image.fromarray(_ToUint8(image))
  for i, box in enumerate(np.asarray(boxes)):
    if scores is not None and scores[i] < min_score_thresh:
      continue
    cls = int(classes[i])
    if scores is None:
      color, label = groundtruth_box_visualization_color, ''
    else:
      color = PIL_COLOR_LIST[cls % len(PIL_COLOR_LIST)]
      label = '' if skip_labels else str(class_id_to_name.get(cls, cls))
      if not skip_scores:
        label = ('%s: %d%%' % (label, int(100 * scores[i]))).strip(': ')
    DrawBoundingBoxOnImage(pil, box, color, line_thickness, label, text_loc)
  return np.asarray(pil)


def TransformBBoxesToTopDown(bboxes, car_to_image_transform=None):
  """7-DOF boxes `[..., 7]` (car frame) → `(cx, cy, w, h, heading)` in top-down pixels
  (ref :291)."""
  t = car_to_image_transform if car_to_image_transform is not None else _CarToImageTransform()
  b = np.asarray(bboxes, np.float32)
  flat = b.reshape(-1, 7)
  out = np.zeros((len(flat), 5), np.float32)
  for i, (x, y, z, dx, dy, _, phi) in enumerate(flat):
    box = transform_util.Box2D(x, y, dx, dy, phi).Apply(t)
    out[i] = [box.center[0], box.center[1], box.width, box.height, box.angle]
  return out.reshape(b.shape[:-1] + (5,))


def DrawBBoxesOnImages(images, bboxes, box_weights, labels, class_id_to_name, groundtruth):
  """Batch version: images `[B,H,W,3]`, top-down boxes `[B,N,5]`, weights `[B,N]` used as
  scores (or as a validity mask for ground truth) (ref :343)."""
  out = []
  for img, bb, w, lab in zip(np.asarray(images), np.asarray(bboxes), np.asarray(box_weights),
                             np.asarray(labels)):
    keep = w > 0
    out.append(VisualizeBoxes(img, bb[keep], lab[keep], None if groundtruth else w[keep],
                              class_id_to_name, min_score_thresh=0.25))
  return np.stack(out) if out else np.zeros((0,) + np.asarray(images).shape[1:], np.uint8)


def DrawTrajectory(image, bboxes, masks, labels, is_groundtruth):
  """Draws a track `[T, 5]` of top-down boxes: first box filled, later ones outlined,
  centres connected (ref :434)."""
  pil = Image.fromarray(_ToUint8(image))
  draw = ImageDraw.Draw(pil)
  color = 'black' if is_groundtruth else 'red'
  prev = None
  for t, (box, m) in enumerate(zip(np.asarray(bboxes), np.asarray(masks))):
    if m <= 0:
      continue
    lab = int(np.asarray(labels).reshape(-1)[min(t, np.asarray(labels).size - 1)])
    c = color if is_groundtruth else PIL_COLOR_LIST[lab % len(PIL_COLOR_LIST)]
    DrawBoundingBoxOnImage(pil, box, c, 2, fill=c if prev is None else None)
    if prev is not None:
      draw.line([prev, (box[0], box[1])], fill=c, width=2)
    prev = (box[0], box[1])
  return np.asarray(pil)


def GetTrajectoryComparison(gt_bboxes, gt_masks, gt_labels, pred_bboxes, pred_masks,
                            pred_labels):
  """Ground-truth vs. predicted tracks side by side on blank top-down canvases → one image
  `[H, 2W, 3]` (ref :560)."""
  blank = np.full((TOP_DOWN_SIZE, TOP_DOWN_SIZE, 3), 255, np.uint8)
  left, right = blank.copy(), blank.copy()
  for bb, m, lab in zip(np.asarray(gt_bboxes), np.asarray(gt_masks), np.asarray(gt_labels)):
    left = DrawTrajectory(left, TransformBBoxesToTopDown(bb), m, lab, True)
  for bb, m, lab in zip(np.asarray(pred_bboxes), np.asarray(pred_masks), np.asarray(pred_labels)):
    right = DrawTrajectory(right, TransformBBoxesToTopDown(bb), m, lab, False)
  return np.concatenate([left, right], 1)
