"""Visual "metrics" for 3-D detection (ref `lingvo/tasks/car/detection_3d_metrics.py`):
they buffer a few decoded scenes and render them as image summaries.

  TopDownVisualizationMetric  lasers + predicted (score-coloured) and ground-truth boxes
  WorldViewer                 sub-sampled point clouds + boxes as a 3-D mesh-style dump
  CameraVisualization         boxes projected into the camera image
"""

from __future__ import annotations

import numpy as np
from PIL import Image
from PIL import ImageDraw

from lingvo_b200.core import metrics as metrics_lib
from lingvo_b200.models.car import summary
from lingvo_b200.models.car import transform_util


def _Png(img):
  import io  # pylint: disable=g-import-not-at-top
  buf = io.BytesIO()
  Image.fromarray(img).save(buf, format='PNG')
  return buf.getvalue()


class TopDownVisualizationMetric(metrics_lib.BaseMetric):
  """ref :31. `Update(decoded_outputs)` expects per-batch arrays `visualization_labels
  [B,N]`, `predicted_bboxes [B,N,5]` (x, y, w, h, heading in car metres), `visualization_
  weights [B,N]`, `gt_bboxes_2d [B,G,5]`, `gt_bboxes_2d_weights`, `labels`, `points_xyz
  [B,P,3]`, `points_padding`, `source_ids`, optional `difficulties`."""

  def __init__(self, top_down_transform=None, image_height=1024, image_width=1024,
               ground_removal_threshold=-1.35, sampler_num_samples=8):
    self._transform = top_down_transform if top_down_transform is not None else \
        transform_util.MakeCarToImageTransform(32.0 / 3.0, image_width / 2,
                                               image_height * 3 / 4, True)
    self._h, self._w = image_height, image_width
    self._ground = ground_removal_threshold
    self._max = sampler_num_samples
    self._samples = []
    self._images = None

  def Update(self, decoded_outputs):
    b = len(decoded_outputs.source_ids)
    for i in range(b):
      if len(self._samples) >= self._max:
        break
      self._samples.append({k: (np.asarray(v)[i] if hasattr(v, '__len__') and len(v) == b
                                else v) for k, v in decoded_outputs.items()})
    self._images = None

  def _XYWHToExtrema(self, bboxes):
    """[N, 5] car-frame boxes → pixel (cx, cy, w, h, heading)."""
    out = np.zeros((len(bboxes), 5), np.float32)
    for i, (x, y, w, h, phi) in enumerate(bboxes):
      box = transform_util.Box2D(x, y, w, h, phi).Apply(self._transform)
      out[i] = [box.center[0], box.center[1], box.width, box.height, box.angle]
    return out

  def _DrawLasers(self, image, points_xyz, points_padding):
    pts = np.asarray(points_xyz, np.float32)[np.asarray(points_padding) < 0.5]
    pts = pts[pts[:, 2] > self._ground]
    hom = np.concatenate([pts[:, :3], np.ones((len(pts), 1), np.float32)], 1) @ self._transform.T
    u, v = hom[:, 0].astype(np.int64), hom[:, 1].astype(np.int64)
    ok = (u >= 0) & (u < self._w) & (v >= 0) & (v < self._h)
    image[v[ok], u[ok]] = (60, 60, 60)
    return image

  def DrawDifficulty(self, image, gt_bboxes, gt_box_weights, difficulties):
    """Writes each ground-truth box's difficulty next to it (ref :241)."""
    pil = Image.fromarray(image)
    draw = ImageDraw.Draw(pil)
    for box, w, d in zip(gt_bboxes, gt_box_weights, difficulties):
      if w > 0:
        draw.text((box[0] + box[2] / 2 + 2, box[1]), str(int(d)), fill='blue')
    return np.asarray(pil)

  def _EvaluateIfNecessary(self, name):
    if self._images is not None:
      return
    self._images = []
    for s in self._samples:
      img = np.full((self._h, self._w, 3), 255, np.uint8)
      img = self._DrawLasers(img, s['points_xyz'], s['points_padding'])
      gt = self._XYWHToExtrema(np.asarray(s['gt_bboxes_2d']))
      gw = np.asarray(s['gt_bboxes_2d_weights'])
      img = summary.VisualizeBoxes(img, gt[gw > 0], np.asarray(s['labels'])[gw > 0], None, {})
      if 'difficulties' in s:
        img = self.DrawDifficulty(img, gt, gw, np.asarray(s['difficulties']))
      pw = np.asarray(s['visualization_weights'])
      pred = self._XYWHToExtrema(np.asarray(s['predicted_bboxes']))
      img = summary.VisualizeBoxes(img, pred[pw > 0], np.asarray(s['visualization_labels'])[pw > 0],
                                   pw[pw > 0], {}, min_score_thresh=0.0, line_thickness=2)
      pil = Image.fromarray(img)
      ImageDraw.Draw(pil).text((6, 6), str(s['source_ids']), fill='black')
      self._images.append(('%s/%s' % (name, s['source_ids']), np.asarray(pil)))

  @property
  def value(self):
    return len(self._samples)

  def Images(self, name):
    self._EvaluateIfNecessary(name)
    return list(self._images)

  def Summary(self, name):
    """[(tag, png bytes)]."""
    return [(t, _Png(img)) for t, img in self.Images(name)]


class WorldViewer(metrics_lib.BaseMetric):
  """Buffers point clouds and boxes for an external 3-D viewer (ref :285): `Summary`
  returns a dict per scene with sub-sampled `points [P,3]`, `colors`, `bboxes [N,7]`,
  `bbox_scores`, `bbox_labels`."""

  def __init__(self, sampler_num_samples=8, max_points=20000):
    self._max, self._max_points = sampler_num_samples, max_points
    self._scenes = []

  def Update(self, decoded_outputs):
    b = len(decoded_outputs.source_ids)
    for i in range(b):
      if len(self._scenes) >= self._max:
        return
      pts = np.asarray(decoded_outputs.points_sampled[i] if 'points_sampled' in decoded_outputs
                       else decoded_outputs.points_xyz[i], np.float32)
      if 'points_padding' in decoded_outputs:
        pts = pts[np.asarray(decoded_outputs.points_padding[i]) < 0.5]
      if len(pts) > self._max_points:
        pts = pts[np.random.RandomState(0).choice(len(pts), self._max_points, replace=False)]
      z = np.clip((pts[:, 2] + 2.0) / 5.0, 0, 1)
      colors = np.stack([z, 1 - z, np.full_like(z, 0.5)], 1)
      self._scenes.append(dict(
          source_id=decoded_outputs.source_ids[i], points=pts, colors=colors,
          bboxes=np.asarray(decoded_outputs.bboxes[i], np.float32),
          bbox_scores=np.asarray(decoded_outputs.bbox_scores[i], np.float32),
          bbox_labels=np.asarray(decoded_outputs.get('bbox_labels',
                                                     np.zeros(len(decoded_outputs.bboxes[i])))[i]
                                 if 'bbox_labels' in decoded_outputs else
                                 np.zeros(len(decoded_outputs.bboxes[i]), np.int32))))

  @property
  def value(self):
    return len(self._scenes)

  def Summary(self, name):
    return [('%s/%s' % (name, s['source_id']), s) for s in self._scenes]


class CameraVisualization(metrics_lib.BaseMetric):
  """Boxes projected into camera images (ref :360). `Update` takes `camera_images
  [B,H,W,3]`, `bbox_corners [B,N,8,2]` pixel corners (top loop then bottom loop),
  `bbox_scores [B,N]`, `source_ids`; optional `bbox_2d [B,N,4]` = ymin,xmin,ymax,xmax."""

  def __init__(self, bbox_score_threshold=0.01, draw_3d_boxes=True, sampler_num_samples=8):
    self._thr, self._3d, self._max = bbox_score_threshold, draw_3d_boxes, sampler_num_samples
    self._samples = []

  def Update(self, decoded_outputs):
    b = len(decoded_outputs.source_ids)
    for i in range(b):
      if len(self._samples) >= self._max:
        return
      self._samples.append({k: np.asarray(v)[i] for k, v in decoded_outputs.items()
                            if hasattr(v, '__len__') and len(v) == b})

  @staticmethod
  def Draw3DBoxes(draw, corners, color):
    top, bot = corners[:4], corners[4:]
    for loop in (top, bot):
      pts = [tuple(p) for p in loop]
      draw.line(pts + [pts[0]], fill=color, width=2)
    for a, c in zip(top, bot):
      draw.line([tuple(a), tuple(c)], fill=color, width=2)

  @staticmethod
  def Draw2DBoxes(draw, box, color):
    ymin, xmin, ymax, xmax = box
    draw.rectangle([xmin, ymin, xmax, ymax], outline=color, width=2)

  @property
  def value(self):
    return len(self._samples)

  def Images(self, name):
    out = []
    for s in self._samples:
      pil = Image.fromarray(summary._ToUint8(s['camera_images']))   # pylint: disable=protected-access
      draw = ImageDraw.Draw(pil)
      for j, score in enumerate(s['bbox_scores']):
        if score < self._thr:
          continue
        color = (int(255 * (1 - score)), int(255 * score), 0)
        if self._3d and 'bbox_corners' in s:
          self.Draw3DBoxes(draw, s['bbox_corners'][j], color)
        elif 'bbox_2d' in s:
          self.Draw2DBoxes(draw, s['bbox_2d'][j], color)
      out.append(('%s/%s' % (name, s['source_ids']), np.asarray(pil)))
    return out

  def Summary(self, name):
    return [(t, _Png(img)) for t, img in self.Images(name)]
