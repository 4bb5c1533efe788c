"""Image decoding + augmentation (ref `lingvo/tasks/milan/image_preprocessor.py`).

Training: random resized crop (area 5–100 %, aspect 3/4–4/3), random flip, brightness /
saturation / contrast jitter; eval: central crop (87.5 %) + resize. Decoding uses PIL when
available (JPEG/PNG bytes); already-decoded uint8 / float arrays pass straight to the
tensor pipeline, which runs on whatever device the input lives on.
"""

from __future__ import annotations

import io
import math

import numpy as np
import torch
import torch.nn.functional as F

from lingvo_b200.core import base_layer


def _DistortBrightnessAndColor(image, gen=None):
  """image [3,H,W] in [0,1] (ref :25)."""
  r = lambda lo, hi: float(torch.empty(1).uniform_(lo, hi, generator=gen))
  image = image + r(-32.0 / 255.0, 32.0 / 255.0)
  gray = image.mean(0, keepdim=True)
  image = gray + (image - gray) * r(0.5, 1.5)                      # saturation
  mean = image.mean()
  image = mean + (image - mean) * r(0.5, 1.5)                      # contrast
  return image.clamp(0.0, 1.0)


def DecodeImage(encoded):
  """bytes / ndarray / tensor → float tensor [3, H, W] in [0, 1]."""
  if isinstance(encoded, (bytes, bytearray)):
    try:
      from PIL import Image  # pylint: disable=g-import-not-at-top
    except ImportError as e:
      raise RuntimeError('decoding image bytes needs PIL; feed decoded arrays instead') from e
    arr = np.asarray(Image.open(io.BytesIO(encoded)).convert('RGB'))
    encoded = arr
  t = torch.as_tensor(np.asarray(encoded) if not isinstance(encoded, torch.Tensor) else encoded)
  if t.dim() == 3 and t.shape[-1] in (1, 3):
    t = t.permute(2, 0, 1)
  if t.dtype == torch.uint8:
    t = t.float() / 255.0
  t = t.float()
  return t.expand(3, -1, -1) if t.shape[0] == 1 else t


class ImagePreprocessor(base_layer.BaseLayer):
  """ref :48."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('output_image_size', 224, 'Output height = width.')
    p.Define('crop_fraction', 0.875, 'Central crop fraction at eval time.')
    p.Define('augment', True, 'Colour jitter during training.')
    p.Define('min_crop_area', 0.05, 'Smallest random-crop area fraction.')
    p.name = 'image_preprocessor'
    return p

  def _Resize(self, image):
    s = self.params.output_image_size
    return F.interpolate(image.unsqueeze(0), size=(s, s), mode='bilinear',
                         align_corners=False).squeeze(0)

  def _PreprocessForTraining(self, image, gen=None):
    p = self.params
    _, h, w = image.shape
    area = h * w
    u = lambda lo, hi: float(torch.empty(1).uniform_(lo, hi, generator=gen))
    crop = None
    for _ in range(10):
      a = area * u(p.min_crop_area, 1.0)
      ratio = math.exp(u(math.log(3 / 4), math.log(4 / 3)))
      cw, ch = int(round(math.sqrt(a * ratio))), int(round(math.sqrt(a / ratio)))
      if 0 < cw <= w and 0 < ch <= h:
        y0 = int(u(0, h - ch + 1e-6))
        x0 = int(u(0, w - cw + 1e-6))
        crop = image[:, y0:y0 + ch, x0:x0 + cw]
        break
    if crop is None:
      crop = image
    out = self._Resize(crop)
    if u(0, 1) < 0.5:
      out = out.flip(-1)
    if p.augment:
      out = _DistortBrightnessAndColor(out, gen)
    return out

  def _PreprocessForEval(self, image):
    p = self.params
    _, h, w = image.shape
    ch, cw = max(1, int(h * p.crop_fraction)), max(1, int(w * p.crop_fraction))
    y0, x0 = (h - ch) // 2, (w - cw) // 2
    return self._Resize(image[:, y0:y0 + ch, x0:x0 + cw])

  def FProp(self, theta, encoded_images):
    """Nested list / array of encoded (or decoded) images with any leading batch dims →
    float tensor `batch_dims + [3, S, S]` scaled to [-1, 1]."""
    del theta
    def _One(e):
      img = DecodeImage(e)
      img = self._PreprocessForEval(img) if self.do_eval else self._PreprocessForTraining(img)
      return img * 2.0 - 1.0
    def _Rec(x):
      if isinstance(x, (list, tuple)):
        return torch.stack([_Rec(e) for e in x])
      if isinstance(x, (np.ndarray, torch.Tensor)) and x.ndim > 3:
        return torch.stack([_Rec(e) for e in x])
      return _One(x)
    return _Rec(encoded_images)
