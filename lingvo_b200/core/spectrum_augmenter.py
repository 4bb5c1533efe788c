"""SpecAugment (ref `lingvo/core/spectrum_augmenter.py`).

Inputs `[B, T, F, C]` + paddings `[B, T]`. Frequency masking (ref :300-420), time
masking with per-utterance max ratio (ref :420-560), optional time warping
(ref :560-700; piecewise-linear resampling around a random anchor), mask
multiplicity (fixed or length-adaptive), all as batched tensor ops on the device
(no per-example Python loops), so augmentation costs a few elementwise kernels.
"""

from __future__ import annotations

import torch

from lingvo_b200.core import base_layer
from lingvo_b200.core import py_utils


class SpectrumAugmenter(base_layer.BaseLayer):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('freq_mask_max_bins', 15, 'Max width of a frequency mask.')
    p.Define('freq_mask_count', 1, 'Number of frequency masks.')
    p.Define('use_dynamic_time_mask_max_frames', False,
             'Max time-mask width = time_mask_max_ratio · length.')
    p.Define('time_mask_max_frames', 50, 'Max width of a time mask (static).')
    p.Define('time_mask_count', 1, 'Number (or max number) of time masks.')
    p.Define('time_mask_max_ratio', 1.0, 'Max fraction of the utterance masked per mask.')
    p.Define('time_masks_per_frame', 0.0, 'Adaptive multiplicity: masks = this · length.')
    p.Define('block_mask_prob', 0.0, 'Kept for parity.')
    p.Define('freq_warp_max_bins', 0, 'Kept for parity.')
    p.Define('time_warp_bound', 'static', 'static | dynamic.')
    p.Define('time_warp_max_frames', 0, 'Max time-warp displacement.')
    p.Define('time_warp_max_ratio', 0.0, 'Max displacement as a fraction of length.')
    p.Define('use_noise', False, 'Fill masks with Gaussian noise instead of zeros.')
    p.Define('gaussian_noise', False, 'Kept for parity.')
    p.Define('unstack', False, 'Kept for parity.')
    p.Define('stack_height', 3, 'Kept for parity.')
    p.Define('domain_ids', [0], 'Kept for parity.')
    p.Define('use_input_dependent_random_seed', False, 'Kept for parity.')
    p.Define('eval_data_in_domains', False, 'Kept for parity.')
    return p

  def _Gen(self, device):
    p = self.params
    if p.random_seed is None:
      return None
    g = torch.Generator(device=device)
    g.manual_seed(int(p.random_seed) + int(py_utils.GetGlobalStep()))
    return g

  def _Rand(self, shape, device, gen):
    return torch.rand(shape, device=device, generator=gen)

  def _Masks1D(self, size, lengths, max_width, count, device, gen, multiplicity=None):
    """Union of `count` random intervals per example → bool [B, size] (True = masked).

    lengths [B] bounds the start so masks fall inside the valid region;
    max_width [B] float; multiplicity [B] (optional) = number of active masks.
    """
    b = lengths.shape[0]
    width = (self._Rand((b, count), device, gen) * (max_width.unsqueeze(1) + 1)).floor()
    width = torch.minimum(width, lengths.unsqueeze(1).float())
    start = (self._Rand((b, count), device, gen) *
             (lengths.unsqueeze(1).float() - width + 1).clamp_min(1)).floor()
    pos = torch.arange(size, device=device).view(1, 1, size).float()
    m = (pos >= start.unsqueeze(-1)) & (pos < (start + width).unsqueeze(-1))
    if multiplicity is not None:
      active = torch.arange(count, device=device).view(1, count) < multiplicity.view(b, 1)
      m = m & active.unsqueeze(-1)
    return m.any(1)

  def _TimeWarp(self, x, lengths, gen):
    """Piecewise-linear warp: a random anchor a ∈ (w, L−w) moves to a + δ."""
    p = self.params
    b, t = x.shape[:2]
    dev = x.device
    lf = lengths.float()
    if p.time_warp_bound == 'dynamic':
      w = (lf * p.time_warp_max_ratio).floor()
    else:
      w = torch.full_like(lf, float(p.time_warp_max_frames))
    w = torch.minimum(w, ((lf - 1) / 2).floor().clamp_min(0))
    anchor = w + self._Rand((b,), dev, gen) * (lf - 2 * w).clamp_min(1)
    delta = (self._Rand((b,), dev, gen) * 2 - 1) * w
    dst_anchor = anchor + delta
    pos = torch.arange(t, device=dev).float().unsqueeze(0)
    left = pos * (anchor / dst_anchor.clamp_min(1e-3)).unsqueeze(1)
    right = anchor.unsqueeze(1) + (pos - dst_anchor.unsqueeze(1)) * (
        (lf - anchor) / (lf - dst_anchor).clamp_min(1e-3)).unsqueeze(1)
    src = torch.where(pos < dst_anchor.unsqueeze(1), left, right)
    src = torch.where(pos < lf.unsqueeze(1), src, pos).clamp(0, t - 1)
    lo = src.floor().long()
    hi = (lo + 1).clamp(max=t - 1)
    frac = (src - lo.float()).view(b, t, *([1] * (x.dim() - 2)))
    idx = lambda i: i.view(b, t, *([1] * (x.dim() - 2))).expand_as(x)
    return x.gather(1, idx(lo)) * (1 - frac) + x.gather(1, idx(hi)) * frac

  def FProp(self, theta, inputs, paddings, domain_ids=None):
    """Returns (augmented inputs, paddings). No-op in eval."""
    p = self.params
    if self.do_eval:
      return inputs, paddings
    b, t, f = inputs.shape[:3]
    dev = inputs.device
    gen = self._Gen(dev)
    lengths = (1.0 - paddings.float()).sum(1)
    x = inputs
    if p.time_warp_max_frames > 0 or p.time_warp_max_ratio > 0:
      x = self._TimeWarp(x, lengths, gen)
    mask = torch.zeros(b, t, f, dtype=torch.bool, device=dev)
    if p.freq_mask_count > 0 and p.freq_mask_max_bins > 0:
      fm = self._Masks1D(f, torch.full((b,), f, device=dev), torch.full(
          (b,), float(p.freq_mask_max_bins), device=dev), p.freq_mask_count, dev, gen)
      mask = mask | fm.unsqueeze(1)
    if p.time_mask_count > 0:
      if p.use_dynamic_time_mask_max_frames:
        max_w = (lengths * p.time_mask_max_ratio).floor()
      else:
        max_w = torch.minimum(torch.full_like(lengths, float(p.time_mask_max_frames)),
                              (lengths * p.time_mask_max_ratio).floor())
      mult = None
      if p.time_masks_per_frame > 0:
        mult = (lengths * p.time_masks_per_frame).floor().clamp(max=p.time_mask_count)
      tm = self._Masks1D(t, lengths.long(), max_w, p.time_mask_count, dev, gen, mult)
      mask = mask | tm.unsqueeze(2)
    mask = mask.view(b, t, f, *([1] * (x.dim() - 3)))
    if p.use_noise:
      noise = torch.randn(x.shape, device=dev, dtype=x.dtype, generator=gen)
      x = torch.where(mask, noise, x)
    else:
      x = x.masked_fill(mask, 0.0)
    return x, paddings
