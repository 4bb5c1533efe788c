"""ASR frontends (ref `lingvo/tasks/asr/frontend.py`).

`MelAsrFrontend` (ref :114): PCM `[B, samples]` → pre-emphasis → framing (Hann
window) → |FFT|² → mel filterbank → log, with optional per-bin normalisation and
frame stacking. Runs as batched `torch.stft` + one GEMM on the device.
"""

from __future__ import annotations

import math

import torch

from lingvo_b200.core import base_layer
from lingvo_b200.core.nested_map import NestedMap


def _NextPowerOfTwo(i):
  return 1 << max(int(i) - 1, 0).bit_length()


class BaseAsrFrontend(base_layer.BaseLayer):
  """FProp(input_batch{src_inputs, paddings}) → NestedMap(src_inputs [B,T,F,1], paddings)."""

  @property
  def config_is_stacked(self):
    return False

  def FProp(self, theta, input_batch):
    raise NotImplementedError


class NullAsrFrontend(BaseAsrFrontend):
  """Pass-through for pre-computed features (ref :96)."""

  def FProp(self, theta, input_batch):
    return input_batch.DeepCopy() if hasattr(input_batch, 'DeepCopy') else input_batch


class MelAsrFrontend(BaseAsrFrontend):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.name = 'frontend'
    p.Define('sample_rate', 16000.0, 'Sample rate (Hz).')
    p.Define('frame_size_ms', 25.0, 'Window size (ms).')
    p.Define('frame_step_ms', 10.0, 'Hop (ms).')
    p.Define('num_bins', 80, 'Mel bins.')
    p.Define('lower_edge_hertz', 125.0, 'Lowest mel edge.')
    p.Define('upper_edge_hertz', 7600.0, 'Highest mel edge.')
    p.Define('preemph', 0.97, 'Pre-emphasis coefficient.')
    p.Define('noise_scale', 8.0, 'Dither std (in 16-bit sample units).')
    p.Define('window_fn', 'HANNING', 'Window function.')
    p.Define('pad_end', False, 'Pad the last partial frame.')
    p.Define('fft_overdrive', True, 'Round the FFT size up to a power of two.')
    p.Define('per_bin_mean', None, 'Per-bin mean for normalisation.')
    p.Define('per_bin_stddev', None, 'Per-bin std for normalisation.')
    p.Define('stack_left_context', 0, 'Frames of left context to stack.')
    p.Define('stack_right_context', 0, 'Frames of right context to stack.')
    p.Define('frame_stride', 1, 'Subsampling after stacking.')
    p.Define('output_floor', 1.0, 'Floor before the log.')
    p.Define('compute_energy', False, 'Kept for parity.')
    p.Define('use_divide_stream', False, 'Kept for parity.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    self._frame_size = int(round(p.sample_rate * p.frame_size_ms / 1000.0)) + 1  # +1: preemph
    self._frame_step = int(round(p.sample_rate * p.frame_step_ms / 1000.0))
    self._fft = _NextPowerOfTwo(self._frame_size - 1) if p.fft_overdrive else self._frame_size - 1
    self._mel = None

  @property
  def config_is_stacked(self):
    p = self.params
    return p.stack_left_context > 0 or p.stack_right_context > 0

  def _MelMatrix(self, device):
    """[fft/2+1, num_bins] triangular HTK-mel filterbank."""
    if self._mel is not None and self._mel.device == device:
      return self._mel
    p = self.params
    n_freq = self._fft // 2 + 1
    hz = torch.linspace(0, p.sample_rate / 2, n_freq)
    mel = lambda f: 1127.0 * torch.log1p(torch.as_tensor(f, dtype=torch.float32) / 700.0)
    edges = torch.linspace(float(mel(p.lower_edge_hertz)), float(mel(p.upper_edge_hertz)),
                           p.num_bins + 2)
    m = mel(hz).unsqueeze(1)
    lo, ce, hi = edges[:-2], edges[1:-1], edges[2:]
    up = (m - lo) / (ce - lo)
    down = (hi - m) / (hi - ce)
    w = torch.clamp(torch.minimum(up, down), min=0.0)
    w[0] = 0.0
    self._mel = w.to(device)
    return self._mel

  def FProp(self, theta, input_batch):
    p = self.params
    pcm = input_batch.src_inputs.float()
    pad = input_batch.paddings.float()
    if pcm.dim() == 3:
      pcm = pcm.squeeze(-1)
    b, n = pcm.shape
    if p.noise_scale > 0 and not self.do_eval:
      pcm = pcm + torch.randn_like(pcm) * p.noise_scale
    size, step = self._frame_size, self._frame_step
    n_frames = max((n - size) // step + 1, 0) if not p.pad_end else -(-n // step)
    need = (n_frames - 1) * step + size
    if need > n:
      pcm = torch.nn.functional.pad(pcm, (0, need - n))
    frames = pcm.unfold(1, size, step)[:, :n_frames]                # [B, T, size]
    frames = frames[..., 1:] - p.preemph * frames[..., :-1]         # pre-emphasis
    win = torch.hann_window(size - 1, periodic=True, device=pcm.device)
    spec = torch.fft.rfft(frames * win, n=self._fft)
    mag = spec.abs()
    mel = torch.matmul(mag, self._MelMatrix(pcm.device))
    feat = torch.log(torch.clamp(mel, min=p.output_floor))
    # frame paddings: a frame is padding if its first sample is padding
    fp = pad[:, ::step][:, :n_frames] if pad.shape[1] == n else pad[:, :n_frames]
    if p.per_bin_mean is not None:
      mean = torch.as_tensor(p.per_bin_mean, device=feat.device, dtype=feat.dtype)
      std = torch.as_tensor(p.per_bin_stddev, device=feat.device, dtype=feat.dtype)
      feat = (feat - mean) / std
    if self.config_is_stacked:
      l, r = p.stack_left_context, p.stack_right_context
      padded = torch.nn.functional.pad(feat, (0, 0, l, r))
      feat = padded.unfold(1, l + r + 1, 1).permute(0, 1, 3, 2).reshape(b, n_frames, -1)
      feat = feat[:, ::p.frame_stride]
      fp = fp[:, ::p.frame_stride]
    feat = feat * (1.0 - fp).unsqueeze(-1)
    return NestedMap(src_inputs=feat.unsqueeze(-1), paddings=fp)
