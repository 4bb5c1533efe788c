"""Insertion-Transformer canvas utilities (ref `lingvo/core/insertion.py`)."""

from __future__ import annotations

import torch

from lingvo_b200.core import base_layer
from lingvo_b200.core.nested_map import NestedMap


def SequenceTrimLastToken(x, x_paddings):
  """Removes the last non-padded token of every row (ref :27)."""
  lens = (1.0 - x_paddings.float()).sum(1).long()
  pos = torch.arange(x.shape[1], device=x.device).unsqueeze(0)
  keep = pos < (lens - 1).clamp_min(0).unsqueeze(1)
  return x * keep.to(x.dtype), 1.0 - keep.float()


def SequenceAppendToken(x, x_paddings, token, extend=False):
  """Writes `token` after the last non-padded position (ref :48)."""
  if extend:
    x = torch.nn.functional.pad(x, (0, 1))
    x_paddings = torch.nn.functional.pad(x_paddings, (0, 1), value=1.0)
  lens = (1.0 - x_paddings.float()).sum(1).long().clamp(max=x.shape[1] - 1)
  x = x.clone()
  pad = x_paddings.clone()
  ar = torch.arange(x.shape[0], device=x.device)
  tok = torch.as_tensor(token, device=x.device, dtype=x.dtype).expand(x.shape[0])
  x[ar, lens] = tok
  pad[ar, lens] = 0.0
  return x, pad


def SequenceConcat(x, x_paddings, y, y_paddings, pad=0):
  """Row-wise concatenation of the non-padded parts of x and y (ref :79)."""
  b, tx = x.shape
  ty = y.shape[1]
  xl = (1.0 - x_paddings.float()).sum(1).long()
  yl = (1.0 - y_paddings.float()).sum(1).long()
  out = torch.full((b, tx + ty), pad, dtype=x.dtype, device=x.device)
  pos = torch.arange(tx + ty, device=x.device).unsqueeze(0)
  from_x = pos < xl.unsqueeze(1)
  from_y = (pos >= xl.unsqueeze(1)) & (pos < (xl + yl).unsqueeze(1))
  xi = pos.clamp(max=tx - 1).expand(b, -1)
  yi = (pos - xl.unsqueeze(1)).clamp(0, ty - 1)
  out = torch.where(from_x, x.gather(1, xi), out)
  out = torch.where(from_y, y.gather(1, yi), out)
  return out, 1.0 - (from_x | from_y).float()


class SymbolInsertionLayer(base_layer.BaseLayer):
  """Samples a partial canvas from a target and the insertion labels that complete it
  (ref :130): each target token is kept with a per-row random rate; the labels are,
  for every canvas slot, the target tokens missing between consecutive kept tokens."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('rollin_policy', 'oracle', 'oracle (sample from the target).')
    p.Define('oracle_policy', 'uniform', 'uniform.')
    p.Define('random_seed', None, 'Seed.') if 'random_seed' not in p else None
    return p

  def FProp(self, theta, x, x_paddings=None, eos_id=1, force_sample_last_token=True):
    """x `[B, T]` target ids → NestedMap(canvas, canvas_indices, canvas_paddings,
    target_indices [N, 3] = (batch, slot, token), target_weights)."""
    b, t = x.shape
    dev = x.device
    if x_paddings is None:
      x_paddings = torch.zeros(b, t, device=dev)
    valid = x_paddings < 0.5
    rate = torch.rand(b, 1, device=dev)
    keep = (torch.rand(b, t, device=dev) < rate) & valid
    if force_sample_last_token:
      lens = valid.sum(1)
      keep[torch.arange(b, device=dev), (lens - 1).clamp_min(0)] = True
    keep = keep & valid
    order = torch.argsort((~keep).to(torch.int32), dim=1, stable=True)
    n_keep = keep.sum(1)
    canvas_idx = order
    pos = torch.arange(t, device=dev).unsqueeze(0)
    canvas_pad = (pos >= n_keep.unsqueeze(1)).float()
    canvas = x.gather(1, canvas_idx) * (1 - canvas_pad).to(x.dtype)
    # slot of a missing token = number of kept tokens before it
    slot = torch.cumsum(keep.to(torch.int64), 1) - keep.to(torch.int64)
    missing = valid & ~keep
    bi = torch.arange(b, device=dev).unsqueeze(1).expand(b, t)
    tgt = torch.stack([bi[missing], slot[missing], x[missing].long()], 1)
    # per-slot uniform weights: 1 / (#missing tokens in the slot)
    key = bi[missing] * (t + 1) + slot[missing]
    _, inv, cnt = torch.unique(key, return_inverse=True, return_counts=True)
    w = 1.0 / cnt[inv].float() if key.numel() else torch.zeros(0, device=dev)
    return NestedMap(canvas=canvas, canvas_indices=canvas_idx, canvas_paddings=canvas_pad,
                     target_indices=tgt, target_weights=w)
