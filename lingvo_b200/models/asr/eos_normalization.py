"""Normalising trailing EOS tokens of label sequences (ref
`lingvo/tasks/asr/eos_normalization.py`).

Datasets disagree on whether a transcript ends with zero, one or several `</s>`;
`NormalizeTrailingEos` rewrites `[B, T]` ids so that each sequence ends with exactly
`need_trailing_eos ? 1 : 0` EOS and is EOS-filled afterwards.
"""

from __future__ import annotations

import numpy as np
import torch


def FillPaddingPos(ids, id_len, padding_value=0):
  """Overwrites positions ≥ id_len[b] with `padding_value` (ref :22)."""
  t = ids.shape[1]
  mask = torch.arange(t, device=ids.device).unsqueeze(0) < id_len.unsqueeze(1)
  return torch.where(mask, ids, torch.full_like(ids, padding_value))


def NormalizeTrailingEos(ids, id_len, need_trailing_eos=True, eos_id=2):
  """→ (new_ids, new_len) (ref :42). Vectorised: strip the run of trailing EOS inside
  the valid prefix, then optionally append one back (capped at T)."""
  t = ids.shape[1]
  pos = torch.arange(t, device=ids.device).unsqueeze(0)
  valid = pos < id_len.unsqueeze(1)
  non_eos = valid & (ids != eos_id)
  # index of the last non-EOS token + 1 = length without trailing EOS
  last = torch.where(non_eos, pos + 1, torch.zeros_like(pos)).max(dim=1).values
  new_len = last
  if need_trailing_eos:
    new_len = (last + 1).clamp(max=t)
  keep = pos < last.unsqueeze(1)
  new_ids = torch.where(keep, ids, torch.full_like(ids, eos_id))
  return new_ids, new_len.to(id_len.dtype)


def NumpyNormalizeTrailingEos(ids: np.ndarray, id_len: np.ndarray, need_trailing_eos=True,
                              eos_id=2):
  """NumPy twin of `NormalizeTrailingEos` (ref :93)."""
  out, out_len = NormalizeTrailingEos(torch.as_tensor(np.asarray(ids)),
                                      torch.as_tensor(np.asarray(id_len)),
                                      need_trailing_eos, eos_id)
  return out.numpy(), out_len.numpy()
