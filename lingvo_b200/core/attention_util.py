"""Attention helpers (ref `lingvo/core/attention_util.py`): blocking utilities for
local attention, relative shift, XL positional logits, k-means for routing attention."""

from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from lingvo_b200.core import base_layer
from lingvo_b200.core import py_utils
from lingvo_b200.core import quant_utils
from lingvo_b200.core.py_utils import WeightInit
from lingvo_b200.core.py_utils import WeightParams


def ConvertToBlocks(x, block_size, padding_val=0.0):
  """[B, T, …] → [B, ⌈T/W⌉, W, …] (ref :25)."""
  b, t = x.shape[:2]
  n = -(-t // block_size)
  pad = n * block_size - t
  if pad:
    x = F.pad(x, (0, 0) * (x.dim() - 2) + (0, pad), value=padding_val)
  return x.reshape(b, n, block_size, *x.shape[2:])


def ExtractBlockContext(x, block_size, left_context, right_context, padding_val=0.0):
  """[B, T, …] → [B, U, W + L − 1 + R, …]: each block with its neighbours (ref :51)."""
  b, t = x.shape[:2]
  n = -(-t // block_size)
  l, r = left_context - 1, right_context
  xp = F.pad(x, (0, 0) * (x.dim() - 2) + (l, n * block_size - t + r), value=padding_val)
  win = block_size + l + r
  idx = (torch.arange(n, device=x.device).unsqueeze(1) * block_size +
         torch.arange(win, device=x.device).unsqueeze(0))
  return xp[:, idx]


def ExtractBlockContextV2(x, block_size, left_context, right_context, padding_val=0.0,
                          paddings=None):
  """`ExtractBlockContext` without constraints between W, L, R, plus the matching paddings
  `[B, U, C]` (1 outside the sequence) when `paddings [B, T]` is given (ref :128)."""
  patches = ExtractBlockContext(x, block_size, left_context, right_context, padding_val)
  if paddings is None:
    return patches, None
  return patches, ExtractBlockContext(paddings, block_size, left_context, right_context, 1.0)


def MakeLocalMask(seq_len, block_size, left_context, right_context, dtype=torch.float32,
                  device=None):
  """[U, W, C] 1 where query w of block u may see context position c (ref :242)."""
  n = -(-seq_len // block_size)
  c = block_size + left_context - 1 + right_context
  q = (torch.arange(n, device=device).view(n, 1, 1) * block_size +
       torch.arange(block_size, device=device).view(1, block_size, 1))
  k = (torch.arange(n, device=device).view(n, 1, 1) * block_size - (left_context - 1) +
       torch.arange(c, device=device).view(1, 1, c))
  ok = (k >= q - (left_context - 1)) & (k <= q + right_context) & (k >= 0) & (k < seq_len) & \
      (q < seq_len)
  return ok.to(dtype)


def RelShift(x):
  """Transformer-XL relative shift: `[B, N, T, 2T−1]` (relative offset −(T−1)…T−1 on the
  last axis) → `[B, N, T, T]` with out[i, j] = x[i, j − i + T − 1] (ref :329)."""
  b, n, t, _ = x.shape
  idx = (torch.arange(t, device=x.device).unsqueeze(0) -
         torch.arange(t, device=x.device).unsqueeze(1) + (t - 1))
  return x.gather(-1, idx.expand(b, n, t, t))


def AttenLogits(query, key, qlayer=None):
  del qlayer
  return torch.einsum('BTNH,BSNH->BNTS', query, key)


def AttenContext(probs, value, qlayer=None):
  del qlayer
  return torch.einsum('BNTS,BSNH->BTNH', probs, value)


class PositionalAttenLogits(quant_utils.QuantizableLayer):
  """Transformer-XL terms (b) and (d): q·R and v·R with relative shift (ref :384)."""

  def AttenLogitsXL(self, query, key, abs_pos_emb, content_bias, positional_bias,
                    skip_term_b=False):
    """Transformer-XL logits (ref :497): query/key [B,T,N,H]; abs_pos_emb [2T−1,N,H] with row
    r the embedding of distance r−(T−1); biases [N,H] → [B,N,T,T]."""
    return self._AttenLogits(query, key, abs_pos_emb, content_bias, positional_bias,
                             skip_term_b)

  @staticmethod
  def RelPositionBias(content, abs_pos_emb, skip_term_b=False):
    """Relative-position logits (ref :392). `content` is [B,T,N,H] (or the bare [N,H] bias when
    `skip_term_b`), `abs_pos_emb` [2T−1,N,H] with row r the embedding of distance r−(T−1).
    out[b,n,i,j] = content[b,i,n] · abs_pos_emb[i−j+T−1, n]; shape [B,N,T,T] ([N,T,T] when
    `skip_term_b`). One gather over the distance axis replaces the pad-and-reshape skew."""
    r = abs_pos_emb.shape[0]
    t = (r + 1) // 2
    ar = torch.arange(t, device=abs_pos_emb.device)
    idx = ar.unsqueeze(1) - ar.unsqueeze(0) + (t - 1)                      # [T(i), T(j)]
    if skip_term_b:
      full = torch.einsum('NH,RNH->NR', content, abs_pos_emb)              # [N, 2T-1]
      return full[:, idx]
    full = torch.einsum('BTNH,RNH->BNTR', content, abs_pos_emb)
    return full.gather(-1, idx.expand(*full.shape[:2], t, t))

  @staticmethod
  def _ValidateBiases(content_bias, positional_bias, n, h):
    for b in (content_bias, positional_bias):
      if b is not None and tuple(b.shape) != (n, h):
        raise ValueError(f'bias shape {tuple(b.shape)} != {(n, h)}')

  def _AttenLogits(self, query, key, abs_pos_emb, content_bias=None, positional_bias=None,
                   skip_term_b=False):
    """term (a)+(c) = (q+u)·k, term (b)+(d) = RelPositionBias(q+v) (ref :448). query/key
    [B,T,N,H] → [B,N,T,T]."""
    b, t, n, h = query.shape
    self._ValidateBiases(content_bias, positional_bias, n, h)
    content = query if content_bias is None else query + content_bias
    term_ac = torch.einsum('BTNH,BSNH->BNTS', content, key)
    if skip_term_b:
      if positional_bias is None:
        return term_ac
      return term_ac + self.RelPositionBias(positional_bias, abs_pos_emb, True).unsqueeze(0)
    pos = query if positional_bias is None else query + positional_bias
    return term_ac + self.RelPositionBias(pos, abs_pos_emb, False)

  def AttenLogitsRPE(self, query, key, abs_pos_emb):
    """Shaw-style relative position logits: no biases (ref :531)."""
    return self._AttenLogits(query, key, abs_pos_emb)

  def AttenLogitsXLOneStep(self, query, key, abs_pos_emb, content_bias, positional_bias,
                           skip_term_b=False):
    """One decode step (ref :562): query [B,N,H], key [S,B,N,H], abs_pos_emb [S,N,H] (all
    sequences at the same time step) or [B,S,N,H] (per-sequence steps) → [S,B,N]."""
    s, b = key.shape[:2]
    _, n, h = query.shape
    key = key.reshape(s, b, n, h)
    self._ValidateBiases(content_bias, positional_bias, n, h)
    term_ac = torch.einsum('BNH,SBNH->SBN', query + content_bias, key)
    synced = abs_pos_emb.dim() == 3
    if not skip_term_b:
      pos = query + positional_bias
      term_bd = torch.einsum('BNH,SNH->SBN' if synced else 'BNH,BSNH->SBN', pos, abs_pos_emb)
    elif synced:
      term_bd = torch.einsum('NH,SNH->SN', positional_bias, abs_pos_emb).unsqueeze(1)
    else:
      term_bd = torch.einsum('NH,BSNH->SBN', positional_bias, abs_pos_emb)
    return term_ac + term_bd

  def AttenLogitsRPEOneStep(self, query, key, abs_pos_emb):
    """One decode step of RPE (ref :628): query [B,N,H], key [S,B,N,H], abs_pos_emb
    [S,1,N,H] → [S,B,N]."""
    s, b = key.shape[:2]
    _, n, h = query.shape
    return torch.einsum('BNH,SBNH->SBN', query, key.reshape(s, b, n, h) + abs_pos_emb)


class KMeansClusteringForAtten(base_layer.BaseLayer):
  """Online spherical k-means over attention heads (routing attention) (ref :656)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('num_clusters', 0, 'Clusters per head.')
    p.Define('num_heads', 1, 'Heads.')
    p.Define('dim_per_head', 0, 'Head dim.')
    p.Define('decay', 0.999, 'EMA decay of the centroids.')
    p.Define('epsilon', 1e-6, 'Normalisation epsilon.')
    p.Define('apply_layer_norm', True, 'Normalise inputs before clustering.')
    return p

  def _CreateLayerVariables(self):
    p = self.params
    self.CreateVariable('means', WeightParams(
        [p.num_heads, p.num_clusters, p.dim_per_head], WeightInit.Gaussian(1.0), p.dtype),
        trainable=False)

  def FProp(self, theta, x, paddings=None, update=False):
    """x [B,L,N,H] → (dists [B,L,N,K], loss). Centroids EMA-updated when `update`."""
    p = self.params
    if p.apply_layer_norm:
      x = F.layer_norm(x, x.shape[-1:], eps=p.epsilon)
    means = F.normalize(theta.means.float(), dim=-1)
    xn = F.normalize(x.float(), dim=-1)
    sim = torch.einsum('BLNH,NKH->BLNK', xn, means)
    dists = 1.0 - sim
    nearest = dists.min(-1)
    w = torch.ones_like(nearest.values) if paddings is None else (
        1.0 - paddings.float()).unsqueeze(-1).expand_as(nearest.values)
    loss = (nearest.values * w).sum() / w.sum().clamp_min(1.0)
    if update and not self.do_eval:
      with torch.no_grad():
        oh = F.one_hot(nearest.indices, p.num_clusters).float() * w.unsqueeze(-1)
        sums = torch.einsum('BLNK,BLNH->NKH', oh, xn)
        cnt = oh.sum((0, 1)).unsqueeze(-1)
        new = torch.where(cnt > 0, sums / cnt.clamp_min(1.0), means)
        self.vars.means.data.mul_(p.decay).add_(new.to(self.vars.means.dtype),
                                                alpha=1 - p.decay)
    return dists, loss


def ComputeSparseAttention(q, k, v, sparsity_indices, paddings=None):
  """q [B,N,T,H] attends only to keys `sparsity_indices[b,n,t,:]` (−1: none) (ref :891)."""
  b, n, t, h = q.shape
  s = k.shape[2]
  idx = sparsity_indices.clamp_min(0)
  kk = k.unsqueeze(2).expand(b, n, t, s, h).gather(3, idx.unsqueeze(-1).expand(*idx.shape, h))
  vv = v.unsqueeze(2).expand(b, n, t, s, h).gather(3, idx.unsqueeze(-1).expand(*idx.shape, h))
  logits = torch.einsum('BNTH,BNTWH->BNTW', q, kk) / math.sqrt(h)
  mask = sparsity_indices < 0
  if paddings is not None:
    mask = mask | (paddings.unsqueeze(1).unsqueeze(2).expand(b, n, t, s).gather(3, idx) > 0)
  logits = logits.masked_fill(mask, -1e30)
  probs = torch.softmax(logits, -1) * (~mask).float()
  return torch.einsum('BNTW,BNTWH->BNTH', probs.to(vv.dtype), vv), probs
