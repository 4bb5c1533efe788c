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

  def AttenLogitsXL(self, content_logits, query, abs_pos_emb, content_bias, positional_bias,
                    skip_term_b=False):
    """content_logits [B,N,T,S]; query [B,T,N,H]; abs_pos_emb [2T−1,N,H] (distance
    T−1 … −(T−1)); biases [N,H]."""
    t = query.shape[1]
    term_d_src = positional_bias if skip_term_b else None
    q = query + (positional_bias if not skip_term_b else 0)
    term_bd = torch.einsum('BTNH,RNH->BNTR', q, abs_pos_emb) if not skip_term_b else \
        torch.einsum('NH,RNH->NR', term_d_src, abs_pos_emb).view(1, -1, 1, 2 * t - 1).expand(
            query.shape[0], -1, t, -1)
    idx = (torch.arange(t, device=query.device).unsqueeze(0) -
           torch.arange(t, device=query.device).unsqueeze(1) + (t - 1))
    term_bd = term_bd.gather(-1, idx.expand(*term_bd.shape[:2], t, t))
    del content_bias
    return content_logits + term_bd


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
