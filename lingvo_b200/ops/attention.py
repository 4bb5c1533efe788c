"""Attention core shared by every attention layer.

`dot_product_attention` takes batch-major `[B, T, N, H]` projections plus an
additive fp32/bf16 bias broadcastable to `[B, N, T, S]` (padding, causal,
segment and relative-position terms are all folded into that one bias, the way
the reference adds them to the logits, ref
`lingvo/core/batch_major_attention.py:943-1044`).

Dispatch:
  * CUDA, no probabilities requested, no logit cap / extra logit → fused
    flash-style kernel (online softmax; logits never reach HBM).
  * otherwise → explicit fp32-softmax path (also the numerics oracle).
"""

from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F


def _ExpandKv(k, n):
  """GQA/MQA: repeat kv heads up to `n` query heads."""
  hk = k.shape[2]
  if hk == n:
    return k
  if hk == 1:
    return k.expand(k.shape[0], k.shape[1], n, k.shape[3])
  return k.repeat_interleave(n // hk, dim=2)


def attention_ref(q, k, v, bias=None, scale=1.0, logit_cap=0.0,
                  extra_logit=None, dropout_prob=0.0, return_probs=False):
  """Explicit softmax(q·kᵀ·scale + bias)·v with fp32 logits."""
  n = q.shape[2]
  k, v = _ExpandKv(k, n), _ExpandKv(v, n)
  logits = torch.einsum('BTNH,BSNH->BNTS', q.float(), k.float())
  if scale != 1.0:
    logits = logits * scale
  if logit_cap and logit_cap > 0:
    logits = logit_cap * torch.tanh(logits / logit_cap)
  if bias is not None:
    logits = logits + bias.float()
  if extra_logit is not None:
    extra = torch.full_like(logits[..., :1], float(extra_logit))
    probs = torch.softmax(torch.cat([logits, extra], -1), -1)[..., :-1]
  else:
    probs = torch.softmax(logits, -1)
  pd = probs.to(v.dtype)
  if dropout_prob:
    pd = F.dropout(pd, dropout_prob, training=True)
  ctx = torch.einsum('BNTS,BSNH->BTNH', pd, v)
  return (ctx, probs) if return_probs else ctx


def dot_product_attention(q, k, v, bias=None, scale=1.0, logit_cap=0.0,
                          extra_logit=None, dropout_prob=0.0,
                          return_probs=False):
  """`[B,T,N,H]` × `[B,S,Nkv,H]` → context `[B,T,N,H]` (and probs if asked)."""
  fused = (q.is_cuda and not return_probs and not logit_cap and
           extra_logit is None and q.dtype in (torch.bfloat16, torch.float16))
  if not fused:
    return attention_ref(q, k, v, bias, scale, logit_cap, extra_logit,
                         dropout_prob, return_probs)
  n = q.shape[2]
  k, v = _ExpandKv(k, n), _ExpandKv(v, n)
  b, t, s = q.shape[0], q.shape[1], k.shape[1]
  mask = None
  if bias is not None:
    mask = bias.to(q.dtype)
    if mask.dim() < 4:
      mask = mask.reshape((1,) * (4 - mask.dim()) + tuple(mask.shape))
    mask = mask.expand(b, n, t, s)
  o = F.scaled_dot_product_attention(
      q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), attn_mask=mask,
      dropout_p=dropout_prob, scale=scale)
  return o.transpose(1, 2)
