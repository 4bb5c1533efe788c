"""FAVOR+ linear attention (Performer) (ref `lingvo/core/favor_attention.py`).

q, k `[B, L, H, D]` are mapped to non-negative random features φ(·) `[B, L, H, M]`;
attention = φ(q)(φ(k)ᵀ v) / φ(q)(φ(k)ᵀ 1). Causal attention uses prefix sums,
computed chunk-wise so the `[L, M, D]` running state never exceeds one chunk.
"""

from __future__ import annotations

import math

import torch


def next_seed(current_seed):  # pylint: disable=invalid-name
  return None if current_seed is None else current_seed + 1


def create_projection_matrix(nb_random_projections, dim, seed=0, scaling=0):  # pylint: disable=invalid-name
  """Block-orthogonal Gaussian matrix `[M, D]`; scaling 0: χ-distributed row norms,
  1: every row has norm √D."""
  g = torch.Generator().manual_seed(int(seed or 0))
  blocks = []
  for _ in range(nb_random_projections // dim):
    q, _ = torch.linalg.qr(torch.randn(dim, dim, generator=g))
    blocks.append(q.t())
  rem = nb_random_projections - (nb_random_projections // dim) * dim
  if rem:
    q, _ = torch.linalg.qr(torch.randn(dim, dim, generator=g))
    blocks.append(q.t()[:rem])
  mat = torch.cat(blocks, 0)
  if scaling == 0:
    mult = torch.randn(nb_random_projections, dim, generator=g).norm(dim=1)
  else:
    mult = math.sqrt(dim) * torch.ones(nb_random_projections)
  return mult.unsqueeze(1) * mat


def relu_kernel_transformation(data, is_query, projection_matrix=None, numerical_stabilizer=0.001):  # pylint: disable=invalid-name
  del is_query
  if projection_matrix is None:
    return torch.relu(data) + numerical_stabilizer
  ratio = 1.0 / math.sqrt(projection_matrix.shape[0])
  return torch.relu(ratio * torch.einsum('blhd,md->blhm', data, projection_matrix.to(data))) + \
      numerical_stabilizer


def softmax_kernel_transformation(data, is_query, projection_matrix=None,  # pylint: disable=invalid-name
                                  numerical_stabilizer=0.000001):
  d = data.shape[-1]
  data = data * (d ** -0.25)
  ratio = 1.0 / math.sqrt(projection_matrix.shape[0])
  dash = torch.einsum('blhd,md->blhm', data, projection_matrix.to(data))
  diag = (data * data).sum(-1, keepdim=True) / 2.0
  mx = dash.amax(-1, keepdim=True) if is_query else dash.amax((1, 3), keepdim=True)
  return ratio * (torch.exp(dash - diag - mx) + numerical_stabilizer)


def cossim_kernel_transformation(data, is_query, projection_matrix=None,  # pylint: disable=invalid-name
                                 numerical_stabilizer=0.0, randomized=True):
  del is_query, numerical_stabilizer
  data = torch.nn.functional.normalize(data, dim=-1)
  if not randomized or projection_matrix is None:
    return data
  ratio = 1.0 / math.sqrt(projection_matrix.shape[0])
  return ratio * torch.einsum('blhd,md->blhm', data, projection_matrix.to(data))


def noncausal_numerator(qs, ks, vs):  # pylint: disable=invalid-name
  """qs, ks `[L,B,H,M]`, vs `[L,B,H,D]` → `[L,B,H,D]`."""
  kvs = torch.einsum('lbhm,lbhd->bhmd', ks, vs)
  return torch.einsum('lbhm,bhmd->lbhd', qs, kvs)


def noncausal_denominator(qs, ks):  # pylint: disable=invalid-name
  return torch.einsum('lbhm,bhm->lbh', qs, ks.sum(0))


def causal_numerator(qs, ks, vs, chunk=128):  # pylint: disable=invalid-name
  """Prefix-sum attention, processed in chunks of `chunk` steps."""
  l = qs.shape[0]
  state = torch.zeros(qs.shape[1], qs.shape[2], qs.shape[3], vs.shape[3],
                      device=qs.device, dtype=qs.dtype)
  outs = []
  for s in range(0, l, chunk):
    q, k, v = qs[s:s + chunk], ks[s:s + chunk], vs[s:s + chunk]
    kv = torch.einsum('lbhm,lbhd->lbhmd', k, v).cumsum(0) + state
    outs.append(torch.einsum('lbhm,lbhmd->lbhd', q, kv))
    state = kv[-1]
  return torch.cat(outs, 0)


def causal_denominator(qs, ks, chunk=128):  # pylint: disable=invalid-name
  l = qs.shape[0]
  state = torch.zeros_like(ks[0])
  outs = []
  for s in range(0, l, chunk):
    kc = ks[s:s + chunk].cumsum(0) + state
    outs.append((qs[s:s + chunk] * kc).sum(-1))
    state = kc[-1]
  return torch.cat(outs, 0)


_ITER_CHUNK_SIZE = 64


def chunked_causal_numerator_func(qs, ks, vs, chunk=None):  # pylint: disable=invalid-name
  """Forward of the causal numerator in chunks → (`[L,B,H,D]` result, last prefix-sum state
  `[B,H,M,D]`) (ref :336). Only one chunk of the `[chunk,B,H,M,D]` prefix sums is alive."""
  chunk = chunk or _ITER_CHUNK_SIZE
  sums = torch.zeros(qs.shape[1], qs.shape[2], qs.shape[3], vs.shape[3], device=qs.device,
                     dtype=qs.dtype)
  outs = []
  for s in range(0, qs.shape[0], chunk):
    kv = torch.einsum('sijk,sijl->sijkl', ks[s:s + chunk], vs[s:s + chunk]).cumsum(0) + sums
    outs.append(torch.einsum('sijkl,sijk->sijl', kv, qs[s:s + chunk]))
    sums = kv[-1]
  return torch.cat(outs, 0), sums


def chunked_causal_numerator_grad(qs, ks, vs, sums, res_grad, chunk=None):  # pylint: disable=invalid-name
  """Backward of the causal numerator (ref :370), walking the sequence from the end: the
  prefix sums are *un-done* from the final state (S_{t-1} = S_t − k_t ⊗ v_t) instead of being
  stored, and the suffix sums G_t = Σ_{s≥t} q_s ⊗ g_s are built alongside.
  dq_t = S_t·g_t, dk_t = G_t·v_t, dv_t = k_t·G_t."""
  chunk = chunk or _ITER_CHUNK_SIZE
  l = qs.shape[0]
  state = sums                                              # S at the end of the chunk
  gsum = torch.zeros_like(sums)                             # G just after the chunk
  dq, dk, dv = [], [], []
  for e in range(l, 0, -chunk):
    s = max(e - chunk, 0)
    q, k, v, g = qs[s:e], ks[s:e], vs[s:e], res_grad[s:e]
    kv = torch.einsum('sijk,sijl->sijkl', k, v)
    start = state - kv.sum(0)                               # S just before the chunk
    pref = kv.cumsum(0) + start                             # S_t inside the chunk
    dq.append(torch.einsum('sijkl,sijl->sijk', pref, g))
    qg = torch.einsum('sijk,sijl->sijkl', q, g)
    suff = qg.flip(0).cumsum(0).flip(0) + gsum              # G_t inside the chunk
    dk.append(torch.einsum('sijkl,sijl->sijk', suff, v))
    dv.append(torch.einsum('sijkl,sijk->sijl', suff, k))
    state, gsum = start, suff[0]
  return torch.cat(dq[::-1], 0), torch.cat(dk[::-1], 0), torch.cat(dv[::-1], 0)


def chunked_causal_denominator_func(qs, ks, chunk=None):  # pylint: disable=invalid-name
  """Forward of the causal normaliser in chunks → (`[L,B,H]`, last key prefix sum) (ref :451)."""
  chunk = chunk or _ITER_CHUNK_SIZE
  sums = torch.zeros_like(ks[0])
  outs = []
  for s in range(0, qs.shape[0], chunk):
    kc = ks[s:s + chunk].cumsum(0) + sums
    outs.append((qs[s:s + chunk] * kc).sum(-1))
    sums = kc[-1]
  return torch.cat(outs, 0), sums


def chunked_causal_denominator_grad(qs, ks, sums, res_grad, chunk=None):  # pylint: disable=invalid-name
  """Backward of the causal normaliser (ref :482): dq_t = g_t·(Σ_{s≤t} k_s),
  dk_t = Σ_{s≥t} g_s·q_s, prefix sums un-done from the final state."""
  chunk = chunk or _ITER_CHUNK_SIZE
  l = qs.shape[0]
  state = sums
  gsum = torch.zeros_like(sums)
  dq, dk = [], []
  for e in range(l, 0, -chunk):
    s = max(e - chunk, 0)
    q, k, g = qs[s:e], ks[s:e], res_grad[s:e].unsqueeze(-1)
    start = state - k.sum(0)
    dq.append((k.cumsum(0) + start) * g)
    suff = (q * g).flip(0).cumsum(0).flip(0) + gsum
    dk.append(suff)
    state, gsum = start, suff[0]
  return torch.cat(dq[::-1], 0), torch.cat(dk[::-1], 0)


class _ChunkedCausalNumerator(torch.autograd.Function):

  @staticmethod
  def forward(ctx, qs, ks, vs):
    out, sums = chunked_causal_numerator_func(qs, ks, vs)
    ctx.save_for_backward(qs, ks, vs, sums)
    return out

  @staticmethod
  def backward(ctx, g):
    qs, ks, vs, sums = ctx.saved_tensors
    return chunked_causal_numerator_grad(qs, ks, vs, sums, g.contiguous())


class _ChunkedCausalDenominator(torch.autograd.Function):

  @staticmethod
  def forward(ctx, qs, ks):
    out, sums = chunked_causal_denominator_func(qs, ks)
    ctx.save_for_backward(qs, ks, sums)
    return out

  @staticmethod
  def backward(ctx, g):
    qs, ks, sums = ctx.saved_tensors
    return chunked_causal_denominator_grad(qs, ks, sums, g.contiguous())


def chunked_causal_numerator(qs, ks, vs):  # pylint: disable=invalid-name
  """Causal numerator whose backward keeps O(chunk·M·D) instead of O(L·M·D) activations
  (ref :432) — what `favor_attention(causal=True)` runs."""
  return _ChunkedCausalNumerator.apply(qs, ks, vs)


def chunked_causal_denominator(qs, ks):  # pylint: disable=invalid-name
  return _ChunkedCausalDenominator.apply(qs, ks)


def favor_attention(query, key, value, paddings, kernel_transformation, causal,  # pylint: disable=invalid-name
                    projection_matrix=None):
  """query/key/value `[B, L, H, D]`, paddings `[B, L]` → `[B, L, H, D]`."""
  qp = kernel_transformation(query, True, projection_matrix)
  kp = kernel_transformation(key, False, projection_matrix)
  if paddings is not None:
    kp = kp * (1.0 - paddings.to(kp.dtype)).unsqueeze(-1).unsqueeze(-1)
  qp, kp, v = qp.transpose(0, 1), kp.transpose(0, 1), value.transpose(0, 1)
  if causal:
    num, den = chunked_causal_numerator(qp, kp, v), chunked_causal_denominator(qp, kp)
  else:
    num, den = noncausal_numerator(qp, kp, v), noncausal_denominator(qp, kp)
  out = num / den.unsqueeze(-1).clamp_min(1e-9)
  return out.transpose(0, 1)
