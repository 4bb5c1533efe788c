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


chunked_causal_numerator = causal_numerator        # ref :432
chunked_causal_denominator = causal_denominator    # ref :532


def favor_attention(query, key, value, paddings, kernel_transformation, causal,  # pylint: disable=invalid-name
                    projection_matrix=None):
  """query/key/value `[B, L, H, D]`, paddings `[B, L]` → `[B, L, H, D]`."""
  qp = kernel_transformation(query, True, projection_matrix)
  kp = kernel_transformation(key, False, projection_matrix)
  if paddings is not None:
    kp = kp * (1.0 - paddings.to(kp.dtype)).unsqueeze(-1).unsqueeze(-1)
  qp, kp, v = qp.transpose(0, 1), kp.transpose(0, 1), value.transpose(0, 1)
  if causal:
    num, den = causal_numerator(qp, kp, v), causal_denominator(qp, kp)
  else:
    num, den = noncausal_numerator(qp, kp, v), noncausal_denominator(qp, kp)
  out = num / den.unsqueeze(-1).clamp_min(1e-9)
  return out.transpose(0, 1)
