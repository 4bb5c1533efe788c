"""GShard layers: gating, MoE dispatch/combine, helpers.

Reference `lingvo/core/gshard_layers.py`:
  * `Top2GatingOnLogits` (:1932-2364) — reproduced verbatim in *math* as the
    dense `GSEC` oracle (`Top2GatingOnLogits`), and re-designed as an
    **index-based** gate (`Top2GatingIndices`) that returns
    `(expert, slot, gate) × 2` per token — what the fused sm_100a gate kernel
    (`ops.moe.top2_gate`) produces. Both implement: softmax over E, top-1,
    second-expert policy (`all` / `sampling` (Gumbel) / `random` threshold),
    exclusive cumsum positions with first-choice priority (:2230, :2294),
    capacity rounding ↑4 (:2062-2076), renorm before/after capacity
    (`legacy_mtf_behavior`), aux loss `mean(density_1·density_1_proxy)·E²`
    (:2250-2255).
  * `ComputeGating` (:2840-2964), `FeedForwardNetworksApplyGating`
    (:2992-3163) — dense einsum oracle `GSEC,GSM->EGCM` … `GSEC,GECM->GSM`.
  * `MoEApplyIndexed` — the B200 path: permute-scatter dispatch, grouped
    tcgen05 expert GEMMs, gated 2-row gather combine; with expert parallelism
    the scatter/gather cross NVLink (`lingvo_b200.parallel.ep`).
  * `HashGatingOnLogits` (:2367), `TokenShufflingOnlogits` (:2496),
    sentence/task-level gating (:3359, :3450), `ReshapeInputLayer` (:1614),
    `StateLayer` (:1186-1430), `CausalDepthwiseConv1DLayer`, `VarLayer`.
"""

from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F

from lingvo_b200 import ops
from lingvo_b200.core import activations
from lingvo_b200.core import base_layer
from lingvo_b200.core import gshard_utils
from lingvo_b200.core import py_utils
from lingvo_b200.core import summary_utils
from lingvo_b200.core.nested_map import NestedMap

Split = gshard_utils.Split
MeshSplit = gshard_utils.MeshSplit
WeightParams = py_utils.WeightParams
WeightInit = py_utils.WeightInit


def EinsumWithModelDim(equation, x, y, model_dim_reshape_segments=None,
                       name=None):
  """Einsum where `M` may have been reshaped to [segments, M/segments]."""
  if model_dim_reshape_segments is None:
    return torch.einsum(equation, x, y)
  if isinstance(model_dim_reshape_segments, int):
    model_dim_reshape_segments = [model_dim_reshape_segments]
  letters = 'NOPQR'[:len(model_dim_reshape_segments)]
  return torch.einsum(equation.replace('M', letters + 'M'), x, y)


def ExpertCapacity(group_size: int, experts_dim: int, expert_capacity_dim: int,
                   capacity_factor: Optional[float]) -> int:
  """C = max(c_dim, ⌊S·cf/E⌋ rounded ↑ to a multiple of 4) (:2062-2076)."""
  if capacity_factor is not None and capacity_factor > 0:
    auto = int((group_size * capacity_factor) / experts_dim)
    if auto == 0:
      auto = 4
    if expert_capacity_dim < auto:
      expert_capacity_dim = auto
      while expert_capacity_dim % 4:
        expert_capacity_dim += 1
  return expert_capacity_dim


def _Gumbel(shape, device, dtype, seeds=None):
  gen = None
  if seeds is not None:
    gen = torch.Generator(device=device)
    gen.manual_seed(((int(seeds[0]) & 0x7fffffff) << 31) ^ int(seeds[1]))
  u = torch.rand(shape, device=device, dtype=torch.float32, generator=gen)
  u = u.clamp_min(1e-20)
  return (-torch.log(-torch.log(u))).to(dtype)


def Top2GatingOnLogits(inputs, paddings, logits, num_devices, experts_dim,
                       expert_capacity_dim, fprop_dtype, use_xla_sharding=True,
                       second_expert_policy='all', second_expert_threshold=0.0,
                       legacy_mtf_behavior=True, capacity_factor=None,
                       importance=None, mask_dtype=None,
                       expert_padding_idx=None, seeds=None):
  """Dense oracle: returns (aux_loss, combine_tensor GSEC, dispatch GSEC)."""
  del inputs, num_devices, use_xla_sharding
  if mask_dtype is None:
    mask_dtype = fprop_dtype
  if expert_padding_idx:
    very_neg = torch.finfo(logits.dtype).max * -0.7
    m = torch.zeros_like(logits, dtype=torch.bool)
    m[..., list(expert_padding_idx)] = True
    logits = torch.where(m, torch.full_like(logits, very_neg), logits)
  raw_gates = torch.softmax(logits, dim=-1).to(fprop_dtype)
  expert_capacity_dim = ExpertCapacity(int(logits.shape[1]), experts_dim,
                                       expert_capacity_dim, capacity_factor)
  summary_utils.scalar('expert_capacity', expert_capacity_dim)

  index_1 = raw_gates.argmax(dim=-1)
  mask_1 = F.one_hot(index_1, experts_dim).to(mask_dtype)
  density_1_proxy = raw_gates
  if importance is not None:
    one = (importance == 1.0)
    mask_1 = mask_1 * one.unsqueeze(-1).to(mask_1.dtype)
    density_1_proxy = density_1_proxy * one.unsqueeze(-1).to(raw_gates.dtype)
  else:
    importance = torch.ones_like(mask_1[..., 0])
    if paddings is not None:
      nonpad = 1.0 - paddings.to(mask_1.dtype)
      mask_1 = mask_1 * nonpad.unsqueeze(-1)
      density_1_proxy = density_1_proxy * nonpad.unsqueeze(-1).to(raw_gates.dtype)
      importance = nonpad
  gate_1 = (raw_gates * mask_1.to(raw_gates.dtype)).sum(-1)
  gates_without_top_1 = raw_gates * (1.0 - mask_1.to(raw_gates.dtype))
  if second_expert_policy == 'sampling':
    noise = _Gumbel(logits.shape, logits.device, logits.dtype, seeds)
    very_neg = torch.finfo(logits.dtype).max * -0.7
    upd = torch.where(mask_1 > 0, torch.full_like(logits, very_neg), logits)
    index_2 = (upd + noise).argmax(dim=-1)
  else:
    index_2 = gates_without_top_1.argmax(dim=-1)
  mask_2 = F.one_hot(index_2, experts_dim).to(mask_dtype)
  if paddings is not None:
    mask_2 = mask_2 * (importance > 0).unsqueeze(-1).to(mask_2.dtype)
  gate_2 = (gates_without_top_1 * mask_2.to(raw_gates.dtype)).sum(-1)
  if legacy_mtf_behavior:
    denom = gate_1 + gate_2 + 1e-9
    gate_1 = gate_1 / denom
    gate_2 = gate_2 / denom

  position_in_expert_1 = torch.cumsum(mask_1, dim=-2) - mask_1
  capacity = float(expert_capacity_dim)
  if legacy_mtf_behavior:
    density_denom = 1.0
  else:
    density_denom = importance.to(fprop_dtype).mean(dim=1)[:, None] + 1e-6
  density_1 = mask_1.to(fprop_dtype).mean(dim=-2) / density_denom
  density_1_proxy = density_1_proxy.mean(dim=-2) / density_denom
  aux_loss = (density_1_proxy * density_1).mean() * (experts_dim * experts_dim)

  def over_capacity(mask, pos, name):
    ge = ((mask * pos) >= capacity).float().sum()
    summary_utils.scalar(name, ge)
    summary_utils.scalar(name + '_ratio', ge / torch.clamp(mask.float().sum(),
                                                           min=1.0))
  over_capacity(mask_1, position_in_expert_1, 'over_capacity_1')
  mask_1 = mask_1 * (position_in_expert_1 < capacity).to(mask_1.dtype)
  position_in_expert_1 = (position_in_expert_1 * mask_1).sum(-1)
  mask_1_count = mask_1.sum(dim=-2)
  mask_1_flat = mask_1.sum(dim=-1)

  if second_expert_policy == 'random':
    gen = None
    if seeds is not None:
      gen = torch.Generator(device=logits.device)
      gen.manual_seed(((int(seeds[0]) & 0x7fffffff) << 31) ^ int(seeds[1]) ^ 7)
    u = torch.rand(gate_2.shape, device=gate_2.device, generator=gen).to(
        gate_2.dtype)
    sampled_2 = u < (gate_2 / max(second_expert_threshold, 1e-9))
    gate_2 = gate_2 * sampled_2.to(gate_2.dtype)
    mask_2 = mask_2 * sampled_2.unsqueeze(-1).to(mask_2.dtype)
  elif second_expert_policy not in ('all', 'sampling'):
    raise ValueError(second_expert_policy)

  position_in_expert_2 = (torch.cumsum(mask_2, dim=-2) - mask_2 +
                          mask_1_count.unsqueeze(-2))
  over_capacity(mask_2, position_in_expert_2, 'over_capacity_2')
  mask_2 = mask_2 * (position_in_expert_2 < capacity).to(mask_2.dtype)
  position_in_expert_2 = (position_in_expert_2 * mask_2).sum(-1)
  mask_2_flat = mask_2.sum(dim=-1)
  gate_1 = gate_1 * mask_1_flat.to(gate_1.dtype)
  gate_2 = gate_2 * mask_2_flat.to(gate_2.dtype)
  if not legacy_mtf_behavior:
    denom = gate_1 + gate_2
    denom = torch.where(denom > 0, denom, torch.ones_like(denom))
    gate_1 = gate_1 / denom
    gate_2 = gate_2 / denom

  def part(gate, mask_flat, index, pos):
    b = F.one_hot(pos.long(), expert_capacity_dim).to(fprop_dtype)
    a = (gate * mask_flat.to(fprop_dtype)).unsqueeze(-1) * F.one_hot(
        index, experts_dim).to(fprop_dtype)
    return torch.einsum('...GSE,...GSC->...GSEC', a, b)

  combine_tensor = (part(gate_1, mask_1_flat, index_1, position_in_expert_1) +
                    part(gate_2, mask_2_flat, index_2, position_in_expert_2))
  dispatch_tensor = (combine_tensor != 0).to(fprop_dtype)
  return aux_loss, combine_tensor, dispatch_tensor


def Top2GatingIndices(logits, paddings, experts_dim, expert_capacity_dim,
                      fprop_dtype=torch.float32, second_expert_policy='all',
                      second_expert_threshold=0.0, legacy_mtf_behavior=True,
                      capacity_factor=None, seeds=None, use_kernel=True):
  """Index-form top-2 gate.

  Args: logits `[G, S, E]` (fp32 recommended), paddings `[G, S]` or None.
  Returns NestedMap:
    index `[2, G, S]` int32 expert ids, pos `[2, G, S]` int32 slot in expert,
    gate `[2, G, S]` fp32 combine weights (0 ⇒ dropped / not dispatched; the
    gate is differentiable w.r.t. logits), aux_loss (scalar), capacity (int).
  Semantically identical to `Top2GatingOnLogits` (tested bit-for-bit on the
  resulting combine tensor).
  """
  g, s, e = logits.shape
  assert e == experts_dim
  cap = ExpertCapacity(s, experts_dim, expert_capacity_dim, capacity_factor)
  raw = torch.softmax(logits.float(), dim=-1)
  nonpad = None if paddings is None else (1.0 - paddings.float())
  index_1 = raw.argmax(-1)
  oh1 = F.one_hot(index_1, e).float()
  proxy = raw
  if nonpad is not None:
    oh1 = oh1 * nonpad.unsqueeze(-1)
    proxy = raw * nonpad.unsqueeze(-1)
  gate_1 = torch.gather(raw, -1, index_1.unsqueeze(-1)).squeeze(-1)
  if nonpad is not None:
    gate_1 = gate_1 * nonpad
  wo1 = raw.scatter(-1, index_1.unsqueeze(-1), 0.0) if nonpad is None else (
      raw * (1.0 - oh1))
  if second_expert_policy == 'sampling':
    noise = _Gumbel(logits.shape, logits.device, torch.float32, seeds)
    very_neg = torch.finfo(torch.float32).max * -0.7
    upd = torch.where(oh1 > 0, torch.full_like(raw, very_neg), logits.float())
    index_2 = (upd + noise).argmax(-1)
  else:
    index_2 = wo1.argmax(-1)
  oh2 = F.one_hot(index_2, e).float()
  if nonpad is not None:
    oh2 = oh2 * (nonpad > 0).float().unsqueeze(-1)
  gate_2 = (wo1 * oh2).sum(-1)
  if legacy_mtf_behavior:
    denom = gate_1 + gate_2 + 1e-9
    gate_1, gate_2 = gate_1 / denom, gate_2 / denom
  pos1_all = torch.cumsum(oh1, dim=1) - oh1
  density_denom = 1.0 if legacy_mtf_behavior else (
      (nonpad if nonpad is not None else torch.ones_like(gate_1)).mean(1)[:, None]
      + 1e-6)
  density_1 = oh1.mean(dim=1) / density_denom
  aux_loss = ((proxy.mean(dim=1) / density_denom) * density_1).mean() * (e * e)
  keep1 = oh1 * (pos1_all < cap).float()
  pos1 = (pos1_all * keep1).sum(-1)
  count1 = keep1.sum(dim=1)
  flat1 = keep1.sum(-1)
  if second_expert_policy == 'random':
    gen = None
    if seeds is not None:
      gen = torch.Generator(device=logits.device)
      gen.manual_seed(((int(seeds[0]) & 0x7fffffff) << 31) ^ int(seeds[1]) ^ 7)
    u = torch.rand(gate_2.shape, device=gate_2.device, generator=gen)
    sampled = (u < gate_2 / max(second_expert_threshold, 1e-9)).float()
    gate_2 = gate_2 * sampled
    oh2 = oh2 * sampled.unsqueeze(-1)
  pos2_all = torch.cumsum(oh2, dim=1) - oh2 + count1.unsqueeze(1)
  keep2 = oh2 * (pos2_all < cap).float()
  pos2 = (pos2_all * keep2).sum(-1)
  flat2 = keep2.sum(-1)
  gate_1 = gate_1 * flat1
  gate_2 = gate_2 * flat2
  if not legacy_mtf_behavior:
    denom = gate_1 + gate_2
    denom = torch.where(denom > 0, denom, torch.ones_like(denom))
    gate_1, gate_2 = gate_1 / denom, gate_2 / denom
  with torch.no_grad():
    over1 = (oh1.sum() - keep1.sum())
    over2 = (oh2.sum() - keep2.sum())
  summary_utils.scalar('expert_capacity', cap)
  summary_utils.scalar('over_capacity_1', over1)
  summary_utils.scalar('over_capacity_2', over2)
  return NestedMap(
      index=torch.stack([index_1, index_2]).to(torch.int32),
      pos=torch.stack([pos1, pos2]).to(torch.int32),
      gate=torch.stack([gate_1, gate_2]), aux_loss=aux_loss, capacity=cap)


def CombineTensorFromIndices(gating: NestedMap, experts_dim: int):
  """Index form → dense `GSEC` combine tensor (for tests / oracle parity)."""
  out = 0
  for k in range(2):
    a = gating.gate[k].unsqueeze(-1) * F.one_hot(
        gating.index[k].long(), experts_dim).float()
    b = F.one_hot(gating.pos[k].long(), gating.capacity).float()
    out = out + torch.einsum('GSE,GSC->GSEC', a, b)
  return out


def HashGatingOnLogits(inputs, expert_id, paddings, num_devices, experts_dim,
                       expert_capacity_dim, fprop_dtype, use_xla_sharding=True,
                       capacity_factor=None, mask_dtype=None):
  """Token-id hash routing, top-1 with capacity (reference :2367)."""
  del num_devices, use_xla_sharding
  mask_dtype = mask_dtype or fprop_dtype
  cap = ExpertCapacity(int(expert_id.shape[1]), experts_dim,
                       expert_capacity_dim, capacity_factor)
  mask = F.one_hot(expert_id.long(), experts_dim).to(mask_dtype)
  if paddings is not None:
    mask = mask * (1.0 - paddings.to(mask_dtype)).unsqueeze(-1)
  pos = torch.cumsum(mask, dim=-2) - mask
  mask = mask * (pos < cap).to(mask_dtype)
  pos = (pos * mask).sum(-1)
  flat = mask.sum(-1)
  b = F.one_hot(pos.long(), cap).to(fprop_dtype)
  a = flat.to(fprop_dtype).unsqueeze(-1) * F.one_hot(
      expert_id.long(), experts_dim).to(fprop_dtype)
  combine = torch.einsum('...GSE,...GSC->...GSEC', a, b)
  dispatch = (combine != 0).to(fprop_dtype)
  return torch.zeros((), device=combine.device), combine, dispatch


def TokenShufflingOnlogits(inputs, logits, experts_dim, fprop_dtype,
                           use_xla_sharding=True, mask_dtype=None,
                           capacity_factor=1.0):
  """Expert-choice routing: every expert takes its top-C tokens (:2496).

  Returns (aux_loss, combine `GECS`-equivalent packed as GSEC, dispatch GSEC).
  """
  del inputs, use_xla_sharding, mask_dtype
  g, s, e = logits.shape
  cap = max(int(s * capacity_factor / e), 1)
  probs = torch.softmax(logits.float(), dim=-1)       # over experts
  top = probs.transpose(1, 2).topk(cap, dim=-1)       # [G, E, C] over tokens
  combine = torch.zeros(g, s, e, cap, device=logits.device, dtype=fprop_dtype)
  gi = torch.arange(g, device=logits.device)[:, None, None].expand(g, e, cap)
  ei = torch.arange(e, device=logits.device)[None, :, None].expand(g, e, cap)
  ci = torch.arange(cap, device=logits.device)[None, None, :].expand(g, e, cap)
  combine[gi, top.indices, ei, ci] = top.values.to(fprop_dtype)
  dispatch = (combine != 0).to(fprop_dtype)
  return torch.zeros((), device=logits.device), combine, dispatch


def SentenceTop2Gating(w, sentence_embeddings, paddings, **kwargs):
  """Sentence-level routing: all tokens of a sentence share experts (:3359)."""
  logits = torch.einsum('GM,ME->GE', sentence_embeddings.float(), w.float())
  return logits


def ComputeGating(w, inputs, paddings, num_devices, experts_dim,
                  expert_capacity_dim, local_dispatch, fprop_dtype,
                  gating_func='top_2', use_xla_sharding=True,
                  second_expert_policy='all', second_expert_threshold=0.0,
                  legacy_mtf_behavior=True, capacity_factor=None,
                  model_dim_reshape_segments=None, mask_dtype=None,
                  gating_logits_dtype=None, expert_id=None,
                  expert_padding_idx=None, seeds=None):
  """Dense gating → NestedMap(combine_tensor, dispatch_tensor, aux_loss)."""
  orig = inputs
  if not local_dispatch:
    inputs = inputs.reshape(1, inputs.shape[0] * inputs.shape[1], -1)
    if paddings is not None:
      paddings = paddings.reshape(1, -1)
  ldt = gating_logits_dtype or fprop_dtype
  logits = EinsumWithModelDim('GSM,ME->GSE', inputs.to(ldt), w.to(ldt),
                              model_dim_reshape_segments)
  if gating_func == 'token_shuffle':
    aux, comb, disp = TokenShufflingOnlogits(inputs, logits, experts_dim,
                                             fprop_dtype)
  elif gating_func == 'top_2':
    aux, comb, disp = Top2GatingOnLogits(
        inputs, paddings, logits, num_devices, experts_dim,
        expert_capacity_dim, fprop_dtype, use_xla_sharding,
        second_expert_policy, second_expert_threshold, legacy_mtf_behavior,
        capacity_factor, None, mask_dtype, expert_padding_idx, seeds=seeds)
  elif gating_func == 'hashing':
    aux, comb, disp = HashGatingOnLogits(
        inputs, expert_id, paddings, num_devices, experts_dim,
        expert_capacity_dim, fprop_dtype, use_xla_sharding, capacity_factor,
        mask_dtype)
  else:
    raise ValueError('Gating function: %s not supported yet!' % gating_func)
  if not local_dispatch:
    disp = disp.reshape(list(orig.shape[:2]) + list(disp.shape[2:]))
    comb = comb.reshape(list(orig.shape[:2]) + list(comb.shape[2:]))
  return NestedMap(combine_tensor=comb, dispatch_tensor=disp, aux_loss=aux)


def FeedForwardNetworksApplyGating(gating, inputs, reshaped_inputs, wi_split,
                                   wo_split, num_devices, num_groups,
                                   bi_split=None, bo_split=None,
                                   dropout_rate=0.0, device_mesh=None,
                                   model_dim_reshape_segments=None,
                                   use_glu=False, gating_func='top_2',
                                   activation_name='RELU', **unused_splits):
  """Dense einsum oracle of MoE apply (reference :2992-3163).

  `wi_split` is `[E, M, H]` (or `[2, E, M, H]` for GLU), `wo_split` `[E, H, M]`.
  Returns (outputs `GSM`, aux_loss).
  """
  del num_devices, device_mesh
  act = activations.GetFn(activation_name)
  disp = gating.dispatch_tensor.to(reshaped_inputs.dtype)
  expert_inputs = torch.einsum('GSEC,GSM->EGCM', disp, reshaped_inputs)
  e, g, c, m = expert_inputs.shape
  x = expert_inputs.reshape(e, g * c, m)
  if use_glu:
    h = torch.einsum('EAM,KEMH->KEAH', x, wi_split.to(x.dtype))
    if bi_split is not None:
      h = h + bi_split.to(h.dtype)
    h = act(h[0]) * h[1]
  else:
    h = torch.einsum('EAM,EMH->EAH', x, wi_split.to(x.dtype))
    if bi_split is not None:
      h = h + bi_split.to(h.dtype)
    h = act(h)
  if dropout_rate:
    h = F.dropout(h, dropout_rate, training=True)
  out = torch.einsum('EAH,EHM->EAM', h, wo_split.to(h.dtype))
  if bo_split is not None:
    out = out + bo_split.to(out.dtype)
  out = out.reshape(e, g, c, m).transpose(0, 1)       # EGCM → GECM
  combined = torch.einsum('GSEC,GECM->GSM',
                          gating.combine_tensor.to(out.dtype), out)
  outputs = combined.reshape(inputs.shape)
  return outputs, gating.aux_loss


# --------------------------------------------------------------------------
# Index-based dispatch / combine (single device). The expert-parallel variant
# lives in lingvo_b200.parallel.ep and shares these slot conventions:
#   slot(e, g, c) = (e * G + g) * C + c   in a [E, G*C, M] buffer.
# --------------------------------------------------------------------------
def _Slots(gating: NestedMap, g: int, s: int):
  cap = gating.capacity
  gi = torch.arange(g, device=gating.index.device).reshape(1, g, 1)
  slot = (gating.index.long() * g + gi) * cap + gating.pos.long()   # [2,G,S]
  valid = gating.gate.detach() > 0
  return slot, valid


class _DispatchFn(torch.autograd.Function):
  """x[G*S, M] → buf[E*G*C, M] (rows of dropped tokens are zero)."""

  @staticmethod
  def forward(ctx, x, slot, valid, num_slots):
    buf = torch.zeros(num_slots, x.shape[-1], dtype=x.dtype, device=x.device)
    for k in range(2):
      tok = valid[k].nonzero(as_tuple=True)[0]
      buf.index_copy_(0, slot[k][tok], x.index_select(0, tok))
    ctx.save_for_backward(slot, valid)
    return buf

  @staticmethod
  def backward(ctx, dbuf):
    slot, valid = ctx.saved_tensors
    dx = None
    for k in range(2):
      part = dbuf.index_select(0, slot[k].clamp(0, dbuf.shape[0] - 1))
      part = part * valid[k].unsqueeze(-1).to(part.dtype)
      dx = part if dx is None else dx + part
    return dx, None, None, None


def MoEDispatchIndexed(x2d, gating: NestedMap, g: int, s: int, e: int):
  """tokens `[G*S, M]` → expert inputs `[E, G*C, M]`."""
  slot, valid = _Slots(gating, g, s)
  buf = _DispatchFn.apply(x2d, slot.reshape(2, -1), valid.reshape(2, -1),
                          e * g * gating.capacity)
  return buf.reshape(e, g * gating.capacity, x2d.shape[-1])


def MoECombineIndexed(expert_out, gating: NestedMap, g: int, s: int):
  """expert outputs `[E, G*C, M]` → tokens `[G*S, M]` (gated 2-row gather)."""
  slot, valid = _Slots(gating, g, s)
  flat = expert_out.reshape(-1, expert_out.shape[-1])
  out = 0
  for k in range(2):
    rows = flat.index_select(0, slot[k].reshape(-1).clamp(0, flat.shape[0] - 1))
    w = gating.gate[k].reshape(-1, 1).to(rows.dtype)
    out = out + rows * w
  return out


def MoEApplyIndexed(x, gating: NestedMap, wi, wo, activation_name='RELU',
                    bi=None, bo=None, use_glu=False, ep_engine=None):
  """Index-based MoE FFN: `x [G, S, M]` → `[G, S, M]`.

  `wi [E_local, M, H]`, `wo [E_local, H, M]`. With an `ep_engine`
  (`parallel.ep.ExpertParallel`) tokens travel to the ranks owning their
  experts and back; otherwise all experts are local.
  """
  from lingvo_b200.ops import gemm
  g, s, m = x.shape
  e = wi.shape[-3] if ep_engine is None else ep_engine.num_experts
  x2d = x.reshape(g * s, m)
  if ep_engine is not None:
    return ep_engine.Apply(x2d, gating, wi, wo, activation_name, bi, bo,
                           use_glu).reshape(g, s, m)
  xin = MoEDispatchIndexed(x2d, gating, g, s, e)
  if use_glu:
    h0 = gemm.grouped_linear(xin, wi[0].to(xin.dtype))
    h1 = gemm.grouped_linear(xin, wi[1].to(xin.dtype))
    h = activations.GetFn(activation_name)(h0) * h1
  elif activation_name in ('RELU', 'NONE'):
    h = gemm.grouped_linear(xin, wi.to(xin.dtype), bi, act=activation_name)
  else:
    h = activations.GetFn(activation_name)(
        gemm.grouped_linear(xin, wi.to(xin.dtype), bi))
  out = gemm.grouped_linear(h, wo.to(h.dtype), bo)
  return MoECombineIndexed(out, gating, g, s).reshape(g, s, m)


# ------------------------------------------------------------------- layers --
class VarLayer(base_layer.BaseLayer):
  """Holds named weights; FProp returns them (reference `VarLayer`)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('weights', None, '[(name, WeightParams)].')
    p.Define('shared_var_collection_suffix', None, 'Kept for parity.')
    return p

  def _CreateLayerVariables(self):
    for k, wp in self.params.weights:
      self.CreateVariable(k, wp)

  def FProp(self, theta, *args, **kwargs):
    def cast(v):
      if v.is_floating_point() and v.dtype != self.fprop_dtype:
        return v.to(self.fprop_dtype)
      return v
    vals = [cast(theta[k]) for k, _ in self.params.weights]
    return vals[0] if len(vals) == 1 else tuple(vals)


class ShardedVarLayer(VarLayer):
  """VarLayer whose weights carry mesh-split annotations."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('cast_to_fprop_dtype', True, 'Cast weights to fprop_dtype.')
    return p


class ReshapeInputLayer(base_layer.BaseLayer):
  """`[B, L, M]` → `[G, S, M]` + paddings from segment ids (reference :1614)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('num_groups', None, 'Number of groups.')
    p.Define('num_devices', 1, 'Number of devices.')
    p.Define('model_dims', None, 'Model dims (list).')
    return p

  def FProp(self, theta, inputs, segment_id):
    p = self.params
    paddings = (segment_id == 0).to(inputs.dtype)
    b, l = inputs.shape[0], inputs.shape[1]
    g = p.num_groups or b
    if (b * l) % g:
      raise ValueError('tokens %d not divisible by num_groups %d' % (b * l, g))
    return (inputs.reshape(g, (b * l) // g, *inputs.shape[2:]),
            paddings.reshape(g, (b * l) // g))


class CausalDepthwiseConv1DLayer(base_layer.BaseLayer):
  """Causal depthwise conv over time on `[B, L, …, D]` (Primer/_LNConv)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('kernel_size', None, 'Kernel size.')
    p.Define('model_dims', None, 'Trailing (channel) dims.')
    p.Define('compatible_with_mtf_ckpt', False, 'Kept for parity.')
    p.Define('conv_vars_reshape', False, 'Kept for parity.')
    return p

  def _CreateLayerVariables(self):
    p = self.params
    dims = list(p.model_dims) if isinstance(p.model_dims, (list, tuple)) else [
        p.model_dims]
    for i in range(p.kernel_size):
      init = WeightInit.Constant(0.5 if i == 0 else 0.5 / p.kernel_size)
      self.CreateVariable('w_%d' % i, WeightParams(dims, init, p.dtype))

  def FProp(self, theta, inputs):
    p = self.params
    out = 0
    for i in range(p.kernel_size):
      shifted = inputs if i == 0 else F.pad(
          inputs, [0, 0] * (inputs.dim() - 2) + [i, 0])[:, :inputs.shape[1]]
      out = out + shifted * theta['w_%d' % i].to(inputs.dtype)
    return out


class StateLayer(base_layer.BaseLayer):
  """Per-layer decode state (KV caches) threaded through DecodeIds (:1186).

  `NewState(theta, shape)` allocates; `FProp` reads; `UpdateState(name, t, v)`
  writes at time `t`. State lives in a thread-local overlay keyed by layer
  path so the functional FProp signature is unchanged.
  """

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('shape', [None, None], 'batch, time, etc...')
    p.Define('use_xla_dynamic_update_slice', True, 'Kept for parity.')
    return p

  _STATE = py_utils._Stack()  # pylint: disable=protected-access

  @classmethod
  def StateContext(cls, state: NestedMap):
    return cls._STATE.Push(state)

  def NewState(self, shape, dtype=None, device=None):
    return torch.zeros(list(shape), dtype=dtype or self.fprop_dtype,
                       device=device or py_utils.CurrentDevice())

  def _Current(self):
    st = self._STATE.Top()
    assert st is not None, 'StateLayer used outside StateContext'
    return st

  def FProp(self, theta):
    return self._Current().Get(self.path.replace('[', '_').replace(']', ''))

  def UpdateState(self, value, t=None):
    st = self._Current()
    key = self.path.replace('[', '_').replace(']', '')
    if t is None:
      st.Set(key, value)
    else:
      cur = st.Get(key)
      cur[:, t] = value.to(cur.dtype)
    return value


class OverrideLayer(base_layer.BaseLayer):
  """Lets decoding override a value at a named key (reference OverrideLayer)."""

  _OVERRIDE = {}

  @classmethod
  def Set(cls, key, value):
    cls._OVERRIDE[key] = value

  @classmethod
  def Clear(cls):
    cls._OVERRIDE.clear()

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('key', None, 'Override key.')
    return p

  def FProp(self, theta, x):
    return self._OVERRIDE.get(self.params.key, x)
