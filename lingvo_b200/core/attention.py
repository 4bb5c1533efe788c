"""Time-major seq2seq attention layers (ref `lingvo/core/attention.py`).

Contract shared by all layers (ref :194-540):
  packed = InitForSourcePacked(theta, source_vecs [T,B,Ds], source_contexts
           [T,B,Dc], source_padding [T,B], source_segment_id=None)
  ctx [Bq,Dc'], probs [Bq,T], state = ComputeContextVectorWithSource(
           theta, packed, query_vec [Bq,Dq], attention_state=None,
           per_step_source_padding=None, query_segment_id=None)
where Bq is a multiple of B (beam search tiles the query batch).

Design: sources are packed ONCE into batch-major `[B, T, …]` tensors with the
projection already applied, so the per-decode-step work is one small GEMM plus
a fused masked softmax; everything is plain tensor code that runs under CUDA
graphs in the decoder loop.
"""

from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from lingvo_b200.core import base_layer
from lingvo_b200.core import py_utils
from lingvo_b200.core import quant_utils
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.core.py_utils import WeightInit
from lingvo_b200.core.py_utils import WeightParams

_NEG = -0.7 * torch.finfo(torch.float32).max


def SafeCumprod(x, dim=-1, exclusive=True):
  """exp(cumsum(log(clip(x)))) — numerically safe cumulative product (:67)."""
  tiny = torch.finfo(x.dtype).tiny
  logs = torch.log(x.clamp(tiny, 1.0))
  c = torch.cumsum(logs, dim)
  if exclusive:
    c = c - logs
  return torch.exp(c)


def MonotonicAttentionProb(p_choose_i, previous_attention, mode):
  """Expected monotonic alignment (Raffel et al. 2017) (:94).

  mode: 'recursive' | 'parallel' | 'hard'. Shapes [B, T].
  """
  if mode == 'hard':
    # previous_attention is one-hot; stay until p_choose says move on.
    p = p_choose_i * torch.cumsum(previous_attention, 1)
    return p * SafeCumprod(1 - p, 1, exclusive=True)
  if mode == 'parallel':
    cp = SafeCumprod(1 - p_choose_i, 1, exclusive=True)
    return p_choose_i * cp * torch.cumsum(
        previous_attention / cp.clamp(1e-10, 1.0), 1)
  if mode == 'recursive':
    b, t = p_choose_i.shape
    out = []
    q = torch.zeros(b, device=p_choose_i.device, dtype=p_choose_i.dtype)
    for j in range(t):
      prev_q = q
      q = (1 - p_choose_i[:, j - 1]) * prev_q + previous_attention[:, j] if j > 0 \
          else previous_attention[:, 0]
      out.append(p_choose_i[:, j] * q)
    return torch.stack(out, 1)
  raise ValueError('unknown mode ' + mode)


def MergeSourcePaddingWithPerStepSourcePadding(source_padding, per_step_source_padding,
                                               tb):
  """[T, B] ∪ [Bq, T] → [T, Bq] (:2870)."""
  t, b = source_padding.shape
  mult = tb // b
  sp = source_padding.unsqueeze(1).expand(t, mult, b).reshape(t, tb) if mult > 1 \
      else source_padding
  if per_step_source_padding is None:
    return sp
  return torch.maximum(sp.float(), per_step_source_padding.t().float())


class BaseAttentionLayer(quant_utils.QuantizableLayer):
  """Common params + the padded softmax / packing protocol (:194)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('atten_dropout_prob', 0.0, 'Dropout on attention weights.')
    p.Define('atten_dropout_deterministic', False, 'Kept for parity.')
    p.Define('packed_input', False, 'Packed (segmented) inputs.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self._source_init_done = False

  def InitForSourcePacked(self, theta, source_vecs, source_contexts,
                          source_padding, source_segment_id=None):
    self._packed_src = self.PackSource(theta, source_vecs, source_contexts,
                                       source_padding, source_segment_id)
    self._source_init_done = True
    return self._packed_src

  def PackSource(self, theta, source_vecs, source_contexts, source_padding,
                 source_segment_id=None):
    raise NotImplementedError

  def ComputeContextVectorWithSource(self, theta, packed_src, query_vec,
                                     attention_state=None,
                                     per_step_source_padding=None,
                                     query_segment_id=None):
    raise NotImplementedError

  def ComputeContextVector(self, theta, query_vec, attention_state=None,
                           per_step_source_padding=None, query_segment_id=None):
    assert self._source_init_done
    return self.ComputeContextVectorWithSource(
        theta, self._packed_src, query_vec, attention_state,
        per_step_source_padding, query_segment_id)

  def GetInitializationSourceState(self):
    assert self._source_init_done
    return self._packed_src

  def SetInitializationSourceState(self, new_init_state):
    self._source_init_done = True
    self._packed_src = new_init_state.DeepCopy() if hasattr(
        new_init_state, 'DeepCopy') else new_init_state

  def ZeroAttentionState(self, source_length, decoder_batch_size):
    return torch.zeros(decoder_batch_size, 0, device=self.Device())

  # -- helpers ------------------------------------------------------------------
  def _Mask(self, packed, bq, per_step_source_padding, query_segment_id):
    """Boolean [Bq, T]: True where the source position must be ignored."""
    pad = packed.source_padding                      # [B, T] batch-major
    b, t = pad.shape
    mult = bq // b
    mask = pad > 0
    if mult > 1:
      # beam search layout: query row i*B + b ↔ source row b (ref tiles on dim 0)
      mask = mask.unsqueeze(0).expand(mult, b, t).reshape(bq, t)
    if per_step_source_padding is not None:
      mask = mask | (per_step_source_padding.reshape(bq, t) > 0)
    if self.params.packed_input and packed.get('source_segment_id') is not None \
        and query_segment_id is not None:
      seg = packed.source_segment_id
      if mult > 1:
        seg = seg.unsqueeze(0).expand(mult, b, t).reshape(bq, t)
      mask = mask | (seg != query_segment_id.reshape(bq, 1))
    return mask

  def _PaddedSoftmax(self, logits, mask):
    """Softmax over T with masked positions removed; all-masked rows → 0 (:408)."""
    logits = logits.float().masked_fill(mask, _NEG)
    probs = torch.softmax(logits, -1)
    return probs * (~mask).to(probs.dtype)

  def _Dropout(self, probs):
    p = self.params
    if p.atten_dropout_prob > 0 and not self.do_eval:
      return F.dropout(probs, p.atten_dropout_prob, training=True)
    return probs

  @staticmethod
  def _TileSource(x, bq):
    b = x.shape[0]
    mult = bq // b
    if mult == 1:
      return x
    return x.unsqueeze(0).expand(mult, *x.shape).reshape(bq, *x.shape[1:])


class AdditiveAttention(BaseAttentionLayer):
  """Bahdanau attention: v·tanh(W·src + U·q) (:547)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('source_dim', 0, 'Source dim.')
    p.Define('query_dim', 0, 'Query dim.')
    p.Define('hidden_dim', 0, 'Hidden dim.')
    p.Define('same_batch_size', False, 'Source and query share the batch.')
    p.params_init = WeightInit.GaussianSqrtDim()
    return p

  def _CreateLayerVariables(self):
    p = self.params
    coll = ['AdditiveAttention_vars']
    self.CreateVariable('source_var', WeightParams(
        [p.source_dim, p.hidden_dim], p.params_init, p.dtype, coll))
    self.CreateVariable('query_var', WeightParams(
        [p.query_dim, p.hidden_dim], p.params_init, p.dtype, coll))
    self.CreateVariable('hidden_var', WeightParams(
        [p.hidden_dim], p.params_init, p.dtype, coll))

  def PackSource(self, theta, source_vecs, source_contexts, source_padding,
                 source_segment_id=None):
    src = torch.matmul(source_vecs.transpose(0, 1), theta.source_var.to(source_vecs.dtype))
    return NestedMap(
        source_vecs=src,                                  # [B, T, H] projected
        source_contexts=source_contexts.transpose(0, 1),  # [B, T, C]
        source_padding=source_padding.transpose(0, 1),
        source_segment_id=None if source_segment_id is None
        else source_segment_id.transpose(0, 1))

  def ZeroAttentionState(self, source_length, decoder_batch_size):
    return torch.zeros(decoder_batch_size, 1, device=self.Device())

  def ComputeContextVectorWithSource(self, theta, packed_src, query_vec,
                                     attention_state=None,
                                     per_step_source_padding=None,
                                     query_segment_id=None):
    bq = query_vec.shape[0]
    q = torch.matmul(query_vec, theta.query_var.to(query_vec.dtype))   # [Bq, H]
    src = self._TileSource(packed_src.source_vecs, bq)
    hid = torch.tanh(src + q.unsqueeze(1))
    logits = torch.matmul(hid, theta.hidden_var.to(hid.dtype))          # [Bq, T]
    mask = self._Mask(packed_src, bq, per_step_source_padding, query_segment_id)
    probs = self._PaddedSoftmax(logits, mask)
    pd = self._Dropout(probs).to(packed_src.source_contexts.dtype)
    ctx = torch.bmm(pd.unsqueeze(1),
                    self._TileSource(packed_src.source_contexts, bq)).squeeze(1)
    return ctx, probs, attention_state


class DotProductAttention(BaseAttentionLayer):
  """Scaled dot-product attention with optional learned per-dim scale (:1015)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('source_dim', 0, 'Source dim.')
    p.Define('query_dim', 0, 'Query dim.')
    p.Define('hidden_dim', 0, 'Hidden dim.')
    p.Define('use_dim_scale', True, 'Learned per-dim scale.')
    p.Define('atten_logit_cap', None, 'tanh cap on logits.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    assert p.source_dim == p.query_dim == p.hidden_dim, (
        'DotProductAttention needs source_dim == query_dim == hidden_dim')

  def _CreateLayerVariables(self):
    p = self.params
    if p.use_dim_scale:
      self.CreateVariable('per_dim_scale', WeightParams(
          [p.hidden_dim], WeightInit.Constant(0.0), p.dtype,
          ['DotProductAttention_vars']))

  def PackSource(self, theta, source_vecs, source_contexts, source_padding,
                 source_segment_id=None):
    return NestedMap(
        source_vecs=source_vecs.transpose(0, 1),
        source_contexts=source_contexts.transpose(0, 1),
        source_padding=source_padding.transpose(0, 1),
        source_segment_id=None if source_segment_id is None
        else source_segment_id.transpose(0, 1))

  def ZeroAttentionState(self, source_length, decoder_batch_size):
    return torch.zeros(decoder_batch_size, 1, device=self.Device())

  def _ScaleQuery(self, theta, q):
    p = self.params
    scale = 1.0 / math.sqrt(p.hidden_dim)
    if p.use_dim_scale:
      return q * (scale * 1.442695041 * F.softplus(theta.per_dim_scale.float())).to(q.dtype)
    return q * scale

  def ComputeContextVectorWithSource(self, theta, packed_src, query_vec,
                                     attention_state=None,
                                     per_step_source_padding=None,
                                     query_segment_id=None):
    p = self.params
    bq = query_vec.shape[0]
    q = self._ScaleQuery(theta, query_vec)
    src = self._TileSource(packed_src.source_vecs, bq)
    logits = torch.bmm(src, q.unsqueeze(-1)).squeeze(-1)               # [Bq, T]
    if p.atten_logit_cap:
      logits = p.atten_logit_cap * torch.tanh(logits / p.atten_logit_cap)
    mask = self._Mask(packed_src, bq, per_step_source_padding, query_segment_id)
    probs = self._PaddedSoftmax(logits, mask)
    pd = self._Dropout(probs).to(packed_src.source_contexts.dtype)
    ctx = torch.bmm(pd.unsqueeze(1),
                    self._TileSource(packed_src.source_contexts, bq)).squeeze(1)
    return ctx, probs, attention_state


class MultiHeadedAttention(BaseAttentionLayer):
  """Projects source/query/context into N heads and runs `inner_atten_params`
  per head (:1425). Variable names follow the reference (`source_proj`,
  `query_proj`, `ctx_proj`, `ctx_post_proj` + `_b`)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('source_dim', 0, 'Source dim.')
    p.Define('query_dim', 0, 'Query dim.')
    p.Define('context_dim', 0, 'Context dim.')
    p.Define('hidden_dim', 0, 'Hidden dim (all heads).')
    p.Define('num_attention_heads', 2, 'Heads.')
    p.Define('use_source_vec_as_attention_value', True,
             'Context = projected source vectors.')
    p.Define('enable_source_proj', True, 'Project sources.')
    p.Define('enable_query_proj', True, 'Project queries.')
    p.Define('inner_atten_params', DotProductAttention.Params(), 'Per-head attention.')
    p.Define('enable_ctx_pre_proj', False, 'Project contexts before attention.')
    p.Define('enable_ctx_post_proj', False, 'Project the attended context.')
    p.Define('ctx_post_proj_dim', 0, 'Output dim of the post projection.')
    p.Define('num_post_proj', 1, 'Kept for parity.')
    p.Define('proj_init', 'default', 'default|uniform|gaussian init of projections.')
    p.Define('attention_head_prob_index', -1, 'Return this head\'s probs (-1: mean).')
    p.Define('use_bias', True, 'Projection biases.')
    p.Define('enable_per_dim_scale', True, 'Per-dim scale in the inner attention.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    assert p.hidden_dim % p.num_attention_heads == 0
    h = p.hidden_dim // p.num_attention_heads
    inner = p.inner_atten_params.Copy().Set(
        source_dim=h, query_dim=h, hidden_dim=h, dtype=p.dtype,
        atten_dropout_prob=p.atten_dropout_prob, packed_input=p.packed_input)
    if 'use_dim_scale' in inner:
      inner.use_dim_scale = p.enable_per_dim_scale
    self.CreateChild('atten', inner)
    if p.use_source_vec_as_attention_value:
      assert not p.enable_ctx_pre_proj

  def _Init(self, dim):
    p = self.params
    if p.proj_init == 'uniform':
      return WeightInit.Uniform(math.sqrt(6.0 / (dim + p.hidden_dim)))
    if p.proj_init == 'gaussian':
      return WeightInit.Gaussian(math.sqrt(2.0 / (dim + p.hidden_dim)))
    return p.params_init

  def _CreateLayerVariables(self):
    p = self.params
    coll = ['MultiHeadedAttention_vars']
    zero = WeightInit.Constant(0.0)

    def _Proj(name, idim, odim):
      self.CreateVariable(name, WeightParams([idim, odim], self._Init(idim), p.dtype, coll))
      if p.use_bias:
        self.CreateVariable(name + '_b', WeightParams([odim], zero, p.dtype, coll))

    if p.enable_source_proj:
      _Proj('source_proj', p.source_dim, p.hidden_dim)
    else:
      assert p.source_dim == p.hidden_dim
    if p.enable_query_proj:
      _Proj('query_proj', p.query_dim, p.hidden_dim)
    else:
      assert p.query_dim == p.hidden_dim
    if p.enable_ctx_pre_proj and not p.use_source_vec_as_attention_value:
      _Proj('ctx_proj', p.context_dim, p.hidden_dim)
    if p.enable_ctx_post_proj:
      _Proj('ctx_post_proj', p.hidden_dim, p.ctx_post_proj_dim)

  @classmethod
  def SetOutputContextDim(cls, p, out_dim):
    p.ctx_post_proj_dim = out_dim

  def _Apply(self, theta, name, x):
    y = torch.matmul(x, theta[name].to(x.dtype))
    if self.params.use_bias:
      y = y + theta[name + '_b'].to(y.dtype)
    return y

  def PackSource(self, theta, source_vecs, source_contexts, source_padding,
                 source_segment_id=None):
    p = self.params
    n = p.num_attention_heads
    t, b = source_vecs.shape[:2]
    src = source_vecs.transpose(0, 1)                       # [B, T, D]
    if p.enable_source_proj:
      src = self._Apply(theta, 'source_proj', src)
    if p.use_source_vec_as_attention_value:
      ctx = src
    else:
      ctx = source_contexts.transpose(0, 1)
      if p.enable_ctx_pre_proj:
        ctx = self._Apply(theta, 'ctx_proj', ctx)
    return NestedMap(
        source_vecs=src.reshape(b, t, n, -1),
        source_contexts=ctx.reshape(b, t, n, -1),
        source_padding=source_padding.transpose(0, 1),
        source_segment_id=None if source_segment_id is None
        else source_segment_id.transpose(0, 1))

  def ExtendSourcePacked(self, theta, new_source_vecs, new_source_contexts,
                         new_source_paddings, new_source_segment_ids,
                         cached_packed_src, t=None):
    """Appends one time step `[B, D]` to a packed source (self-attention decode)."""
    step = self.PackSource(
        theta, new_source_vecs.unsqueeze(0), new_source_contexts.unsqueeze(0),
        new_source_paddings.unsqueeze(0),
        None if new_source_segment_ids is None else new_source_segment_ids.unsqueeze(0))
    if t is None:
      cat = lambda a, b: b if a is None else torch.cat([a, b], 1)
      return NestedMap(
          source_vecs=cat(cached_packed_src.get('source_vecs'), step.source_vecs),
          source_contexts=cat(cached_packed_src.get('source_contexts'), step.source_contexts),
          source_padding=cat(cached_packed_src.get('source_padding'), step.source_padding),
          source_segment_id=None if step.source_segment_id is None else cat(
              cached_packed_src.get('source_segment_id'), step.source_segment_id))
    out = cached_packed_src.DeepCopy() if hasattr(cached_packed_src, 'DeepCopy') else cached_packed_src
    for k in ('source_vecs', 'source_contexts', 'source_padding', 'source_segment_id'):
      if step.get(k) is not None and out.get(k) is not None:
        v = out[k].clone()
        v[:, t] = step[k][:, 0]
        out[k] = v
    return out

  def ZeroAttentionState(self, source_length, decoder_batch_size):
    return self.atten.ZeroAttentionState(
        source_length, decoder_batch_size * self.params.num_attention_heads)

  def ComputeContextVectorWithSource(self, theta, packed_src, query_vec,
                                     attention_state=None,
                                     per_step_source_padding=None,
                                     query_segment_id=None,
                                     atten_idx=None):
    p = self.params
    n = p.num_attention_heads
    bq = query_vec.shape[0]
    q = self._Apply(theta, 'query_proj', query_vec) if p.enable_query_proj else query_vec
    q = q.reshape(bq, n, -1)
    inner = self.atten
    if isinstance(inner, DotProductAttention):
      q = inner._ScaleQuery(theta.atten, q)  # pylint: disable=protected-access
      src = self._TileSource(packed_src.source_vecs, bq)           # [Bq,T,N,H]
      logits = torch.einsum('BTNH,BNH->BNT', src.float(), q.float())
      if inner.params.atten_logit_cap:
        c = inner.params.atten_logit_cap
        logits = c * torch.tanh(logits / c)
    else:
      # Additive inner attention, vectorised over heads.
      src = self._TileSource(packed_src.source_vecs, bq)
      sp = torch.einsum('BTNH,HK->BTNK', src, theta.atten.source_var.to(src.dtype))
      qp = torch.einsum('BNH,HK->BNK', q, theta.atten.query_var.to(q.dtype))
      hid = torch.tanh(sp + qp.unsqueeze(1))
      logits = torch.einsum('BTNK,K->BNT', hid.float(), theta.atten.hidden_var.float())
    mask = self._Mask(packed_src, bq, per_step_source_padding, query_segment_id)
    probs = self._PaddedSoftmax(logits, mask.unsqueeze(1))           # [Bq,N,T]
    ctxs = self._TileSource(packed_src.source_contexts, bq)
    ctx = torch.einsum('BNT,BTNH->BNH', self._Dropout(probs).to(ctxs.dtype), ctxs)
    ctx = ctx.reshape(bq, -1)
    if p.enable_ctx_post_proj:
      ctx = self._Apply(theta, 'ctx_post_proj', ctx)
    if p.attention_head_prob_index >= 0:
      out_probs = probs[:, p.attention_head_prob_index]
    else:
      out_probs = probs.mean(1)
    return ctx, out_probs, attention_state

  def ComputeContextVectorWithAttenProbs(self, theta, packed_context, atten_probs):
    """Context from externally supplied probs `[Bq, N, T]` (:2215)."""
    p = self.params
    bq = atten_probs.shape[0]
    ctxs = self._TileSource(packed_context, bq)
    ctx = torch.einsum('BNT,BTNH->BNH', atten_probs.to(ctxs.dtype), ctxs).reshape(bq, -1)
    if p.enable_ctx_post_proj:
      ctx = self._Apply(theta, 'ctx_post_proj', ctx)
    return ctx

  def PackCachedSource(self, theta, source_vecs, source_contexts, source_padding,
                       source_segment_id=None):
    return self.PackSource(theta, source_vecs, source_contexts, source_padding,
                           source_segment_id)

  ComputeContextVectorWithCachedSource = ComputeContextVectorWithSource


class LocationSensitiveAttention(BaseAttentionLayer):
  """Additive attention + conv features of the previous alignment (:2334)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('source_dim', 0, 'Source dim.')
    p.Define('query_dim', 0, 'Query dim.')
    p.Define('hidden_dim', 0, 'Hidden dim.')
    p.Define('location_filter_size', 0, 'Odd conv filter width.')
    p.Define('location_num_filters', 0, 'Number of location filters.')
    p.Define('same_batch_size', False, 'Kept for parity.')
    p.Define('location_features', ['PREV_PROBS'], 'PREV_PROBS and/or CUMULATIVE_PROBS.')
    p.params_init = WeightInit.GaussianSqrtDim()
    return p

  def _CreateLayerVariables(self):
    p = self.params
    assert p.location_filter_size % 2 == 1
    coll = ['LocationSensitiveAttention_vars']
    self.CreateVariable('source_var', WeightParams(
        [p.source_dim, p.hidden_dim], p.params_init, p.dtype, coll))
    self.CreateVariable('query_var', WeightParams(
        [p.query_dim, p.hidden_dim], p.params_init, p.dtype, coll))
    self.CreateVariable('hidden_var', WeightParams(
        [p.hidden_dim], p.params_init, p.dtype, coll))
    self.CreateVariable('location_filter_var', WeightParams(
        [p.location_filter_size, len(p.location_features), p.location_num_filters],
        WeightInit.Uniform(0.05), p.dtype, coll))
    self.CreateVariable('location_var', WeightParams(
        [p.location_num_filters, p.hidden_dim], WeightInit.Uniform(0.05), p.dtype, coll))

  PackSource = AdditiveAttention.PackSource

  def ZeroAttentionState(self, source_length, decoder_batch_size):
    p = self.params
    st = torch.zeros(decoder_batch_size, len(p.location_features), source_length,
                     device=self.Device())
    st[:, :, 0] = 1.0
    return st

  def ComputeContextVectorWithSource(self, theta, packed_src, query_vec,
                                     attention_state=None,
                                     per_step_source_padding=None,
                                     query_segment_id=None):
    p = self.params
    bq = query_vec.shape[0]
    q = torch.matmul(query_vec, theta.query_var.to(query_vec.dtype))
    src = self._TileSource(packed_src.source_vecs, bq)                # [Bq,T,H]
    w = theta.location_filter_var.permute(2, 1, 0).to(attention_state.dtype)  # [F,C,K]
    loc = F.conv1d(attention_state, w, padding=p.location_filter_size // 2)   # [Bq,F,T]
    loc = torch.matmul(loc.transpose(1, 2), theta.location_var.to(loc.dtype))  # [Bq,T,H]
    hid = torch.tanh(src + q.unsqueeze(1) + loc.to(src.dtype))
    logits = torch.matmul(hid, theta.hidden_var.to(hid.dtype))
    mask = self._Mask(packed_src, bq, per_step_source_padding, query_segment_id)
    probs = self._PaddedSoftmax(logits, mask)
    feats = []
    for i, f in enumerate(p.location_features):
      if f == 'PREV_PROBS':
        feats.append(probs)
      elif f == 'CUMULATIVE_PROBS':
        feats.append(attention_state[:, i] + probs)
      else:
        raise ValueError(f)
    new_state = torch.stack(feats, 1).to(attention_state.dtype)
    pd = self._Dropout(probs).to(packed_src.source_contexts.dtype)
    ctx = torch.bmm(pd.unsqueeze(1),
                    self._TileSource(packed_src.source_contexts, bq)).squeeze(1)
    return ctx, probs, new_state


class MonotonicAttention(BaseAttentionLayer):
  """Soft monotonic alignment with energy-function noise (:2900)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('source_dim', 0, 'Source dim.')
    p.Define('query_dim', 0, 'Query dim.')
    p.Define('hidden_dim', 0, 'Hidden dim.')
    p.Define('pre_sigmoid_noise', 0.0, 'Std of noise added before the sigmoid.')
    p.Define('hidden_bias_init', -1, 'Initial scalar bias of the energy.')
    p.params_init = WeightInit.GaussianSqrtDim()
    return p

  def _CreateLayerVariables(self):
    p = self.params
    coll = ['MonotonicAttention_vars']
    self.CreateVariable('source_var', WeightParams(
        [p.source_dim, p.hidden_dim], p.params_init, p.dtype, coll))
    self.CreateVariable('query_var', WeightParams(
        [p.query_dim, p.hidden_dim], p.params_init, p.dtype, coll))
    self.CreateVariable('energy_bias_var', WeightParams(
        [p.hidden_dim], WeightInit.Constant(0.0), p.dtype, coll))
    self.CreateVariable('hidden_var', WeightParams(
        [p.hidden_dim], p.params_init, p.dtype, coll))
    self.CreateVariable('hidden_scale_var', WeightParams(
        [], WeightInit.Constant(1.0 / math.sqrt(p.hidden_dim)), p.dtype, coll))
    self.CreateVariable('hidden_bias_var', WeightParams(
        [], WeightInit.Constant(float(p.hidden_bias_init)), p.dtype, coll))

  PackSource = AdditiveAttention.PackSource

  def ZeroAttentionState(self, source_length, decoder_batch_size):
    emit = torch.zeros(decoder_batch_size, source_length, device=self.Device())
    emit[:, 0] = 1.0
    return NestedMap(emit_probs=emit)

  def ComputeProbabilities(self, theta, src, mask, query_vec, previous_attention):
    p = self.params
    q = torch.matmul(query_vec, theta.query_var.to(query_vec.dtype))
    hid = torch.tanh(src + q.unsqueeze(1) + theta.energy_bias_var.to(src.dtype))
    v = theta.hidden_var.float()
    v = theta.hidden_scale_var.float() * v / v.norm().clamp_min(1e-12)
    logits = torch.matmul(hid.float(), v) + theta.hidden_bias_var.float()
    if p.pre_sigmoid_noise > 0 and not self.do_eval:
      logits = logits + p.pre_sigmoid_noise * torch.randn_like(logits)
    p_choose = torch.sigmoid(logits) * (~mask).float()
    if self.do_eval:
      p_choose = (p_choose > 0.5).float()
      return MonotonicAttentionProb(p_choose, previous_attention, 'hard')
    return MonotonicAttentionProb(p_choose, previous_attention, 'parallel')

  def ComputeContextVectorWithSource(self, theta, packed_src, query_vec,
                                     attention_state=None,
                                     per_step_source_padding=None,
                                     query_segment_id=None):
    bq = query_vec.shape[0]
    src = self._TileSource(packed_src.source_vecs, bq)
    mask = self._Mask(packed_src, bq, per_step_source_padding, query_segment_id)
    probs = self.ComputeProbabilities(theta, src, mask, query_vec,
                                      attention_state.emit_probs)
    ctxs = self._TileSource(packed_src.source_contexts, bq)
    ctx = torch.bmm(probs.to(ctxs.dtype).unsqueeze(1), ctxs).squeeze(1)
    return ctx, probs, NestedMap(emit_probs=probs)


class GmmMonotonicAttention(BaseAttentionLayer):
  """Graves-style GMM attention over encoder positions (:3267)."""

  @classmethod
  def Params(cls):
    from lingvo_b200.core import layers  # pylint: disable=g-import-not-at-top
    p = super().Params()
    p.Define('source_dim', 0, 'Source dim.')
    p.Define('query_dim', 0, 'Query dim.')
    p.Define('hidden_dim', 128, 'Hidden dim of the MLP.')
    p.Define('max_offset', -1, 'Max step size (-1: exp).')
    p.Define('num_mixtures', 5, 'Mixture components.')
    p.Define('normalize_probs', False, 'Renormalise over unpadded positions.')
    p.Define('gmm_mlp_tpl', layers.FeedForwardNet.Params(), 'MLP template.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    self.CreateChild('GMM', p.gmm_mlp_tpl.Copy().Set(
        input_dim=p.query_dim, hidden_layer_dims=[p.hidden_dim, p.num_mixtures * 3],
        activation=['SIGMOID', 'NONE']))

  def PackSource(self, theta, source_vecs, source_contexts, source_padding,
                 source_segment_id=None):
    return NestedMap(source_vecs=source_vecs.transpose(0, 1),
                     source_contexts=source_contexts.transpose(0, 1),
                     source_padding=source_padding.transpose(0, 1),
                     source_segment_id=None)

  def ZeroAttentionState(self, source_length, decoder_batch_size):
    p = self.params
    st = torch.zeros(decoder_batch_size, p.num_mixtures, 4, device=self.Device())
    st[:, :, 2] = 1.0   # variance
    return st

  def ComputeContextVectorWithSource(self, theta, packed_src, query_vec,
                                     attention_state=None,
                                     per_step_source_padding=None,
                                     query_segment_id=None):
    p = self.params
    bq = query_vec.shape[0]
    k = p.num_mixtures
    out = self.GMM.FProp(theta.GMM, query_vec).float().reshape(bq, k, 3)
    prior_l, offset_l, var_l = out.unbind(-1)
    prev_pos = attention_state[:, :, 0]
    step = (p.max_offset * torch.sigmoid(offset_l) if p.max_offset > 0
            else F.softplus(offset_l))
    pos = prev_pos + step
    var = F.softplus(var_l) + 1e-4
    prior = torch.softmax(prior_l, -1)
    t = packed_src.source_vecs.shape[1]
    enc = torch.arange(t, device=query_vec.device).float().view(1, 1, t)
    # Discretised Gaussian mass on [j-0.5, j+0.5).
    std = var.sqrt().unsqueeze(-1)
    cdf = lambda x: 0.5 * (1 + torch.erf((x - pos.unsqueeze(-1)) / (std * math.sqrt(2))))
    probs = (prior.unsqueeze(-1) * (cdf(enc + 0.5) - cdf(enc - 0.5))).sum(1)
    mask = self._Mask(packed_src, bq, per_step_source_padding, query_segment_id)
    probs = probs * (~mask).float()
    if p.normalize_probs:
      probs = probs / probs.sum(-1, keepdim=True).clamp_min(1e-12)
    ctxs = self._TileSource(packed_src.source_contexts, bq)
    ctx = torch.bmm(probs.to(ctxs.dtype).unsqueeze(1), ctxs).squeeze(1)
    new_state = torch.stack([pos, step, var, prior], -1)
    return ctx, probs, new_state


class MergerLayer(base_layer.BaseLayer):
  """Combines several context vectors: mean | sum | atten | concat | weighted_sum |
  gated_avg (:3608)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('merger_op', None, 'mean|atten|concat|sum|weighted_sum|gated_avg.')
    p.Define('attention_tpl', AdditiveAttention.Params(), 'For merger_op=atten.')
    p.Define('pre_proj_input_dims', None, 'Project every source to a common dim.')
    p.Define('pre_proj_output_dims', None, 'Output dims of the pre-projections.')
    p.Define('proj_tpl', None, 'Projection template (ProjectionLayer).')
    p.Define('source_dim', 0, 'Source dim (atten).')
    p.Define('query_dim', 0, 'Query dim (atten).')
    p.Define('hidden_dim', 0, 'Hidden dim (atten).')
    p.Define('num_sources', 0, 'Number of sources (weighted_sum, gated_avg).')
    p.Define('gated_avg_tpl', None, 'GatedAverageLayer params.')
    return p

  def __init__(self, params):
    from lingvo_b200.core import layers  # pylint: disable=g-import-not-at-top
    super().__init__(params)
    p = self.params
    assert p.merger_op in ('mean', 'atten', 'concat', 'sum', 'weighted_sum', 'gated_avg')
    if p.merger_op == 'atten':
      self.CreateChild('atten', p.attention_tpl.Copy().Set(
          source_dim=p.source_dim, query_dim=p.query_dim, hidden_dim=p.hidden_dim))
    if p.pre_proj_input_dims:
      tpl = p.proj_tpl or layers.ProjectionLayer.Params().Set(batch_norm=False)
      self.CreateChildren('pre_proj', [
          tpl.Copy().Set(name='pre_proj_%d' % i, input_dim=i_d, output_dim=o_d)
          for i, (i_d, o_d) in enumerate(zip(p.pre_proj_input_dims,
                                             p.pre_proj_output_dims))])
    if p.merger_op == 'gated_avg':
      tpl = p.gated_avg_tpl or layers.GatedAverageLayer.Params()
      self.CreateChild('gated_average', tpl.Copy().Set(
          num_nodes=p.source_dim, num_inputs=p.num_sources))

  def _CreateLayerVariables(self):
    p = self.params
    if p.merger_op == 'weighted_sum':
      self.CreateVariable('sum_weight', WeightParams(
          [p.num_sources], WeightInit.Constant(1.0 / max(p.num_sources, 1)), p.dtype))

  def FProp(self, theta, inputs, query_vec=None):
    p = self.params
    n = len(inputs)
    if p.pre_proj_input_dims:
      inputs = [self.pre_proj[i].FProp(theta.pre_proj[i], x) for i, x in enumerate(inputs)]
    if p.merger_op == 'mean':
      return sum(inputs) / n
    if p.merger_op == 'sum':
      return sum(inputs)
    if p.merger_op == 'concat':
      return torch.cat(inputs, -1)
    if p.merger_op == 'weighted_sum':
      w = torch.softmax(theta.sum_weight.float(), 0)
      return sum(w[i].to(x.dtype) * x for i, x in enumerate(inputs))
    if p.merger_op == 'gated_avg':
      return self.gated_average.FProp(theta.gated_average, inputs)
    # atten: sources stacked on the time axis.
    src = torch.stack(inputs, 0)                          # [n, B, D]
    pad = torch.zeros(n, src.shape[1], device=src.device)
    packed = self.atten.PackSource(theta.atten, src, src, pad)
    ctx, _, _ = self.atten.ComputeContextVectorWithSource(theta.atten, packed, query_vec)
    return ctx


class MultiSourceAttention(BaseAttentionLayer):
  """One attention per named source + a MergerLayer (:3856)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('source_atten_tpls', None, 'List of (source_key, attention params).')
    p.Define('source_dim', 0, 'Default source dim.')
    p.Define('query_dim', 0, 'Query dim.')
    p.Define('primary_source_key', 'source_0', 'Probs/state come from this source.')
    p.Define('atten_merger_tpl', MergerLayer.Params().Set(merger_op='sum'), 'Merger.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    self._keys = [k for k, _ in p.source_atten_tpls]
    for k, tpl in p.source_atten_tpls:
      t = tpl.Copy()
      if 'query_dim' in t and not t.query_dim:
        t.query_dim = p.query_dim
      if 'source_dim' in t and not t.source_dim:
        t.source_dim = p.source_dim
      self.CreateChild('atten_%s' % k, t)
    self.CreateChild('atten_merger', p.atten_merger_tpl)

  def PackSource(self, theta, source_vecs, source_contexts, source_padding,
                 source_segment_id=None):
    return NestedMap({
        k: getattr(self, 'atten_%s' % k).PackSource(
            theta['atten_%s' % k], source_vecs[k], source_contexts[k],
            source_padding[k], source_segment_id[k] if source_segment_id else None)
        for k in self._keys})

  def ZeroAttentionState(self, source_seq_length, decoder_batch_size):
    return NestedMap({
        k: getattr(self, 'atten_%s' % k).ZeroAttentionState(
            source_seq_length[k], decoder_batch_size) for k in self._keys})

  def ComputeContextVectorWithSource(self, theta, packed_src, query_vec,
                                     attention_state=None,
                                     per_step_source_padding=None,
                                     query_segment_id=None):
    p = self.params
    ctxs, probs, states = [], None, NestedMap()
    for k in self._keys:
      c, pr, st = getattr(self, 'atten_%s' % k).ComputeContextVectorWithSource(
          theta['atten_%s' % k], packed_src[k], query_vec,
          attention_state[k] if attention_state is not None else None,
          per_step_source_padding, query_segment_id)
      ctxs.append(c)
      states[k] = st
      if k == p.primary_source_key:
        probs = pr
    return self.atten_merger.FProp(theta.atten_merger, ctxs, query_vec), probs, states
