"""Time-major Transformer building blocks (ref `lingvo/core/layers_with_attention.py`).

`TransformerAttentionLayer` (:85), `TransformerFeedForwardLayer` (:529),
`MoEFeedforwardLayer` (:740), `HybridFeedforwardLayer` (:788),
`TransformerLayer` (:1334), `TransformerLayerWithMultitaskAdapters` (:2192),
`SelfAttentiveLayer` (:2984), `StochasticResidualLayer` (:32).

Inputs are `[time, batch, dim]`. The self-attention of a whole sequence is
computed batch-major through the fused kernel (`ops.attention`) using the same
projection variables (`source_proj`, `query_proj`, `ctx_post_proj`) the
step-wise `attention.MultiHeadedAttention` owns, so FProp is O(1) kernel
launches in T while ExtendStep stays incremental.
"""

from __future__ import annotations

import torch
import torch.nn.functional as F

from lingvo_b200.core import attention
from lingvo_b200.core import base_layer
from lingvo_b200.core import layers
from lingvo_b200.core import py_utils
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.core.py_utils import WeightInit
from lingvo_b200.core.py_utils import WeightParams
from lingvo_b200.ops import attention as attention_ops

_NEG = -0.7 * torch.finfo(torch.float32).max


class StochasticResidualLayer(base_layer.BaseLayer):
  """x + drop_path(f(x)); in eval f(x) is scaled by the survival prob (:32)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('residual_weight', 1.0, 'Residual weight.')
    p.Define('survival_prob', 1.0, 'Probability of keeping the residual branch.')
    return p

  def FProp(self, theta, x, y):
    p = self.params
    if self.do_eval:
      return x + p.residual_weight * p.survival_prob * y
    keep = (torch.rand((), device=y.device) < p.survival_prob).to(y.dtype)
    return x + p.residual_weight * keep * y


class TransformerAttentionLayer(base_layer.BaseLayer):
  """LN → multi-headed attention → dropout → residual, time-major (:85)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('source_dim', 0, 'Query/source dim.')
    p.Define('context_dim', 0, 'Aux (cross-attention) source dim.')
    p.Define('atten_hidden_dim', 0, 'Attention hidden dim.')
    p.Define('num_attention_heads', 8, 'Heads.')
    p.Define('is_masked', False, 'Masked self-attention.')
    p.Define('mask_ngram_order', 0, 'For mask_type=ngram.')
    p.Define('mask_type', 'future', 'future | eye | ngram.')
    p.Define('ln_tpl', layers.LayerNorm.Params(), 'LN template.')
    p.Define('atten_tpl', attention.MultiHeadedAttention.Params().Set(
        use_source_vec_as_attention_value=False, enable_ctx_post_proj=True,
        enable_ctx_pre_proj=True), 'Attention template.')
    p.Define('atten_dropout_prob', 0.0, 'Attention dropout.')
    p.Define('residual_dropout_prob', 0.0, 'Residual dropout.')
    p.Define('residual_dropout_tpl', layers.DropoutLayer.Params(), 'Dropout tpl.')
    p.Define('packed_input', False, 'Packed input.')
    p.Define('add_unnormalized_input', False, 'Residual on the raw input.')
    p.Define('residual_function', None, 'Gated residual layer params.')
    p.Define('pre_layer_norm', True, 'Pre-LN.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    assert p.source_dim
    hid = p.atten_hidden_dim or p.source_dim
    ctx = p.context_dim or p.source_dim
    self.CreateChild('atten', p.atten_tpl.Copy().Set(
        source_dim=ctx, query_dim=p.source_dim, context_dim=ctx, hidden_dim=hid,
        ctx_post_proj_dim=p.source_dim, num_attention_heads=p.num_attention_heads,
        atten_dropout_prob=p.atten_dropout_prob, packed_input=p.packed_input))
    self.CreateChild('layer_norm', p.ln_tpl.Copy().Set(input_dim=p.source_dim))
    self.CreateChild('residual_dropout', p.residual_dropout_tpl.Copy().Set(
        keep_prob=1.0 - p.residual_dropout_prob))
    if p.residual_function is not None:
      self.CreateChild('residual_function', p.residual_function.Copy().Set(
          input_dim=p.source_dim))

  def _StepMask(self, t, device):
    """[T, T] True where query i may NOT see key j."""
    p = self.params
    i = torch.arange(t, device=device).unsqueeze(1)
    j = torch.arange(t, device=device).unsqueeze(0)
    if p.mask_type == 'future':
      return j > i
    if p.mask_type == 'eye':
      return j == i
    if p.mask_type == 'ngram':
      return (j > i) | (j <= i - p.mask_ngram_order)
    raise ValueError(p.mask_type)

  def _Residual(self, theta, base, ctx):
    p = self.params
    ctx = self.residual_dropout.FProp(theta.residual_dropout, ctx)
    if p.residual_function is None:
      return base + ctx
    return self.residual_function.FProp(theta.residual_function, base, ctx)

  def FProp(self, theta, query_vec, source_paddings, source_vecs=None,
            query_segment_id=None, source_segment_id=None, context_vecs=None,
            **kwargs):
    """query_vec [T, B, D]; source_vecs [S, B, D] (None ⇒ self-attention).
    Returns (output [T, B, D], probs [T, B, S])."""
    p = self.params
    a, th = self.atten, theta.atten
    unnormalized = query_vec
    q_in = self.layer_norm.FProp(theta.layer_norm, query_vec) if p.pre_layer_norm \
        else query_vec
    if source_vecs is None:
      source_vecs, source_segment_id = q_in, query_segment_id
    if context_vecs is None:
      context_vecs = source_vecs
    t, b = q_in.shape[:2]
    s = source_vecs.shape[0]
    n = p.num_attention_heads
    packed = a.PackSource(th, source_vecs, context_vecs, source_paddings,
                          source_segment_id)
    qp = a._Apply(th, 'query_proj', q_in.transpose(0, 1)) if a.params.enable_query_proj \
        else q_in.transpose(0, 1)                    # pylint: disable=protected-access
    qh = qp.reshape(b, t, n, -1)
    if isinstance(a.atten, attention.DotProductAttention):
      qh = a.atten._ScaleQuery(th.atten, qh)        # pylint: disable=protected-access
    mask = (packed.source_padding > 0).view(b, 1, 1, s)
    if p.is_masked:
      mask = mask | self._StepMask(t, q_in.device).view(1, 1, t, s)
    if p.packed_input and query_segment_id is not None:
      mask = mask | (query_segment_id.t().view(b, 1, t, 1) !=
                     packed.source_segment_id.view(b, 1, 1, s))
    bias = mask.float() * _NEG
    drop = 0.0 if self.do_eval else p.atten_dropout_prob
    ctx, probs = attention_ops.dot_product_attention(
        qh, packed.source_vecs, packed.source_contexts, bias, 1.0,
        dropout_prob=drop, return_probs=True)
    ctx = ctx.reshape(b, t, -1)
    if a.params.enable_ctx_post_proj:
      ctx = a._Apply(th, 'ctx_post_proj', ctx)       # pylint: disable=protected-access
    ctx = ctx.transpose(0, 1)
    base = unnormalized if p.add_unnormalized_input else q_in
    out = self._Residual(theta, base, ctx)
    if not p.pre_layer_norm:
      out = self.layer_norm.FProp(theta.layer_norm, out)
    return out, probs.mean(1).transpose(0, 1)       # [T, B, S]

  def ExtendStep(self, theta, query_vec, prefix_state, t=None):
    """query_vec [B, D]; prefix_state {key,value: [t, B, D]} grows by one."""
    p = self.params
    assert p.is_masked
    a, th = self.atten, theta.atten
    unnormalized = query_vec
    q_in = self.layer_norm.FProp(theta.layer_norm, query_vec) if p.pre_layer_norm \
        else query_vec
    if t is None:
      key = torch.cat([prefix_state.key, q_in.unsqueeze(0)], 0)
      value = torch.cat([prefix_state.value, q_in.unsqueeze(0)], 0)
      pad = torch.zeros(key.shape[0], key.shape[1], device=q_in.device)
    else:
      key = prefix_state.key.clone()
      value = prefix_state.value.clone()
      key[t] = q_in
      value[t] = q_in
      pad = (torch.arange(key.shape[0], device=q_in.device) > t).float().unsqueeze(1) \
          .expand(key.shape[0], key.shape[1])
    packed = a.PackSource(th, key, value, pad)
    ctx, probs, _ = a.ComputeContextVectorWithSource(th, packed, q_in)
    base = unnormalized if p.add_unnormalized_input else q_in
    out = self._Residual(theta, base, ctx)
    if not p.pre_layer_norm:
      out = self.layer_norm.FProp(theta.layer_norm, out)
    return out, probs, NestedMap(key=key, value=value)


class TransformerMultiSourceAttentionLayer(TransformerAttentionLayer):
  """Cross-attention over several named sources (:480)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('num_source', 0, 'Number of sources.')
    p.Define('primary_source_index', 0, 'Which source returns probs.')
    p.Define('multi_source_atten', attention.MultiSourceAttention.Params(), 'Wrapper.')
    return p

  def FProp(self, theta, query_vec, source_paddings, source_vecs=None, **kwargs):
    outs, probs = [], None
    for i in range(self.params.num_source):
      k = 'source_%d' % i
      o, pr = super().FProp(theta, query_vec, source_paddings[k], source_vecs[k])
      outs.append(o)
      if i == self.params.primary_source_index:
        probs = pr
    return sum(outs) / len(outs), probs


class TransformerFeedForwardLayer(base_layer.BaseLayer):
  """LN → FFN(hidden) → dropout → residual (:529). Works on [..., D] inputs."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('input_dim', 0, 'Input dim.')
    p.Define('output_dim', 0, 'Output dim (0 ⇒ input_dim).')
    p.Define('hidden_dim', 0, 'Hidden dim.')
    p.Define('num_tasks', 0, 'Kept for parity.')
    p.Define('ln_tpl', layers.LayerNorm.Params(), 'LN template.')
    p.Define('activation', 'RELU', 'Activation (GATED_* use a gated unit).')
    p.Define('residual_weight', 1.0, 'Weight of f(x) in the residual add.')
    p.Define('fflayer_tpl', layers.FeedForwardNet.Params().Set(
        activation=['RELU', 'NONE']), 'FFN template.')
    p.Define('res_proj_tpl', layers.ProjectionLayer.Params().Set(batch_norm=True),
             'Residual re-projection when dims differ.')
    p.Define('residual_dropout_prob', 0.0, 'Residual dropout.')
    p.Define('residual_dropout_tpl', layers.DropoutLayer.Params(), 'Dropout tpl.')
    p.Define('relu_dropout_prob', 0.0, 'Hidden dropout.')
    p.Define('add_skip_connection', True, 'Residual add.')
    p.Define('pre_layer_norm', True, 'Pre-LN.')
    p.Define('primer_hybrid_norm', False, 'Pre- and post-LN.')
    p.Define('residual_droppath_prob', 0.0, 'Stochastic depth.')
    p.Define('use_block_diagonal_matmul_pl', False, 'Kept for parity.')
    p.Define('num_blocks_pl', 1, 'Kept for parity.')
    p.Define('memory_augmentation', False, 'Kept for parity.')
    p.Define('memory', None, 'Kept for parity.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    assert p.input_dim and p.hidden_dim
    odim = p.output_dim or p.input_dim
    self._odim = odim
    act = p.activation
    self._gated = isinstance(act, str) and act.startswith('GATED_')
    ff = p.fflayer_tpl.Copy().Set(
        input_dim=p.input_dim, hidden_layer_dims=[p.hidden_dim, odim],
        dropout=[layers.DropoutLayer.Params().Set(keep_prob=1.0 - p.relu_dropout_prob),
                 layers.DropoutLayer.Params().Set(keep_prob=1.0)])
    if self._gated:
      ff.activation = [act[len('GATED_'):], 'NONE']
      self.CreateChild('gate', layers.ProjectionLayer.Params().Set(
          input_dim=p.input_dim, output_dim=p.hidden_dim, activation='NONE',
          batch_norm=False, has_bias=True))
    else:
      ff.activation = [act, 'NONE']
    self.CreateChild('fflayer', ff)
    ln_name = 'pre_layer_norm' if p.primer_hybrid_norm else 'layer_norm'
    self.CreateChild(ln_name, p.ln_tpl.Copy().Set(
        input_dim=p.input_dim if p.pre_layer_norm else odim))
    if p.primer_hybrid_norm:
      self.CreateChild('post_layer_norm', p.ln_tpl.Copy().Set(input_dim=odim))
    self.CreateChild('residual_dropout', p.residual_dropout_tpl.Copy().Set(
        keep_prob=1.0 - p.residual_dropout_prob))
    if p.add_skip_connection and odim != p.input_dim:
      self.CreateChild('res_proj_layer', p.res_proj_tpl.Copy().Set(
          input_dim=p.input_dim, output_dim=odim, activation='NONE'))
    if p.residual_droppath_prob:
      self.CreateChild('residual_droppath', StochasticResidualLayer.Params().Set(
          residual_weight=p.residual_weight,
          survival_prob=1.0 - p.residual_droppath_prob))

  @property
  def output_dim(self):
    return self._odim

  @classmethod
  def NumOutputNodes(cls, p):
    return p.output_dim or p.input_dim

  def FProp(self, theta, inputs, paddings=None, tasks=None):
    p = self.params
    del tasks
    inputs = self._CastToFPropDtype(inputs)
    if p.pre_layer_norm:
      ln = self.pre_layer_norm if p.primer_hybrid_norm else self.layer_norm
      lt = theta.pre_layer_norm if p.primer_hybrid_norm else theta.layer_norm
      normed = ln.FProp(lt, inputs)
    else:
      normed = inputs
    pad = paddings.unsqueeze(-1) if paddings is not None else None
    if self._gated:
      ff, ft = self.fflayer, theta.fflayer
      h = ff.fc[0].FProp(ft.fc[0], normed) * self.gate.FProp(theta.gate, normed)
      h = ff.dropout[0].FProp(ft.dropout[0], h)
      h = ff.fc[1].FProp(ft.fc[1], h)
      if pad is not None:
        h = py_utils.ApplyPadding(pad, h)
    else:
      h = self.fflayer.FProp(theta.fflayer, normed, pad)
    if p.primer_hybrid_norm:
      h = self.post_layer_norm.FProp(theta.post_layer_norm, h)
    h = self.residual_dropout.FProp(theta.residual_dropout, h)
    if p.add_skip_connection:
      res = inputs
      if hasattr(self, 'res_proj_layer'):
        res = self.res_proj_layer.FProp(theta.res_proj_layer, inputs)
      if p.residual_droppath_prob:
        h = self.residual_droppath.FProp(theta.residual_droppath, res, h)
      else:
        h = res + h * p.residual_weight
    if not p.pre_layer_norm:
      h = self.layer_norm.FProp(theta.layer_norm, h)
    return h


class ReshapedTransformerFeedForwardLayer(TransformerFeedForwardLayer):
  """Inputs arrive `[..., N, D/N]`; flattened, processed, reshaped back (:1246)."""

  def FProp(self, theta, inputs, paddings=None, tasks=None):
    shp = inputs.shape
    out = super().FProp(theta, inputs.reshape(*shp[:-2], -1), paddings, tasks)
    return out.reshape(*shp[:-2], shp[-2], -1)


class TransformerFeedForwardLayerWithTaskId(TransformerFeedForwardLayer):
  """FFN whose projections are selected per task id (:bma 7720)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('use_task_ids', False, 'Select per-task weights.')
    return p


class MoEFeedforwardLayer(base_layer.BaseLayer):
  """GShard MoE layer behind the time-major FFN interface (:740)."""

  @classmethod
  def Params(cls):
    from lingvo_b200.core import gshard_builder  # pylint: disable=g-import-not-at-top
    p = super().Params()
    p.Define('moe_builder_p', gshard_builder.MoEBuilder.Params(), 'MoE builder.')
    p.Define('fflayer_residual_weight', 0.5, 'Residual weight.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    b = p.moe_builder_p.Instantiate()
    self.CreateChild('moe_fflayer', b.EncoderLayer(
        p.name, b.MoE(p.name), residual_weight=p.fflayer_residual_weight))

  def FProp(self, theta, inputs, paddings):
    seg = (1.0 - paddings).to(torch.int32)
    moe_in = NestedMap(vec=inputs, segment_id=seg, segment_pos=torch.zeros_like(seg))
    out = self.moe_fflayer.FProp(theta.moe_fflayer, moe_in)
    ctx = py_utils.AuxLossContext.Current()
    if ctx is None:
      raise ValueError('MoEFeedforwardLayer needs an AuxLossContext.')
    ctx.AddLoss(out.aux_loss)
    return out.vec


class HybridFeedforwardLayer(base_layer.BaseLayer):
  """Named sub-FFNs, chosen by `sub_key` at FProp time (:788)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('sub', None, 'Dict name → FFN layer params.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    self._keys = sorted(p.sub.keys())
    for k in self._keys:
      self.CreateChild(k, p.sub[k])

  def FProp(self, theta, inputs, paddings, sub_key=None):
    k = sub_key or self._keys[0]
    return getattr(self, k).FProp(theta[k], inputs, paddings)


class TransformerLayer(base_layer.BaseLayer):
  """Time-major Transformer block: self-atten [→ aux-atten] → FFN (:1334)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('source_dim', 0, 'Input dim.')
    p.Define('output_dim', 0, 'Output dim.')
    p.Define('tr_atten_tpl', TransformerAttentionLayer.Params().Set(
        num_attention_heads=8), 'Self-attention tpl.')
    p.Define('tr_post_ln_tpl', None, 'Optional output LN.')
    p.Define('tr_fflayer_tpl', TransformerFeedForwardLayer.Params().Set(hidden_dim=2048),
             'FFN tpl.')
    p.Define('has_aux_atten', False, 'Cross attention.')
    p.Define('tr_aux_atten_tpl', None, 'Cross-attention tpl (defaults to tr_atten_tpl).')
    p.Define('mask_self_atten', False, 'Masked self-attention.')
    p.Define('packed_input', False, 'Packed input.')
    p.Define('is_decoder', False, 'Kept for parity.')
    p.Define('num_aux_atten_post_proj', 1, 'Kept for parity.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    assert p.source_dim
    self.CreateChild('self_atten', p.tr_atten_tpl.Copy().Set(
        source_dim=p.source_dim, is_masked=p.mask_self_atten,
        packed_input=p.packed_input))
    if p.has_aux_atten:
      aux = (p.tr_aux_atten_tpl or p.tr_atten_tpl).Copy().Set(
          source_dim=p.source_dim, is_masked=False, packed_input=p.packed_input)
      self.CreateChild('atten', aux)
    self.CreateChild('fflayer', p.tr_fflayer_tpl.Copy().Set(
        input_dim=p.source_dim, output_dim=p.output_dim or p.source_dim))
    if p.tr_post_ln_tpl is not None:
      self.CreateChild('layer_norm', p.tr_post_ln_tpl.Copy().Set(
          input_dim=p.output_dim or p.source_dim))

  @property
  def output_dim(self):
    return self.fflayer.output_dim

  @classmethod
  def NumOutputNodes(cls, p):
    return p.output_dim or p.source_dim

  def FProp(self, theta, source_vecs, source_paddings, aux_vecs=None,
            aux_paddings=None, source_segment_id=None, aux_segment_id=None,
            **kwargs):
    p = self.params
    out, probs = self.self_atten.FProp(
        theta.self_atten, source_vecs, source_paddings,
        query_segment_id=source_segment_id)
    if p.has_aux_atten:
      assert aux_vecs is not None
      out, probs = self.atten.FProp(
          theta.atten, out, aux_paddings, aux_vecs,
          query_segment_id=source_segment_id, source_segment_id=aux_segment_id)
    out = self.fflayer.FProp(theta.fflayer, out, source_paddings)
    if p.tr_post_ln_tpl is not None:
      out = self.layer_norm.FProp(theta.layer_norm, out)
    return out, probs

  def ExtendStep(self, theta, source_vecs, prefix_states, aux_vecs=None,
                 aux_paddings=None, t=None, **kwargs):
    """source_vecs [B, D] → (out [B, D], probs [B, S], new prefix state)."""
    p = self.params
    out, probs, state = self.self_atten.ExtendStep(
        theta.self_atten, source_vecs, prefix_states, t)
    if p.has_aux_atten:
      o, pr = self.atten.FProp(theta.atten, out.unsqueeze(0), aux_paddings, aux_vecs)
      out, probs = o.squeeze(0), pr.squeeze(0)
    out = self.fflayer.FProp(theta.fflayer, out.unsqueeze(0),
                             torch.zeros(1, out.shape[0], device=out.device)).squeeze(0)
    if p.tr_post_ln_tpl is not None:
      out = self.layer_norm.FProp(theta.layer_norm, out)
    return out, probs, state


class TransformerLayerWithMultitaskAdapters(TransformerLayer):
  """TransformerLayer followed by per-task residual adapters (:2192)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('adapter_tpl', layers.MultitaskAdapterLayer.Params(), 'Adapter tpl.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    self.CreateChild('adapters', p.adapter_tpl.Copy().Set(
        input_dim=p.output_dim or p.source_dim))

  def FProp(self, theta, source_vecs, source_paddings, aux_vecs=None,
            aux_paddings=None, source_segment_id=None, aux_segment_id=None,
            source_task_id=None, **kwargs):
    out, probs = super().FProp(theta, source_vecs, source_paddings, aux_vecs,
                               aux_paddings, source_segment_id, aux_segment_id)
    out = self.adapters.FProp(theta.adapters, out, source_task_id)
    return out, probs


class SelfAttentiveLayer(base_layer.BaseLayer):
  """Structured self-attentive sentence embedding (Lin et al. 2017) (:2984):
  A = softmax(W2·tanh(W1·Hᵀ)); M = A·H, plus the ‖AAᵀ − I‖² penalty."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('num_heads', 5, 'Attention hops.')
    p.Define('input_dim', 0, 'Input dim.')
    p.Define('hidden_dim', 0, 'Hidden dim.')
    p.Define('penalty_coef', 1.0, 'Coefficient of the orthogonality penalty.')
    p.Define('penalty_terms', [1.0, 1.0], 'Kept for parity.')
    return p

  def _CreateLayerVariables(self):
    p = self.params
    self.CreateVariable('w1', WeightParams([p.input_dim, p.hidden_dim], p.params_init, p.dtype))
    self.CreateVariable('w2', WeightParams([p.hidden_dim, p.num_heads], p.params_init, p.dtype))

  def FProp(self, theta, inputs, paddings=None):
    """inputs [B, T, D] → ([B, N, D], penalty scalar)."""
    p = self.params
    hid = torch.tanh(torch.matmul(inputs, theta.w1.to(inputs.dtype)))
    logits = torch.matmul(hid, theta.w2.to(hid.dtype)).transpose(1, 2).float()
    if paddings is not None:
      logits = logits.masked_fill(paddings.unsqueeze(1) > 0, _NEG)
    a = torch.softmax(logits, -1)
    out = torch.bmm(a.to(inputs.dtype), inputs)
    eye = torch.eye(p.num_heads, device=a.device)
    pen = ((torch.bmm(a, a.transpose(1, 2)) - eye) ** 2).sum((1, 2)).mean()
    return out, p.penalty_coef * pen
