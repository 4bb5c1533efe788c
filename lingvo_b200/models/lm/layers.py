"""Language-model layers (ref `lingvo/tasks/lm/layers.py`).

`BaseLanguageModel` contract (ref :40-140):
  FProp(theta, inputs [T,B] ids (or activations), paddings [T,B], state0,
        labels=NestedMap(class_ids [T,B], class_weights [T,B]))
    → (xent_output NestedMap(logits?, log_probs, total_xent, avg_xent,
       total_weight, per_example_xent), state1)
  zero_state(theta, batch_size), Logits(...), Step/ExtendStep for decoding.

RNN LMs run through `rnn_layers.StackedFRNNLayerByLayer` (input GEMM hoisted out
of the time loop); Transformer LMs use the batch-major stack with the fused
attention path and the fused LM-head cross-entropy kernel.
"""

from __future__ import annotations

import math

import torch

from lingvo_b200.core import base_layer
from lingvo_b200.core import batch_major_attention as bma
from lingvo_b200.core import layers
from lingvo_b200.core import py_utils
from lingvo_b200.core import rnn_cell
from lingvo_b200.core import rnn_layers
from lingvo_b200.core.nested_map import NestedMap


def get_basic_rnn_lm_cell_params(dim, hidden=None):
  return rnn_cell.LSTMCellSimple.Params().Set(
      num_input_nodes=dim, num_output_nodes=dim, num_hidden_nodes=hidden or 0)


class BaseLanguageModel(base_layer.BaseLayer):
  """Abstract LM (ref :40)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('vocab_size', 0, 'Vocabulary size.')
    return p

  @classmethod
  def UpdateTargetVocabSize(cls, p, vocab_size, wpm_model=None):
    """Sets the vocabulary size on LM params (and, in subclasses, on their softmax /
    embedding) (ref :40)."""
    del wpm_model
    p.vocab_size = vocab_size
    for name in ('softmax', 'emb'):
      if name in p and p.Get(name) is not None:
        sub = p.Get(name)
        for key in ('num_classes', 'vocab_size'):
          if key in sub:
            sub.Set(**{key: vocab_size})
    return p

  def zero_state(self, theta, batch_size):
    raise NotImplementedError

  def FProp(self, theta, inputs, paddings, state0, labels=None, direct_features=None):
    raise NotImplementedError

  def GetFeedDict(self):
    """Optional extra inputs of the LM keyed by name (ref :149)."""
    return {}

  def CombineStates(self, state0, state1, switch_cond):
    """Per batch element: `state0` where `switch_cond [batch]` is true, else `state1`
    (ref :153) — used by beam search / shallow fusion to roll back rejected steps. The
    default works for any state whose tensors have the batch in dim 0."""
    def Pick(a, b):
      cond = switch_cond.reshape([-1] + [1] * (a.dim() - 1))
      return torch.where(cond, a, b)
    return py_utils.Transform(Pick, state0, state1) if not isinstance(state0, NestedMap) \
        else state0.Pack([Pick(a, b) for a, b in zip(state0.Flatten(), state1.Flatten())])

  def Logits(self, theta, inputs, paddings, *args, **kwargs):
    xent, _ = self.FProp(theta, inputs, paddings, *args, **kwargs)
    return xent.logits

  @classmethod
  def StepOutputDimension(cls, params):
    raise NotImplementedError

  def _Xent(self, softmax, theta_softmax, acts, labels):
    """acts [T,B,D] + labels → xent NestedMap with [T,B] tensors."""
    t, b = acts.shape[:2]
    flat = acts.reshape(t * b, -1)
    if labels is None:
      logits = softmax.Logits(theta_softmax, flat)
      return NestedMap(logits=logits.reshape(t, b, -1))
    ids = labels.class_ids.reshape(t * b, 1) if 'class_ids' in labels else None
    probs = labels.get('class_probabilities')
    out = softmax.FProp(
        theta_softmax, flat, labels.class_weights.reshape(t * b, 1),
        class_ids=ids,
        class_probabilities=None if probs is None else probs.reshape(t * b, -1))
    res = NestedMap(total_xent=out.total_xent, avg_xent=out.avg_xent,
                    total_weight=out.total_weight)
    if out.get('logits') is not None:
      res.logits = out.logits.reshape(t, b, -1)
    if out.get('log_probs') is not None:
      res.log_probs = out.log_probs.reshape(t, b, -1)
    if out.get('per_example_xent') is not None:
      res.per_example_xent = out.per_example_xent.reshape(t, b)
    if out.get('per_example_argmax') is not None:
      res.per_example_argmax = out.per_example_argmax.reshape(t, b)
    return res


def ComputeXentOutput(softmax_layer, softmax_theta, activations, labels, num_samples=1):
  """Softmax cross entropy of `[time, batch · num_samples, dim]` activations (ref :233):
  without labels only the logits; `labels.class_ids` / `class_probabilities` (+
  `class_weights`) are tiled `num_samples` times along the batch."""
  t, b = activations.shape[:2]
  if labels is None:
    logits = softmax_layer.Logits(softmax_theta, activations.reshape(t * b, -1))
    return NestedMap(logits=logits.reshape(t, b, -1))
  tile = (lambda x: x.repeat(1, num_samples)) if num_samples > 1 else (lambda x: x)
  flat = activations.reshape(t * b, -1)
  weights = tile(labels.class_weights).reshape(t * b, 1)
  if 'class_ids' in labels:
    return softmax_layer.FProp(softmax_theta, flat, weights,
                               class_ids=tile(labels.class_ids).reshape(t * b, 1))
  assert 'class_probabilities' in labels
  probs = labels.class_probabilities
  if num_samples > 1:
    probs = probs.repeat(1, num_samples, 1)
  return softmax_layer.FProp(softmax_theta, flat, weights,
                             class_probabilities=probs.reshape(t * b, -1))


class NullLm(BaseLanguageModel):
  """Uniform LM (all-zero logits) (ref :142)."""

  def zero_state(self, theta, batch_size):
    return NestedMap(m=torch.zeros(batch_size, 0))

  def FProp(self, theta, inputs, paddings, state0=None, labels=None, direct_features=None):
    p = self.params
    t, b = inputs.shape[:2]
    logits = torch.zeros(t, b, p.vocab_size, device=inputs.device)
    out = NestedMap(logits=logits, log_probs=torch.log_softmax(logits, -1))
    if labels is not None:
      w = labels.class_weights.float()
      per = torch.full((t, b), math.log(p.vocab_size), device=inputs.device)
      out.per_example_xent = per
      out.total_xent = (per * w).sum()
      out.total_weight = w.sum()
      out.avg_xent = out.total_xent / out.total_weight.clamp_min(1e-8)
    return out, state0


class RnnLmNoEmbedding(BaseLanguageModel):
  """Stacked RNN over pre-embedded inputs + softmax (ref :190)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('rnns', rnn_layers.StackedFRNNLayerByLayer.Params(), 'The RNN stack.')
    p.Define('softmax', layers.SimpleFullSoftmax.Params(), 'Softmax.')
    p.Define('output_dropout_prob', 0.0, 'Dropout on the RNN output.')
    p.Define('direct_features_dim', 0, 'Extra features concatenated before softmax.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    assert p.softmax.input_dim == p.rnns.num_output_nodes + p.direct_features_dim
    self.CreateChild('rnns', p.rnns)
    self.CreateChild('softmax', p.softmax.Copy().Set(num_classes=p.vocab_size))
    self.CreateChild('output_dropout', layers.DropoutLayer.Params().Set(
        keep_prob=1.0 - p.output_dropout_prob))

  def zero_state(self, theta, batch_size):
    return self.rnns.zero_state(theta.rnns, batch_size)

  @classmethod
  def StepOutputDimension(cls, params):
    return NestedMap(logits=params.vocab_size, last_hidden=params.softmax.input_dim)

  def FProp(self, theta, inputs, paddings, state0=None, labels=None, direct_features=None):
    pad3 = paddings.unsqueeze(-1) if paddings.dim() == 2 else paddings
    acts, state1 = self.rnns.FProp(theta.rnns, inputs, pad3, state0)
    acts = self.output_dropout.FProp(theta.output_dropout, acts)
    if direct_features is not None:
      acts = torch.cat([acts, direct_features.to(acts.dtype)], -1)
    out = self._Xent(self.softmax, theta.softmax, acts, labels)
    out.last_hidden = acts
    return out, state1


class RnnLm(RnnLmNoEmbedding):
  """Embedding + RnnLmNoEmbedding (ref :320)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('emb', layers.SimpleEmbeddingLayer.Params(), 'Embedding.')
    p.Define('embedding_dropout_keep_prob', 1.0, 'Embedding dropout keep prob.')
    p.Define('embedding_dropout_seed', None, 'Kept for parity.')
    return p

  @classmethod
  def CommonParams(cls, vocab_size, emb_dim=1024, num_layers=2, rnn_dims=2048,
                   rnn_hidden_dims=0, residual_start=1, softmax_max_alloc=None):
    p = cls.Params()
    p.vocab_size = vocab_size
    p.emb.Set(vocab_size=vocab_size, embedding_dim=emb_dim)
    p.rnns.Set(num_layers=num_layers, num_input_nodes=emb_dim, num_output_nodes=rnn_dims,
               skip_start=residual_start,
               cell_tpl=rnn_cell.LSTMCellSimple.Params().Set(
                   num_hidden_nodes=rnn_hidden_dims))
    p.softmax.Set(input_dim=rnn_dims, num_classes=vocab_size)
    if softmax_max_alloc:
      p.softmax.chunk_size = max(1, int(softmax_max_alloc / vocab_size))
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    assert p.emb.vocab_size == p.vocab_size
    self.CreateChild('emb', p.emb)
    self.CreateChild('emb_dropout', layers.DropoutLayer.Params().Set(
        keep_prob=p.embedding_dropout_keep_prob))

  def FProp(self, theta, inputs, paddings, state0=None, labels=None, direct_features=None):
    ids = inputs.long()
    acts = self.emb.EmbLookup(theta.emb, ids)
    acts = self.emb_dropout.FProp(theta.emb_dropout, acts)
    return super().FProp(theta, acts, paddings, state0, labels, direct_features)


class ConditionalRnnLm(RnnLmNoEmbedding):
  """RNN LM whose every input step is the token embedding concatenated with a
  per-sequence condition vector (ref :679)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('condition_dim', 128, 'Size of the condition vector.')
    p.Define('emb', layers.SimpleEmbeddingLayer.Params(), 'Embedding.')
    p.Define('embedding_dropout_keep_prob', 1.0, 'Embedding dropout keep prob.')
    p.Define('embedding_dropout_seed', None, 'Kept for parity.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    assert p.emb.vocab_size == p.vocab_size
    assert p.emb.embedding_dim + p.condition_dim == p.rnns.num_input_nodes, (
        'rnn input = embedding ⊕ condition')
    self.CreateChild('emb', p.emb)
    self.CreateChild('emb_dropout', layers.DropoutLayer.Params().Set(
        keep_prob=p.embedding_dropout_keep_prob))

  def FProp(self, theta, inputs, paddings, state0, condition, labels=None,
            direct_features=None):
    """inputs/paddings [T,B]; condition [B, condition_dim]."""
    p = self.params
    assert condition.shape == (paddings.shape[1], p.condition_dim)
    acts = self.emb.EmbLookup(theta.emb, inputs.long())
    cond = condition.to(acts.dtype).unsqueeze(0).expand(acts.shape[0], -1, -1)
    acts = self.emb_dropout.FProp(theta.emb_dropout, torch.cat([acts, cond], -1))
    return super().FProp(theta, acts, paddings, state0, labels, direct_features)


class MoeLm(BaseLanguageModel):
  """Mixture of RNN-LM experts (ref :763): one RNN stack predicts a soft domain
  assignment per step; `number_of_experts` further stacks run in parallel and their
  outputs are mixed by that assignment, then fed to a merging LM (or a plain softmax)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('emb', layers.SimpleEmbeddingLayer.Params(), 'Embedding.')
    p.Define('shared_emb', True, 'One embedding for the gate and all experts.')
    p.Define('add_postgating_rnn', True, 'Merge with an RNN LM (else a softmax only).')
    p.Define('rnns', rnn_layers.StackedFRNNLayerByLayer.Params(), 'RNN stack template.')
    p.Define('number_of_experts', 7, 'Number of experts.')
    p.Define('merge', RnnLmNoEmbedding.Params(), 'The LM applied to the mixed features.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    n = 1 + p.number_of_experts
    assert p.emb.vocab_size == p.vocab_size
    assert p.emb.embedding_dim == p.rnns.num_input_nodes
    if p.shared_emb:
      self.CreateChild('emb', p.emb)
    else:
      self.CreateChildren('emb', [p.emb.Copy().Set(name='emb_%d' % i) for i in range(n)])
    self.CreateChildren('rnns', [p.rnns.Copy().Set(name='rnns_%d' % i) for i in range(n)])
    dim = p.rnns.num_output_nodes
    self.CreateChild('domain_predictor_softmax', layers.SimpleFullSoftmax.Params().Set(
        input_dim=dim, num_classes=p.number_of_experts))
    if p.add_postgating_rnn:
      assert p.merge.vocab_size == p.vocab_size
      self.CreateChild('merge', p.merge)
    else:
      self.CreateChild('output_softmax', layers.SimpleFullSoftmax.Params().Set(
          input_dim=dim, num_classes=p.vocab_size))

  def zero_state(self, theta, batch_size):
    st = NestedMap(rnns=[r.zero_state(theta.rnns[i], batch_size)
                         for i, r in enumerate(self.rnns)])
    if self.params.add_postgating_rnn:
      st.merge = self.merge.zero_state(theta.merge, batch_size)
    return st

  def FProp(self, theta, inputs, paddings, state0, labels=None, direct_features=None):
    p = self.params
    ids = inputs.long()
    t, b = ids.shape
    pad3 = paddings.unsqueeze(-1)
    n = 1 + p.number_of_experts
    if p.shared_emb:
      embs = [self.emb.EmbLookup(theta.emb, ids)] * n
    else:
      embs = [self.emb[i].EmbLookup(theta.emb[i], ids) for i in range(n)]
    acts, state1 = [], NestedMap(rnns=[])
    for i in range(n):
      a, st = self.rnns[i].FProp(theta.rnns[i], embs[i], pad3, state0.rnns[i])
      acts.append(a)
      state1.rnns.append(st)
    gate_logits = self.domain_predictor_softmax.Logits(
        theta.domain_predictor_softmax, acts[0].reshape(t * b, -1))
    gating = torch.softmax(gate_logits.float(), -1).reshape(t, b, -1).to(acts[0].dtype)
    experts = torch.stack(acts[1:], 2)                      # [T,B,E,D]
    mixed = (gating.unsqueeze(-1) * experts).sum(2)
    if p.add_postgating_rnn:
      out, state1.merge = self.merge.FProp(theta.merge, mixed, paddings, state0.merge, labels)
    else:
      out = self._Xent(self.output_softmax, theta.output_softmax, mixed, labels)
    out.gating = gating
    return out, state1


class TransformerLmNoEmbedding(BaseLanguageModel):
  """Causal Transformer stack over pre-embedded `[T,B,D]` inputs (ref :560)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('position_emb', layers.PositionalEmbeddingLayer.Params(), 'Positions.')
    p.Define('model_dim', 512, 'Model dim.')
    p.Define('num_trans_layers', 6, 'Layers.')
    p.Define('trans_tpl', bma.TransformerLayer.Params(), 'Layer template.')
    p.Define('input_dropout_prob', 0.0, 'Input dropout.')
    p.Define('residual_dropout_prob', 0.0, 'Residual dropout.')
    p.Define('atten_dropout_prob', 0.0, 'Attention dropout.')
    p.Define('relu_dropout_prob', 0.0, 'FFN dropout.')
    p.Define('softmax', layers.SimpleFullSoftmax.Params(), 'Softmax.')
    p.Define('num_atten_heads', 8, 'Heads.')
    p.Define('hidden_dim', 2048, 'FFN hidden dim.')
    p.Define('packed_input', False, 'Packed inputs.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    self.CreateChild('position_emb', p.position_emb.Copy().Set(embedding_dim=p.model_dim))
    self.CreateChild('input_dropout', layers.DropoutLayer.Params().Set(
        keep_prob=1.0 - p.input_dropout_prob))
    tpl = p.trans_tpl.Copy()
    for a in (tpl.tr_atten_tpl, tpl.tr_self_atten_tpl):
      if a is not None:
        a.atten_tpl.return_atten_probs = False      # stay on the fused kernel
    self.CreateChild('stack', bma.StackedTransformerLayers.Params().Set(
        num_layers=p.num_trans_layers, mdl_dim=p.model_dim, hidden_dim=p.hidden_dim,
        num_atten_heads=p.num_atten_heads, dropout_prob=p.residual_dropout_prob,
        mask_self_atten=True, packed_input=p.packed_input, final_layer_norm=True,
        transformer_layer_params_tpl=tpl))
    self.CreateChild('softmax', p.softmax.Copy().Set(
        input_dim=p.model_dim, num_classes=p.vocab_size))

  def zero_state(self, theta, batch_size):
    return NestedMap()

  @classmethod
  def StepOutputDimension(cls, params):
    return NestedMap(logits=params.vocab_size, last_hidden=params.model_dim)

  def FProp(self, theta, inputs, paddings, state0=None, labels=None, direct_features=None,
            segment_ids=None, segment_pos=None):
    p = self.params
    t, b, _ = inputs.shape
    x = inputs.transpose(0, 1)                         # [B,T,D]
    pad = paddings.transpose(0, 1)
    if segment_pos is not None:
      pos = self.position_emb.FPropWithPosition(theta.position_emb, segment_pos.transpose(0, 1))
    else:
      pos = self.position_emb.FProp(theta.position_emb, t).unsqueeze(0)
    x = self.input_dropout.FProp(theta.input_dropout, x + pos.to(x.dtype))
    seg_mask = None
    if p.packed_input and segment_ids is not None:
      seg_mask = bma.CausalSegmentMask(segment_ids.transpose(0, 1))
    out, _ = self.stack.FProp(theta.stack, x, pad, segment_mask=seg_mask)
    acts = out.transpose(0, 1)
    res = self._Xent(self.softmax, theta.softmax, acts, labels)
    res.last_hidden = acts
    return res, state0


class TransformerLm(TransformerLmNoEmbedding):
  """Token embedding (scaled by √D) + TransformerLmNoEmbedding (ref :760)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('emb', layers.SimpleEmbeddingLayer.Params(), 'Embedding.')
    return p

  @classmethod
  def CommonParams(cls, model_dim, hidden_dim, num_heads, num_layers, learning_rate=None,
                   warmup_steps=None, vocab_size=0, input_dropout_prob=0.0,
                   residual_dropout_prob=0.1, atten_dropout_prob=0.0,
                   relu_dropout_prob=0.0, softmax_max_alloc=None):
    del learning_rate, warmup_steps
    p = cls.Params()
    p.vocab_size = vocab_size
    p.model_dim = model_dim
    p.hidden_dim = hidden_dim
    p.num_atten_heads = num_heads
    p.num_trans_layers = num_layers
    p.input_dropout_prob = input_dropout_prob
    p.residual_dropout_prob = residual_dropout_prob
    p.atten_dropout_prob = atten_dropout_prob
    p.relu_dropout_prob = relu_dropout_prob
    p.emb.Set(vocab_size=vocab_size, embedding_dim=model_dim)
    p.softmax.Set(input_dim=model_dim, num_classes=vocab_size)
    if softmax_max_alloc:
      p.softmax.chunk_size = max(1, int(softmax_max_alloc / vocab_size))
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    assert p.emb.embedding_dim == p.model_dim
    self.CreateChild('emb', p.emb.Copy().Set(vocab_size=p.vocab_size))

  def FProp(self, theta, inputs, paddings, state0=None, labels=None, direct_features=None,
            segment_ids=None, segment_pos=None):
    p = self.params
    acts = self.emb.EmbLookup(theta.emb, inputs.long()) * (p.model_dim ** 0.5)
    return super().FProp(theta, acts, paddings, state0, labels, direct_features,
                         segment_ids, segment_pos)


class GPipeTransformerLm(TransformerLm):
  """TransformerLm whose layer stack is split into `num_splits` pipeline cells
  (ref :1010); with a `parallel.pp.PipelineEngine` the cells run on separate ranks."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('num_splits', 1, 'Pipeline stages.')
    p.Define('num_micro_batches', 1, 'Micro-batches.')
    return p
