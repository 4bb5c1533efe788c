"""MT decoders (ref `lingvo/tasks/mt/decoder.py`).

`MTBaseDecoder` (ref :34): softmax + label smoothing + loss/metric assembly.
`MTDecoderV1` (ref :398): RNMT attention decoder (FRNNWithAttention + stacked
LSTMs, context fed to every layer and to the softmax).
`TransformerDecoder` (ref :1219): masked self-attention + cross-attention stack,
with incremental `ExtendStep` for beam search.
"""

from __future__ import annotations

import torch

from lingvo_b200.core import attention
from lingvo_b200.core import base_decoder
from lingvo_b200.core import batch_major_attention as bma
from lingvo_b200.core import layers
from lingvo_b200.core import py_utils
from lingvo_b200.core import rnn_cell
from lingvo_b200.core import rnn_layers
from lingvo_b200.core.nested_map import NestedMap


class MTBaseDecoder(base_decoder.BaseBeamSearchDecoder):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('label_smoothing', None, 'Label smoother params.')
    p.Define('softmax', layers.SimpleFullSoftmax.Params(), 'Softmax.')
    p.Define('per_word_avg_loss', False, 'Average the loss per word (else per sentence).')
    p.Define('per_example_tensors', False, 'Emit per-example tensors.')
    p.Define('token_normalized_per_seq_loss', False, 'Kept for parity.')
    p.Define('use_prev_atten_ctx', False, 'Kept for parity.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    if p.label_smoothing is not None:
      self.CreateChild('smoother', p.label_smoothing.Copy().Set(
          num_classes=p.softmax.num_classes))

  @classmethod
  def UpdateTargetVocabSize(cls, p, vocab_size, wpm_model=None):
    p.softmax.num_classes = vocab_size
    return p

  def _FPropSoftmax(self, theta, softmax_input, target_labels, target_weights,
                    target_paddings, target_segment_ids=None):
    """softmax_input [T,B,D]; labels/weights/paddings [T,B] → (metrics, per_seq)."""
    p = self.params
    t, b, d = softmax_input.shape
    w = target_weights.float()
    kwargs = dict(class_ids=target_labels.reshape(-1, 1).long())
    if p.label_smoothing is not None:
      probs = self.smoother.FProp(theta.smoother, target_paddings, target_labels.long(),
                                  target_ids=None)
      kwargs = dict(class_probabilities=probs.reshape(t * b, -1))
    out = self.softmax.FProp(theta.softmax, softmax_input.reshape(t * b, d),
                             w.reshape(-1, 1), **kwargs)
    per_tok = out.per_example_xent.reshape(t, b)
    seq_xent = (per_tok * w).sum(0)
    num_words = w.sum().clamp_min(1e-8)
    if p.per_word_avg_loss:
      loss, loss_w = out.total_xent / num_words, num_words
    else:
      loss = out.total_xent / float(b)
      loss_w = torch.tensor(float(b), device=w.device)
    correct = ((out.per_example_argmax.reshape(t, b) == target_labels).float() * w).sum() \
        if out.get('per_example_argmax') is not None else torch.zeros((), device=w.device)
    metrics = NestedMap(
        loss=(loss, loss_w), log_pplx=(out.total_xent / num_words, num_words),
        fraction_of_correct_next_step_preds=(correct / num_words, num_words),
        num_predictions=(num_words, 1.0))
    return metrics, NestedMap(per_sequence_xent=seq_xent)

  def ComputeLoss(self, theta, predictions, targets):
    lab = targets.labels.t()
    w = targets.weights.t()
    pad = targets.paddings.t()
    return self._FPropSoftmax(theta, predictions.softmax_input, lab, w, pad)


class MTDecoderV1(MTBaseDecoder):
  """RNMT decoder."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('emb', layers.SimpleEmbeddingLayer.Params(), 'Embedding.')
    p.Define('source_dim', 1024, 'Encoder output dim.')
    p.Define('attention', attention.AdditiveAttention.Params(), 'Attention.')
    p.Define('atten_rnn_cell_tpl', rnn_cell.LSTMCellSimple.Params(), 'Attention RNN cell.')
    p.Define('rnn_cell_tpl', rnn_cell.LSTMCellSimple.Params(), 'Upper RNN cells.')
    p.Define('rnn_cell_dim', 1024, 'RNN cell dim.')
    p.Define('rnn_layers', 8, 'Total decoder RNN layers.')
    p.Define('residual_start', 2, 'First residual layer.')
    p.Define('atten_rnn_cls', rnn_layers.FRNNWithAttention, 'Attention RNN class.')
    p.Define('feed_attention_context_vec_to_softmax', False, 'Concat context to softmax in.')
    p.Define('dropout_prob', 0.0, 'Dropout.')
    p.Define('cc_schedule', None, 'Kept for parity.')
    p.Define('init_step_ids', False, 'Kept for parity.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    self.CreateChild('emb', p.emb)
    self.CreateChild('dropout', layers.DropoutLayer.Params().Set(keep_prob=1.0 - p.dropout_prob))
    atten = p.attention.Copy().Set(source_dim=p.source_dim, query_dim=p.rnn_cell_dim)
    if 'context_dim' in atten:
      atten.context_dim = p.source_dim
    cell = p.atten_rnn_cell_tpl.Copy().Set(
        num_input_nodes=p.emb.embedding_dim + p.source_dim, num_output_nodes=p.rnn_cell_dim)
    self.CreateChild('frnn_with_atten', p.atten_rnn_cls.Params().Set(
        cell=cell, attention=atten, output_prev_atten_ctx=False, use_zero_atten_state=True,
        atten_context_dim=p.source_dim, packed_input=p.packed_input))
    rnns = []
    for i in range(1, p.rnn_layers):
      rnns.append(rnn_layers.FRNN.Params().Set(
          name='frnn_%d' % i, packed_input=p.packed_input,
          cell=p.rnn_cell_tpl.Copy().Set(num_input_nodes=p.rnn_cell_dim + p.source_dim,
                                         num_output_nodes=p.rnn_cell_dim)))
    self.CreateChildren('frnn', rnns)
    sm_in = p.rnn_cell_dim + (p.source_dim if p.feed_attention_context_vec_to_softmax else 0)
    self.CreateChild('softmax', p.softmax.Copy().Set(input_dim=sm_in))

  def _Upper(self, theta, xs, ctx, pad, states=None, step=False):
    p = self.params
    new_states = []
    for i, r in enumerate(self.frnn):
      inp = torch.cat([xs, ctx], -1)
      if step:
        st, _ = r.cell.FProp(theta.frnn[i].cell, states[i],
                             NestedMap(act=[inp], padding=pad))
        ys = r.cell.GetOutput(st)
        new_states.append(st)
      else:
        ys, _ = r.FProp(theta.frnn[i], self.dropout.FProp(theta.dropout, inp), pad)
      xs = xs + ys if i + 1 >= p.residual_start else ys
    return xs, new_states

  def ComputePredictions(self, theta, encoder_outputs, targets):
    p = self.params
    ids = targets.ids.t().long()
    pad = targets.paddings.t().float().unsqueeze(-1)
    emb = self.dropout.FProp(theta.dropout, self.emb.EmbLookup(theta.emb, ids))
    ctx, xs, probs, _ = self.frnn_with_atten.FProp(
        theta.frnn_with_atten, encoder_outputs.encoded, encoder_outputs.padding, emb, pad)
    xs, _ = self._Upper(theta, xs, ctx, pad)
    sm_in = torch.cat([xs, ctx], -1) if p.feed_attention_context_vec_to_softmax else xs
    sm_in = self.dropout.FProp(theta.dropout, sm_in)
    return NestedMap(softmax_input=sm_in, attention=NestedMap(probs=probs))

  # -- beam search callbacks ------------------------------------------------------
  def _InitBeamSearchStateCallback(self, theta, encoder_outputs, num_hyps_per_beam):
    p = self.params
    src_b = encoder_outputs.encoded.shape[1]
    n = src_b * num_hyps_per_beam
    fa = self.frnn_with_atten
    packed = fa.InitForSourcePacked(theta.frnn_with_atten, encoder_outputs.encoded,
                                    encoder_outputs.padding)
    encoder_outputs.packed_src = packed
    st = fa.zero_state(theta.frnn_with_atten, encoder_outputs.encoded, packed, n)
    upper = [r.zero_state(theta.frnn[i], n) for i, r in enumerate(self.frnn)]
    s_len = encoder_outputs.encoded.shape[0]
    dev = encoder_outputs.encoded.device
    init = NestedMap(log_probs=torch.zeros(n, p.softmax.num_classes, device=dev),
                     atten_probs=torch.zeros(n, s_len, device=dev))
    return init, NestedMap(atten=st, upper=upper)

  def _PreBeamSearchStepCallback(self, theta, encoder_outputs, step_ids, states,
                                 num_hyps_per_beam, cur_step):
    p = self.params
    n = step_ids.shape[0]
    emb = self.emb.EmbLookup(theta.emb, step_ids.squeeze(1).long())
    pad = torch.zeros(n, 1, device=emb.device)
    st = self.frnn_with_atten.Step(theta.frnn_with_atten, encoder_outputs.packed_src,
                                   states.atten, emb, pad)
    xs = self.frnn_with_atten.cell.GetOutput(st.rnn)
    xs, upper = self._Upper(theta, xs, st.atten, pad, states.upper, step=True)
    sm_in = torch.cat([xs, st.atten], -1) if p.feed_attention_context_vec_to_softmax else xs
    logits = self.softmax.Logits(theta.softmax, sm_in)
    return (NestedMap(log_probs=torch.log_softmax(logits.float(), -1),
                      atten_probs=st.atten_probs), NestedMap(atten=st, upper=upper))


class TransformerDecoder(MTBaseDecoder):
  """Transformer decoder."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('token_emb', layers.SimpleEmbeddingLayer.Params(), 'Token embedding.')
    p.Define('shared_emb', None, 'Kept for parity.')
    p.Define('position_emb', layers.PositionalEmbeddingLayer.Params(), 'Positions.')
    p.Define('source_dim', 512, 'Encoder dim.')
    p.Define('model_dim', 512, 'Model dim.')
    p.Define('num_trans_layers', 6, 'Layers.')
    p.Define('trans_tpl', bma.TransformerDecoderLayer.Params(), 'Layer template.')
    p.Define('input_dropout_prob', 0.0, 'Input dropout.')
    p.Define('is_transparent', False, 'Kept for parity.')
    p.Define('final_layer_norm', True, 'LN before the softmax.')
    p.Define('hidden_dim', 2048, 'FFN hidden dim.')
    p.Define('num_atten_heads', 8, 'Heads.')
    p.Define('residual_dropout_prob', 0.0, 'Residual dropout.')
    p.softmax.num_classes = 32000
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    self.CreateChild('token_emb', p.token_emb.Copy().Set(embedding_dim=p.model_dim))
    self.CreateChild('position_emb', p.position_emb.Copy().Set(embedding_dim=p.model_dim))
    self.CreateChild('input_dropout', layers.DropoutLayer.Params().Set(
        keep_prob=1.0 - p.input_dropout_prob))
    tpl = p.trans_tpl.Copy()
    tpl.aux_atten_input_dim = p.source_dim
    tpl.tr_self_atten_tpl = tpl.tr_self_atten_tpl or tpl.tr_atten_tpl.Copy()
    tpl.tr_self_atten_tpl.atten_tpl.return_atten_probs = False
    self.CreateChild('stack', bma.StackedTransformerLayers.Params().Set(
        num_layers=p.num_trans_layers, mdl_dim=p.model_dim, hidden_dim=p.hidden_dim,
        num_atten_heads=p.num_atten_heads, dropout_prob=p.residual_dropout_prob,
        mask_self_atten=True, has_aux_atten=True, packed_input=p.packed_input,
        final_layer_norm=p.final_layer_norm, transformer_layer_params_tpl=tpl))
    self.CreateChild('softmax', p.softmax.Copy().Set(input_dim=p.model_dim))

  def _Embed(self, theta, ids, t0=None):
    p = self.params
    x = self.token_emb.EmbLookup(theta.token_emb, ids.long()) * (p.model_dim ** 0.5)
    t = ids.shape[1]
    if t0 is None:
      pos = self.position_emb.FProp(theta.position_emb, t).unsqueeze(0)
    else:
      pos = self.position_emb.FProp(theta.position_emb, t0 + t)[t0:t0 + t].unsqueeze(0)
    return x + pos.to(x.dtype)

  def ComputePredictions(self, theta, encoder_outputs, targets):
    x = self.input_dropout.FProp(theta.input_dropout, self._Embed(theta, targets.ids))
    aux = encoder_outputs.encoded.transpose(0, 1)
    aux_pad = encoder_outputs.padding.t()
    out, _ = self.stack.FProp(theta.stack, x, targets.paddings.float(), aux, aux_pad)
    return NestedMap(softmax_input=out.transpose(0, 1))

  def _InitBeamSearchStateCallback(self, theta, encoder_outputs, num_hyps_per_beam):
    p = self.params
    src_b = encoder_outputs.encoded.shape[1]
    n = src_b * num_hyps_per_beam
    dev = encoder_outputs.encoded.device
    t_max = p.target_seq_len
    aux = encoder_outputs.encoded.transpose(0, 1)              # [B,S,D]
    # hyp index = hyp_id * src_b + beam → tile sources along dim 0
    encoder_outputs.aux_tiled = aux.repeat(num_hyps_per_beam, 1, 1)
    encoder_outputs.aux_pad_tiled = encoder_outputs.padding.t().repeat(num_hyps_per_beam, 1)
    cache = self.stack.InitStates(theta.stack, n, t_max)
    # caches are [T, n, N, H]: make dim 0 the hyp dim for the helper's re-ordering
    cache = cache.Transform(lambda x: x.transpose(0, 1).contiguous())
    init = NestedMap(log_probs=torch.zeros(n, p.softmax.num_classes, device=dev),
                     atten_probs=torch.zeros(n, aux.shape[1], device=dev))
    return init, NestedMap(cache=cache, time_step=torch.zeros(n, dtype=torch.int64, device=dev))

  def _PreBeamSearchStepCallback(self, theta, encoder_outputs, step_ids, states,
                                 num_hyps_per_beam, cur_step):
    x = self._Embed(theta, step_ids, t0=cur_step)
    cache = states.cache.Transform(lambda c: c.transpose(0, 1))
    out, new_cache = self.stack.ExtendStep(
        theta.stack, x, encoder_outputs.aux_tiled, encoder_outputs.aux_pad_tiled, cache,
        cur_step)
    logits = self.softmax.Logits(theta.softmax, out.squeeze(1))
    n = step_ids.shape[0]
    new_cache = new_cache.Transform(lambda c: c.transpose(0, 1).contiguous())
    atten = torch.full((n, encoder_outputs.aux_tiled.shape[1]),
                       1.0 / encoder_outputs.aux_tiled.shape[1], device=logits.device)
    return (NestedMap(log_probs=torch.log_softmax(logits.float(), -1), atten_probs=atten),
            NestedMap(cache=new_cache, time_step=states.time_step + 1))


class TransformerBatchMajorDecoder(TransformerDecoder):
  """Batch-major Transformer decoder (ref :2361). `TransformerDecoder` here already runs
  batch-major over the fused attention kernels with a pre-allocated KV cache; this class
  carries the reference's extra knobs."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('input_data_format', 'TBC', "Encoder output layout: 'TBC' or 'BTC'.")
    p.Define('prediction_data_format', 'TBC', "Layout of softmax_input: 'TBC' or 'BTC'.")
    p.Define('use_fused_layernorm', False, 'Kept for parity (LN is always fused).')
    p.Define('use_fast_softmax', False, 'Kept for parity.')
    return p

  def ComputePredictions(self, theta, encoder_outputs, targets):
    p = self.params
    if p.input_data_format == 'BTC':
      encoder_outputs = NestedMap(encoded=encoder_outputs.encoded.transpose(0, 1),
                                  padding=encoder_outputs.padding.t())
    out = super().ComputePredictions(theta, encoder_outputs, targets)
    if p.prediction_data_format == 'BTC':
      out.softmax_input = out.softmax_input.transpose(0, 1)
    return out

  def ComputeLoss(self, theta, predictions, targets):
    if self.params.prediction_data_format == 'BTC':
      predictions = NestedMap(softmax_input=predictions.softmax_input.transpose(0, 1))
    return super().ComputeLoss(theta, predictions, targets)


class TransformerXDecoder(MTBaseDecoder):
  """Decoder for XEnDec (ref :2935): target embeddings of two sentences can be
  interpolated (`other_targets`, `lambdas`), the cross-attention distribution is returned
  (it drives the label mixing ratios), and the loss accepts soft target distributions."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('token_emb', layers.SimpleEmbeddingLayer.Params(), 'Token embedding.')
    p.Define('position_emb', layers.PositionalEmbeddingLayer.Params(), 'Positions.')
    p.Define('source_dim', 512, 'Encoder dim.')
    p.Define('model_dim', 512, 'Model dim.')
    p.Define('num_trans_layers', 6, 'Layers.')
    p.Define('trans_tpl', None, 'Time-major TransformerLayer template.')
    p.Define('input_dropout_prob', 0.0, 'Input dropout.')
    p.Define('hidden_dim', 2048, 'FFN hidden dim.')
    p.Define('num_atten_heads', 8, 'Heads.')
    p.softmax.num_classes = 32000
    return p

  def __init__(self, params):
    super().__init__(params)
    from lingvo_b200.core import layers_with_attention as lwa
    from lingvo_b200.models.mt import layers as mt_layers
    p = self.params
    self.CreateChild('token_emb', p.token_emb.Copy().Set(embedding_dim=p.model_dim))
    self.CreateChild('position_emb', p.position_emb.Copy().Set(embedding_dim=p.model_dim))
    self.CreateChild('input_dropout', layers.DropoutLayer.Params().Set(
        keep_prob=1.0 - p.input_dropout_prob))
    tpl = (p.trans_tpl or lwa.TransformerLayer.Params()).Copy()
    tpl.tr_atten_tpl.num_attention_heads = p.num_atten_heads
    tpl.tr_fflayer_tpl.hidden_dim = p.hidden_dim
    self.CreateChild('stack', mt_layers.TransformerStack.Params().Set(
        model_dim=p.model_dim, num_transformer_layers=p.num_trans_layers,
        transformer_tpl=tpl, ln_output=True, has_aux_attention=True, mask_self_atten=True))
    self.CreateChild('softmax', p.softmax.Copy().Set(input_dim=p.model_dim))

  def ComputePredictions(self, theta, encoder_outputs, targets, other_targets=None,
                         lambdas=None):
    p = self.params
    ids = targets.ids.long()
    emb = targets.get('embs')
    if emb is None or other_targets is None:
      emb = self.token_emb.EmbLookup(theta.token_emb, ids)
    pad = targets.paddings.float()
    if other_targets is not None:
      other = other_targets.get('embs')
      if other is None:
        other = self.token_emb.EmbLookup(theta.token_emb, other_targets.ids.long())
      emb = lambdas[0].unsqueeze(-1).to(emb.dtype) * emb + \
          lambdas[1].unsqueeze(-1).to(emb.dtype) * other
      pad = (pad + other_targets.paddings.float() - 1.0).clamp(0.0, 1.0)
    orig = emb
    x = emb * (p.model_dim ** 0.5)
    pos = self.position_emb.FProp(theta.position_emb, ids.shape[1]).unsqueeze(0)
    x = self.input_dropout.FProp(theta.input_dropout, x + pos.to(x.dtype)).transpose(0, 1)
    out, _, _, probs = self.stack.FProp(
        theta.stack, x, pad.t(), aux_vecs=encoder_outputs.encoded,
        aux_paddings=encoder_outputs.padding, return_atten_probs=True)
    # probs: [T, B, S] cross-attention of the last layer → [B, T, S]
    return NestedMap(softmax_input=out, attention=NestedMap(probs=probs.transpose(0, 1)),
                     source_embs=encoder_outputs.get('embedded_inputs'), target_embs=orig)

  def ComputeLoss(self, theta, predictions, targets, target_probs=None):
    """Hard-label loss, or cross entropy against `target_probs [B,T,V]` (mixed labels)."""
    p = self.params
    x = predictions.softmax_input                      # [T,B,D]
    t, b, d = x.shape
    w = targets.weights.t().float()
    lab = targets.labels.t().long()
    logits = self.softmax.Logits(theta.softmax, x.reshape(t * b, d)).float()
    logp = torch.log_softmax(logits, -1)
    hard = torch.nn.functional.one_hot(lab.reshape(-1), logits.shape[-1]).float()
    if target_probs is None:
      tp = hard
      if p.label_smoothing is not None:
        tp = self.smoother.FProp(theta.smoother, targets.paddings.t(), lab,
                                 target_ids=None).reshape(t * b, -1)
    else:
      tp = target_probs.transpose(0, 1).reshape(t * b, -1)
    per_tok = -(tp * logp).sum(-1).reshape(t, b)
    total = (per_tok * w).sum()
    num_words = w.sum().clamp_min(1e-8)
    if p.per_word_avg_loss:
      loss, loss_w = total / num_words, num_words
    else:
      loss, loss_w = total / float(b), torch.tensor(float(b), device=w.device)
    correct = ((logits.argmax(-1).reshape(t, b) == lab).float() * w).sum()
    metrics = NestedMap(
        loss=(loss, loss_w), log_pplx=(total / num_words, num_words),
        fraction_of_correct_next_step_preds=(correct / num_words, num_words),
        num_predictions=(num_words, 1.0))
    per_seq = NestedMap(
        per_sequence_xent=(per_tok * w).sum(0),
        reshape_probs=torch.softmax(logits, -1).reshape(t, b, -1).transpose(0, 1).detach(),
        target_hard_probs=hard.reshape(t, b, -1).transpose(0, 1))
    return metrics, per_seq

  # decode: greedy/beam through full re-computation of the prefix (XEnDec is a training
  # recipe; serving uses TransformerDecoder with the same weights layout).
  def _InitBeamSearchStateCallback(self, theta, encoder_outputs, num_hyps_per_beam):
    p = self.params
    src_b = encoder_outputs.encoded.shape[1]
    n = src_b * num_hyps_per_beam
    dev = encoder_outputs.encoded.device
    encoder_outputs.enc_tiled = encoder_outputs.encoded.repeat(1, num_hyps_per_beam, 1)
    encoder_outputs.pad_tiled = encoder_outputs.padding.repeat(1, num_hyps_per_beam)
    init = NestedMap(log_probs=torch.zeros(n, p.softmax.num_classes, device=dev),
                     atten_probs=torch.zeros(n, encoder_outputs.encoded.shape[0], device=dev))
    prefix = torch.zeros(n, p.target_seq_len, dtype=torch.int64, device=dev)
    return init, NestedMap(prefix=prefix)

  def _PreBeamSearchStepCallback(self, theta, encoder_outputs, step_ids, states,
                                 num_hyps_per_beam, cur_step):
    prefix = states.prefix.clone()
    prefix[:, cur_step] = step_ids.squeeze(1)
    ids = prefix[:, :cur_step + 1]
    enc = NestedMap(encoded=encoder_outputs.enc_tiled, padding=encoder_outputs.pad_tiled)
    pred = self.ComputePredictions(
        theta, enc, NestedMap(ids=ids, paddings=torch.zeros_like(ids, dtype=torch.float32)))
    logits = self.softmax.Logits(theta.softmax, pred.softmax_input[-1])
    return (NestedMap(log_probs=torch.log_softmax(logits.float(), -1),
                      atten_probs=pred.attention.probs[:, -1]),
            NestedMap(prefix=prefix))


class InsertionDecoder(base_decoder.BaseBeamSearchDecoder):
  """Insertion Transformer / KERMIT decoder (ref :2179): a bidirectional Transformer over
  the current canvas whose softmax at slot *i* scores the token to insert after canvas
  position *i*. The token embedding holds 2× the vocabulary so that source-side canvas
  tokens (offset by `softmax.num_classes`) are distinguishable from target-side ones."""

  @classmethod
  def Params(cls):
    from lingvo_b200.core import layers_with_attention as lwa
    p = super().Params()
    p.Define('token_emb', layers.SimpleEmbeddingLayer.Params(), 'Token embedding (2·V rows).')
    p.Define('position_emb', layers.PositionalEmbeddingLayer.Params(), 'Positions.')
    p.Define('model_dim', 1024, 'Model dim.')
    p.Define('num_trans_layers', 6, 'Layers.')
    p.Define('trans_tpl', lwa.TransformerLayer.Params(), 'Layer template.')
    p.Define('softmax', layers.SimpleFullSoftmax.Params(), 'Softmax.')
    p.Define('input_dropout_prob', 0.0, 'Input dropout.')
    p.token_emb.vocab_size = 32000 * 2
    p.trans_tpl.tr_atten_tpl.num_attention_heads = 8
    p.trans_tpl.tr_fflayer_tpl.hidden_dim = 4096
    p.softmax.num_classes = 32000
    p.target_seq_len = 300
    return p

  @classmethod
  def UpdateTargetVocabSize(cls, p, vocab_size, wpm_model=None):
    p.softmax.num_classes = vocab_size
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    assert p.token_emb.vocab_size % p.softmax.num_classes == 0
    self.CreateChild('token_emb', p.token_emb.Copy().Set(embedding_dim=p.model_dim))
    self.CreateChild('position_emb', p.position_emb.Copy().Set(embedding_dim=p.model_dim))
    self.CreateChild('input_dropout', layers.DropoutLayer.Params().Set(
        keep_prob=1.0 - p.input_dropout_prob))
    self.CreateChildren('trans', [
        p.trans_tpl.Copy().Set(name='trans_layer_%d' % i, source_dim=p.model_dim,
                               packed_input=p.packed_input, has_aux_atten=False,
                               mask_self_atten=False)
        for i in range(p.num_trans_layers)])
    self.CreateChild('softmax', p.softmax.Copy().Set(input_dim=p.model_dim))

  def ComputePredictions(self, theta, encoder_outputs, targets):
    """targets.ids/paddings `[B, C]` = the canvas → outputs `[B, C, D]`."""
    assert encoder_outputs is None
    p = self.params
    ids = targets.ids.long()
    x = self.token_emb.EmbLookup(theta.token_emb, ids) * (p.model_dim ** 0.5)
    x = x + self.position_emb.FProp(theta.position_emb, ids.shape[1]).unsqueeze(0).to(x.dtype)
    x = self.input_dropout.FProp(theta.input_dropout, x).transpose(0, 1)
    pad = targets.paddings.float().t()
    for i, layer in enumerate(self.trans):
      x, _ = layer.FProp(theta.trans[i], x, pad)
    return NestedMap(outputs=x.transpose(0, 1))

  def ComputeLoss(self, theta, predictions, targets=None):
    """−Σ w · log p(token | slot) over `predictions.tgt.target_indices [N,3]` =
    (batch, slot, token) with weights `target_weights [N]`."""
    out = predictions.outputs
    b, c, d = out.shape
    logits = self.softmax.Logits(theta.softmax, out.reshape(b * c, d)).reshape(b, c, -1)
    logp = torch.log_softmax(logits.float(), -1)
    idx = predictions.tgt.target_indices.long()
    picked = logp[idx[:, 0], idx[:, 1], idx[:, 2]]
    loss = -(picked * predictions.tgt.target_weights.float()).sum() / float(b)
    return ({'loss': (loss, torch.tensor(float(b), device=out.device))},
            {'log_probs': logp, 'logits': logits})
