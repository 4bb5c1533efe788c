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
from lingvo_b200.core import summary_utils
from lingvo_b200.core.nested_map import NestedMap


class MTBaseDecoder(base_decoder.BaseBeamSearchDecoder):
  """Softmax, label smoothing and loss / metric assembly shared by the MT decoders
  (ref :34)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('label_smoothing', None, 'Label smoother params.')
    p.Define('softmax', layers.SimpleFullSoftmax.Params(), 'Softmax.')
    p.Define('per_word_avg_loss', False, 'Average the loss per word (else per sentence).')
    p.Define('unidi_rnn_type', 'func', 'func: FRNN (the only flavour here).')
    p.Define('feed_attention_context_vec_to_softmax', False,
             'Concatenate the attention context to the RNN output before the softmax.')
    p.Define('per_example_tensors', False, 'Return per-example tensors.')
    p.Define('token_normalized_per_seq_loss', False, 'Deprecated; unused.')
    p.softmax.num_classes = 32000
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    if p.label_smoothing is not None:
      self.CreateChild('smoother', p.label_smoothing.Copy().Set(
          name='smoother', num_classes=p.softmax.num_classes))

  @classmethod
  def UpdateTargetVocabSize(cls, p, vocab_size, wpm_model=None):
    p.softmax.num_classes = vocab_size
    return p

  # -- loss ------------------------------------------------------------------------------------
  def _ComputeXentLoss(self, theta, softmax_input, target_labels, target_weights,
                       target_paddings, target_segment_ids=None, time_axis=0):
    """softmax_input `[T, B, D]` (time_axis 0) or `[B, T, D]` (1) → the softmax layer's
    output NestedMap for the flattened tokens (ref :81)."""
    del target_segment_ids
    p = self.params
    flat = softmax_input.reshape(-1, softmax_input.shape[-1])
    kwargs = dict(class_ids=target_labels.reshape(-1, 1).long())
    if p.label_smoothing is not None:
      if time_axis == 0:
        probs = self.smoother.FProp(theta.smoother, target_paddings.t(),
                                    target_labels.t().long(), target_ids=None).transpose(0, 1)
      else:
        probs = self.smoother.FProp(theta.smoother, target_paddings, target_labels.long(),
                                    target_ids=None)
      kwargs = dict(class_probabilities=probs.reshape(-1, p.softmax.num_classes))
    return self.softmax.FProp(theta.softmax, flat, target_weights.float().reshape(-1, 1),
                              **kwargs)

  def _ComputeSoftmaxMetrics(self, xent_loss, target_labels, target_weights,
                             target_segment_ids=None, time_axis=0):
    """→ (metrics, per-example tensors) (ref :134). Per-sequence averaging divides by the
    number of *sentences* — for packed inputs Σ_rows max(segment_id), not the row count."""
    p = self.params
    w = target_weights.float()
    per_tok = xent_loss.per_example_xent.reshape(w.shape)
    num_words = w.sum().clamp_min(1e-8)
    per_seq = (per_tok * w).sum(time_axis)
    if p.per_word_avg_loss:
      loss, loss_w = xent_loss.total_xent / num_words, num_words
    else:
      if p.packed_input:
        if target_segment_ids is None:
          raise AssertionError('Need target segment ids for normalizing loss when training '
                               'with packed inputs.')
        num_samples = target_segment_ids.max(time_axis).values.sum().to(per_seq.dtype)
        loss = per_seq.sum() / num_samples
      else:
        loss = per_seq.mean()
      loss_w = torch.tensor(float(per_seq.shape[0]), device=w.device)
    metrics = NestedMap(loss=(loss, loss_w),
                        log_pplx=(xent_loss.total_xent / num_words, num_words),
                        num_predictions=(num_words, 1.0))
    per_example = NestedMap(per_sequence_xent=per_seq)
    if p.per_example_tensors:
      per_example.per_example_loss = per_tok
      per_example.per_sequence_loss = per_seq
      per_example.loss = per_seq
      if xent_loss.get('logits') is not None:
        per_example.logits = xent_loss.logits.reshape(tuple(w.shape) + (-1,))
        per_example.log_probs = torch.log_softmax(per_example.logits.float(), -1)
    argmax = xent_loss.get('per_example_argmax')
    if argmax is not None:
      correct = ((argmax.reshape(w.shape) == target_labels).float() * w).sum()
      metrics.fraction_of_correct_next_step_preds = (correct / num_words, num_words)
    return metrics, per_example

  def _FPropSoftmax(self, theta, softmax_input, target_labels, target_weights,
                    target_paddings, target_segment_ids=None, time_axis=0):
    xent = self._ComputeXentLoss(theta, softmax_input, target_labels, target_weights,
                                 target_paddings, target_segment_ids, time_axis)
    return self._ComputeSoftmaxMetrics(xent, target_labels, target_weights,
                                       target_segment_ids, time_axis)

  def ComputeLoss(self, theta, predictions, targets):
    seg = targets.segment_ids.t() if self.params.packed_input else None
    if isinstance(predictions, NestedMap):
      predictions = predictions.softmax_input
    return self._FPropSoftmax(theta, predictions, targets.labels.t(), targets.weights.t(),
                              targets.paddings.t(), seg)

  # -- helpers ---------------------------------------------------------------------------------
  def _TruncateTargetSequence(self, targets):
    """Cuts `[batch, time]` targets to the longest real sequence of the batch (ref :283)."""
    targets = targets.Pack(targets.Flatten())
    max_len = int(torch.round((1.0 - targets.paddings.float()).sum(1).max()))
    summary_utils.scalar('max_seq_length', max_len)
    assert bool((targets.paddings[:, max_len:] > 0.5).all())
    py_utils.AssertIdShape([None, None], list(targets.ids.shape), list(targets.labels.shape),
                           list(targets.paddings.shape), list(targets.weights.shape))
    for k in ('ids', 'labels', 'weights', 'paddings'):
      targets[k] = targets[k][:, :max_len]
    return targets

  def _AddAttenProbsSummary(self, source_paddings, targets, atten_probs):
    """atten_probs: list of `[tgt_len, tgt_batch, src_len]` (ref :315)."""
    if not summary_utils._ShouldAddSummary():   # pylint: disable=protected-access
      return
    self._AddAttenProbsImageSummary(source_paddings, targets, atten_probs)
    self._AddAttenProbsHistogramSummary(atten_probs)

  def _AddAttenProbsHistogramSummary(self, atten_probs):
    for i, probs in enumerate(atten_probs):
      summary_utils.histogram('atten{}'.format(i + 1), probs.detach())

  def _AddAttenProbsImageSummary(self, source_paddings, targets, atten_probs):
    summary_utils.AddAttentionSummary(
        'decoder_example', [a.detach() for a in atten_probs], source_paddings.float(),
        targets.paddings.t().float(), max_outputs=1)

  def _ExpandToNumHyps(self, source_enc_len, num_hyps_per_beam):
    """[3, 2, 1] with 2 hyps → [3, 2, 1, 3, 2, 1] (target batch is hyp-major) (ref :381)."""
    return source_enc_len.repeat(num_hyps_per_beam)


class MTDecoderV1(MTBaseDecoder):
  """RNMT decoder (ref :398): attention LSTM + stacked LSTMs with the attention context fed
  to every layer, residuals from `residual_start`, optional clipping-cap schedule,
  embedding / output projections, shared embedding-softmax, sentence-aligned and
  single-token fast beam search."""

  _FLOAT_DTYPE_MAX_SCALER = 0.7

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('emb', layers.SimpleEmbeddingLayer.Params(), 'Embedding.')
    p.Define('source_dim', 1024, 'Encoder output dim.')
    p.Define('attention', attention.AdditiveAttention.Params(), 'Attention.')
    p.Define('atten_rnn_cell_tpl', rnn_cell.LSTMCellSimple.Params(), 'Attention RNN cell.')
    p.Define('emb_projection_tpl', None,
             'ProjectionLayer params: when the embedding dim differs from rnn_cell_dim the '
             'embeddings are projected up (and, with a shared softmax, the output down).')
    p.Define('rnn_cell_tpl', rnn_cell.LSTMCellSimple.Params(), 'Upper RNN cells.')
    p.Define('rnn_cell_dim', 1024, 'RNN cell dim.')
    p.Define('rnn_layers', 8, 'Total decoder RNN layers.')
    p.Define('residual_start', 2, 'First residual layer.')
    p.Define('atten_rnn_cls', rnn_layers.FRNNWithAttention, 'Attention RNN class.')
    p.Define('use_prev_atten_ctx', False,
             'Upper layers get the PREVIOUS step\'s attention context (else the current).')
    p.Define('dropout_prob', 0.0, 'Dropout.')
    p.Define('use_zero_atten_state', False,
             'Zero initial attention context instead of attending with a zero query.')
    p.Define('cc_schedule', None, 'Clipping-cap schedule params (quantization-aware training).')
    p.Define('init_step_ids', False,
             'Beam search starts from the first target id (e.g. a target-language token) '
             'instead of <s>.')
    p.Define('force_alignment', False,
             'Multi-sentence inputs: hypotheses must contain as many sentences as the source '
             '(needs sentence_boundary_token_id and encoder_outputs.num_sentences).')
    p.Define('sentence_boundary_token_id', None, 'Token id separating sentences.')
    p.Define('single_token_fast_decode', False,
             'Sources of length ≤ 1 finish decoding in one step (used for padding inputs).')
    p.Define('zero_token_embs_first_time_step', False,
             'The first step sees a zero embedding instead of emb(<s>).')
    p.Define('use_sigmoid_activation', False, 'log-sigmoid instead of log-softmax scores.')
    p.emb.vocab_size = 32000
    p.attention.hidden_dim = 1024
    p.target_seq_len = 300
    p.beam_search.length_normalization = 0.2
    p.beam_search.coverage_penalty = 0.2
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    if p.force_alignment and p.sentence_boundary_token_id is None:
      raise ValueError('When p.force_alignment is set, must specify '
                       'p.sentence_boundary_token_id.')
    self._share_sm_emb = p.softmax.cls is layers.SharedSoftmaxLayer
    if p.cc_schedule is not None:
      self.CreateChild('cc_schedule', p.cc_schedule)
    else:
      self.cc_schedule = None
    if not self._share_sm_emb:
      assert p.emb.vocab_size == p.softmax.num_classes, (p.emb.vocab_size,
                                                         p.softmax.num_classes)
      self.CreateChild('emb', p.emb)
    emb_dim = ((p.softmax.embedding_dim or p.softmax.input_dim) if self._share_sm_emb
               else p.emb.embedding_dim)
    self._project_emb = bool(p.emb_projection_tpl) and emb_dim != p.rnn_cell_dim
    self._project_out = self._project_emb and self._share_sm_emb
    self.CreateChild('dropout', layers.DropoutLayer.Params().Set(keep_prob=1.0 - p.dropout_prob))
    atten = p.attention.Copy().Set(source_dim=p.source_dim, query_dim=p.rnn_cell_dim)
    if 'packed_input' in atten:
      atten.packed_input = p.packed_input
    ctx_dim = p.source_dim
    if atten.Get('enable_ctx_post_proj') if 'enable_ctx_post_proj' in atten else False:
      ctx_dim = atten.ctx_post_proj_dim
    elif 'context_dim' in atten:
      atten.context_dim = p.source_dim
    self._ctx_dim = ctx_dim
    in_dim = (p.rnn_cell_dim if self._project_emb else emb_dim) + ctx_dim
    cell = p.atten_rnn_cell_tpl.Copy().Set(
        name='atten_rnn', num_input_nodes=in_dim, num_output_nodes=p.rnn_cell_dim,
        reset_cell_state=p.packed_input)
    self.CreateChild('frnn_with_atten', p.atten_rnn_cls.Params().Set(
        cell=cell, attention=atten, output_prev_atten_ctx=p.use_prev_atten_ctx,
        use_zero_atten_state=p.use_zero_atten_state, atten_context_dim=ctx_dim,
        packed_input=p.packed_input))
    rnns = []
    for i in range(1, p.rnn_layers):
      rnns.append(rnn_layers.FRNN.Params().Set(
          name='frnn_%d' % i, packed_input=p.packed_input,
          cell=p.rnn_cell_tpl.Copy().Set(num_input_nodes=p.rnn_cell_dim + ctx_dim,
                                         num_output_nodes=p.rnn_cell_dim,
                                         reset_cell_state=p.packed_input)))
    self.CreateChildren('frnn', rnns)
    if p.feed_attention_context_vec_to_softmax:
      assert not (self._share_sm_emb or self._project_emb or self._project_out)
      sm_in = p.rnn_cell_dim + ctx_dim
    else:
      sm_in = p.rnn_cell_dim
    if self._project_emb:
      self._CreateProjection(p.emb_projection_tpl, 'emb_proj', emb_dim, p.rnn_cell_dim)
    if self._project_out:
      sm_in = emb_dim
      self._CreateProjection(p.emb_projection_tpl, 'out_proj', p.rnn_cell_dim, emb_dim)
    sm = p.softmax.Copy()
    if not (self._share_sm_emb and sm.embedding_dim):
      sm.input_dim = sm_in
    self.CreateChild('softmax', sm)

  def _CreateProjection(self, proj_tpl, name, input_dim, output_dim):
    assert proj_tpl.cls is layers.ProjectionLayer
    self.CreateChild(name, proj_tpl.Copy().Set(name=name, input_dim=input_dim,
                                               output_dim=output_dim))

  # -- small pieces ------------------------------------------------------------------------------
  def ApplyDropout(self, x_in):
    p = self.params
    assert 0 <= p.dropout_prob < 1.0
    if self.do_eval or p.dropout_prob == 0.0:
      return x_in
    return torch.nn.functional.dropout(x_in, p.dropout_prob, training=True)

  def ApplyClipping(self, theta, x):
    if self.cc_schedule is not None:
      return self.cc_schedule.ApplyClipping(theta.cc_schedule, x)
    return x

  def _ZeroOutFirstTimeStep(self, token_embs, batch=None, time=None):
    """`[time, batch, dim]` embeddings with step 0 zeroed (ref :657)."""
    del batch, time
    mask = torch.ones(token_embs.shape[0], 1, 1, device=token_embs.device,
                      dtype=token_embs.dtype)
    mask[0] = 0.0
    return token_embs * mask

  def _EmbLookup(self, theta, ids):
    if self._share_sm_emb:
      return self.softmax.EmbLookup(theta.softmax, ids)
    return self.emb.EmbLookup(theta.emb, ids)

  def AddExtraDecodingInfo(self, encoder_outputs, targets):
    if self.params.init_step_ids:
      encoder_outputs['init_step_ids'] = targets.ids[:, 0]
    return encoder_outputs

  # -- training ----------------------------------------------------------------------------------
  def _Upper(self, theta, xs, ctx, pad, states=None, step=False, segment_id=None):
    p = self.params
    new_states = []
    for i, r in enumerate(self.frnn):
      inp = torch.cat([xs, ctx], -1)
      if step:
        st, _ = r.cell.FProp(theta.frnn[i].cell, states[i], NestedMap(
            act=[inp], padding=pad, reset_mask=torch.ones_like(pad)))
        ys = r.cell.GetOutput(st)
        new_states.append(st)
      else:
        kw = {'segment_id': segment_id} if (segment_id is not None and p.packed_input) else {}
        ys, _ = r.FProp(theta.frnn[i], inp, pad, **kw)
        ys = self.ApplyDropout(ys)
      if i + 1 >= p.residual_start:
        xs = self.ApplyClipping(theta, xs + ys)
      else:
        xs = ys
      if not step:
        summary_utils.histogram('layer_out_%s' % i, xs.detach())
    return xs, new_states

  def ComputePredictions(self, theta, encoder_outputs, targets):
    """→ NestedMap(softmax_input `[T, B, D]`, attention.probs `[B, T, S]`, source_enc_len)."""
    p = self.params
    ids = targets.ids.t().long()
    pad = targets.paddings.t().float().unsqueeze(-1)
    seg = targets.segment_ids.t().float().unsqueeze(-1) if p.packed_input else None
    emb = self._EmbLookup(theta, ids)
    if p.zero_token_embs_first_time_step:
      emb = self._ZeroOutFirstTimeStep(emb)
    emb = self.ApplyClipping(theta, emb)
    summary_utils.histogram('input_emb', emb.detach())
    emb = self.ApplyDropout(emb)
    if self._project_emb:
      emb = self.ApplyClipping(theta, self.emb_proj.FProp(theta.emb_proj, emb))
    kw = {}
    if p.packed_input:
      kw = dict(src_segment_id=encoder_outputs.get('segment_id'), segment_id=seg)
    ctx, xs, probs, _ = self.frnn_with_atten.FProp(
        theta.frnn_with_atten, encoder_outputs.encoded, encoder_outputs.padding, emb, pad, **kw)
    self._AddAttenProbsSummary(encoder_outputs.padding, targets, [probs])
    ctx = self.ApplyClipping(theta, ctx)
    summary_utils.histogram('atten_ctxs', ctx.detach())
    xs, _ = self._Upper(theta, xs, ctx, pad, segment_id=seg)
    if p.feed_attention_context_vec_to_softmax:
      xs = torch.cat([xs, ctx], -1)
    if self._project_out:
      xs = self.ApplyClipping(theta, self.out_proj.FProp(theta.out_proj, xs))
    return NestedMap(softmax_input=xs, attention=NestedMap(probs=probs.transpose(0, 1)),
                     source_enc_len=(1.0 - encoder_outputs.padding.float()).sum(0))

  # -- beam search -----------------------------------------------------------------------------------
  def _InitDecoder(self, theta, encoder_outputs, num_hyps):
    """→ (rnn_states, atten_context, atten_probs, atten_states); stores `packed_src`."""
    fa = self.frnn_with_atten
    packed = fa.InitForSourcePacked(theta.frnn_with_atten, encoder_outputs.encoded,
                                    encoder_outputs.padding)
    encoder_outputs.packed_src = packed
    st = fa.zero_state(theta.frnn_with_atten, encoder_outputs.encoded, packed, num_hyps)
    rnn_states = [st.rnn] + [r.zero_state(theta.frnn[i], num_hyps)
                             for i, r in enumerate(self.frnn)]
    return rnn_states, st.atten, st.atten_probs, st.atten_state

  def _DecodeStep(self, theta, encoder_outputs, embs, step_paddings, prev_atten_context,
                  rnn_states, prev_atten_states):
    """One step → (cur_atten_context, atten_probs, new_rnn_states, step_out, atten_states)."""
    p = self.params
    if self._project_emb:
      embs = self.ApplyClipping(theta, self.emb_proj.FProp(theta.emb_proj, embs))
    fa = self.frnn_with_atten
    s0, _ = fa.cell.FProp(theta.frnn_with_atten.cell, rnn_states[0], NestedMap(
        act=[torch.cat([embs, prev_atten_context.to(embs.dtype)], 1)], padding=step_paddings,
        reset_mask=torch.ones_like(step_paddings)))
    rnn_out = fa.cell.GetOutput(s0)
    cur_ctx, probs, atten_states = fa.atten.ComputeContextVectorWithSource(
        theta.frnn_with_atten.atten, encoder_outputs.packed_src, rnn_out,
        attention_state=prev_atten_states)
    ctx = prev_atten_context if p.use_prev_atten_ctx else cur_ctx
    xs, upper = self._Upper(theta, rnn_out, ctx.to(rnn_out.dtype), step_paddings,
                            rnn_states[1:], step=True)
    step_out = torch.cat([xs, ctx.to(xs.dtype)], 1) if p.feed_attention_context_vec_to_softmax \
        else xs
    if self._project_out:
      step_out = self.ApplyClipping(theta, self.out_proj.FProp(theta.out_proj, step_out))
    return cur_ctx, probs, [s0] + upper, step_out, atten_states

  def _InitBeamSearchStateCallback(self, theta, encoder_outputs, num_hyps_per_beam):
    p = self.params
    num_beams = encoder_outputs.padding.shape[1]
    n = num_beams * num_hyps_per_beam
    rnn_states, ctx, probs, atten_states = self._InitDecoder(theta, encoder_outputs, n)
    dev = encoder_outputs.encoded.device
    init = NestedMap(log_probs=torch.zeros(n, p.softmax.num_classes, device=dev),
                     atten_probs=probs)
    if p.init_step_ids and 'init_step_ids' in encoder_outputs:
      init.step_ids = self._ExpandToNumHyps(encoder_outputs.init_step_ids,
                                            num_hyps_per_beam).unsqueeze(1)
    states = NestedMap(time_step=torch.zeros((), dtype=torch.int64, device=dev),
                       rnn_states=rnn_states, atten_context=ctx, atten_probs=probs,
                       atten_states=atten_states)
    if p.force_alignment:
      states.num_sentences = torch.ones(n, dtype=torch.int32, device=dev)
    return init, states

  def _PreBeamSearchStepCallback(self, theta, encoder_outputs, step_ids, states,
                                 num_hyps_per_beam, cur_step):
    p = self.params
    n = step_ids.shape[0]
    embs = self._EmbLookup(theta, step_ids.reshape(-1).long())
    if p.zero_token_embs_first_time_step and int(cur_step) == 0:
      embs = torch.zeros_like(embs)
    embs = self.ApplyClipping(theta, embs)
    pad = torch.zeros(n, 1, device=embs.device, dtype=embs.dtype)
    ctx, probs, rnn_states, step_out, atten_states = self._DecodeStep(
        theta, encoder_outputs, embs, pad, states.atten_context, states.rnn_states,
        states.atten_states)
    probs = probs.reshape(states.atten_probs.shape)
    logits = self.softmax.Logits(theta.softmax, step_out).float()
    log_probs = torch.nn.functional.logsigmoid(logits) if p.use_sigmoid_activation \
        else torch.log_softmax(logits, -1)
    if p.force_alignment:
      if 'num_sentences' not in encoder_outputs:
        raise ValueError('Model does not support p.force_alignment as key "num_sentences" '
                         'is missing from encoder_outputs.')
      log_probs = self._ForceAlignment(
          log_probs, encoder_outputs['num_sentences'].repeat(num_hyps_per_beam),
          states.num_sentences)
    if p.single_token_fast_decode:
      single = (1.0 - encoder_outputs.padding.float()).sum(0) <= 1.0
      if bool(single.any()):
        log_probs = self._UpdateLogitsForSingleTokenFastDecode(log_probs, single,
                                                               num_hyps_per_beam)
    results = NestedMap(log_probs=log_probs,
                        atten_probs=states.atten_probs if p.use_prev_atten_ctx else probs)
    new_states = NestedMap(time_step=states.time_step + 1, rnn_states=rnn_states,
                           atten_context=ctx, atten_probs=probs, atten_states=atten_states)
    if p.force_alignment:
      new_states.num_sentences = states.num_sentences
    return results, new_states

  def _PostBeamSearchStepCallback(self, theta, encoder_outputs, new_step_ids, states):
    p = self.params
    if p.force_alignment:
      add = (new_step_ids.reshape(-1) == p.sentence_boundary_token_id)
      states.num_sentences = states.num_sentences + add.to(states.num_sentences.dtype)
    return states

  def _ForceAlignment(self, log_probs, source_num_sentences, hyp_num_sentences):
    """EOS is forbidden while the hypothesis has fewer sentences than the source; the
    sentence-boundary token once it has as many (ref :1144)."""
    p = self.params
    neg = -self._FLOAT_DTYPE_MAX_SCALER * torch.finfo(log_probs.dtype).max
    out = log_probs.clone()
    eos, boundary = p.target_eos_id, p.sentence_boundary_token_id
    out[:, eos] = torch.where(source_num_sentences > hyp_num_sentences,
                              torch.full_like(out[:, eos], neg), log_probs[:, eos])
    out[:, boundary] = torch.where(source_num_sentences <= hyp_num_sentences,
                                   torch.full_like(out[:, boundary], neg),
                                   log_probs[:, boundary])
    return out

  def _UpdateLogitsForSingleTokenFastDecode(self, log_probs, is_single_token,
                                            num_hyps_per_beam):
    """Rows of single-token sources: all mass on EOS (ref :1189)."""
    neg = -self._FLOAT_DTYPE_MAX_SCALER * torch.finfo(log_probs.dtype).max
    forced = torch.full_like(log_probs, neg)
    forced[:, self.params.target_eos_id] = 0.0
    rows = is_single_token.repeat(num_hyps_per_beam).unsqueeze(1)
    return torch.where(rows, forced, log_probs)


class TransformerDecoder(MTBaseDecoder):
  """Transformer decoder (ref :1219): token + position (+ task) embeddings, masked
  self-attention / cross-attention stack (fused attention kernels, pre-allocated KV cache for
  `ExtendStep`), optional transparent per-layer source encodings."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('token_emb', layers.SimpleEmbeddingLayer.Params(), 'Token embedding.')
    p.Define('shared_emb', None, 'Shared embedding-softmax params (replaces token_emb and '
             'softmax).')
    p.Define('position_emb', layers.PositionalEmbeddingLayer.Params(), 'Positions.')
    p.Define('source_dim', 512, 'Encoder dim.')
    p.Define('model_dim', 512, 'Model dim.')
    p.Define('num_trans_layers', 6, 'Layers.')
    p.Define('trans_tpl', bma.TransformerDecoderLayer.Params(), 'Layer template.')
    p.Define('input_dropout_prob', 0.0, 'Input dropout.')
    p.Define('is_transparent', False,
             'Source encodings are `[time, batch, source_dim, num_trans_layers]`: layer i '
             'attends to slice i (transparent encoder).')
    p.Define('add_multiheaded_attention_scalar_summary', False,
             'Scalar summaries of the attention entropy per layer.')
    p.Define('ln_tpl', layers.LayerNorm.Params(), 'Layer norm template.')
    p.Define('ln_output', False, 'Layer-normalise the decoder output (alias of '
             'final_layer_norm).')
    p.Define('final_layer_norm', True, 'LN before the softmax.')
    p.Define('task_emb', None, 'Task embedding params: added to every target position.')
    p.Define('init_step_ids', False, 'Beam search starts from targets.ids[:, 0].')
    p.Define('use_lang_dependent_atten', False, 'Unsupported: one cross-attention for all '
             'languages.')
    p.Define('zero_token_embs_first_time_step', False,
             'The first step sees a zero token embedding.')
    p.Define('ln_input', None, 'LayerNorm params applied to the embedded inputs.')
    p.Define('hidden_dim', 2048, 'FFN hidden dim.')
    p.Define('num_atten_heads', 8, 'Heads.')
    p.Define('residual_dropout_prob', 0.0, 'Residual dropout.')
    p.softmax.num_classes = 32000
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    assert not p.use_lang_dependent_atten, 'language-dependent attention is not supported'
    self._share_sm_emb = p.shared_emb is not None
    if self._share_sm_emb:
      self.CreateChild('softmax', p.shared_emb.Copy().Set(name='softmax'))
    else:
      self.CreateChild('token_emb', p.token_emb.Copy().Set(embedding_dim=p.model_dim))
      self.CreateChild('softmax', p.softmax.Copy().Set(input_dim=p.model_dim))
    self.CreateChild('position_emb', p.position_emb.Copy().Set(embedding_dim=p.model_dim))
    if p.task_emb is not None:
      self.CreateChild('task_emb', p.task_emb.Copy().Set(embedding_dim=p.model_dim))
    if p.ln_input is not None:
      self.CreateChild('layer_norm_input', p.ln_input.Copy().Set(
          name='decoder_ln_input', input_dim=p.model_dim))
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
        final_layer_norm=p.final_layer_norm or p.ln_output, transformer_layer_params_tpl=tpl))

  def _TokenEmb(self, theta, ids):
    if self._share_sm_emb:
      return self.softmax.EmbLookup(theta.softmax, ids.long())
    return self.token_emb.EmbLookup(theta.token_emb, ids.long())

  def _ZeroOutFirstTimeStep(self, token_embs, batch=None, target_time=None):
    """`[batch, time, dim]` embeddings with step 0 zeroed (ref :1433)."""
    del batch, target_time
    mask = torch.ones(1, token_embs.shape[1], 1, device=token_embs.device,
                      dtype=token_embs.dtype)
    mask[:, 0] = 0.0
    return token_embs * mask

  def _Embed(self, theta, ids, t0=None, task_ids=None, segment_pos=None):
    """ids `[B, T]` → `[B, T, D]`; `t0`: position of the first column (incremental decode)."""
    p = self.params
    x = self._TokenEmb(theta, ids) * (p.model_dim ** 0.5)
    if p.zero_token_embs_first_time_step:
      if t0 is None:
        x = self._ZeroOutFirstTimeStep(x)
      elif int(t0) == 0:
        x = torch.zeros_like(x)
    t = ids.shape[1]
    if p.packed_input and segment_pos is not None:
      pos = self.position_emb.FPropWithPosition(theta.position_emb, segment_pos)
    elif t0 is None:
      pos = self.position_emb.FProp(theta.position_emb, t).unsqueeze(0)
    else:
      pos = self.position_emb.FProp(theta.position_emb, t0 + t)[t0:t0 + t].unsqueeze(0)
    x = x + pos.to(x.dtype)
    if p.task_emb is not None and task_ids is not None:
      x = x + self.task_emb.EmbLookup(theta.task_emb, task_ids.long())
    if p.ln_input is not None:
      x = self.layer_norm_input.FProp(theta.layer_norm_input, x)
    return x

  def _Sources(self, encoder_outputs):
    """→ (aux `[B, S, D]` or a per-layer list for transparent encoders, paddings `[B, S]`)."""
    enc = encoder_outputs.encoded
    if self.params.is_transparent:
      assert enc.dim() == 4 and enc.shape[-1] == self.params.num_trans_layers, enc.shape
      aux = [enc[..., i].transpose(0, 1) for i in range(enc.shape[-1])]
    else:
      aux = enc.transpose(0, 1)
    return aux, encoder_outputs.padding.t()

  def AddExtraDecodingInfo(self, encoder_outputs, targets):
    p = self.params
    if p.task_emb is not None:
      encoder_outputs['target_task_ids'] = targets.task_ids[:, 0]
    if p.init_step_ids:
      encoder_outputs['init_step_ids'] = targets.ids[:, 0]
    return encoder_outputs

  def _FProp(self, theta, encoder_outputs, targets):
    p = self.params
    x = self._Embed(theta, targets.ids, task_ids=targets.get('task_ids'),
                    segment_pos=targets.get('segment_pos'))
    x = self.input_dropout.FProp(theta.input_dropout, x)
    aux, aux_pad = self._Sources(encoder_outputs)
    seg_mask = aux_seg_mask = None
    if p.packed_input:
      seg_mask = bma.SegmentMask(targets.segment_ids, targets.segment_ids, dtype=x.dtype)
      aux_seg_mask = bma.SegmentMask(targets.segment_ids, encoder_outputs.segment_id.t(),
                                     dtype=x.dtype)
    out, _ = self.stack.FProp(theta.stack, x, targets.paddings.float(), aux, aux_pad,
                              segment_mask=seg_mask, aux_segment_mask=aux_seg_mask)
    return NestedMap(softmax_input=out.transpose(0, 1),
                     source_enc_len=(1.0 - encoder_outputs.padding.float()).sum(0))

  def ComputePredictions(self, theta, encoder_outputs, targets):
    return self._FProp(theta, encoder_outputs, targets)

  def ExtendStep(self, theta, encoder_outputs, new_ids, t, prefix_states):
    """One incremental step for ids `[B]` at position `t` (ref :1629) → (softmax input
    `[B, D]`, new prefix states). `prefix_states` comes from `InitPrefixStates`."""
    x = self._Embed(theta, new_ids.reshape(-1, 1), t0=int(t),
                    task_ids=encoder_outputs.get('target_task_ids'))
    aux, aux_pad = self._Sources(encoder_outputs)
    out, new_states = self.stack.ExtendStep(theta.stack, x, aux, aux_pad, prefix_states, int(t))
    return out.squeeze(1), new_states

  def InitPrefixStates(self, theta, batch, max_len=None):
    return self.stack.InitStates(theta.stack, batch, max_len or self.params.target_seq_len)

  def _InitBeamSearchStateCallback(self, theta, encoder_outputs, num_hyps_per_beam):
    p = self.params
    src_b = encoder_outputs.encoded.shape[1]
    n = src_b * num_hyps_per_beam
    dev = encoder_outputs.encoded.device
    t_max = p.target_seq_len
    aux, aux_pad = self._Sources(encoder_outputs)
    # hyp index = hyp_id * src_b + beam → tile sources along dim 0
    tile = lambda a: a.repeat(num_hyps_per_beam, 1, 1)
    encoder_outputs.aux_tiled = [tile(a) for a in aux] if isinstance(aux, list) else tile(aux)
    encoder_outputs.aux_pad_tiled = aux_pad.repeat(num_hyps_per_beam, 1)
    cache = self.stack.InitStates(theta.stack, n, t_max)
    # caches are [T, n, N, H]: make dim 0 the hyp dim for the helper's re-ordering
    cache = cache.Transform(lambda x: x.transpose(0, 1).contiguous())
    s_len = aux_pad.shape[1]
    init = NestedMap(log_probs=torch.zeros(n, p.softmax.num_classes, device=dev),
                     atten_probs=torch.zeros(n, s_len, device=dev))
    if p.init_step_ids and 'init_step_ids' in encoder_outputs:
      init.step_ids = self._ExpandToNumHyps(encoder_outputs.init_step_ids,
                                            num_hyps_per_beam).unsqueeze(1)
    if p.task_emb is not None and 'target_task_ids' in encoder_outputs:
      encoder_outputs.task_ids_tiled = self._ExpandToNumHyps(
          encoder_outputs.target_task_ids, num_hyps_per_beam).unsqueeze(1)
    return init, NestedMap(cache=cache, time_step=torch.zeros(n, dtype=torch.int64, device=dev))

  def _PreBeamSearchStepCallback(self, theta, encoder_outputs, step_ids, states,
                                 num_hyps_per_beam, cur_step):
    x = self._Embed(theta, step_ids, t0=cur_step,
                    task_ids=encoder_outputs.get('task_ids_tiled'))
    cache = states.cache.Transform(lambda c: c.transpose(0, 1))
    out, new_cache = self.stack.ExtendStep(
        theta.stack, x, encoder_outputs.aux_tiled, encoder_outputs.aux_pad_tiled, cache,
        cur_step)
    logits = self.softmax.Logits(theta.softmax, out.squeeze(1))
    n = step_ids.shape[0]
    new_cache = new_cache.Transform(lambda c: c.transpose(0, 1).contiguous())
    s_len = encoder_outputs.aux_pad_tiled.shape[1]
    atten = torch.full((n, s_len), 1.0 / s_len, device=logits.device)
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
