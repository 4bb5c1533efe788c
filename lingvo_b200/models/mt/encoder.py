"""MT encoders (ref `lingvo/tasks/mt/encoder.py`).

All encoders map `input_batch(ids [B,T], paddings [B,T])` to
`NestedMap(encoded [T,B,D], padding [T,B], segment_id [T,B] | None)`.
"""

from __future__ import annotations

import torch

from lingvo_b200.core import base_layer
from lingvo_b200.core import batch_major_attention as bma
from lingvo_b200.core import layers
from lingvo_b200.core import rnn_cell
from lingvo_b200.core import rnn_layers
from lingvo_b200.core.nested_map import NestedMap


class MTEncoderBiRNN(base_layer.BaseLayer):
  """Embedding + stacked bidirectional LSTMs with residuals (RNMT+ encoder, ref :343)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('emb', layers.SimpleEmbeddingLayer.Params(), 'Embedding.')
    p.Define('lstm_tpl', rnn_cell.LayerNormalizedLSTMCellSimple.Params(), 'Cell tpl.')
    p.Define('proj_tpl', layers.ProjectionLayer.Params(), 'Final projection tpl.')
    p.Define('lstm_cell_size', 512, 'Cell size per direction.')
    p.Define('num_lstm_layers', 6, 'Bi-LSTM layers.')
    p.Define('dropout_prob', 0.0, 'Dropout.')
    p.Define('residual_start', 2, 'First layer with a residual connection.')
    p.Define('encoder_out_dim', 1024, 'Output dim.')
    p.Define('bidi_rnn_type', 'func', 'Kept for parity.')
    p.Define('packed_input', False, 'Packed inputs.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    self.CreateChild('emb', p.emb)
    self.CreateChild('dropout', layers.DropoutLayer.Params().Set(
        keep_prob=1.0 - p.dropout_prob))
    rnns = []
    for i in range(p.num_lstm_layers):
      idim = p.emb.embedding_dim if i == 0 else 2 * p.lstm_cell_size
      cell = p.lstm_tpl.Copy().Set(num_input_nodes=idim, num_output_nodes=p.lstm_cell_size)
      rnns.append(rnn_layers.BidirectionalFRNN.Params().Set(
          name='bidi_%d' % i, fwd=cell.Copy(), bak=cell.Copy(), packed_input=p.packed_input))
    self.CreateChildren('rnn', rnns)
    self.CreateChild('final_proj', p.proj_tpl.Copy().Set(
        input_dim=2 * p.lstm_cell_size, output_dim=p.encoder_out_dim, activation='NONE',
        batch_norm=False, has_bias=True))

  def _ComputeInputs(self, theta, ids_tm, input_batch):
    return self.emb.EmbLookup(theta.emb, ids_tm.long())

  def FProp(self, theta, input_batch):
    p = self.params
    ids = input_batch.ids.t()
    pad = input_batch.paddings.t().float().unsqueeze(-1)
    seg = input_batch.segment_ids.t().unsqueeze(-1) if p.packed_input else None
    xs = self.dropout.FProp(theta.dropout, self._ComputeInputs(theta, ids, input_batch))
    for i, r in enumerate(self.rnn):
      ys = r.FProp(theta.rnn[i], xs, pad, segment_id=seg)
      ys = self.dropout.FProp(theta.dropout, ys)
      xs = xs + ys if (i >= p.residual_start and xs.shape == ys.shape) else ys
    out = self.final_proj.FProp(theta.final_proj, xs)
    out = out * (1.0 - pad)
    return NestedMap(encoded=out, padding=pad.squeeze(-1),
                     segment_id=seg.squeeze(-1) if seg is not None else None)


class MTEncoderV1(base_layer.BaseLayer):
  """GNMT-v1 encoder (ref :33): embedding → one bidirectional LSTM layer → a stack of
  unidirectional LSTM layers with residual connections from the third layer on."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('emb', layers.SimpleEmbeddingLayer.Params(), 'Embedding.')
    p.Define('lstm_tpl', rnn_cell.LSTMCellSimple.Params(), 'Cell tpl.')
    p.Define('lstm_tpl_uni', None, 'Override tpl for the unidirectional layers.')
    p.Define('lstm_tpl_bidi', None, 'Override tpl for the bidirectional layer.')
    p.Define('lstm_cell_size', 1024, 'Cell size.')
    p.Define('num_lstm_layers', 8, 'Total RNN layers (1 bidi + N-1 uni).')
    p.Define('dropout_prob', 0.0, 'Dropout.')
    p.Define('unidi_rnn_type', 'func', 'Kept for parity.')
    p.Define('bidi_rnn_type', 'func', 'Kept for parity.')
    p.Define('cc_schedule', None, 'Clipping-cap schedule (quantised training).')
    p.Define('packed_input', False, 'Packed inputs.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    assert p.num_lstm_layers >= 2
    self.CreateChild('emb', p.emb)
    self.CreateChild('dropout', layers.DropoutLayer.Params().Set(
        keep_prob=1.0 - p.dropout_prob))
    bidi = (p.lstm_tpl_bidi or p.lstm_tpl).Copy().Set(
        num_input_nodes=p.emb.embedding_dim, num_output_nodes=p.lstm_cell_size)
    self.CreateChild('rnn_bidi', rnn_layers.BidirectionalFRNN.Params().Set(
        fwd=bidi.Copy(), bak=bidi.Copy(), packed_input=p.packed_input))
    unis = []
    for i in range(1, p.num_lstm_layers):
      cell = (p.lstm_tpl_uni or p.lstm_tpl).Copy().Set(
          num_input_nodes=2 * p.lstm_cell_size if i == 1 else p.lstm_cell_size,
          num_output_nodes=p.lstm_cell_size)
      unis.append(rnn_layers.FRNN.Params().Set(name='uni_%d' % i, cell=cell,
                                               packed_input=p.packed_input))
    self.CreateChildren('rnn_uni', unis)

  def FProp(self, theta, input_batch):
    p = self.params
    ids = input_batch.ids.t()
    pad = input_batch.paddings.t().float().unsqueeze(-1)
    seg = input_batch.segment_ids.t().unsqueeze(-1) if p.packed_input else None
    xs = self.dropout.FProp(theta.dropout, self.emb.EmbLookup(theta.emb, ids.long()))
    xs = self.rnn_bidi.FProp(theta.rnn_bidi, xs, pad, segment_id=seg)
    for i, r in enumerate(self.rnn_uni):
      ys, _ = r.FProp(theta.rnn_uni[i], self.dropout.FProp(theta.dropout, xs), pad,
                      segment_id=seg)
      xs = xs + ys if (i >= 1 and xs.shape == ys.shape) else ys   # residual from layer 3
    return NestedMap(encoded=xs * (1.0 - pad), padding=pad.squeeze(-1),
                     segment_id=seg.squeeze(-1) if seg is not None else None)


class MTEncoderUniRNN(base_layer.BaseLayer):
  """Stack of unidirectional LSTMs with residuals (ref :201); supports carrying the
  recurrent state across calls (`state0`) and the transparent multi-output mode."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('emb', layers.SimpleEmbeddingLayer.Params(), 'Embedding.')
    p.Define('lstm_tpl', rnn_cell.LSTMCellSimple.Params(), 'Cell tpl.')
    p.Define('lstm_cell_size', 512, 'Cell size.')
    p.Define('num_lstm_layers', 8, 'Layers.')
    p.Define('dropout_prob', 0.0, 'Dropout.')
    p.Define('residual_start', 2, 'First layer with a residual connection.')
    p.Define('unidi_rnn_type', 'func', 'Kept for parity.')
    p.Define('cc_schedule', None, 'Clipping-cap schedule.')
    p.Define('is_transparent', False, 'Emit learned mergers of all layer outputs.')
    p.Define('transparent_merger_tpl',
             layers.WeightedSumLayer.Params().Set(add_weight_summaries=True), 'Merger tpl.')
    p.Define('packed_input', False, 'Packed inputs.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    self.CreateChild('emb', p.emb)
    self.CreateChild('dropout', layers.DropoutLayer.Params().Set(
        keep_prob=1.0 - p.dropout_prob))
    rnns = []
    for i in range(p.num_lstm_layers):
      cell = p.lstm_tpl.Copy().Set(
          num_input_nodes=p.emb.embedding_dim if i == 0 else p.lstm_cell_size,
          num_output_nodes=p.lstm_cell_size)
      rnns.append(rnn_layers.FRNN.Params().Set(name='rnn_%d' % i, cell=cell,
                                               packed_input=p.packed_input))
    self.CreateChildren('rnn', rnns)
    if p.is_transparent:
      self.CreateChild('transparent_merger', p.transparent_merger_tpl.Copy().Set(
          num_sources=p.num_lstm_layers))

  def zero_state(self, theta, batch_size):
    return NestedMap(rnn=[r.zero_state(theta.rnn[i], batch_size)
                          for i, r in enumerate(self.rnn)])

  def FProp(self, theta, input_batch, state0=None):
    p = self.params
    ids = input_batch.ids.t()
    pad = input_batch.paddings.t().float().unsqueeze(-1)
    xs = self.dropout.FProp(theta.dropout, self.emb.EmbLookup(theta.emb, ids.long()))
    outs, states = [], []
    for i, r in enumerate(self.rnn):
      ys, st = r.FProp(theta.rnn[i], xs, pad,
                       state0=state0.rnn[i] if state0 is not None else None)
      ys = self.dropout.FProp(theta.dropout, ys)
      xs = xs + ys if (i >= p.residual_start and xs.shape == ys.shape) else ys
      outs.append(xs)
      states.append(st)
    enc = self.transparent_merger.FProp(theta.transparent_merger, outs) \
        if p.is_transparent else xs
    return NestedMap(encoded=enc * (1.0 - pad), padding=pad.squeeze(-1), segment_id=None,
                     state=NestedMap(rnn=states))


class TransformerEncoder(base_layer.BaseLayer):
  """Token + positional embedding → N Transformer layers → LN (ref :538)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('token_emb', layers.SimpleEmbeddingLayer.Params(), 'Token embedding.')
    p.Define('shared_emb', None, 'Shared embedding/softmax params (optional).')
    p.Define('position_emb', layers.PositionalEmbeddingLayer.Params(), 'Positions.')
    p.Define('model_dim', 512, 'Model dim.')
    p.Define('transformer_stack', bma.StackedTransformerLayers.Params(), 'Stack.')
    p.Define('input_dropout_prob', 0.0, 'Input dropout.')
    p.Define('packed_input', False, 'Packed inputs.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    self.CreateChild('token_emb', p.token_emb.Copy().Set(embedding_dim=p.model_dim))
    self.CreateChild('position_emb', p.position_emb.Copy().Set(embedding_dim=p.model_dim))
    self.CreateChild('input_dropout', layers.DropoutLayer.Params().Set(
        keep_prob=1.0 - p.input_dropout_prob))
    stack = p.transformer_stack.Copy().Set(mdl_dim=p.model_dim, packed_input=p.packed_input,
                                           final_layer_norm=True, mask_self_atten=False)
    tpl = stack.transformer_layer_params_tpl
    for a in (tpl.tr_atten_tpl, tpl.tr_self_atten_tpl):
      if a is not None:
        a.atten_tpl.return_atten_probs = False
    self.CreateChild('transformer_stack', stack)

  def FProp(self, theta, input_batch):
    p = self.params
    ids, pad = input_batch.ids.long(), input_batch.paddings.float()
    t = ids.shape[1]
    x = self.token_emb.EmbLookup(theta.token_emb, ids) * (p.model_dim ** 0.5)
    if p.packed_input:
      pos = self.position_emb.FPropWithPosition(theta.position_emb, input_batch.segment_pos)
      seg_mask = bma.SegmentMask(input_batch.segment_ids, input_batch.segment_ids)
    else:
      pos = self.position_emb.FProp(theta.position_emb, t).unsqueeze(0)
      seg_mask = None
    x = self.input_dropout.FProp(theta.input_dropout, x + pos.to(x.dtype))
    out, _ = self.transformer_stack.FProp(theta.transformer_stack, x, pad,
                                          segment_mask=seg_mask)
    return NestedMap(encoded=out.transpose(0, 1), padding=pad.t(),
                     segment_id=input_batch.segment_ids.t() if p.packed_input else None)


class MTEncoderBiRNNPrecomputedEmbedding(MTEncoderBiRNN):
  """Bi-RNN encoder fed with externally computed embeddings `input_batch.embs [B,T,D]`
  (ref :528)."""

  def _ComputeInputs(self, theta, ids_tm, input_batch):
    return input_batch.embs.transpose(0, 1)


class TransformerBatchMajorEncoder(TransformerEncoder):
  """Batch-major Transformer encoder (ref :836). The stack here is batch-major already
  (`bma.StackedTransformerLayers` over the fused attention kernels); this variant adds the
  reference's knobs: optional final layer norm and an `output_data_format` switch."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('final_layer_norm', False, 'LN on the encoder output.')
    p.Define('output_data_format', 'TBC', "'TBC' (time-major) or 'BTC'.")
    p.Define('input_dropout_tpl', layers.DropoutLayer.Params(), 'Kept for parity.')
    p.Define('use_fused_layernorm', False, 'Kept for parity (LN is always fused).')
    return p

  def __init__(self, params):
    params = params.Copy()
    params.transformer_stack = params.transformer_stack.Copy().Set(
        final_layer_norm=params.final_layer_norm)
    super().__init__(params)

  def FProp(self, theta, input_batch):
    out = super().FProp(theta, input_batch)
    if self.params.output_data_format == 'BTC':
      out.encoded = out.encoded.transpose(0, 1)
      out.padding = out.padding.t()
    return out


class TransformerXEncoder(TransformerEncoder):
  """Encoder that can interpolate the embeddings of two sentences (XEnDec, ref :1034):
  `emb = λ0 · emb(batch) + λ1 · emb(interpolation_batch)`, paddings intersected."""

  def FProp(self, theta, input_batch, interpolation_batch=None, lambdas=None):
    p = self.params
    ids, pad = input_batch.ids.long(), input_batch.paddings.float()
    t = ids.shape[1]
    emb = input_batch.get('embs')
    if emb is None or interpolation_batch is None:
      emb = self.token_emb.EmbLookup(theta.token_emb, ids)
    if interpolation_batch is not None:
      other = interpolation_batch.get('embs')
      if other is None:
        other = self.token_emb.EmbLookup(theta.token_emb, interpolation_batch.ids.long())
      emb = lambdas[0].unsqueeze(-1).to(emb.dtype) * emb + \
          lambdas[1].unsqueeze(-1).to(emb.dtype) * other
      pad = (pad + interpolation_batch.paddings.float() - 1.0).clamp(0.0, 1.0)
    orig = emb
    x = emb * (p.model_dim ** 0.5)
    pos = self.position_emb.FProp(theta.position_emb, t).unsqueeze(0)
    x = self.input_dropout.FProp(theta.input_dropout, x + pos.to(x.dtype))
    out, _ = self.transformer_stack.FProp(theta.transformer_stack, x, pad)
    return NestedMap(encoded=out.transpose(0, 1), padding=pad.t(), segment_id=None,
                     embedded_inputs=orig)
