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

  def FProp(self, theta, input_batch):
    p = self.params
    ids = input_batch.ids.t()
    pad = input_batch.paddings.t().float().unsqueeze(-1)
    seg = input_batch.segment_ids.t().unsqueeze(-1) if p.packed_input else None
    xs = self.dropout.FProp(theta.dropout, self.emb.EmbLookup(theta.emb, ids.long()))
    for i, r in enumerate(self.rnn):
      ys = r.FProp(theta.rnn[i], xs, pad, segment_id=seg)
      ys = self.dropout.FProp(theta.dropout, ys)
      xs = xs + ys if (i >= p.residual_start and xs.shape == ys.shape) else ys
    out = self.final_proj.FProp(theta.final_proj, xs)
    out = out * (1.0 - pad)
    return NestedMap(encoded=out, padding=pad.squeeze(-1),
                     segment_id=seg.squeeze(-1) if seg is not None else None)


MTEncoderV1 = MTEncoderBiRNN   # ref :33 (GNMT v1 encoder; same building blocks)


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


TransformerBatchMajorEncoder = TransformerEncoder   # ref :836 (already batch-major inside)
