"""LAS attention decoder (ref `lingvo/tasks/asr/decoder.py:48,1378`).

Embedding → attention LSTM (`FRNNWithAttention`, context from the previous step
fed as input) → (N−1) LSTMs that also see the context → softmax over
[rnn_out ; context]. Beam search reuses the same step function.
"""

from __future__ import annotations

import torch

from lingvo_b200.core import attention
from lingvo_b200.core import base_decoder
from lingvo_b200.core import layers
from lingvo_b200.core import rnn_cell
from lingvo_b200.core import rnn_layers
from lingvo_b200.core.nested_map import NestedMap


class AsrDecoderBase(base_decoder.BaseBeamSearchDecoder):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('dropout_prob', 0.0, 'Dropout.')
    p.Define('emb', layers.EmbeddingLayer.Params(), 'Embedding.')
    p.Define('emb_dim', 0, 'Embedding dim.')
    p.Define('label_smoothing', None, 'Label smoother.')
    p.Define('rnn_cell_tpl', rnn_cell.LSTMCellSimple.Params(), 'RNN cell template.')
    p.Define('rnn_cell_dim', 0, 'RNN cell dim.')
    p.Define('rnn_cell_hidden_dim', 0, 'RNN hidden dim (projection).')
    p.Define('attention', attention.AdditiveAttention.Params(), 'Attention.')
    p.Define('softmax', layers.SimpleFullSoftmax.Params(), 'Softmax.')
    p.Define('softmax_uses_attention', True, 'Concat context to the softmax input.')
    p.Define('source_dim', 0, 'Encoder output dim.')
    p.Define('atten_context_dim', 0, 'Context dim (0: source_dim).')
    p.Define('first_rnn_input_dim', 0, 'Kept for parity.')
    p.Define('rnn_layers', 1, 'Decoder RNN layers.')
    p.Define('residual_start', 0, 'First residual layer (0: none).')
    p.Define('fusion', None, 'LM fusion params (kept for parity).')
    p.Define('parallel_iterations', 30, 'Kept for parity.')
    p.Define('per_token_avg_loss', True, 'Average the loss per token.')
    p.Define('token_normalized_per_seq_loss', False, 'Kept for parity.')
    p.Define('min_ground_truth_prob', 1.0, 'Scheduled sampling: P(ground truth).')
    p.Define('min_prob_step', 1e6, 'Scheduled sampling ramp end.')
    p.Define('prob_decay_start_step', 1e4, 'Scheduled sampling ramp start.')
    p.Define('use_while_loop_based_unrolling', False, 'Kept for parity.')
    p.Define('logit_types', {'logits': 1.0}, 'Kept for parity.')
    p.Define('use_unnormalized_logits_as_log_probs', True, 'Kept for parity.')
    p.Define('contextualizer', None, 'Kept for parity.')
    p.Define('focal_loss_alpha', None, 'Focal loss α.')
    p.Define('focal_loss_gamma', None, 'Focal loss γ.')
    p.target_seq_len = 300
    return p

  @classmethod
  def UpdateTargetVocabSize(cls, p, vocab_size, wpm_model=None):
    p.emb.vocab_size = vocab_size
    p.softmax.num_classes = vocab_size
    return p


class AsrDecoder(AsrDecoderBase):

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    ctx = p.atten_context_dim or p.source_dim
    self._ctx = ctx
    self.CreateChild('emb', p.emb.Copy().Set(embedding_dim=p.emb_dim))
    self.CreateChild('dropout', layers.DropoutLayer.Params().Set(keep_prob=1 - p.dropout_prob))
    atten = p.attention.Copy().Set(source_dim=p.source_dim, query_dim=p.rnn_cell_dim)
    cell = p.rnn_cell_tpl.Copy().Set(num_input_nodes=p.emb_dim + ctx,
                                     num_output_nodes=p.rnn_cell_dim,
                                     num_hidden_nodes=p.rnn_cell_hidden_dim)
    self.CreateChild('atten_rnn', rnn_layers.FRNNWithAttention.Params().Set(
        cell=cell, attention=atten, use_zero_atten_state=True, atten_context_dim=ctx,
        packed_input=p.packed_input))
    rnns = []
    for i in range(1, p.rnn_layers):
      rnns.append(rnn_layers.FRNN.Params().Set(
          name='rnn_%d' % i, cell=p.rnn_cell_tpl.Copy().Set(
              num_input_nodes=p.rnn_cell_dim + ctx, num_output_nodes=p.rnn_cell_dim,
              num_hidden_nodes=p.rnn_cell_hidden_dim)))
    self.CreateChildren('rnn', rnns)
    sm_in = p.rnn_cell_dim + (ctx if p.softmax_uses_attention else 0)
    self.CreateChild('softmax', p.softmax.Copy().Set(input_dim=sm_in))
    if p.label_smoothing is not None:
      self.CreateChild('smoother', p.label_smoothing.Copy().Set(
          num_classes=p.softmax.num_classes))

  def ComputePredictions(self, theta, encoder_outputs, targets):
    p = self.params
    ids = targets.ids.t().long()
    pad = targets.paddings.t().float().unsqueeze(-1)
    emb = self.dropout.FProp(theta.dropout, self.emb.EmbLookup(theta.emb, ids))
    ctx, xs, probs, _ = self.atten_rnn.FProp(
        theta.atten_rnn, encoder_outputs.encoded, encoder_outputs.padding, emb, pad)
    for i, r in enumerate(self.rnn):
      ys, _ = r.FProp(theta.rnn[i], torch.cat([xs, ctx], -1), pad)
      xs = xs + ys if (p.residual_start and i + 1 >= p.residual_start) else ys
    sm_in = torch.cat([xs, ctx], -1) if p.softmax_uses_attention else xs
    return NestedMap(softmax_input=self.dropout.FProp(theta.dropout, sm_in),
                     attention=NestedMap(probs=probs))

  def ComputeLoss(self, theta, predictions, targets):
    p = self.params
    x = predictions.softmax_input
    t, b, d = x.shape
    lab = targets.labels.t().long()
    w = targets.weights.t().float()
    kwargs = dict(class_ids=lab.reshape(-1, 1))
    if p.label_smoothing is not None:
      probs = self.smoother.FProp(theta.smoother, targets.paddings.t(), lab, target_ids=None)
      kwargs = dict(class_probabilities=probs.reshape(t * b, -1))
    out = self.softmax.FProp(theta.softmax, x.reshape(t * b, d), w.reshape(-1, 1), **kwargs)
    n = w.sum().clamp_min(1e-8)
    loss = out.total_xent / n if p.per_token_avg_loss else out.total_xent / float(b)
    correct = ((out.per_example_argmax.reshape(t, b) == lab).float() * w).sum() \
        if out.get('per_example_argmax') is not None else torch.zeros((), device=w.device)
    metrics = NestedMap(loss=(loss, n), log_pplx=(out.total_xent / n, n),
                        fraction_of_correct_next_step_preds=(correct / n, n))
    return metrics, NestedMap()

  # -- beam search ----------------------------------------------------------------
  def _InitBeamSearchStateCallback(self, theta, encoder_outputs, num_hyps_per_beam):
    p = self.params
    n = encoder_outputs.encoded.shape[1] * num_hyps_per_beam
    fa = self.atten_rnn
    packed = fa.InitForSourcePacked(theta.atten_rnn, encoder_outputs.encoded,
                                    encoder_outputs.padding)
    encoder_outputs.packed_src = packed
    st = fa.zero_state(theta.atten_rnn, encoder_outputs.encoded, packed, n)
    upper = [r.zero_state(theta.rnn[i], n) for i, r in enumerate(self.rnn)]
    dev = encoder_outputs.encoded.device
    init = NestedMap(log_probs=torch.zeros(n, p.softmax.num_classes, device=dev),
                     atten_probs=torch.zeros(n, encoder_outputs.encoded.shape[0], device=dev))
    return init, NestedMap(atten=st, upper=upper)

  def _PreBeamSearchStepCallback(self, theta, encoder_outputs, step_ids, states,
                                 num_hyps_per_beam, cur_step):
    p = self.params
    n = step_ids.shape[0]
    emb = self.emb.EmbLookup(theta.emb, step_ids.squeeze(1).long())
    pad = torch.zeros(n, 1, device=emb.device)
    st = self.atten_rnn.Step(theta.atten_rnn, encoder_outputs.packed_src, states.atten,
                             emb, pad)
    xs = self.atten_rnn.cell.GetOutput(st.rnn)
    upper = []
    for i, r in enumerate(self.rnn):
      s1, _ = r.cell.FProp(theta.rnn[i].cell, states.upper[i],
                           NestedMap(act=[torch.cat([xs, st.atten], -1)], padding=pad))
      ys = r.cell.GetOutput(s1)
      xs = xs + ys if (p.residual_start and i + 1 >= p.residual_start) else ys
      upper.append(s1)
    sm_in = torch.cat([xs, st.atten], -1) if p.softmax_uses_attention else xs
    logits = self.softmax.Logits(theta.softmax, sm_in)
    return (NestedMap(log_probs=torch.log_softmax(logits.float(), -1),
                      atten_probs=st.atten_probs), NestedMap(atten=st, upper=upper))
