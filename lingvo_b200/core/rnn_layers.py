"""Sequence-level RNN layers (ref `lingvo/core/rnn_layers.py`).

All tensors are time-major: inputs `[T, B, D]`, paddings `[T, B, 1]`.

`FRNN` is the workhorse: it asks the cell for `ProjectInput` (one tensor-core
GEMM over the whole `[T·B, D]` input) and then loops only over the recurrent
half of the step — the reference gets the same effect for a few cells only via
`LayerNormalizedLSTMCellLean`-style "lean" code paths; here it is the protocol
of every cell (see `rnn_cell.py`).
"""

from __future__ import annotations

import torch

from lingvo_b200.core import attention as attention_lib
from lingvo_b200.core import base_layer
from lingvo_b200.core import layers
from lingvo_b200.core import py_utils
from lingvo_b200.core import quant_utils
from lingvo_b200.core import recurrent
from lingvo_b200.core import rnn_cell
from lingvo_b200.core.nested_map import NestedMap


def GeneratePackedInputResetMask(segment_id, is_reverse=False):
  """1 where step t continues step t∓1's segment, 0 at segment starts (ref :27).

  segment_id `[T, B, 1]` → mask `[T, B, 1]` float.
  """
  seg = segment_id
  if is_reverse:
    nxt = torch.cat([seg[1:], seg[-1:]], 0)
    same = (seg == nxt)
    same[-1] = False
  else:
    prev = torch.cat([seg[:1], seg[:-1]], 0)
    same = (seg == prev)
    same[0] = False
  return same.to(torch.float32)


class IdentitySeqLayer(base_layer.BaseLayer):
  """Pass-through with the FRNN signature (ref :55)."""

  def zero_state(self, theta, batch_size):
    return NestedMap()

  def FPropFullSequence(self, theta, inputs, paddings):
    del theta, paddings
    return inputs

  def FProp(self, theta, inputs, *args, **kwargs):
    return inputs


def _Pad3(paddings):
  return paddings if paddings.dim() == 3 else paddings.unsqueeze(-1)


def _RunCell(cell, theta, inputs, paddings, state0, reverse=False, reset_mask=None):
  """Shared FRNN driver. Returns (outputs [T,B,Dout], final_state)."""
  t = inputs.shape[0]
  paddings = _Pad3(paddings)
  hoist = hasattr(cell, '_Step') and type(cell).FProp is rnn_cell.RNNCell.FProp
  xw = cell.ProjectInput(theta, inputs) if hoist else None
  order = range(t - 1, -1, -1) if reverse else range(t)
  outs = [None] * t
  state = state0
  for i in order:
    if hoist:
      st = state
      if reset_mask is not None and cell.params.reset_cell_state:
        st = st.Transform(lambda x, m=reset_mask[i]: x * m.to(x.dtype))
      state = cell._Step(theta, st, xw[i], paddings[i])   # pylint: disable=protected-access
    else:
      step_in = NestedMap(act=[inputs[i]], padding=paddings[i])
      if reset_mask is not None:
        step_in.reset_mask = reset_mask[i]
      state, _ = cell.FProp(theta, state, step_in)
    outs[i] = cell.GetOutput(state)
  return torch.stack(outs, 0), state


class RNN(base_layer.BaseLayer):
  """Statically unrolled RNN (ref :69)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('cell', rnn_cell.LSTMCellSimple.Params(), 'Cell params.')
    p.Define('sequence_length', 0, 'Kept for parity (length comes from the input).')
    p.Define('reverse', False, 'Process the sequence backwards.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self.CreateChild('cell', self.params.cell)

  def zero_state(self, theta, batch_size):
    return self.cell.zero_state(theta.cell, batch_size)

  def FProp(self, theta, inputs, paddings, state0=None):
    if isinstance(inputs, (list, tuple)):
      inputs = torch.stack(list(inputs), 0)
      paddings = torch.stack(list(paddings), 0)
    if state0 is None:
      state0 = self.zero_state(theta, inputs.shape[1])
    return _RunCell(self.cell, theta.cell, inputs, paddings, state0,
                    self.params.reverse)


class FRNN(base_layer.BaseLayer):
  """Functional RNN over a whole sequence (ref :365)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('cell', rnn_cell.LSTMCellSimple.Params(), 'Cell params.')
    p.Define('reverse', False, 'Process the sequence backwards.')
    p.Define('packed_input', False, 'Reset state at segment boundaries.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    cell_p = p.cell.Copy()
    if p.packed_input:
      cell_p.reset_cell_state = True
    self.CreateChild('cell', cell_p)

  @property
  def rnn_cell(self):
    return self.cell

  def zero_state(self, theta, batch_size):
    return self.cell.zero_state(theta.cell, batch_size)

  def FProp(self, theta, inputs, paddings, state0=None, segment_id=None):
    p = self.params
    if state0 is None:
      state0 = self.zero_state(theta, inputs.shape[1])
    reset = None
    if p.packed_input:
      assert segment_id is not None
      reset = GeneratePackedInputResetMask(_Pad3(segment_id), p.reverse)
    return _RunCell(self.cell, theta.cell, inputs, paddings, state0, p.reverse, reset)


class BidirectionalFRNN(base_layer.BaseLayer):
  """Forward + backward FRNN, outputs concatenated (ref :487)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('fwd', rnn_cell.LSTMCellSimple.Params(), 'Forward cell.')
    p.Define('bak', rnn_cell.LSTMCellSimple.Params(), 'Backward cell.')
    p.Define('rnn', FRNN.Params(), 'FRNN template.')
    p.Define('packed_input', False, 'Packed inputs.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    self.CreateChild('fwd_rnn', p.rnn.Copy().Set(
        cell=p.fwd, reverse=False, packed_input=p.packed_input))
    self.CreateChild('bak_rnn', p.rnn.Copy().Set(
        cell=p.bak, reverse=True, packed_input=p.packed_input))

  def FProp(self, theta, inputs, paddings, segment_id=None):
    f, _ = self.fwd_rnn.FProp(theta.fwd_rnn, inputs, paddings, segment_id=segment_id)
    b, _ = self.bak_rnn.FProp(theta.bak_rnn, inputs, paddings, segment_id=segment_id)
    return torch.cat([f, b], -1)


class BidirectionalRNN(base_layer.BaseLayer):
  """Statically unrolled bidirectional RNN (ref :592)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('fwd', rnn_cell.LSTMCellSimple.Params(), 'Forward cell.')
    p.Define('bak', rnn_cell.LSTMCellSimple.Params(), 'Backward cell.')
    p.Define('sequence_length', 0, 'Kept for parity.')
    p.Define('rnn', RNN.Params(), 'RNN template.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    self.CreateChild('fwd_rnn', p.rnn.Copy().Set(cell=p.fwd, reverse=False))
    self.CreateChild('bak_rnn', p.rnn.Copy().Set(cell=p.bak, reverse=True))

  def FProp(self, theta, inputs, paddings):
    f, _ = self.fwd_rnn.FProp(theta.fwd_rnn, inputs, paddings)
    b, _ = self.bak_rnn.FProp(theta.bak_rnn, inputs, paddings)
    return torch.cat([f, b], -1)


class BidirectionalRNNV2(base_layer.BaseLayer):
  """Bidirectional RNN unrolled over a fixed `sequence_length` (ref :659): shorter inputs are
  padded (activations with 0, paddings with 1) up to `sequence_length`, the wrapped
  `BidirectionalRNN` runs on the fixed-length sequence — one static unroll that captures
  into a single CUDA graph whatever the batch's true length — and the output is cut back."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('fwd', rnn_cell.LSTMCellSimple.Params(), 'Forward cell.')
    p.Define('bak', rnn_cell.LSTMCellSimple.Params(), 'Backward cell.')
    p.Define('sequence_length', 0, 'Sequence length.')
    p.Define('packed_input', False, 'Not supported (as in the reference).')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    assert not p.packed_input, 'Packed input is not supported by BidirectionalRNNV2'
    self.CreateChild('brnn', BidirectionalRNN.Params().Set(
        name='%s_brnn' % p.name, fwd=p.fwd.Copy(), bak=p.bak.Copy(),
        sequence_length=p.sequence_length))

  @staticmethod
  def _PadSequenceToLength(x, length, pad_value):
    t = x.shape[0]
    assert t <= length, 'sequence of %d steps exceeds sequence_length=%d' % (t, length)
    if t == length:
      return x
    return torch.cat([x, x.new_full((length - t,) + tuple(x.shape[1:]), pad_value)], 0)

  def FProp(self, theta, inputs, paddings):
    """inputs `[T, B, D]`, paddings `[T, B, 1]` → `[T, B, fwd+bak dims]`."""
    p = self.params
    seq_len = paddings.shape[0]
    length = p.sequence_length or seq_len
    x = self._PadSequenceToLength(inputs, length, 0.0)
    pad = self._PadSequenceToLength(_Pad3(paddings), length, 1.0)
    return self.brnn.FProp(theta.brnn, x, pad)[:seq_len]


class StackedRNNBase(base_layer.BaseLayer):
  """Shared params of the stacked variants (ref :153)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('num_layers', 1, 'Number of layers.')
    p.Define('skip_start', 1, 'First layer with a residual connection.')
    p.Define('num_input_nodes', 0, 'Input width.')
    p.Define('num_output_nodes', 0, 'Output width.')
    p.Define('packed_input', False, 'Packed inputs.')
    p.Define('cell_tpl', rnn_cell.LSTMCellSimple.Params(), 'Cell template(s).')
    p.Define('dropout', layers.DropoutLayer.Params(), 'Dropout between layers.')
    return p

  def _CellParams(self, i, idim, odim):
    p = self.params
    tpl = p.cell_tpl[i] if isinstance(p.cell_tpl, (list, tuple)) else p.cell_tpl
    return tpl.Copy().Set(num_input_nodes=idim, num_output_nodes=odim)


class StackedFRNNLayerByLayer(StackedRNNBase, quant_utils.QuantizableLayer):
  """N FRNNs, residual from `skip_start`, dropout on each layer's input (ref :193)."""

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    rnns = []
    for i in range(p.num_layers):
      idim = p.num_input_nodes if i == 0 else p.num_output_nodes
      rnns.append(FRNN.Params().Set(
          name='frnn_%d' % i, cell=self._CellParams(i, idim, p.num_output_nodes),
          packed_input=p.packed_input))
    self.CreateChildren('rnn', rnns)
    self.CreateChild('dropout', p.dropout)

  def zero_state(self, theta, batch_size):
    return NestedMap(rnn=[r.zero_state(theta.rnn[i], batch_size)
                          for i, r in enumerate(self.rnn)])

  def FProp(self, theta, inputs, paddings, state0=None, segment_id=None):
    p = self.params
    xs = inputs
    finals = []
    for i, r in enumerate(self.rnn):
      s0 = state0.rnn[i] if state0 is not None else None
      ys, final = r.FProp(theta.rnn[i], self.dropout.FProp(theta.dropout, xs),
                          paddings, s0, segment_id=segment_id)
      finals.append(final)
      xs = xs + ys if (i >= p.skip_start and xs.shape == ys.shape) else ys
    return xs, NestedMap(rnn=finals)


def _StackedFPropFullSequence(self, theta, inputs, paddings):
  """Outputs of the whole sequence only (no final state) (ref :287, :361)."""
  out = self.FProp(theta, inputs, paddings)
  return out[0] if isinstance(out, tuple) else out


StackedFRNNLayerByLayer.FPropFullSequence = _StackedFPropFullSequence


class StackedBiFRNNLayerByLayer(StackedRNNBase, quant_utils.QuantizableLayer):
  """N bidirectional FRNNs (ref :291); each direction has `num_output_nodes/2`."""

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    assert p.num_output_nodes % 2 == 0
    rnns = []
    for i in range(p.num_layers):
      idim = p.num_input_nodes if i == 0 else p.num_output_nodes
      cell = self._CellParams(i, idim, p.num_output_nodes // 2)
      rnns.append(BidirectionalFRNN.Params().Set(
          name='bifrnn_%d' % i, fwd=cell.Copy(), bak=cell.Copy(),
          packed_input=p.packed_input))
    self.CreateChildren('rnn', rnns)
    self.CreateChild('dropout', p.dropout)

  def FProp(self, theta, inputs, paddings, segment_id=None):
    p = self.params
    xs = inputs
    for i, r in enumerate(self.rnn):
      ys = r.FProp(theta.rnn[i], self.dropout.FProp(theta.dropout, xs), paddings,
                   segment_id=segment_id)
      xs = xs + ys if (i >= p.skip_start and xs.shape == ys.shape) else ys
    return xs


class FRNNWithAttention(base_layer.BaseLayer):
  """RNN whose input at step t is [x_t ; context_{t-1}], followed by attention
  over a packed source (ref :756). The decoder core of RNMT / LAS."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('cell', rnn_cell.LSTMCellSimple.Params(), 'Cell params.')
    p.Define('attention', attention_lib.AdditiveAttention.Params(), 'Attention params.')
    p.Define('output_prev_atten_ctx', False, 'Emit context_{t-1} instead of context_t.')
    p.Define('use_zero_atten_state', False, 'Zero initial attention state/context.')
    p.Define('atten_context_dim', 0, 'Context width (needed for zero state).')
    p.Define('packed_input', False, 'Packed inputs.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    cell_p = p.cell.Copy()
    if p.packed_input:
      cell_p.reset_cell_state = True
    self.CreateChild('cell', cell_p)
    self.CreateChild('atten', p.attention.Copy().Set(packed_input=p.packed_input))

  @property
  def rnn_cell(self):
    return self.cell

  @property
  def attention(self):
    return self.atten

  def InitForSourcePacked(self, theta, src_encs, src_enc_padding, src_contexts=None,
                          src_segment_id=None):
    if src_contexts is None:
      src_contexts = src_encs
    pad = src_enc_padding.squeeze(-1) if src_enc_padding.dim() == 3 else src_enc_padding
    seg = None
    if src_segment_id is not None:
      seg = src_segment_id.squeeze(-1) if src_segment_id.dim() == 3 else src_segment_id
    return self.atten.InitForSourcePacked(theta.atten, src_encs, src_contexts, pad, seg)

  def zero_state(self, theta, src_encs, packed_src, batch_size):
    """Initial (rnn state, attention context/probs/state)."""
    p = self.params
    s_len = src_encs.shape[0]
    rnn = self.cell.zero_state(theta.cell, batch_size)
    atten_state = self.atten.ZeroAttentionState(s_len, batch_size)
    if p.use_zero_atten_state:
      ctx = torch.zeros(batch_size, p.atten_context_dim or src_encs.shape[-1],
                        device=src_encs.device, dtype=src_encs.dtype)
      probs = torch.zeros(batch_size, s_len, device=src_encs.device)
    else:
      ctx, probs, atten_state = self.atten.ComputeContextVectorWithSource(
          theta.atten, packed_src, self.cell.GetOutput(rnn).to(src_encs.dtype),
          atten_state)
    return NestedMap(rnn=rnn, atten=ctx, atten_probs=probs, atten_state=atten_state)

  def Step(self, theta, packed_src, state0, x_t, padding_t, reset_mask=None,
           query_segment_id=None):
    """One decoder step; returns the new state NestedMap."""
    step_in = NestedMap(act=[torch.cat([x_t, state0.atten.to(x_t.dtype)], -1)],
                        padding=padding_t)
    if reset_mask is not None:
      step_in.reset_mask = reset_mask
    rnn1, _ = self.cell.FProp(theta.cell, state0.rnn, step_in)
    ctx, probs, astate = self.atten.ComputeContextVectorWithSource(
        theta.atten, packed_src, self.cell.GetOutput(rnn1), state0.atten_state,
        query_segment_id=query_segment_id)
    return NestedMap(rnn=rnn1, atten=ctx, atten_probs=probs, atten_state=astate)

  def FProp(self, theta, src_encs, src_enc_padding, inputs, paddings,
            src_contexts=None, state0=None, src_segment_id=None, segment_id=None):
    """Returns (atten_context [T,B,C], rnn_output [T,B,D], atten_probs [T,B,S],
    final_state)."""
    p = self.params
    packed = self.InitForSourcePacked(theta, src_encs, src_enc_padding, src_contexts,
                                      src_segment_id)
    t, b = inputs.shape[:2]
    paddings = _Pad3(paddings)
    if state0 is None:
      state0 = self.zero_state(theta, src_encs, packed, b)
    reset = None
    if p.packed_input and segment_id is not None:
      reset = GeneratePackedInputResetMask(_Pad3(segment_id))
    state = state0
    ctxs, outs, probs = [], [], []
    for i in range(t):
      prev_ctx = state.atten
      qseg = None
      if p.packed_input and segment_id is not None:
        qseg = _Pad3(segment_id)[i].squeeze(-1)
      state = self.Step(theta, packed, state, inputs[i], paddings[i],
                        reset[i] if reset is not None else None, qseg)
      ctxs.append(prev_ctx if p.output_prev_atten_ctx else state.atten)
      outs.append(self.cell.GetOutput(state.rnn))
      probs.append(state.atten_probs)
    return torch.stack(ctxs, 0), torch.stack(outs, 0), torch.stack(probs, 0), state


StackedBiFRNNLayerByLayer.FPropFullSequence = _StackedFPropFullSequence


def _FrnnAttenInitAttention(self, theta, src_encs, src_paddings, src_contexts=None,
                            src_segment_id=None):
  """Alias of `InitForSourcePacked` under the reference's name (ref :806)."""
  return self.InitForSourcePacked(theta, src_encs, src_paddings, src_contexts, src_segment_id)


def _FrnnAttenResetAttenState(self, theta, state, inputs):
  """Packed inputs: zero the attention context / probs / state where a new segment starts
  (`inputs.reset_mask` is 0 there) (ref :898)."""
  del theta
  m = inputs.reset_mask
  state.atten = m.to(state.atten.dtype) * state.atten
  if isinstance(state.atten_state, NestedMap):
    if 'inner' not in state.atten_state:
      raise ValueError('Unknown .atten_state, expecting field "inner": %s' % state.atten_state)
    state.atten_state.inner = m.to(state.atten_state.inner.dtype) * state.atten_state.inner
  elif isinstance(state.atten_state, torch.Tensor) and state.atten_state.numel():
    state.atten_state = m.to(state.atten_state.dtype) * state.atten_state
  state.atten_probs = m.to(state.atten_probs.dtype) * state.atten_probs
  return state


def _FrnnAttenAccumulateStates(self, theta, src_encs, src_paddings, inputs, paddings,
                               src_contexts=None, state0=None, src_segment_id=None,
                               segment_id=None):
  """The recurrence only (ref :911) → (accumulated states stacked over time, final state,
  side info for `PostProcessStates`). Splitting FProp this way lets a caller reuse the raw
  per-step states (attention states of monotonic / location-aware attention, …)."""
  p = self.params
  packed = self.InitForSourcePacked(theta, src_encs, src_paddings, src_contexts,
                                    src_segment_id)
  t, b = inputs.shape[:2]
  paddings = _Pad3(paddings)
  if state0 is None:
    state0 = self.zero_state(theta, src_encs, packed, b)
  else:
    assert not p.packed_input, 'packed input is only supported with default initial states.'
  reset = torch.zeros_like(paddings)
  if p.packed_input and segment_id is not None:
    reset = GeneratePackedInputResetMask(_Pad3(segment_id))
  state = state0
  steps = []
  for i in range(t):
    qseg = None
    if p.packed_input and segment_id is not None:
      qseg = _Pad3(segment_id)[i].squeeze(-1)
    state = self.Step(theta, packed, state, inputs[i], paddings[i],
                      reset[i] if p.packed_input else None, qseg)
    steps.append(state)
  flat = [s.Flatten() for s in steps]
  acc = steps[0].Pack([torch.stack([f[k] for f in flat], 0) if isinstance(
      flat[0][k], torch.Tensor) else flat[0][k] for k in range(len(flat[0]))])
  return acc, state, NestedMap(state0=state0, reset_mask=reset)


def _FrnnAttenPostProcessStates(self, acc_state, side_info):
  """→ (attention context `[T, B, C]`, rnn output `[T, B, D]`, attention probs `[T, B, S]`)
  (ref :1013); with `output_prev_atten_ctx` the contexts are shifted right by one step
  (step 0 gets the initial context; packed inputs restart at segment boundaries)."""
  p = self.params
  ctx = acc_state.atten
  if p.output_prev_atten_ctx:
    ctx = torch.cat([side_info.state0.atten.unsqueeze(0).to(ctx.dtype), ctx[:-1]], 0)
    if p.packed_input:
      ctx = ctx * side_info.reset_mask.to(ctx.dtype)
  return ctx, self.cell.GetOutput(acc_state.rnn), acc_state.atten_probs


FRNNWithAttention.InitAttention = _FrnnAttenInitAttention
FRNNWithAttention.reset_atten_state = _FrnnAttenResetAttenState
FRNNWithAttention.AccumulateStates = _FrnnAttenAccumulateStates
FRNNWithAttention.PostProcessStates = _FrnnAttenPostProcessStates


class MultiSourceFRNNWithAttention(base_layer.BaseLayer):
  """FRNNWithAttention over several named sources; per-source contexts are merged
  by `atten_merger` (ref :1121)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('cell', rnn_cell.LSTMCellSimple.Params(), 'Cell params.')
    p.Define('attention_tpl', attention_lib.AdditiveAttention.Params(), 'Attention tpl.')
    p.Define('atten_merger', attention_lib.MergerLayer.Params().Set(merger_op='sum'),
             'Merger params.')
    p.Define('source_names', None, 'List of source names.')
    p.Define('share_attention', False, 'One attention for all sources.')
    p.Define('source_name_to_attention_params', None, 'Per-source attention params.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    self.CreateChild('cell', p.cell)
    if p.share_attention:
      self.CreateChild('attentions', p.attention_tpl.Copy().Set(name='atten_shared'))
    else:
      attens = []
      for n in p.source_names:
        tpl = (p.source_name_to_attention_params or {}).get(n, p.attention_tpl)
        attens.append(tpl.Copy().Set(name='atten_%s' % n))
      self.CreateChildren('attentions', attens)
    self.CreateChild('atten_merger', p.atten_merger)

  def _Atten(self, theta, i):
    if self.params.share_attention:
      return self.attentions, theta.attentions
    return self.attentions[i], theta.attentions[i]

  def FProp(self, theta, src_encs, src_paddings, inputs, paddings):
    """src_encs / src_paddings: NestedMap keyed by source name."""
    p = self.params
    packed = []
    for i, n in enumerate(p.source_names):
      a, th = self._Atten(theta, i)
      pad = src_paddings[n]
      pad = pad.squeeze(-1) if pad.dim() == 3 else pad
      packed.append(a.PackSource(th, src_encs[n], src_encs[n], pad))
    t, b = inputs.shape[:2]
    paddings = _Pad3(paddings)
    rnn = self.cell.zero_state(theta.cell, b)
    ctx_dim = src_encs[p.source_names[0]].shape[-1]
    ctx = torch.zeros(b, ctx_dim, device=inputs.device, dtype=inputs.dtype)
    ctxs, outs = [], []
    for s in range(t):
      rnn, _ = self.cell.FProp(theta.cell, rnn, NestedMap(
          act=[torch.cat([inputs[s], ctx], -1)], padding=paddings[s]))
      q = self.cell.GetOutput(rnn)
      per_src = []
      for i in range(len(p.source_names)):
        a, th = self._Atten(theta, i)
        c, _, _ = a.ComputeContextVectorWithSource(th, packed[i], q)
        per_src.append(c)
      ctx = self.atten_merger.FProp(theta.atten_merger, per_src, q)
      ctxs.append(ctx)
      outs.append(q)
    return torch.stack(ctxs, 0), torch.stack(outs, 0)


class BidirectionalFRNNQuasi(base_layer.BaseLayer):
  """Quasi-RNN: conv-produced gates + `QRNNPoolingCell` recurrence, both
  directions (ref :1365)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('fwd', rnn_cell.QRNNPoolingCell.Params(), 'Forward pooling cell.')
    p.Define('bak', rnn_cell.QRNNPoolingCell.Params(), 'Backward pooling cell.')
    p.Define('rnn', FRNN.Params(), 'FRNN template.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    self.CreateChild('fwd_rnn', p.rnn.Copy().Set(cell=p.fwd, reverse=False))
    self.CreateChild('bak_rnn', p.rnn.Copy().Set(cell=p.bak, reverse=True))

  def FProp(self, theta, fwd_gates, bak_gates, paddings):
    f, _ = self.fwd_rnn.FProp(theta.fwd_rnn, fwd_gates, paddings)
    b, _ = self.bak_rnn.FProp(theta.bak_rnn, bak_gates, paddings)
    return torch.cat([f, b], -1)
