"""LstmFRNN: FRNN specialised to LSTM cells (ref `lingvo/core/lstm_frnn_layer.py`).

The reference hand-writes the LSTM backward to save memory; here the input GEMM
is hoisted for the whole sequence (see `rnn_cell.ProjectInput`) and the time loop
is rematerialised in chunks (`recurrent.Recurrent(remat_steps=…)`).
"""

from __future__ import annotations

import torch

from lingvo_b200.core import base_layer
from lingvo_b200.core import recurrent
from lingvo_b200.core import rnn_cell
from lingvo_b200.core.nested_map import NestedMap


class LSTMCellExt:
  """Mixin for LSTM cells: the input GEMM of a whole sequence in one call, then per-step
  updates that only add the recurrent projection (ref :27)."""

  def ProjectInputSequence(self, theta, inputs):
    """inputs.act: list of `[T, B, D_i]` → `[T, B, gates·H]` (bias included)."""
    assert isinstance(inputs.act, (list, tuple))
    x = inputs.act[0] if len(inputs.act) == 1 else torch.cat(list(inputs.act), -1)
    return self.ProjectInput(theta, x)

  def _MixWithProjectedInput(self, theta, state0, inputs):
    p = self.params
    w_h = theta.wm[p.num_input_nodes:].to(inputs.dtype)
    return inputs + torch.matmul(state0.m.to(inputs.dtype), w_h)

  def FPropWithProjectedInput(self, theta, state0, inputs):
    """inputs: NestedMap(proj_inputs `[B, gates·H]`, padding `[B, 1]`[, reset_mask]).
    Equivalent to `FProp` on the un-projected step input."""
    if self.params.reset_cell_state:
      state0 = self._ResetState(state0.DeepCopy(), inputs)
    return self._Step(theta, state0, inputs.proj_inputs, inputs.padding, inputs), NestedMap()


class LSTMCellSimpleExt(rnn_cell.LSTMCellSimple, LSTMCellExt):
  """LSTMCellSimple + sequence-level input projection (ref :123)."""


class LayerNormalizedLSTMCellSimpleExt(rnn_cell.LayerNormalizedLSTMCellSimple, LSTMCellExt):
  """LayerNormalizedLSTMCellSimple + sequence-level input projection (ref :131)."""


class LayerNormalizedLSTMCellLeanExt(rnn_cell.LayerNormalizedLSTMCellLean, LSTMCellExt):
  """LayerNormalizedLSTMCellLean + sequence-level input projection (ref :140)."""


class LstmFRNN(base_layer.BaseLayer):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('cell', LSTMCellSimpleExt.Params(), 'LSTM cell params.')
    p.Define('reverse', False, 'Process the sequence backwards.')
    p.Define('packed_input', False, 'Packed inputs.')
    p.Define('remat_steps', 16, 'Rematerialisation chunk (0: keep all activations).')
    return p

  def __init__(self, params):
    super().__init__(params)
    self.CreateChild('cell', self.params.cell)

  def zero_state(self, theta, batch_size):
    return self.cell.zero_state(theta.cell, batch_size)

  def FProp(self, theta, inputs, paddings, state0=None):
    p = self.params
    if state0 is None:
      state0 = self.zero_state(theta, inputs.shape[1])
    pad = paddings if paddings.dim() == 3 else paddings.unsqueeze(-1)
    xw = self.cell.ProjectInput(theta.cell, inputs)
    if p.reverse:
      xw, pad = torch.flip(xw, [0]), torch.flip(pad, [0])

    def cell_fn(th, state, inp):
      return self.cell._Step(th, state, inp.xw, inp.padding), NestedMap()  # pylint: disable=protected-access
    acc, final = recurrent.Recurrent(theta.cell, state0, NestedMap(xw=xw, padding=pad),
                                     cell_fn, remat_steps=p.remat_steps)
    out = acc.m
    if p.reverse:
      out = torch.flip(out, [0])
    return out, final
