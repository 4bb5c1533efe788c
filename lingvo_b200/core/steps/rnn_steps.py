"""RNN steps (ref `lingvo/core/steps/rnn_steps.py`)."""
import torch

from lingvo_b200.core import rnn_cell
from lingvo_b200.core import step
from lingvo_b200.core.nested_map import NestedMap


class RnnStep(step.Step):
  """One RNN cell as a Step: output = cell output, state = cell state (ref :27)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('cell', rnn_cell.LSTMCellSimple.Params(), 'Cell params.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self.CreateChild('cell', self.params.cell)

  def PrepareExternalInputs(self, theta, external_inputs):
    return external_inputs or NestedMap()

  def ZeroState(self, theta, prepared_inputs, batch_size):
    return self.cell.zero_state(theta.cell, batch_size)

  def FProp(self, theta, prepared_inputs, step_inputs, padding, state0):
    st, _ = self.cell.FProp(theta.cell, state0, NestedMap(act=list(step_inputs.inputs),
                                                          padding=padding))
    return NestedMap(output=self.cell.GetOutput(st)), st


class RnnStackStep(step.Step):
  """Stack of RnnSteps; step inputs (+ optional `context`) feed every layer (ref :90)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('rnn_cell_tpl', rnn_cell.LSTMCellSimple.Params(), 'Cell template.')
    p.Define('external_input_dim', 0, 'Dim of the per-step context fed to every layer.')
    p.Define('step_input_dim', 0, 'Input dim of layer 0.')
    p.Define('context_input_dim', 0, 'Dim of the external context.')
    p.Define('rnn_cell_dim', 0, 'Cell dim.')
    p.Define('rnn_cell_hidden_dim', 0, 'Cell hidden dim.')
    p.Define('rnn_layers', 1, 'Number of layers.')
    p.Define('residual_start', -1, 'First residual layer.')
    p.Define('residual_stride', 1, 'Residual stride.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    subs = []
    for i in range(p.rnn_layers):
      idim = (p.step_input_dim if i == 0 else p.rnn_cell_dim) + p.context_input_dim
      subs.append(RnnStep.Params().Set(name='rnn_%d' % i, cell=p.rnn_cell_tpl.Copy().Set(
          num_input_nodes=idim, num_output_nodes=p.rnn_cell_dim,
          num_hidden_nodes=p.rnn_cell_hidden_dim)))
    self.CreateChildren('sub', subs)

  def PrepareExternalInputs(self, theta, external_inputs):
    return external_inputs or NestedMap()

  def ZeroState(self, theta, prepared_inputs, batch_size):
    return NestedMap(sub=[s.ZeroState(theta.sub[i], NestedMap(), batch_size)
                          for i, s in enumerate(self.sub)])

  def FProp(self, theta, prepared_inputs, step_inputs, padding, state0):
    p = self.params
    x = step_inputs.inputs[0] if len(step_inputs.inputs) == 1 else torch.cat(
        list(step_inputs.inputs), -1)
    ctx = step_inputs.get('context')
    states = []
    for i, s in enumerate(self.sub):
      ins = [x] + ([ctx] if ctx is not None else [])
      o, st = s.FProp(theta.sub[i], NestedMap(), NestedMap(inputs=ins), padding, state0.sub[i])
      y = o.output
      if p.residual_start >= 0 and i >= p.residual_start and \
          (i - p.residual_start) % p.residual_stride == 0 and y.shape == x.shape:
        y = y + x
      x = y
      states.append(st)
    return NestedMap(output=x), NestedMap(sub=states)
