"""Step abstraction: process sequences one time step at a time
(ref `lingvo/core/step.py`).

  prepared = step.PrepareExternalInputs(theta, external_inputs)
  state = step.ZeroState(theta, prepared, batch_size)
  for t: out, state = step.FProp(theta, prepared, step_inputs_t, padding_t, state)

`StatelessLayerStep` (ref :240) adapts any layer, `StackStep` (ref :300) chains
steps with optional residuals, `ParallelStep` (ref :420) runs steps side by side,
`IteratorStep` (ref :560) feeds a pre-computed sequence one frame per call,
`RecurrentStepWrapper` (ref :620) runs a step over a whole sequence.
"""

from __future__ import annotations

import torch

from lingvo_b200.core import base_layer
from lingvo_b200.core import py_utils
from lingvo_b200.core.nested_map import NestedMap


class Step(base_layer.BaseLayer):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('adaptive_task_ids', None, 'Kept for parity.')
    p.Define('adaptive_task_dim', 0, 'Kept for parity.')
    return p

  def _StepChildren(self):
    for name, child in self.children.items():
      yield name, child

  def PrepareExternalInputs(self, theta, external_inputs):
    external_inputs = external_inputs or NestedMap()
    packed = NestedMap(external_inputs)
    for name, child in self._StepChildren():
      sub_in = external_inputs.get(name, NestedMap())
      if isinstance(child, (list, tuple)):
        outs = [c.PrepareExternalInputs(theta[name][i], sub_in)
                for i, c in enumerate(child) if isinstance(c, Step)]
        if outs:
          packed[name] = outs
      elif isinstance(child, Step):
        packed[name] = child.PrepareExternalInputs(theta[name], sub_in)
    return packed

  def ZeroState(self, theta, prepared_inputs, batch_size):
    state0 = NestedMap()
    for name, child in self._StepChildren():
      if isinstance(child, (list, tuple)):
        outs = [c.ZeroState(theta[name][i], prepared_inputs[name][i], batch_size)
                for i, c in enumerate(child) if isinstance(c, Step)]
        if outs:
          state0[name] = outs
      elif isinstance(child, Step):
        state0[name] = child.ZeroState(theta[name], prepared_inputs.get(name, NestedMap()),
                                       batch_size)
    return state0

  def FProp(self, theta, prepared_inputs, step_inputs, padding, state0):
    """→ (output NestedMap(output=…), state1)."""
    raise NotImplementedError(type(self))


class StatelessLayerStep(Step):
  """Wraps a stateless layer: output = layer.FProp(*step_inputs.inputs) (ref :240)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('layer', None, 'Params of the wrapped layer.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self.CreateChild('layer', self.params.layer)

  def FProp(self, theta, prepared_inputs, step_inputs, padding, state0):
    args = {k: v for k, v in step_inputs.items() if k != 'inputs'}
    out = self.layer.FProp(theta.layer, *step_inputs.inputs, **args)
    return NestedMap(output=out), state0


class StackStep(Step):
  """Feeds each sub-step's output to the next; optional residual connections
  starting at `residual_start` every `residual_stride` layers (ref :300)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('sub', [], 'List of step params.')
    p.Define('residual_start', -1, 'First layer with a residual (-1: none).')
    p.Define('residual_stride', 1, 'Residual every n layers.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    subs = []
    for i, sp in enumerate(p.sub):
      sp = sp.Copy()
      sp.name = sp.name or 'sub%d' % i
      subs.append(sp)
    self.CreateChildren('sub', subs)

  def PrepareExternalInputs(self, theta, external_inputs):
    external_inputs = external_inputs or NestedMap()
    return NestedMap(sub=[s.PrepareExternalInputs(theta.sub[i], external_inputs.get(
        s.params.name, NestedMap())) for i, s in enumerate(self.sub)])

  def ZeroState(self, theta, prepared_inputs, batch_size):
    return NestedMap(sub=[s.ZeroState(theta.sub[i], prepared_inputs.sub[i], batch_size)
                          for i, s in enumerate(self.sub)])

  def FProp(self, theta, prepared_inputs, step_inputs, padding, state0):
    p = self.params
    state1 = NestedMap(sub=[])
    inputs = list(step_inputs.inputs)
    extra = {k: v for k, v in step_inputs.items() if k != 'inputs'}
    res_in = None
    for i, s in enumerate(self.sub):
      if p.residual_start >= 0 and i >= p.residual_start and \
          (i - p.residual_start) % p.residual_stride == 0:
        res_in = inputs[0]
      else:
        res_in = None if (p.residual_start < 0 or i < p.residual_start) else res_in
      out, st = s.FProp(theta.sub[i], prepared_inputs.sub[i],
                        NestedMap(inputs=inputs, **extra), padding, state0.sub[i])
      y = out.output
      if res_in is not None and isinstance(y, torch.Tensor) and y.shape == res_in.shape \
          and (i - p.residual_start + 1) % p.residual_stride == 0:
        y = y + res_in
      inputs = [y]
      state1.sub.append(st)
    return NestedMap(output=inputs[0]), state1


class ParallelStep(Step):
  """Runs all sub-steps on the same input; outputs concatenated on the last dim (ref :420)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('sub', [], 'List of step params.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self.CreateChildren('sub', [sp.Copy().Set(name=sp.name or 'sub%d' % i)
                                for i, sp in enumerate(self.params.sub)])

  def PrepareExternalInputs(self, theta, external_inputs):
    external_inputs = external_inputs or NestedMap()
    return NestedMap(sub=[s.PrepareExternalInputs(theta.sub[i], external_inputs.get(
        s.params.name, NestedMap())) for i, s in enumerate(self.sub)])

  def ZeroState(self, theta, prepared_inputs, batch_size):
    return NestedMap(sub=[s.ZeroState(theta.sub[i], prepared_inputs.sub[i], batch_size)
                          for i, s in enumerate(self.sub)])

  def FProp(self, theta, prepared_inputs, step_inputs, padding, state0):
    outs, states = [], []
    for i, s in enumerate(self.sub):
      o, st = s.FProp(theta.sub[i], prepared_inputs.sub[i], step_inputs, padding, state0.sub[i])
      outs.append(o.output)
      states.append(st)
    return NestedMap(output=torch.cat(outs, -1)), NestedMap(sub=states)


class IteratorStep(Step):
  """Each call emits the next frame of a `[B, T, …]` external sequence (ref :560)."""

  def PrepareExternalInputs(self, theta, external_inputs):
    return external_inputs

  def ZeroState(self, theta, prepared_inputs, batch_size):
    return NestedMap(t=0)

  def FProp(self, theta, prepared_inputs, step_inputs, padding, state0):
    t = state0.t
    out = prepared_inputs.Transform(lambda x: x[:, t])
    return NestedMap(output=out), NestedMap(t=t + 1)


class RecurrentStepWrapper(base_layer.BaseLayer):
  """Runs a step over `[T, B, …]` inputs (ref :620)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('step', None, 'Step params.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self.CreateChild('step', self.params.step)

  def PrepareExternalInputs(self, theta, external_inputs):
    return self.step.PrepareExternalInputs(theta.step, external_inputs)

  def ZeroState(self, theta, prepared_inputs, batch_size):
    return self.step.ZeroState(theta.step, prepared_inputs, batch_size)

  def FProp(self, theta, prepared_inputs, inputs, padding, state0):
    """inputs: NestedMap of `[T, B, …]`; padding `[T, B, 1]` → (outputs [T,…], final state)."""
    t = padding.shape[0]
    state = state0
    outs = []
    for i in range(t):
      step_in = inputs.Transform(lambda x: x[i])
      o, state = self.step.FProp(theta.step, prepared_inputs, step_in, padding[i], state)
      outs.append(o)
    flat = [torch.stack([o.Flatten()[k] for o in outs]) for k in range(len(outs[0].Flatten()))]
    return outs[0].Pack(flat), state
