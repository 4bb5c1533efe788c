"""Step abstraction: process sequences one time step at a time
(ref `lingvo/core/step.py`).

  prepared = step.PrepareExternalInputs(theta, external_inputs)
  state = step.ZeroState(theta, prepared, batch_size)
  for t: out, state = step.FProp(theta, prepared, step_inputs_t, padding_t, state)

`StatelessLayerStep` (ref :168) adapts any layer, `StackStep` (ref :212) chains
steps with optional residuals, `ParallelStep` (ref :341) runs steps side by side,
`GraphStep` (ref :404) wires sub-steps into a data-flow graph with string signatures,
`IteratorStep` (ref :572) feeds a pre-computed sequence one frame per call,
`RecurrentStepWrapper` (ref :660) runs a step over a whole sequence.
"""

from __future__ import annotations

import collections

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

  def StatelessInference(self, theta, prepared_inputs, step_inputs, padding):
    """Like FProp, but `step_inputs` carries everything the step needs (no recurrent
    state); returns only the outputs (ref :143)."""
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
    del prepared_inputs, state0
    args = {k: v for k, v in step_inputs.items() if k != 'inputs'}
    ins = step_inputs.inputs
    ins = list(ins) if isinstance(ins, (list, tuple)) else [ins]
    out = self.layer.FProp(theta.layer, *ins, **args)
    return (out if isinstance(out, NestedMap) and 'output' in out else NestedMap(output=out),
            NestedMap())

  def StatelessInference(self, theta, prepared_inputs, step_inputs, padding):
    return self.FProp(theta, prepared_inputs, step_inputs, padding, NestedMap())[0]


class StackStep(Step):
  """A stack of steps (ref :212). Each sub-step takes `NestedMap(inputs=[…])` and returns
  `NestedMap(output=tensor)`; the output of layer n-1 is the input of layer n.

  Three ways to feed the stack: `step_inputs.inputs` reach only the lowest layer,
  `step_inputs.context` (optional) is appended to every layer's inputs, and the prepared
  external inputs are visible to every layer and constant over time.

  Residuals: for i >= residual_start >= 0,
    output[i] = output[i - residual_stride] + sub[i](output[i - 1]),   output[-1] = input.
  """

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('sub', [], 'List of step params.')
    p.Define('residual_start', -1, 'First layer with a residual (-1: none).')
    p.Define('residual_stride', 1, 'Number of layers each residual connection skips.')
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
    external_inputs = external_inputs if external_inputs is not None else NestedMap()
    return NestedMap(sub=[s.PrepareExternalInputs(theta.sub[i], external_inputs)
                          for i, s in enumerate(self.sub)])

  def ZeroState(self, theta, prepared_inputs, batch_size):
    return NestedMap(sub=[s.ZeroState(theta.sub[i], prepared_inputs.sub[i], batch_size)
                          for i, s in enumerate(self.sub)])

  def FProp(self, theta, prepared_inputs, step_inputs, padding, state0):
    p = self.params
    state1 = NestedMap(sub=[])
    inputs = list(step_inputs.inputs)
    residual_inputs = [inputs[0] if len(inputs) == 1 else torch.cat(inputs, 1)]
    additional = [step_inputs.context] if 'context' in step_inputs else []
    output = None
    for i, s in enumerate(self.sub):
      sub_out, st = s.FProp(theta.sub[i], prepared_inputs.sub[i],
                            NestedMap(inputs=inputs + additional), padding, state0.sub[i])
      state1.sub.append(st)
      output = sub_out.output
      if i >= p.residual_start >= 0:
        src = i + 1 - p.residual_stride
        assert 0 <= src < len(residual_inputs), (i, p.residual_start, p.residual_stride)
        output = output + residual_inputs[src]
      residual_inputs.append(output)
      inputs = [output]
    return NestedMap(output=output), state1


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
    external_inputs = external_inputs if external_inputs is not None else NestedMap()
    return NestedMap(sub=[s.PrepareExternalInputs(theta.sub[i], external_inputs)
                          for i, s in enumerate(self.sub)])

  def ZeroState(self, theta, prepared_inputs, batch_size):
    return NestedMap(sub=[s.ZeroState(theta.sub[i], prepared_inputs.sub[i], batch_size)
                          for i, s in enumerate(self.sub)])

  def FProp(self, theta, prepared_inputs, step_inputs, padding, state0):
    outs, states = [], []
    for i, s in enumerate(self.sub):
      o, st = s.FProp(theta.sub[i], prepared_inputs.sub[i], step_inputs, padding, state0.sub[i])
      outs.append(o.output)
      states.append(st)
    return NestedMap(output=torch.cat(outs, 1)), NestedMap(sub=states)


SubStep = collections.namedtuple('SubStep', ['signature', 'external_signature', 'params'])


class GraphStep(Step):
  """Sub-steps connected by a data-flow graph (ref :404) — `builder_layers.GraphLayer` for
  steps. `p.sub` is a list of `SubStep(signature, external_signature, params)`:

    * signature: `'<one input expression>-><output name>'` in `GraphSignature` syntax. The
      input expression is evaluated over the names `step_inputs`, `prepared_inputs` and the
      outputs of earlier sub-steps and becomes that sub-step's `step_inputs`; state0/state1
      are threaded automatically.
    * external_signature: `'external_inputs.<path>'` (or None): which part of this step's
      external inputs the sub-step's `PrepareExternalInputs` receives.
    * params: the sub-step params.

  e.g. `SubStep('(inputs=[rnn.output,step_inputs.context])->atten', 'external_inputs.src', ap)`.
  No cycles; every output name is unique. `p.output_signature` selects the step's output.
  """

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('output_signature', '', 'Signature of the step output.')
    p.Define('sub', [], 'A list of SubSteps.')
    p.Define('dict_type', NestedMap, 'Type of nested dicts.')
    return p

  _Seq = collections.namedtuple('_Seq', ['name', 'signature', 'external_signature', 'step'])

  def __init__(self, params):
    from lingvo_b200.core import builder_layers   # pylint: disable=g-import-not-at-top
    super().__init__(params)
    p = self.params
    assert p.name
    self._seq = []
    produced = set()
    for i, (signature, external_signature, sub_params) in enumerate(p.sub):
      assert signature, 'sub-step %d has no signature' % i
      sig = builder_layers.GraphSignature(signature)
      assert len(sig.inputs) == 1, signature
      assert sig.outputs, signature
      assert sig.outputs[0] not in produced, 'output %r is produced twice' % sig.outputs[0]
      produced.add(sig.outputs[0])
      external_sig = None
      if external_signature:
        external_sig = builder_layers.GraphSignature(
            external_signature if '->' in external_signature else external_signature + '->')
        assert len(external_sig.inputs) == 1 and not external_sig.outputs, external_signature
      sub_params = sub_params.Copy()
      if not sub_params.name:
        sub_params.name = '%s_%02d' % (sig.outputs[0], i)
      self.CreateChild(sub_params.name, sub_params)
      self._seq.append(GraphStep._Seq(sub_params.name, sig, external_sig,
                                      self.children[sub_params.name]))
    osig = p.output_signature
    self.output_signature = builder_layers.GraphSignature(osig if '->' in osig else osig + '->')
    self._graph_tensors_cls = builder_layers.GraphTensors

  @staticmethod
  def _Gather(graph_tensors, sig_inputs):
    return NestedMap(inputs=sig_inputs).Transform(graph_tensors.GetTensor).inputs[0]

  def PrepareExternalInputs(self, theta, external_inputs):
    """→ NestedMap keyed by sub-step name (empty map for steps without externals)."""
    gt = self._graph_tensors_cls()
    gt.StoreTensor('external_inputs', external_inputs)
    prepared = NestedMap()
    for seq in self._seq:
      if seq.external_signature is not None:
        sub_ext = self._Gather(gt, seq.external_signature.inputs)
        prepared[seq.name] = seq.step.PrepareExternalInputs(theta[seq.name], sub_ext)
      else:
        prepared[seq.name] = NestedMap()
    return prepared

  def ZeroState(self, theta, prepared_inputs, batch_size):
    return NestedMap({seq.name: seq.step.ZeroState(theta[seq.name], prepared_inputs[seq.name],
                                                   batch_size) for seq in self._seq})

  def FProp(self, theta, prepared_inputs, step_inputs, padding, state0):
    gt = self._graph_tensors_cls()
    gt.StoreTensor('prepared_inputs', prepared_inputs)
    gt.StoreTensor('step_inputs', step_inputs)
    state1 = NestedMap()
    for seq in self._seq:
      external = prepared_inputs[seq.name] if seq.external_signature is not None else None
      sub_in = self._Gather(gt, seq.signature.inputs)
      out, st = seq.step.FProp(theta[seq.name], external, sub_in, padding, state0[seq.name])
      gt.StoreTensor(seq.signature.outputs[0], out)
      state1[seq.name] = st
    return self._Gather(gt, self.output_signature.inputs), state1


class IteratorStep(Step):
  """Steps through the time axis `p.axis` of the external tensors: each call emits the
  `[batch, …]` slice at the current time (ref :572). `step_inputs` is unused."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('axis', 1, 'The time dimension of the tensors.')
    return p

  def PrepareExternalInputs(self, theta, external_inputs):
    return external_inputs

  def ZeroState(self, theta, prepared_inputs, batch_size):
    return NestedMap(t=0)

  def FProp(self, theta, prepared_inputs, step_inputs, padding, state0):
    del theta, step_inputs, padding
    t, axis = state0.t, self.params.axis
    if isinstance(t, torch.Tensor):
      out = prepared_inputs.Transform(
          lambda x: x.index_select(axis, t.reshape(1).to(x.device)).squeeze(axis))
    else:
      out = prepared_inputs.Transform(lambda x: x.select(axis, t))
    return out, NestedMap(t=t + 1)


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
    """Runs the step over every time step (ref :690).

    inputs: NestedMap of `[T, B, …]`; padding `[T, B(, 1)]`.
    Returns (outputs, states): the per-step outputs and recurrent states stacked on a new
    leading time axis (non-tensor state leaves, e.g. python counters, keep their last value).
    The loop is a plain host loop: every step launches the same kernels, so under
    `GraphedTrainStep` the whole unrolled sequence is one CUDA graph.
    """
    t = padding.shape[0]
    state = state0
    outs, states = [], []
    for i in range(t):
      step_in = inputs.Transform(lambda x, i=i: x[i])
      o, state = self.step.FProp(theta.step, prepared_inputs, step_in, padding[i], state)
      outs.append(o)
      states.append(state)

    def _Stack(maps):
      flats = [m.Flatten() for m in maps]
      merged = []
      for k in range(len(flats[0])):
        vals = [f[k] for f in flats]
        merged.append(torch.stack(vals) if isinstance(vals[0], torch.Tensor) else vals[-1])
      return maps[0].Pack(merged)

    return _Stack(outs), _Stack(states)
