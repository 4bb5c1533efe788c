"""Attention steps (ref `lingvo/core/steps/attention_steps.py`)."""
import torch

from lingvo_b200.core import attention
from lingvo_b200.core import step
from lingvo_b200.core.nested_map import NestedMap


class AttentionStep(step.Step):
  """Attention over a packed external source; the query is the step input (ref :30).

  external_inputs: NestedMap(src [T,B,D], context [T,B,C] (optional), padding [T,B]).
  """

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('atten', attention.AdditiveAttention.Params(), 'Attention params.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self.CreateChild('atten', self.params.atten)

  def PrepareExternalInputs(self, theta, external_inputs):
    ctx = external_inputs.get('context')
    if ctx is None:
      ctx = external_inputs.src
    packed = self.atten.PackSource(theta.atten, external_inputs.src, ctx,
                                   external_inputs.padding)
    return NestedMap(packed_src=packed, src_len=external_inputs.src.shape[0],
                     ctx_dim=ctx.shape[-1])

  def ZeroState(self, theta, prepared_inputs, batch_size):
    dev = self.Device()
    return NestedMap(
        atten_context=torch.zeros(batch_size, prepared_inputs.ctx_dim, device=dev),
        atten_probs=torch.zeros(batch_size, prepared_inputs.src_len, device=dev),
        atten_state=self.atten.ZeroAttentionState(prepared_inputs.src_len, batch_size))

  def FProp(self, theta, prepared_inputs, step_inputs, padding, state0):
    q = step_inputs.inputs[0] if len(step_inputs.inputs) == 1 else torch.cat(
        list(step_inputs.inputs), -1)
    ctx, probs, st = self.atten.ComputeContextVectorWithSource(
        theta.atten, prepared_inputs.packed_src, q, state0.atten_state)
    return (NestedMap(output=ctx, context=ctx, probs=probs),
            NestedMap(atten_context=ctx, atten_probs=probs, atten_state=st))


class AttentionBlockStep(step.Step):
  """query step (e.g. RNN) + attention + optional combine: the LAS/RNMT decoder block
  (ref :150). The previous context is fed back to the query step."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('query_generator', None, 'Step producing the query from [inputs; prev ctx].')
    p.Define('attention', AttentionStep.Params(), 'Attention step.')
    p.Define('attention_combiner', None, 'Optional step combining query and context.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    self.CreateChild('query_generator', p.query_generator)
    self.CreateChild('attention', p.attention)
    if p.attention_combiner is not None:
      self.CreateChild('attention_combiner', p.attention_combiner)

  def PrepareExternalInputs(self, theta, external_inputs):
    return NestedMap(
        attention=self.attention.PrepareExternalInputs(theta.attention,
                                                       external_inputs.attention),
        query_generator=self.query_generator.PrepareExternalInputs(
            theta.query_generator, external_inputs.get('query_generator', NestedMap())))

  def ZeroState(self, theta, prepared_inputs, batch_size):
    return NestedMap(
        attention=self.attention.ZeroState(theta.attention, prepared_inputs.attention,
                                           batch_size),
        query_generator=self.query_generator.ZeroState(
            theta.query_generator, prepared_inputs.query_generator, batch_size))

  def FProp(self, theta, prepared_inputs, step_inputs, padding, state0):
    prev_ctx = state0.attention.atten_context
    q_in = NestedMap(inputs=list(step_inputs.inputs) + [prev_ctx.to(step_inputs.inputs[0].dtype)])
    q_out, q_state = self.query_generator.FProp(theta.query_generator,
                                                prepared_inputs.query_generator, q_in, padding,
                                                state0.query_generator)
    a_out, a_state = self.attention.FProp(theta.attention, prepared_inputs.attention,
                                          NestedMap(inputs=[q_out.output]), padding,
                                          state0.attention)
    out = a_out.context
    if self.params.attention_combiner is not None:
      c_out, _ = self.attention_combiner.FProp(
          theta.attention_combiner, NestedMap(),
          NestedMap(inputs=[q_out.output, a_out.context]), padding, NestedMap())
      out = c_out.output
    return (NestedMap(output=out, context=a_out.context, probs=a_out.probs,
                      query=q_out.output),
            NestedMap(attention=a_state, query_generator=q_state))
