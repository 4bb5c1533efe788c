"""Step library tests (CPU)."""

import torch

from lingvo_b200.core import attention
from lingvo_b200.core import layers
from lingvo_b200.core import rnn_cell
from lingvo_b200.core import rnn_layers
from lingvo_b200.core import step
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.core.steps import attention_steps
from lingvo_b200.core.steps import embedding_steps
from lingvo_b200.core.steps import rnn_steps


def test_rnn_stack_step_matches_frnn_stack():
  cell = rnn_cell.LSTMCellSimple.Params()
  sp = rnn_steps.RnnStackStep.Params().Set(name='stack', rnn_cell_tpl=cell, step_input_dim=6,
                                           rnn_cell_dim=6, rnn_layers=2)
  wrapper = step.RecurrentStepWrapper.Params().Set(name='w', step=sp).Instantiate()
  x = torch.randn(5, 3, 6)
  pad = torch.zeros(5, 3, 1)
  prepared = wrapper.PrepareExternalInputs(wrapper.theta, NestedMap())
  s0 = wrapper.ZeroState(wrapper.theta, prepared, 3)
  out, _ = wrapper.FProp(wrapper.theta, prepared, NestedMap(inputs=[x]), pad, s0)
  # oracle: same cells driven by the FRNN layers
  ys = x
  for i in range(2):
    f = rnn_layers.FRNN.Params().Set(name='f%d' % i, cell=cell.Copy().Set(
        num_input_nodes=6, num_output_nodes=6)).Instantiate()
    ys, _ = f.FProp(NestedMap(cell=wrapper.theta.step.sub[i].cell), ys, pad)
  torch.testing.assert_close(out.output, ys, atol=1e-5, rtol=1e-5)


def test_attention_block_step():
  q = rnn_steps.RnnStep.Params().Set(name='q', cell=rnn_cell.LSTMCellSimple.Params().Set(
      num_input_nodes=4 + 8, num_output_nodes=5))
  blk = attention_steps.AttentionBlockStep.Params().Set(
      name='blk', query_generator=q,
      attention=attention_steps.AttentionStep.Params().Set(
          atten=attention.AdditiveAttention.Params().Set(source_dim=8, query_dim=5,
                                                         hidden_dim=7))).Instantiate()
  src = torch.randn(6, 2, 8)
  pad = torch.zeros(6, 2)
  prepared = blk.PrepareExternalInputs(blk.theta, NestedMap(
      attention=NestedMap(src=src, padding=pad)))
  st = blk.ZeroState(blk.theta, prepared, 2)
  for _ in range(3):
    out, st = blk.FProp(blk.theta, prepared, NestedMap(inputs=[torch.randn(2, 4)]),
                        torch.zeros(2, 1), st)
  assert out.output.shape == (2, 8) and out.probs.shape == (2, 6)
  torch.testing.assert_close(out.probs.sum(-1), torch.ones(2), atol=1e-5, rtol=1e-5)


def test_embedding_and_stateless_steps():
  e = embedding_steps.StatefulEmbeddingStep.Params().Set(
      name='e', target_vocab_size=10, embedding_dim=4, num_prev_tokens=2).Instantiate()
  st = e.ZeroState(e.theta, NestedMap(), 3)
  for t in range(3):
    out, st = e.FProp(e.theta, NestedMap(), NestedMap(inputs=[torch.tensor([1, 2, 3])]),
                      None, st)
  assert out.output.shape == (3, 4) and st.t == 3 and st.prev.shape == (3, 2)
  s = step.StatelessLayerStep.Params().Set(
      name='s', layer=layers.FCLayer.Params().Set(input_dim=4, output_dim=2)).Instantiate()
  o, _ = s.FProp(s.theta, NestedMap(), NestedMap(inputs=[torch.randn(3, 4)]), None, NestedMap())
  assert o.output.shape == (3, 2)


def _Fc(name, i, o):
  return step.StatelessLayerStep.Params().Set(
      name=name, layer=layers.FCLayer.Params().Set(input_dim=i, output_dim=o, activation='NONE'))


def test_stack_step_residuals_follow_the_stride():
  """output[i] = output[i - stride] + sub[i](output[i - 1]) for i >= residual_start."""
  p = step.StackStep.Params().Set(name='stack', sub=[_Fc('a', 4, 4), _Fc('b', 4, 4),
                                                    _Fc('c', 4, 4)],
                                  residual_start=1, residual_stride=2)
  s = p.Instantiate()
  x = torch.randn(3, 4)
  prepared = s.PrepareExternalInputs(s.theta, NestedMap())
  st0 = s.ZeroState(s.theta, prepared, 3)
  out, st1 = s.FProp(s.theta, prepared, NestedMap(inputs=[x]), torch.zeros(3, 1), st0)
  f = [lambda v, i=i: s.sub[i].layer.FPropDefaultTheta(v) for i in range(3)]
  o0 = f[0](x)
  o1 = f[1](o0) + x            # residual_inputs[1 + 1 - 2] = the stack input
  o2 = f[2](o1) + o0           # residual_inputs[2 + 1 - 2] = output[0]
  torch.testing.assert_close(out.output, o2, atol=1e-5, rtol=1e-5)
  assert len(st1.sub) == 3


def test_stack_step_context_reaches_every_layer():
  cell = rnn_cell.LSTMCellSimple.Params()
  subs = [rnn_steps.RnnStep.Params().Set(name='r%d' % i, cell=cell.Copy().Set(
      num_input_nodes=5 + 2, num_output_nodes=5)) for i in range(2)]
  s = step.StackStep.Params().Set(name='stack', sub=subs).Instantiate()
  prepared = s.PrepareExternalInputs(s.theta, NestedMap())
  st = s.ZeroState(s.theta, prepared, 3)
  x, ctx = torch.randn(3, 5), torch.randn(3, 2)
  out, st1 = s.FProp(s.theta, prepared, NestedMap(inputs=[x], context=ctx), torch.zeros(3, 1), st)
  out2, _ = s.FProp(s.theta, prepared, NestedMap(inputs=[x], context=ctx * 2),
                    torch.zeros(3, 1), st)
  assert out.output.shape == (3, 5)
  assert not torch.allclose(out.output, out2.output)


def test_parallel_step_concatenates():
  s = step.ParallelStep.Params().Set(name='par', sub=[_Fc('a', 4, 2), _Fc('b', 4, 3)]).Instantiate()
  prepared = s.PrepareExternalInputs(s.theta, NestedMap())
  st = s.ZeroState(s.theta, prepared, 5)
  x = torch.randn(5, 4)
  out, _ = s.FProp(s.theta, prepared, NestedMap(inputs=[x]), None, st)
  want = torch.cat([s.sub[0].layer.FPropDefaultTheta(x), s.sub[1].layer.FPropDefaultTheta(x)], 1)
  torch.testing.assert_close(out.output, want)


def test_iterator_step_walks_the_time_axis():
  it = step.IteratorStep.Params().Set(name='it', axis=1).Instantiate()
  seq = NestedMap(a=torch.arange(24.0).reshape(2, 3, 4), b=torch.arange(6).reshape(2, 3))
  prepared = it.PrepareExternalInputs(it.theta, seq)
  st = it.ZeroState(it.theta, prepared, 2)
  for t in range(3):
    out, st = it.FProp(it.theta, prepared, None, None, st)
    assert torch.equal(out.a, seq.a[:, t]) and torch.equal(out.b, seq.b[:, t])
  tm = step.IteratorStep.Params().Set(name='it0', axis=0).Instantiate()
  out, _ = tm.FProp(tm.theta, seq, None, None, NestedMap(t=torch.tensor(1)))
  assert torch.equal(out.a, seq.a[1])


def test_graph_step_wires_iterator_rnn_and_attention():
  """A decoder-style graph: an iterator feeds an RNN whose output queries attention; the
  graph output exposes both (ref step_test.py GraphStep cases)."""
  sub = [
      step.SubStep('(inputs=[step_inputs.extra])->unused_probe', None, _Fc('probe', 3, 3)),
      step.SubStep('step_inputs->iter', 'external_inputs.seq',
                   step.IteratorStep.Params().Set(name='iter', axis=1)),
      step.SubStep('(inputs=[iter.x,step_inputs.extra])->rnn', None,
                   rnn_steps.RnnStep.Params().Set(name='rnn', cell=rnn_cell.LSTMCellSimple.Params(
                   ).Set(num_input_nodes=4 + 3, num_output_nodes=5))),
      step.SubStep('(inputs=[rnn.output])->atten', 'external_inputs.memory',
                   attention_steps.AttentionStep.Params().Set(
                       name='atten', atten=attention.AdditiveAttention.Params().Set(
                           source_dim=8, query_dim=5, hidden_dim=7))),
  ]
  g = step.GraphStep.Params().Set(
      name='graph', sub=sub, output_signature='(rnn=rnn.output,ctx=atten.context)').Instantiate()
  b, t = 2, 3
  ext = NestedMap(seq=NestedMap(x=torch.randn(b, t, 4)),
                  memory=NestedMap(src=torch.randn(6, b, 8), padding=torch.zeros(6, b)))
  prepared = g.PrepareExternalInputs(g.theta, ext)
  assert sorted(prepared.keys()) == ['atten', 'iter', 'probe', 'rnn']
  assert len(prepared.probe) == 0 and 'packed_src' in prepared.atten
  st = g.ZeroState(g.theta, prepared, b)
  extra = torch.randn(b, 3)
  # manual oracle with the same children
  rnn_st = g.rnn.ZeroState(g.theta.rnn, NestedMap(), b)
  att_st = g.atten.ZeroState(g.theta.atten, prepared.atten, b)
  for i in range(t):
    out, st = g.FProp(g.theta, prepared, NestedMap(extra=extra), torch.zeros(b, 1), st)
    r, rnn_st = g.rnn.FProp(g.theta.rnn, None, NestedMap(inputs=[ext.seq.x[:, i], extra]),
                            torch.zeros(b, 1), rnn_st)
    a, att_st = g.atten.FProp(g.theta.atten, prepared.atten, NestedMap(inputs=[r.output]),
                              torch.zeros(b, 1), att_st)
    torch.testing.assert_close(out.rnn, r.output)
    torch.testing.assert_close(out.ctx, a.context)
  assert st.iter.t == t
  out.ctx.sum().backward()
  assert g.rnn.vars.Flatten()[0].grad is not None


def test_graph_step_rejects_duplicate_outputs():
  import pytest
  p = step.GraphStep.Params().Set(name='g', output_signature='a', sub=[
      step.SubStep('step_inputs->a', None, _Fc('x', 2, 2)),
      step.SubStep('step_inputs->a', None, _Fc('y', 2, 2))])
  with pytest.raises(AssertionError):
    p.Instantiate()


def test_recurrent_step_wrapper_accumulates_states():
  cell = rnn_cell.LSTMCellSimple.Params().Set(num_input_nodes=4, num_output_nodes=3)
  w = step.RecurrentStepWrapper.Params().Set(
      name='w', step=rnn_steps.RnnStep.Params().Set(name='r', cell=cell)).Instantiate()
  prepared = w.PrepareExternalInputs(w.theta, NestedMap())
  s0 = w.ZeroState(w.theta, prepared, 2)
  x = torch.randn(5, 2, 4)
  out, states = w.FProp(w.theta, prepared, NestedMap(inputs=[x]), torch.zeros(5, 2, 1), s0)
  assert out.output.shape == (5, 2, 3)
  assert states.m.shape == (5, 2, 3) and states.c.shape[0] == 5
  torch.testing.assert_close(states.m, out.output)
