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
