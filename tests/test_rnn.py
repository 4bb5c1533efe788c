"""RNN cell / layer tests (CPU): shapes, padding carry-over, hoisted FRNN ≡
step-by-step FProp, packed-input reset, bidirectional symmetry, LSTM vs torch."""

import pytest
import torch

from lingvo_b200.core import attention
from lingvo_b200.core import py_utils
from lingvo_b200.core import recurrent
from lingvo_b200.core import rnn_cell
from lingvo_b200.core import rnn_layers
from lingvo_b200.core.nested_map import NestedMap

CELLS = [
    (rnn_cell.LSTMCellSimple, {}),
    (rnn_cell.LSTMCellSimple, dict(num_hidden_nodes=10, couple_input_forget_gates=True)),
    (rnn_cell.LayerNormalizedLSTMCellSimple, {}),
    (rnn_cell.LayerNormalizedLSTMCellLean, {}),
    (rnn_cell.DoubleProjectionLSTMCell, dict(num_input_hidden_nodes=4, num_hidden_nodes=8)),
    (rnn_cell.WeightNormalizedLSTMCellSimple, {}),
    (rnn_cell.SRUCell, {}),
    (rnn_cell.SRUCell, dict(couple_input_forget_gates=False, pointwise_peephole=True,
                            apply_layer_norm=True)),
    (rnn_cell.GRUCell, {}),
    (rnn_cell.LayerNormalizedLSTMCell, {}),
    (rnn_cell.NormalizedLSTMCellSimple, dict(enable_lstm_bias=False)),
    (rnn_cell.LSTMCellGrouped, dict(num_groups=2, num_shuffle_shards=2)),
]


def _Cell(cls, kw, idim=6, odim=6):
  p = cls.Params().Set(name='cell', num_input_nodes=idim, num_output_nodes=odim, **kw)
  p.params_init = py_utils.WeightInit.Uniform(0.5)
  return p


@pytest.mark.parametrize('cls,kw', CELLS)
def test_cell_step_and_padding(cls, kw):
  cell = _Cell(cls, kw).Instantiate()
  b = 3
  s0 = cell.zero_state(cell.theta, b)
  s0 = s0.Transform(lambda x: x + 0.1)
  x = torch.randn(b, 6)
  pad = torch.tensor([[0.], [1.], [0.]])
  s1, _ = cell.FProp(cell.theta, s0, NestedMap(act=[x], padding=pad))
  assert s1.m.shape == (b, 6)
  # padded row keeps its state exactly
  torch.testing.assert_close(s1.m[1], s0.m[1])
  torch.testing.assert_close(s1.c[1], s0.c[1])
  assert not torch.allclose(s1.m[0], s0.m[0])


@pytest.mark.parametrize('cls,kw', CELLS[:11])
def test_frnn_equals_stepwise(cls, kw):
  p = rnn_layers.FRNN.Params().Set(name='frnn', cell=_Cell(cls, kw))
  l = p.Instantiate()
  t, b = 5, 2
  x = torch.randn(t, b, 6)
  pad = torch.zeros(t, b, 1)
  pad[3:, 1] = 1.0
  out, final = l.FPropDefaultTheta(x, pad)
  state = l.zero_state(l.theta, b)
  outs = []
  for i in range(t):
    state, _ = l.cell.FProp(l.theta.cell, state, NestedMap(act=[x[i]], padding=pad[i]))
    outs.append(l.cell.GetOutput(state))
  torch.testing.assert_close(out, torch.stack(outs), atol=1e-5, rtol=1e-5)
  torch.testing.assert_close(final.m, state.m, atol=1e-5, rtol=1e-5)
  out.sum().backward()
  assert all(v.grad is not None for v in l.vars.Flatten())


def test_lstm_matches_torch_lstm():
  cell_p = _Cell(rnn_cell.LSTMCellSimple, dict(cell_value_cap=None), idim=4, odim=5)
  l = rnn_layers.FRNN.Params().Set(name='f', cell=cell_p).Instantiate()
  ref = torch.nn.LSTM(4, 5)
  wm, b = l.theta.cell.wm.detach(), l.theta.cell.b.detach()
  # ours: gate order (g, i, f, o) along columns; torch: (i, f, g, o) rows
  def reorder(w):
    g, i, f, o = w.chunk(4, -1)
    return torch.cat([i, f, g, o], -1)
  with torch.no_grad():
    ref.weight_ih_l0.copy_(reorder(wm[:4]).t())
    ref.weight_hh_l0.copy_(reorder(wm[4:]).t())
    ref.bias_ih_l0.copy_(reorder(b))
    ref.bias_hh_l0.zero_()
  x = torch.randn(7, 3, 4)
  out, _ = l.FPropDefaultTheta(x, torch.zeros(7, 3, 1))
  want, _ = ref(x)
  torch.testing.assert_close(out, want, atol=1e-5, rtol=1e-5)


def test_reverse_and_bidirectional():
  cell = _Cell(rnn_cell.LSTMCellSimple, {})
  fwd = rnn_layers.FRNN.Params().Set(name='f', cell=cell).Instantiate()
  bak = rnn_layers.FRNN.Params().Set(name='f', cell=cell, reverse=True).Instantiate()
  x = torch.randn(4, 2, 6)
  pad = torch.zeros(4, 2, 1)
  o_f, _ = fwd.FPropDefaultTheta(x, pad)
  o_b, _ = bak.FProp(fwd.theta, torch.flip(x, [0]), pad)
  torch.testing.assert_close(o_f, torch.flip(o_b, [0]), atol=1e-6, rtol=1e-6)
  bi = rnn_layers.BidirectionalFRNN.Params().Set(name='bi', fwd=cell, bak=cell).Instantiate()
  assert bi.FPropDefaultTheta(x, pad).shape == (4, 2, 12)


def test_packed_input_resets_state():
  cell = _Cell(rnn_cell.LSTMCellSimple, {})
  l = rnn_layers.FRNN.Params().Set(name='f', cell=cell, packed_input=True).Instantiate()
  x = torch.randn(6, 1, 6)
  pad = torch.zeros(6, 1, 1)
  seg = torch.tensor([1, 1, 1, 2, 2, 2]).view(6, 1, 1)
  out, _ = l.FPropDefaultTheta(x, pad, segment_id=seg)
  plain = rnn_layers.FRNN.Params().Set(name='f', cell=cell).Instantiate()
  second, _ = plain.FProp(l.theta, x[3:], pad[3:])
  torch.testing.assert_close(out[3:], second, atol=1e-6, rtol=1e-6)


def test_stacked_layers():
  p = rnn_layers.StackedFRNNLayerByLayer.Params().Set(
      name='s', num_layers=3, num_input_nodes=6, num_output_nodes=6,
      cell_tpl=rnn_cell.LSTMCellSimple.Params())
  l = p.Instantiate()
  out, st = l.FPropDefaultTheta(torch.randn(4, 2, 6), torch.zeros(4, 2, 1))
  assert out.shape == (4, 2, 6) and len(st.rnn) == 3
  p = rnn_layers.StackedBiFRNNLayerByLayer.Params().Set(
      name='sb', num_layers=2, num_input_nodes=6, num_output_nodes=8,
      cell_tpl=rnn_cell.LSTMCellSimple.Params())
  assert p.Instantiate().FPropDefaultTheta(
      torch.randn(4, 2, 6), torch.zeros(4, 2, 1)).shape == (4, 2, 8)


def test_frnn_with_attention():
  cell = _Cell(rnn_cell.LSTMCellSimple, {}, idim=6 + 8, odim=6)
  att = attention.AdditiveAttention.Params().Set(source_dim=8, query_dim=6, hidden_dim=5)
  l = rnn_layers.FRNNWithAttention.Params().Set(
      name='fa', cell=cell, attention=att).Instantiate()
  src = torch.randn(7, 2, 8)
  src_pad = torch.zeros(7, 2)
  src_pad[5:, 0] = 1
  x = torch.randn(4, 2, 6)
  ctx, out, probs, final = l.FPropDefaultTheta(src, src_pad, x, torch.zeros(4, 2, 1))
  assert ctx.shape == (4, 2, 8) and out.shape == (4, 2, 6) and probs.shape == (4, 2, 7)
  assert float(probs[:, 0, 5:].abs().max()) == 0
  torch.testing.assert_close(probs.sum(-1), torch.ones(4, 2), atol=1e-5, rtol=1e-5)


def test_recurrent_remat_matches_plain():
  cell = _Cell(rnn_cell.LSTMCellSimple, {}).Instantiate()
  x = torch.randn(6, 2, 6)
  pad = torch.zeros(6, 2, 1)
  inputs = NestedMap(act=[x], padding=pad)
  s0 = cell.zero_state(cell.theta, 2)

  def cell_fn(theta, state, inp):
    return cell.FProp(theta, state, NestedMap(act=inp.act, padding=inp.padding))

  def run(remat):
    for v in cell.vars.Flatten():
      v.grad = None
    acc, final = recurrent.Recurrent(cell.theta, s0, inputs, cell_fn, remat_steps=remat)
    acc.m.sum().backward()
    return acc.m.detach().clone(), [v.grad.clone() for v in cell.vars.Flatten()]
  a0, g0 = run(0)
  a1, g1 = run(2)
  torch.testing.assert_close(a0, a1)
  for x0, x1 in zip(g0, g1):
    torch.testing.assert_close(x0, x1, atol=1e-6, rtol=1e-5)


def test_qrnn_pooling():
  cell = rnn_cell.QRNNPoolingCell.Params().Set(
      name='q', num_input_nodes=12, num_output_nodes=4, pooling_formula='fo').Instantiate()
  s0 = cell.zero_state(cell.theta, 2)
  s1, _ = cell.FProp(cell.theta, s0, NestedMap(act=[torch.randn(2, 12)],
                                               padding=torch.zeros(2, 1)))
  assert s1.m.shape == (2, 4)


def test_layer_normalized_lstm_cell_packs_bias_and_scale_in_b():
  """`b` = [4 gate biases | 4 LN scales (offset 1)]; matches a hand-written LN-LSTM."""
  cell = _Cell(rnn_cell.LayerNormalizedLSTMCell, dict(forget_gate_bias=0.5)).Instantiate()
  assert cell.vars.b.shape == (8 * 6,)
  with torch.no_grad():
    cell.vars.b.copy_(torch.randn(48) * 0.1)
  b = 3
  s0 = cell.zero_state(cell.theta, b).Transform(lambda x: x + 0.2)
  x = torch.randn(b, 6)
  s1, _ = cell.FProp(cell.theta, s0, NestedMap(act=[x], padding=torch.zeros(b, 1)))
  z = (torch.cat([x, s0.m], 1) @ cell.vars.wm).reshape(b, 4, 6)
  z = (z - z.mean(-1, keepdim=True)) * torch.rsqrt(z.var(-1, unbiased=False, keepdim=True) + 1e-8)
  z = z * (cell.vars.b[24:].reshape(4, 6) + 1.0) + cell.vars.b[:24].reshape(4, 6)
  i_i, i_g, f_g, o_g = z.unbind(1)
  c = torch.sigmoid(f_g + 0.5) * s0.c + torch.sigmoid(i_g) * torch.tanh(i_i)
  torch.testing.assert_close(s1.c, c.clamp(-10, 10), atol=1e-5, rtol=1e-5)
  torch.testing.assert_close(s1.m, torch.sigmoid(o_g) * torch.tanh(c), atol=1e-5, rtol=1e-5)


def test_layer_normalized_lstm_cell_cc_schedule_caps_the_cell():
  from lingvo_b200.core import quant_utils
  p = _Cell(rnn_cell.LayerNormalizedLSTMCell, dict(
      cc_schedule=quant_utils.LinearClippingCapSchedule.Params().Set(
          start_step=0, end_step=10, start_cap=5.0, end_cap=0.05)))
  cell = p.Instantiate()
  py_utils.SetGlobalStep(10)
  try:
    s0 = cell.zero_state(cell.theta, 2).Transform(lambda x: x + 3.0)
    s1, _ = cell.FProp(cell.theta, s0, NestedMap(act=[torch.randn(2, 6)],
                                                 padding=torch.zeros(2, 1)))
    assert float(s1.c.abs().max()) <= 0.05 + 1e-6
  finally:
    py_utils.SetGlobalStep(0)


def test_normalized_lstm_cell_uses_per_gate_norm_layers():
  from lingvo_b200.core import layers
  cell = _Cell(rnn_cell.NormalizedLSTMCellSimple, dict(enable_lstm_bias=False)).Instantiate()
  assert sorted(k for k in cell.children.keys() if k.startswith('norm_')) == [
      'norm_f_g', 'norm_i_g', 'norm_i_i', 'norm_o_g']
  assert isinstance(cell.norm_i_i, layers.LayerNorm)
  with pytest.raises(AssertionError):
    _Cell(rnn_cell.NormalizedLSTMCellSimple, {}).Instantiate()       # bias must be off
  b = 2
  s0 = cell.zero_state(cell.theta, b).Transform(lambda x: x + 0.1)
  x = torch.randn(b, 6)
  s1, _ = cell.FProp(cell.theta, s0, NestedMap(act=[x], padding=torch.zeros(b, 1)))
  z = (torch.cat([x, s0.m], 1) @ cell.vars.wm).chunk(4, -1)
  n = [getattr(cell, 'norm_' + g).FPropDefaultTheta(v) for g, v in
       zip(('i_i', 'i_g', 'f_g', 'o_g'), z)]
  c = torch.sigmoid(n[2]) * s0.c + torch.sigmoid(n[1]) * torch.tanh(n[0])
  torch.testing.assert_close(s1.c, c, atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize('method', ['add', 'concat'])
def test_embedding_augmented_cell(method):
  from lingvo_b200.core.steps import embedding_steps
  emb_dim = 6 if method == 'add' else 2
  p = _Cell(rnn_cell.EmbeddingAugmentedLayerNormalizedLSTMCellSimple, dict(
      inject_emb_method=method,
      emb=embedding_steps.StatefulEmbeddingStep.Params().Set(
          name='emb', target_vocab_size=11, embedding_dim=emb_dim, num_prev_tokens=1)))
  cell = p.Instantiate()
  b = 3
  state = cell.zero_state(cell.theta, b)
  assert 'emb' in state and state.emb.prev.shape == (b, 1)
  x = torch.randn(b, 6)
  if method == 'concat':
    x[:, -emb_dim:] = 0.0
  ids = torch.tensor([1, 2, 3])
  s1, _ = cell.FProp(cell.theta, state, NestedMap(act=[x], padding=torch.zeros(b, 1), ids=ids))
  s1b, _ = cell.FProp(cell.theta, state, NestedMap(act=[x], padding=torch.zeros(b, 1),
                                                   ids=torch.tensor([4, 5, 6])))
  assert s1.m.shape == (b, 6) and not torch.allclose(s1.m, s1b.m)   # ids matter
  assert s1.emb.t == 1 and torch.equal(s1.emb.prev[:, -1], ids)
  # equals the plain LN cell fed with act + (padded) embedding
  e, _ = cell.emb.FProp(cell.theta.emb, None, NestedMap(inputs=[ids]), None, state.emb)
  e = e.output
  if method == 'concat':
    e = torch.nn.functional.pad(e, (6 - emb_dim, 0))
  ref, _ = rnn_cell.LayerNormalizedLSTMCellSimple.FProp(
      cell, cell.theta, NestedMap(m=state.m, c=state.c),
      NestedMap(act=[x + e], padding=torch.zeros(b, 1)))
  torch.testing.assert_close(s1.m, ref.m)
  # a second step sees the first step's id as "previous token"
  s2, _ = cell.FProp(cell.theta, s1, NestedMap(act=[x], padding=torch.zeros(b, 1), ids=ids))
  assert s2.emb.t == 2


def test_bidirectional_rnn_v2_pads_to_sequence_length():
  fwd = _Cell(rnn_cell.LSTMCellSimple, {})
  p = rnn_layers.BidirectionalRNNV2.Params().Set(name='bi', fwd=fwd, bak=fwd.Copy(),
                                                 sequence_length=8)
  l = p.Instantiate()
  t, b = 5, 2
  x = torch.randn(t, b, 6)
  pad = torch.zeros(t, b, 1)
  pad[4:, 1] = 1.0
  out = l.FPropDefaultTheta(x, pad)
  assert out.shape == (t, b, 12)
  ref = l.brnn.FProp(l.theta.brnn, x, pad)          # unpadded run of the same weights
  torch.testing.assert_close(out, ref, atol=1e-6, rtol=1e-5)
  with pytest.raises(AssertionError):
    l.FPropDefaultTheta(torch.randn(9, b, 6), torch.zeros(9, b, 1))


# ----------------------------------------------------------------------- recurrent.py --
def _GruLikeCell(theta, state0, inputs):
  h = torch.tanh(inputs.x @ theta.w + state0.h @ theta.u)
  pad = inputs.padding
  h = torch.where(pad > 0, state0.h, h)
  return NestedMap(h=h), NestedMap(pre=h * 2.0)


def _RecInputs(t=6, b=3, d=4):
  g = torch.Generator().manual_seed(0)
  theta = NestedMap(w=torch.randn(d, d, generator=g).requires_grad_(True),
                    u=(torch.randn(d, d, generator=g) * 0.3).requires_grad_(True))
  x = torch.randn(t, b, d, generator=g).requires_grad_(True)
  pad = torch.zeros(t, b, 1)
  pad[4:, 0] = 1.0
  return theta, NestedMap(h=torch.zeros(b, d)), NestedMap(x=x, padding=pad)


def test_recurrent_acc_extras_stop_fn_and_single_step():
  theta, s0, inp = _RecInputs()
  acc, final, ex = recurrent.Recurrent(theta, s0, inp, _GruLikeCell, return_acc_extras=True)
  assert acc.h.shape == (6, 3, 4) and ex.pre.shape == (6, 3, 4)
  torch.testing.assert_close(acc.h[-1], final.h)
  torch.testing.assert_close(ex.pre, acc.h * 2.0)
  torch.testing.assert_close(acc.h[5, 0], acc.h[3, 0])            # padded rows carry state
  # stop after 3 steps: the tail repeats the last computed state
  acc2, final2 = recurrent.Recurrent(theta, s0, inp, _GruLikeCell,
                                     stop_fn=lambda t, th, st: t >= 3)
  torch.testing.assert_close(acc2.h[:3], acc.h[:3])
  torch.testing.assert_close(acc2.h[5], acc.h[2])
  torch.testing.assert_close(final2.h, acc.h[2])
  one = NestedMap(x=inp.x[:1], padding=inp.padding[:1])
  a1, f1 = recurrent.Recurrent(theta, s0, one, _GruLikeCell, remat_steps=2)
  torch.testing.assert_close(a1.h[0], acc.h[0])


def test_recurrent_custom_cell_grad_matches_autograd():
  theta, s0, inp = _RecInputs()
  acc, _ = recurrent.Recurrent(theta, s0, inp, _GruLikeCell)
  acc.h.pow(2).sum().backward()
  want = (theta.w.grad.clone(), theta.u.grad.clone(), inp.x.grad.clone())
  theta.w.grad = theta.u.grad = inp.x.grad = None
  calls = []

  def CellGrad(th, state0, inputs, extras, dstate1):
    calls.append(1)
    pre = inputs.x @ th.w + state0.h @ th.u
    h = torch.tanh(pre)
    live = (inputs.padding <= 0).float()
    dpre = dstate1.h * live * (1 - h * h)
    dth = NestedMap(w=inputs.x.t() @ dpre, u=state0.h.t() @ dpre)
    dst = NestedMap(h=dpre @ th.u.t() + dstate1.h * (1 - live))
    din = NestedMap(x=dpre @ th.w.t(), padding=None)
    return dth, dst, din, None

  acc2, _ = recurrent.Recurrent(theta, s0, inp, _GruLikeCell, cell_grad=CellGrad)
  torch.testing.assert_close(acc2.h, acc.h.detach())
  acc2.h.pow(2).sum().backward()
  assert len(calls) == 6
  torch.testing.assert_close(theta.w.grad, want[0], atol=1e-5, rtol=1e-4)
  torch.testing.assert_close(theta.u.grad, want[1], atol=1e-5, rtol=1e-4)
  torch.testing.assert_close(inp.x.grad, want[2], atol=1e-5, rtol=1e-4)


def test_recurrent_accumulators_and_step_seeds():
  from lingvo_b200.core import base_layer
  from lingvo_b200.core import py_utils

  class Counter(base_layer.Accumulator):

    def DefaultValue(self):
      return torch.zeros(())

  class Holder(base_layer.BaseLayer):

    def __init__(self, params):
      super().__init__(params)
      self.RegisterAccumulator('count', Counter())

  layer = Holder.Params().Set(name='holder').Instantiate()
  seeds = []

  def Cell(theta, state0, inputs):
    acc = layer.accumulators.count
    acc.Update(acc.GetValue() + inputs.x.sum() * 0 + 1.0)
    seeds.append(py_utils.GetStepSeed())
    return _GruLikeCell(theta, state0, inputs)

  theta, s0, inp = _RecInputs()
  py_utils.ResetStepSeed(100)
  acc, final = recurrent.Recurrent(theta, s0, inp, Cell, accumulator_layer=layer)
  assert 'accumulators' not in final and 'accumulators' not in acc
  assert float(layer.accumulators.count.GetValue()) == 6.0        # carried to the layer
  assert seeds == list(range(100, 106)) and py_utils.GetStepSeed() == 106
  # rematerialised: the re-run in backward sees the same per-step seeds
  seeds.clear()
  py_utils.ResetStepSeed(100)
  acc2, _ = recurrent.Recurrent(theta, s0, inp, Cell, remat_steps=2)
  acc2.h.sum().backward()
  assert seeds[:6] == list(range(100, 106)) and sorted(set(seeds)) == list(range(100, 106))
  assert len(seeds) == 12
  torch.testing.assert_close(acc2.h, acc.h)
  with pytest.raises(ValueError):
    bad = NestedMap(s0)
    bad.accumulators = NestedMap()
    recurrent.Recurrent(theta, bad, inp, Cell, accumulator_layer=layer)
  py_utils.ResetStepSeed(0)


def test_stacked_recurrent_matches_layer_by_layer():
  theta, s0, inp = _RecInputs()
  g = torch.Generator().manual_seed(9)
  theta2 = NestedMap(w=torch.randn(4, 4, generator=g), u=torch.randn(4, 4, generator=g) * 0.3)
  out_fn = lambda st: NestedMap(x=st.h)
  acc, finals = recurrent.StackedRecurrent(
      [None, None], [_GruLikeCell, _GruLikeCell], [None, None], [out_fn, out_fn], [None, None],
      [theta, theta2], [s0, s0], inp)
  a1, f1 = recurrent.Recurrent(theta, s0, inp, _GruLikeCell)
  a2, f2 = recurrent.Recurrent(theta2, s0, NestedMap(x=a1.h, padding=inp.padding), _GruLikeCell)
  torch.testing.assert_close(acc.x, a2.h)
  torch.testing.assert_close(finals[0].h, f1.h)
  torch.testing.assert_close(finals[1].h, f2.h)
  none_acc, _ = recurrent.StackedRecurrent(
      [None, None], [_GruLikeCell, _GruLikeCell], [None, None], [out_fn, out_fn], [None, None],
      [theta, theta2], [s0, s0], inp, unused_acc_state=True)
  assert none_acc is None


def test_frnn_with_attention_accumulate_and_postprocess_equal_fprop():
  from lingvo_b200.core import attention as attention_lib
  for prev in (False, True):
    p = rnn_layers.FRNNWithAttention.Params().Set(
        name='fa', output_prev_atten_ctx=prev, use_zero_atten_state=True,
        atten_context_dim=5, random_seed=8,
        cell=rnn_cell.LSTMCellSimple.Params().Set(name='c', num_input_nodes=3 + 5,
                                                  num_output_nodes=6),
        attention=attention_lib.AdditiveAttention.Params().Set(
            source_dim=5, query_dim=6, hidden_dim=4))
    layer = p.Instantiate()
    src, spad = torch.randn(4, 2, 5), torch.zeros(4, 2)
    x, pad = torch.randn(3, 2, 3), torch.zeros(3, 2, 1)
    ctx, out, probs, final = layer.FProp(layer.theta, src, spad, x, pad)
    acc, final2, side = layer.AccumulateStates(layer.theta, src, spad, x, pad)
    ctx2, out2, probs2 = layer.PostProcessStates(acc, side)
    torch.testing.assert_close(ctx2, ctx)
    torch.testing.assert_close(out2, out)
    torch.testing.assert_close(probs2, probs)
    torch.testing.assert_close(final2.atten, final.atten)
    assert acc.atten_probs.shape == (3, 2, 4)
  packed = layer.InitAttention(layer.theta, src, spad)
  assert packed is not None
  st = NestedMap(atten=torch.ones(2, 5), atten_probs=torch.ones(2, 4), atten_state=torch.ones(2, 1))
  st = layer.reset_atten_state(layer.theta, st, NestedMap(reset_mask=torch.tensor([[0.0], [1.0]])))
  assert st.atten[0].sum() == 0 and st.atten[1].sum() == 5 and st.atten_probs[0].sum() == 0
  with pytest.raises(ValueError):
    layer.reset_atten_state(layer.theta, NestedMap(
        atten=torch.ones(2, 5), atten_probs=torch.ones(2, 4),
        atten_state=NestedMap(other=torch.ones(2))), NestedMap(reset_mask=torch.ones(2, 1)))
  stack = rnn_layers.StackedFRNNLayerByLayer.Params().Set(
      name='st', num_layers=2, num_input_nodes=3, num_output_nodes=3,
      cell_tpl=rnn_cell.LSTMCellSimple.Params().Set(
          num_input_nodes=3, num_output_nodes=3)).Instantiate()
  full = stack.FPropFullSequence(stack.theta, x, pad)
  assert full.shape == (3, 2, 3)
