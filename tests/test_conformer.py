"""Conv-with-time-padding and Conformer tests (CPU)."""

import pytest
import torch

from lingvo_b200.core import bn_layers
from lingvo_b200.core import conformer_layer
from lingvo_b200.core import conv_layers_with_time_padding as conv_lib
from lingvo_b200.core import layers
from lingvo_b200.core import py_utils
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.ops import conv as conv_ops


def test_output_padding_stride():
  pad = torch.tensor([[0., 0, 0, 0, 1, 1]])
  out = conv_lib.ComputeConvOutputPadding(pad, 3, 2)
  assert out.shape == (1, 3)
  assert out[0, 2] == 1 and out[0, 0] == 0


def test_conv2d_with_padding_ignores_padded_frames():
  p = conv_lib.Conv2DLayerWithPadding.Params().Set(
      name='c', filter_shape=(3, 3, 2, 4), filter_stride=(1, 1))
  l = p.Instantiate()
  x = torch.randn(2, 8, 5, 2)
  pad = torch.zeros(2, 8)
  pad[0, 5:] = 1
  y, yp = l.FPropDefaultTheta(x, pad)
  assert y.shape == (2, 8, 5, 4) and torch.equal(yp, pad)
  x2 = x.clone()
  x2[0, 5:] = 100.0               # garbage in padded frames must not leak
  y2, _ = l.FPropDefaultTheta(x2, pad)
  torch.testing.assert_close(y, y2)
  assert float(y[0, 5:].abs().max()) == 0


def test_causal_depthwise_is_causal_and_streams():
  p = conv_lib.CausalDepthwiseConv2DLayer.Params().Set(name='dw', filter_shape=(4, 1, 6, 1))
  l = p.Instantiate()
  x = torch.randn(2, 10, 1, 6)
  pad = torch.zeros(2, 10)
  y, _ = l.FPropDefaultTheta(x, pad)
  x2 = x.clone()
  x2[:, 7:] += 1.0
  y2, _ = l.FPropDefaultTheta(x2, pad)
  torch.testing.assert_close(y[:, :7], y2[:, :7])     # future does not affect the past
  st = l.zero_state(2)
  outs = []
  for i in range(0, 10, 5):
    o, _, st = l.StreamStep(l.theta, x[:, i:i + 5], pad[:, i:i + 5], st)
    outs.append(o)
  torch.testing.assert_close(torch.cat(outs, 1), y, atol=1e-6, rtol=1e-5)


def test_normalized_depthwise_weights_sum_to_one():
  p = conv_lib.NormalizedDepthwiseConv2DLayer.Params().Set(
      name='n', filter_shape=(5, 1, 2, 1), weight_tiling_factor=3)
  l = p.Instantiate()
  w = l._GetWeight(l.theta)
  assert w.shape == (5, 1, 6, 1)
  torch.testing.assert_close(w.sum(0), torch.ones(1, 6, 1))
  y, _ = l.FPropDefaultTheta(torch.randn(2, 7, 1, 6), torch.zeros(2, 7))
  assert y.shape == (2, 7, 1, 6)


def test_global_pooling():
  l = conv_lib.GlobalPoolingLayer.Params().Set(name='g', pooling_type='AVG').Instantiate()
  x = torch.ones(1, 4, 2, 3)
  x[0, 2:] = 50.0
  pad = torch.tensor([[0., 0, 1, 1]])
  out, op = l.FPropDefaultTheta(x, pad)
  torch.testing.assert_close(out, torch.ones(1, 1, 1, 3))


@pytest.mark.parametrize('causal', [False, True])
@pytest.mark.parametrize('k', [3, 8])
def test_glu_dwconv_ref_matches_layer_path(causal, k):
  d = 6
  cls = conv_lib.CausalDepthwiseConv2DLayer if causal else conv_lib.DepthwiseConv2DLayer
  l = cls.Params().Set(name='dw', filter_shape=(k, 1, d, 1)).Instantiate()
  proj = torch.randn(2, 9, 2 * d)
  pad = torch.zeros(2, 9)
  pad[1, 6:] = 1
  gated, act = proj.chunk(2, -1)
  want, _ = l.FPropDefaultTheta((act * torch.sigmoid(gated)).unsqueeze(2), pad)
  got = conv_ops.glu_dwconv1d_ref(proj, l.theta.w.reshape(k, d), pad, causal)
  torch.testing.assert_close(got, want.squeeze(2), atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize('norm', ['bn', 'ln', 'gn'])
def test_lconv_layer(norm):
  p = conformer_layer.LConvLayer.CommonParams(input_dim=8, kernel_size=3)
  p.name = 'lconv'
  if norm == 'ln':
    p.conv_norm_layer_tpl = layers.LayerNorm.Params()
  elif norm == 'gn':
    p.conv_norm_layer_tpl = bn_layers.GroupNormLayer.Params().Set(num_groups=2)
  l = p.Instantiate()
  x = torch.randn(2, 7, 8)
  pad = torch.zeros(2, 7)
  pad[1, 5:] = 1
  y, yp = l.FPropDefaultTheta(x, pad)
  assert y.shape == x.shape and torch.equal(yp, pad)
  y.sum().backward()


@pytest.mark.parametrize('order', ['mhsa_before_conv', 'conv_before_mhsa', 'mhsa', 'conv'])
def test_conformer_layer_orders(order):
  p = conformer_layer.ConformerLayer.CommonParams(
      input_dim=8, atten_num_heads=2 if order != 'conv' else None,
      kernel_size=3 if order != 'mhsa' else None, fflayer_hidden_dim=16,
      layer_order=order, use_relative_atten=(order != 'conv'))
  p.name = 'conf'
  p.params_init = py_utils.WeightInit.Xavier(1.0)
  if p.lconv_tpl is not None:
    p.lconv_tpl.conv_norm_layer_tpl = layers.LayerNorm.Params()
  l = p.Instantiate()
  x = torch.randn(2, 6, 8)
  pad = torch.zeros(2, 6)
  pad[0, 4:] = 1
  out = l.FPropDefaultTheta(NestedMap(features=x, paddings=pad))
  assert out.features.shape == x.shape
  assert float(out.features[0, 4:].abs().max()) == 0
  out.features.sum().backward()
  missing = [v.var_name for v in l.vars.Flatten() if v.grad is None and v.requires_grad]
  assert not missing, missing


def test_causal_conformer_stream_step_matches_fprop():
  p = conformer_layer.ConformerLayer.CommonParams(
      input_dim=8, atten_num_heads=2, atten_left_context=3, atten_right_context=0,
      kernel_size=3, fflayer_hidden_dim=16, is_causal=True, use_relative_atten=False)
  p.name = 'conf'
  p.lconv_tpl.conv_norm_layer_tpl = layers.LayerNorm.Params()
  p.trans_atten_tpl.atten_tpl.return_atten_probs = False
  l = p.Instantiate()
  x = torch.randn(2, 8, 8)
  pad = torch.zeros(2, 8)
  full = l.FPropDefaultTheta(NestedMap(features=x, paddings=pad)).features
  st = l.zero_state(2)
  outs = []
  for i in range(0, 8, 2):
    o, st = l.StreamStep(l.theta, NestedMap(features=x[:, i:i + 2], paddings=pad[:, i:i + 2]), st)
    outs.append(o.features)
  torch.testing.assert_close(torch.cat(outs, 1), full, atol=1e-4, rtol=1e-4)


def _ConfParams(**kw):
  p = conformer_layer.ConformerLayer.CommonParams(
      input_dim=8, atten_num_heads=2, kernel_size=3, fflayer_hidden_dim=16,
      use_relative_atten=False, **kw)
  p.name = 'conf'
  p.params_init = py_utils.WeightInit.Xavier(1.0)
  p.lconv_tpl.conv_norm_layer_tpl = layers.LayerNorm.Params()
  return p


def _ConfIn(b=2, t=6):
  x = torch.randn(b, t, 8)
  pad = torch.zeros(b, t)
  pad[0, 4:] = 1
  return NestedMap(features=x, paddings=pad)


def test_conformer_moe_ffn_reports_aux_loss_and_optional_start_ffn():
  from lingvo_b200.core import gshard_builder
  torch.manual_seed(0)
  p = _ConfParams()
  moe = conformer_layer.ConformerLayer.ConfigMoEParams(
      tpl=gshard_builder.MoEBuilder.Params().Set(e_dim=2, c_dim=8, num_devices=1,
                                                 moe_mode='dense'),
      input_dim=8, hidden_dim=16, activation='SWISH', residual_weight=0.5, dropout_prob=0.)
  p.fflayer_end_tpl = moe
  p.fflayer_start_tpl = None
  l = p.Instantiate()
  assert not l.has_fflayer_start and 'fflayer_end_moe' in l.children
  assert 'fflayer_start' not in l.children
  inp = _ConfIn()
  out = l.FPropDefaultTheta(inp)
  assert out.features.shape == inp.features.shape and out.aux_loss.dim() == 0
  inp.aux_loss = torch.ones(2)
  out2 = l.FPropDefaultTheta(inp)
  torch.testing.assert_close(out2.aux_loss, 1.0 + out.aux_loss.expand(2))
  (out2.features.sum() + out2.aux_loss.sum()).backward()
  assert l.vars.fflayer_end_moe.Flatten()[0].grad is not None
  with pytest.raises(ValueError):
    conformer_layer.ConformerLayer.ConfigMoEParams(
        tpl=gshard_builder.MoEBuilder.Params(), input_dim=8, hidden_dim=16,
        activation='RELU', residual_weight=1.0, dropout_prob=0.)
  assert conformer_layer.ConformerLayer.Stride(p) == 1
  assert conformer_layer.ConformerLayer.RightContext(
      NestedMap(atten_tpl=NestedMap(right_context=3), cls=None)) == 3


@pytest.mark.parametrize('pos', ['block_sequential', 'block_parallel', 'ff_sequential',
                                 'ff_parallel'])
def test_conformer_adapter_positions(pos):
  torch.manual_seed(0)
  p = _ConfParams()
  p.adapter_tpl = layers.MultitaskAdapterLayer.Params().Set(
      num_tasks=3, bottleneck_dim=4, data_format='BTC')
  p.adapter_pos = pos
  p.fflayer_task_ids = ''
  l = p.Instantiate()
  names = [n for n in l.children if 'adapter' in n]
  assert names == (['adapter'] if pos.startswith('block') else
                   ['fflayer_start_adapter', 'fflayer_end_adapter'])
  inp = _ConfIn()
  inp.task_ids = torch.tensor([0, 2])
  out = l.FPropDefaultTheta(inp)
  assert out.features.shape == inp.features.shape
  # Zero-initialised up-projection ⇒ the adapter starts as the identity of the block.
  base = _ConfParams().Instantiate()
  for dst, src in zip(base.vars.Flatten(), [v for v in l.vars.Flatten()
                                             if 'adapter' not in v.var_name]):
    dst.data.copy_(src.data)
  ref = base.FPropDefaultTheta(_ConfInLike(inp))
  up = [v for v in l.vars.Flatten() if 'adapter' in v.var_name and 'up' in v.var_name]
  if up and all(float(v.abs().max()) == 0 for v in up):
    torch.testing.assert_close(out.features, ref.features, atol=1e-5, rtol=1e-5)
  out.features.sum().backward()
  assert any(v.grad is not None for v in l.vars.Flatten() if 'adapter' in v.var_name)
  with pytest.raises(ValueError):
    bad = p.Copy().Set(adapter_pos='nowhere')
    bad.Instantiate()


def _ConfInLike(inp):
  return NestedMap(features=inp.features.clone(), paddings=inp.paddings.clone())


def test_conformer_weight_sharing_and_remat_match_plain():
  torch.manual_seed(0)
  p = _ConfParams(fflayer_weight_sharing=True)
  l = p.Instantiate()
  assert 'fflayer_end' not in l.children
  inp = _ConfIn()
  out = l.FPropDefaultTheta(inp)
  rp = p.Copy().Set(remat=True)
  r = rp.Instantiate()
  for dst, src in zip(r.vars.Flatten(), l.vars.Flatten()):
    dst.data.copy_(src.data)
  out_r = r.FPropDefaultTheta(_ConfInLike(inp))
  torch.testing.assert_close(out_r.features, out.features)
  out.features.sum().backward()
  out_r.features.sum().backward()
  for a, b in zip(l.vars.Flatten(), r.vars.Flatten()):
    torch.testing.assert_close(a.grad, b.grad, atol=1e-5, rtol=1e-5)


def test_lconv_stream_step_carries_cumulative_group_norm_state():
  torch.manual_seed(0)
  p = conformer_layer.LConvLayer.CommonParams(input_dim=8, kernel_size=3, is_causal=True)
  p.conv_norm_layer_tpl = bn_layers.GroupNormLayer.Params().Set(
      num_groups=2, cumulative=True, input_rank=4)
  p.name = 'lconv'
  p.params_init = py_utils.WeightInit.Xavier(1.0)
  l = p.Instantiate()
  x, pad = torch.randn(2, 8, 8), torch.zeros(2, 8)
  with torch.no_grad():
    full, _ = l.FPropDefaultTheta(x, pad)
    st = l.zero_state(2)
    assert 'norm_state' in st
    outs = []
    for i in range(0, 8, 2):
      o, _, st = l.StreamStep(l.theta, x[:, i:i + 2], pad[:, i:i + 2], st)
      outs.append(o)
  torch.testing.assert_close(torch.cat(outs, 1), full, atol=1e-4, rtol=1e-4)
  assert l._ApplyActivation(x, 'NONE') is x
