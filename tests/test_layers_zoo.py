"""Per-class tests of `core/layers.py` (ref `lingvo/core/layers_test.py`): every layer runs
FProp against a plain-PyTorch oracle of the same math (and backward where it has weights);
hard-coded values are the reference's goldens where they do not depend on TF's RNG."""

import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from lingvo_b200.core import cluster_factory
from lingvo_b200.core import layers
from lingvo_b200.core import py_utils
from lingvo_b200.core.nested_map import NestedMap


def _Make(cls, **kw):
  name = kw.pop('name', cls.__name__.lower())
  layer = cls.Params().Set(name=name, **kw).Instantiate()
  layer.InstantiateVariables()
  return layer


def _GradsFlow(layer, out):
  out = out if isinstance(out, torch.Tensor) else out[0]
  out.float().sum().backward()
  vs = [v for v in layer.vars.Flatten() if v.requires_grad]
  assert vs and all(v.grad is not None for v in vs), [v.var_name for v in vs if v.grad is None]


# ------------------------------------------------------------------------- identity --
def test_identity_layer_nested_map():
  l = _Make(layers.IdentityLayer)
  x = NestedMap(a=torch.ones(2), b=NestedMap(c=torch.zeros(3)))
  y = l.FPropDefaultTheta(x)
  assert y.a is x.a and y.b.c is x.b.c
  a, b = l.FPropDefaultTheta(torch.ones(1), torch.zeros(1))
  assert float(a) == 1 and float(b) == 0


# ---------------------------------------------------------------------------- convs --
def _ConvOracle(x, w, stride, dilation=(1, 1), depthwise=False):
  """TF 'SAME' conv on NHWC input with HWIO (or HWI·mult for depthwise) filters."""
  kh, kw, cin, cout = w.shape
  n, h, wd, _ = x.shape
  eff_h, eff_w = (kh - 1) * dilation[0] + 1, (kw - 1) * dilation[1] + 1
  ph = max((-(-h // stride[0]) - 1) * stride[0] + eff_h - h, 0)
  pw = max((-(-wd // stride[1]) - 1) * stride[1] + eff_w - wd, 0)
  xp = F.pad(x.permute(0, 3, 1, 2), (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
  if depthwise:
    wt = w.permute(2, 3, 0, 1).reshape(cin * cout, 1, kh, kw)
    y = F.conv2d(xp, wt, stride=stride, dilation=dilation, groups=cin)
  else:
    y = F.conv2d(xp, w.permute(3, 2, 0, 1), stride=stride, dilation=dilation)
  return y.permute(0, 2, 3, 1)


@pytest.mark.parametrize('stride,dilation', [((1, 1), (1, 1)), ((2, 2), (1, 1)),
                                             ((1, 1), (2, 2))])
def test_conv2d_layer_matches_torch_conv(stride, dilation):
  l = _Make(layers.Conv2DLayer, filter_shape=(3, 3, 3, 5), filter_stride=stride,
            dilation_rate=dilation, batch_norm=False, activation='NONE', bias=True)
  x = torch.randn(2, 8, 6, 3)
  pad = torch.zeros(2, 8)
  y, out_pad = l.FPropDefaultTheta(x, pad)
  want = _ConvOracle(x, l.vars.w.detach(), stride, dilation) + l.vars.b.detach()
  torch.testing.assert_close(y, want, atol=1e-4, rtol=1e-4)
  assert out_pad.shape == (2, want.shape[1])
  assert tuple(l.OutShape((2, 8, 6, 3))) == tuple(want.shape)
  _GradsFlow(l, y)


def test_conv2d_layer_paddings_zero_the_output_and_shrink_with_stride():
  l = _Make(layers.Conv2DLayer, filter_shape=(3, 3, 1, 4), filter_stride=(2, 2),
            batch_norm=False, activation='RELU')
  x = torch.randn(2, 9, 4, 1)
  pad = torch.zeros(2, 9)
  pad[1, 4:] = 1.0
  y, out_pad = l.FPropDefaultTheta(x, pad)
  assert y.shape == (2, 5, 2, 4) and out_pad.shape == (2, 5)
  assert float(out_pad[1, 2:].min()) == 1.0 and float(out_pad[0].max()) == 0.0
  assert float(y[1, 3:].abs().max()) == 0.0               # padded frames stay zero
  assert (y >= 0).all()


def test_depthwise_and_separable_conv():
  d = _Make(layers.DepthwiseConv2DLayer, filter_shape=(3, 3, 4, 2), batch_norm=False,
            activation='NONE')
  x = torch.randn(2, 7, 5, 4)
  y, _ = d.FPropDefaultTheta(x, torch.zeros(2, 7))
  want = _ConvOracle(x, d.vars.w.detach(), (1, 1), depthwise=True)
  torch.testing.assert_close(y, want, atol=1e-4, rtol=1e-4)
  assert y.shape[-1] == 8
  s = _Make(layers.SeparableConv2DLayer, filter_shape=(3, 3, 4, 6), depth_multiplier=2,
            batch_norm=False, activation='NONE')
  ys, _ = s.FPropDefaultTheta(x, torch.zeros(2, 7))
  assert ys.shape == (2, 7, 5, 6)
  _GradsFlow(s, ys)
  assert len(s.vars.Flatten()) >= 2                       # depthwise + pointwise weights


def test_causal_conv_does_not_look_ahead():
  l = _Make(layers.Conv2DLayer, filter_shape=(3, 1, 2, 2), causal_convolution=True,
            batch_norm=False, activation='NONE')
  x = torch.randn(1, 10, 1, 2)
  y, _ = l.FPropDefaultTheta(x, torch.zeros(1, 10))
  x2 = x.clone()
  x2[:, 6:] += 5.0
  y2, _ = l.FPropDefaultTheta(x2, torch.zeros(1, 10))
  torch.testing.assert_close(y[:, :6], y2[:, :6])
  assert not torch.allclose(y[:, 6:], y2[:, 6:])


def test_conv_weight_norm_and_batch_norm():
  l = _Make(layers.Conv2DLayer, filter_shape=(3, 3, 2, 4), weight_norm=True, batch_norm=True,
            activation='RELU')
  x = torch.randn(4, 6, 6, 2)
  y, _ = l.FPropDefaultTheta(x, torch.zeros(4, 6))
  assert y.shape == (4, 6, 6, 4) and 'g' in l.vars
  _GradsFlow(l, y)
  # BN in training mode: per-channel pre-activation is standardised → about half the units fire
  frac = float((y > 0).float().mean())
  assert 0.3 < frac < 0.7


def test_conv2d_no_padding_and_deconv():
  l = _Make(layers.Conv2DLayerNoPadding, filter_shape=(3, 3, 2, 5), filter_stride=(1, 1),
            padding='VALID', use_bias=True)
  x = torch.randn(2, 8, 8, 2)
  y = l.FPropDefaultTheta(x)
  want = F.conv2d(x.permute(0, 3, 1, 2), l.vars.w.detach().permute(3, 2, 0, 1)).permute(
      0, 2, 3, 1) + l.vars.b.detach()
  torch.testing.assert_close(y, want, atol=1e-4, rtol=1e-4)
  d = _Make(layers.DeconvLayer, filter_shape=(2, 2, 3, 5), filter_stride=(2, 2))
  z = d.FPropDefaultTheta(torch.randn(2, 4, 4, 5))
  assert z.shape == (2, 8, 8, 3)
  assert tuple(d.OutShape((2, 4, 4, 5))) == (2, 8, 8, 3)
  _GradsFlow(d, z)


def test_conv_set_layer_concatenates_its_filters():
  l = _Make(layers.ConvSetLayer, cnn_tpl=layers.Conv2DLayer.Params().Set(
      batch_norm=False, activation='NONE'), filter_shapes=[(3, 3, 2, 4), (5, 5, 2, 3)])
  x = torch.randn(2, 8, 6, 2)
  y, pad = l.FPropDefaultTheta(x, torch.zeros(2, 8))
  assert y.shape == (2, 8, 6, 7) and pad.shape == (2, 8)


# -------------------------------------------------------------------------- pooling --
def test_pooling_layer_max_and_avg():
  x = torch.arange(2 * 6 * 4 * 1, dtype=torch.float32).reshape(2, 6, 4, 1)
  mp = _Make(layers.PoolingLayer, window_shape=(2, 2), window_stride=(2, 2), pooling_type='MAX')
  y, pad = mp.FPropDefaultTheta(x, torch.zeros(2, 6))
  want = F.max_pool2d(x.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
  torch.testing.assert_close(y, want)
  assert pad.shape == (2, 3) and tuple(mp.OutShape((2, 6, 4, 1))) == (2, 3, 2, 1)
  ap = _Make(layers.PoolingLayer, window_shape=(2, 2), window_stride=(2, 2), pooling_type='AVG')
  ya, _ = ap.FPropDefaultTheta(x, torch.zeros(2, 6))
  torch.testing.assert_close(ya, F.avg_pool2d(x.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1))
  # padded frames must not leak into the max of the preceding window
  pad_in = torch.zeros(2, 6)
  pad_in[0, 3:] = 1
  yp, pp = mp.FPropDefaultTheta(x, pad_in)
  assert float(yp[0, 1].max()) <= float(x[0, 2].max())
  assert pp[0].tolist() == [0, 0, 1]


def test_blur_pool_is_a_normalised_low_pass_subsampler():
  l = _Make(layers.BlurPoolLayer, blur_filter='B5', subsample_type='1D', input_channels=3)
  x = torch.ones(2, 12, 1, 3)
  y, pad = l.FPropDefaultTheta(x, torch.zeros(2, 12))
  assert y.shape == (2, 6, 1, 3) and pad.shape == (2, 6)
  torch.testing.assert_close(y[:, 1:-1], torch.ones(2, 4, 1, 3))   # DC gain 1 in the interior
  alt = torch.tensor([1.0, -1.0] * 6).view(1, 12, 1, 1).expand(2, 12, 1, 3).contiguous()
  ya, _ = l.FPropDefaultTheta(alt, torch.zeros(2, 12))
  assert float(ya[:, 1:-1].abs().max()) < 1e-5                     # Nyquist is removed


# ----------------------------------------------------------------------- projection --
@pytest.mark.parametrize('act', ['RELU', 'TANH', 'SIGMOID', 'NONE', 'GELU', 'SWISH'])
def test_projection_layer_activations(act):
  l = _Make(layers.ProjectionLayer, input_dim=5, output_dim=4, activation=act, has_bias=True,
            bias_init=0.1)
  x = torch.randn(3, 7, 5)
  y = l.FPropDefaultTheta(x)
  z = x @ l.vars.w.detach() + l.vars.b.detach()
  fn = dict(RELU=torch.relu, TANH=torch.tanh, SIGMOID=torch.sigmoid, NONE=lambda v: v,
            GELU=F.gelu, SWISH=F.silu)[act]
  tol = 2e-3 if act == 'GELU' else 1e-5                    # tanh-approximate GELU allowed
  torch.testing.assert_close(y, fn(z), atol=tol, rtol=tol)
  _GradsFlow(l, y)


def test_projection_layer_paddings_weight_norm_and_batch_norm():
  l = _Make(layers.ProjectionLayer, input_dim=4, output_dim=3, weight_norm=True,
            batch_norm=False, activation='NONE')
  x = torch.randn(2, 5, 4)
  pad = torch.zeros(2, 5, 1)
  pad[0, 3:] = 1
  y = l.FPropDefaultTheta(x, pad)
  assert float(y[0, 3:].abs().max()) == 0.0
  w = l.vars.w.detach()
  wn = w / w.norm(dim=0, keepdim=True) * (1.0 + l.vars.g.detach())
  torch.testing.assert_close(y[1], x[1] @ wn, atol=1e-5, rtol=1e-5)
  bn = _Make(layers.ProjectionLayer, input_dim=4, output_dim=3, batch_norm=True,
             activation='NONE')
  yb = bn.FPropDefaultTheta(torch.randn(64, 4) * 3 + 1)
  assert abs(float(yb.mean())) < 0.1 and abs(float(yb.std()) - 1.0) < 0.1


def test_projection_layer_block_diagonal():
  l = _Make(layers.ProjectionLayer, input_dim=6, output_dim=4, use_block_diagonal_matmul=True,
            bd_num_blocks=2, activation='NONE')
  assert l.vars.w.shape == (2, 3, 2)
  x = torch.randn(5, 6)
  y = l.FPropDefaultTheta(x)
  w = l.vars.w.detach()
  want = torch.cat([x[:, :3] @ w[0], x[:, 3:] @ w[1]], -1)
  torch.testing.assert_close(y, want, atol=1e-5, rtol=1e-5)


def test_fc_layer_and_feed_forward_net():
  fc = _Make(layers.FCLayer, input_dim=4, output_dim=3)
  x = torch.randn(2, 4)
  torch.testing.assert_close(fc.FPropDefaultTheta(x),
                             torch.relu(x @ fc.vars.w.detach() + fc.vars.b.detach()))
  net = _Make(layers.FeedForwardNet, input_dim=4, hidden_layer_dims=[8, 3],
              activation=['RELU', 'NONE'], dropout=layers.DropoutLayer.Params())
  y = net.FPropDefaultTheta(x)
  ws = [c for c in net.fc]
  h = torch.relu(x @ ws[0].vars.w.detach() + ws[0].vars.b.detach())
  want = h @ ws[1].vars.w.detach() + ws[1].vars.b.detach()
  torch.testing.assert_close(y, want, atol=1e-5, rtol=1e-5)
  _GradsFlow(net, y)
  skip = _Make(layers.FeedForwardNet, input_dim=4, hidden_layer_dims=[4, 4],
               activation='TANH', skip_connections=['ResNet', 'ResNet'])
  assert skip.FPropDefaultTheta(x).shape == (2, 4)


def test_multitask_projection_and_ffn_select_task_weights():
  l = _Make(layers.MultitaskProjectionEinsumLayer, input_dim=4, output_dim=3, num_tasks=3,
            activation='NONE', has_bias=True)
  x = torch.randn(5, 4)
  tasks = torch.tensor([0, 2, 1, 2, 0])
  y = l.FPropDefaultTheta(x, tasks)
  w, b = l.vars.w.detach(), l.vars.b.detach()
  want = torch.stack([x[i] @ w[tasks[i]] + b[tasks[i]] for i in range(5)])
  torch.testing.assert_close(y, want, atol=1e-5, rtol=1e-5)
  onehot = F.one_hot(tasks, 3).float()
  torch.testing.assert_close(l.FPropDefaultTheta(x, onehot), want, atol=1e-5, rtol=1e-5)
  net = _Make(layers.MultitaskFeedForwardNet, input_dim=4, hidden_layer_dims=[6, 2],
              num_tasks=3, activation=['RELU', 'NONE'])
  z = net.FPropDefaultTheta(x, tasks)
  assert z.shape == (5, 2)
  z2 = net.FPropDefaultTheta(x, torch.tensor([1, 2, 1, 2, 0]))
  assert not torch.allclose(z[0], z2[0]) and torch.allclose(z[1:], z2[1:])
  # scalar task, per-time-step tasks and both einsum orders (ref py_utils.MultiTaskProjection)
  xt = torch.randn(2, 3, 4)
  per_step = torch.tensor([[0, 1, 2], [2, 2, 0]])
  a = py_utils.MultiTaskProjection(w, b, xt, per_step)
  bb = py_utils.MultiTaskProjection(w, b, xt, per_step, einsum_order='multiply_and_select')
  want_t = torch.stack([torch.stack([xt[i, j] @ w[per_step[i, j]] + b[per_step[i, j]]
                                     for j in range(3)]) for i in range(2)])
  torch.testing.assert_close(a, want_t, atol=1e-5, rtol=1e-5)
  torch.testing.assert_close(bb, want_t, atol=1e-5, rtol=1e-5)
  torch.testing.assert_close(py_utils.MultiTaskProjection(w, b, xt, torch.tensor(1)),
                             xt @ w[1] + b[1], atol=1e-5, rtol=1e-5)
  with pytest.raises(ValueError):
    py_utils.MultiTaskProjection(w, b, xt, per_step, einsum_order='nope')


# ------------------------------------------------------------------ StackingOverTime --
@pytest.mark.parametrize('pad_with_left_frame', [True, False])
def test_stacking_over_time_reference_golden(pad_with_left_frame):
  l = _Make(layers.StackingOverTime, left_context=2, right_context=0, stride=2,
            pad_with_left_frame=pad_with_left_frame)
  assert l.window_size == 3
  x = torch.tensor([[[1, 1], [2, 2], [3, 3], [4, 4], [5, 5], [6, 6]],
                    [[7, 7], [8, 8], [0, 0], [0, 0], [0, 0], [0, 0]]], dtype=torch.float32)
  pad = torch.tensor([[[0], [0], [0], [0], [0], [0]], [[0], [0], [1], [1], [1], [1]]],
                     dtype=torch.float32)
  y, yp = l.FProp(x, pad)
  if pad_with_left_frame:
    want = [[[1, 1, 1, 1, 1, 1], [1, 1, 2, 2, 3, 3], [3, 3, 4, 4, 5, 5]],
            [[7, 7, 7, 7, 7, 7], [7, 7, 8, 8, 0, 0], [0, 0, 0, 0, 0, 0]]]
  else:
    want = [[[0, 0, 0, 0, 1, 1], [1, 1, 2, 2, 3, 3], [3, 3, 4, 4, 5, 5]],
            [[0, 0, 0, 0, 7, 7], [7, 7, 8, 8, 0, 0], [0, 0, 0, 0, 0, 0]]]
  np.testing.assert_allclose(y.numpy(), want)
  np.testing.assert_allclose(yp.numpy(), [[[0], [0], [0]], [[0], [0], [1]]])


def test_stacking_over_time_lengths_sums_and_unstack():
  l = _Make(layers.StackingOverTime, left_context=0, right_context=1, stride=2)
  assert l.window_size == 2
  x = torch.randn(2, 21, 16)
  lens = torch.tensor([9, 14])
  pad = (torch.arange(21).unsqueeze(0) >= lens.unsqueeze(1)).float().unsqueeze(-1)
  x = x * (1 - pad)
  y, yp = l.FProp(x, pad)
  np.testing.assert_array_equal((1 - yp).sum((1, 2)).numpy(), [5, 7])     # ref :3549
  torch.testing.assert_close(y.sum((1, 2)), x.sum((1, 2)), atol=1e-4, rtol=1e-4)
  ident = _Make(layers.StackingOverTime, left_context=0, right_context=0, stride=1)
  xi = torch.tensor([[[1.], [2.], [3.], [4.], [5.]]])
  yi, _ = ident.FProp(xi, torch.zeros(1, 5, 1))
  torch.testing.assert_close(yi, xi)
  back = l.Unstack(y)
  torch.testing.assert_close(back[:, :20], x[:, :20], atol=1e-6, rtol=1e-6)


# ----------------------------------------------------------------------- embeddings --
@pytest.mark.parametrize('cls,kw', [
    (layers.SingleShardEmbeddingLayer, {}),
    (layers.EmbeddingLayer, dict(max_num_shards=2)),
    (layers.SimpleEmbeddingLayer, {}),
    (layers.SimpleEmbeddingLayer, dict(use_matmul=True)),
    (layers.EinsumEmbeddingLayer, {}),
])
def test_embedding_layers_lookup_scale_and_grad(cls, kw):
  l = _Make(cls, vocab_size=10, embedding_dim=4, **kw)
  ids = torch.tensor([[1, 3, 9], [0, 3, 3]])
  y = l.EmbLookup(l.theta, ids)
  assert y.shape == (2, 3, 4)
  torch.testing.assert_close(y[0, 1], y[1, 1])
  torch.testing.assert_close(y[1, 1], y[1, 2])
  assert not torch.allclose(y[0, 0], y[0, 2])
  _GradsFlow(l, y)
  s = _Make(cls, name='scaled', vocab_size=10, embedding_dim=4, scale_sqrt_depth=True, **kw)
  ys = s.EmbLookup(s.theta, ids)
  base = s.EmbLookup(s.theta, ids) / 2.0
  torch.testing.assert_close(ys, base * 2.0)
  # scale_sqrt_depth multiplies the table rows by sqrt(dim) = 2
  table = torch.cat([v.detach().reshape(-1, 4) for v in s.vars.Flatten()], 0)
  assert any(torch.allclose(ys[0, 0], row * 2.0, atol=1e-5) for row in table)


def test_one_hot_embedding_layer():
  l = _Make(layers.OneHotEmbeddingLayer, vocab_size=5, embedding_dim=5, uncertainty=0.0)
  y = l.EmbLookup(l.theta, torch.tensor([[0, 4]]))
  np.testing.assert_allclose(y.numpy(), [[[1, 0, 0, 0, 0], [0, 0, 0, 0, 1]]])
  u = _Make(layers.OneHotEmbeddingLayer, vocab_size=5, embedding_dim=5, uncertainty=0.2)
  yu = u.EmbLookup(u.theta, torch.tensor([2]))
  np.testing.assert_allclose(yu.numpy(), [[0.05, 0.05, 0.8, 0.05, 0.05]], atol=1e-6)


def test_positional_embedding_layer_reference_golden():
  l = _Make(layers.PositionalEmbeddingLayer, min_timescale=1, max_timescale=7, embedding_dim=4)
  want = [[0., 0., 1., 1.],
          [0.84147096, 0.14237173, 0.54030228, 0.98981327],
          [0.90929741, 0.28184283, -0.41614676, 0.95946062],
          [0.14112, 0.4155719, -0.9899925, 0.90956032],
          [-0.7568025, 0.54083425, -0.65364361, 0.84112918],
          [-0.95892417, 0.65507787, 0.28366217, 0.75556135],
          [-0.27941549, 0.75597537, 0.96017027, 0.65460002],
          [0.65698659, 0.84147096, 0.7539022, 0.54030228],
          [0.98935831, 0.90982294, -0.14550003, 0.41499668],
          [0.41211855, 0.9596386, -0.91113025, 0.28123617],
          [-0.54402113, 0.98990309, -0.83907151, 0.14174587]]   # ref layers_test.py:4229
  np.testing.assert_allclose(l.FPropDefaultTheta(11).numpy(), want, atol=1e-5)
  pos = torch.tensor([[0, 1, 2, 3], [0, 1, 0, 1]])
  yp = l.FPropWithPosition(l.theta, pos)
  np.testing.assert_allclose(yp[1, 2].numpy(), want[0], atol=1e-5)
  np.testing.assert_allclose(yp[0, 3].numpy(), want[3], atol=1e-5)
  sc = _Make(layers.PositionalEmbeddingLayer, min_timescale=1, max_timescale=7, embedding_dim=4,
             trainable_scaling=True, trainable_scaling_init=3.0)
  np.testing.assert_allclose(sc.FPropDefaultTheta(11).detach().numpy(),
                             3.0 * np.array(want), atol=1e-4)


def test_learnable_relative_and_sinusoidal_positional_embeddings():
  l = _Make(layers.LearnablePositionalEmbeddingLayer, embedding_dim=6, max_pos=16)
  y = l.FPropDefaultTheta(5)
  assert y.shape == (5, 6)
  yp = l.FPropWithPosition(l.theta, torch.tensor([[3, 0]]))
  torch.testing.assert_close(yp[0, 0], y[3])
  _GradsFlow(l, y)
  r = _Make(layers.RelativePositionalEmbeddingLayer, radius=3, dim=4)
  d = torch.tensor([[-10, -3, 0, 2, 7]])
  e = r.FPropDefaultTheta(d)
  assert e.shape == (1, 5, 4)
  torch.testing.assert_close(e[0, 0], e[0, 1])                      # clipped at −radius
  torch.testing.assert_close(e[0, 4], r.FPropDefaultTheta(torch.tensor([3]))[0])
  s = _Make(layers.SinusoidalPositionalEmbeddingLayer, embedding_dim=6)
  ys = s.FPropDefaultTheta(7)
  assert ys.shape == (7, 6)
  torch.testing.assert_close((ys[:, 0::2] ** 2 + ys[:, 1::2] ** 2),
                             torch.ones(7, 3), atol=1e-5, rtol=1e-5)   # sin² + cos² pairs


def test_rotary_positional_embedding_is_a_rotation_and_relative():
  l = _Make(layers.RotaryPositionalEmbeddingLayer, embedding_dim=8)
  q = torch.randn(2, 6, 3, 8)
  k = torch.randn(2, 6, 3, 8)
  rq, rk = l.FPropDefaultTheta(q), l.FPropDefaultTheta(k)
  torch.testing.assert_close(rq.norm(dim=-1), q.norm(dim=-1), atol=1e-5, rtol=1e-5)
  # the q·k score depends only on the relative offset: shift both positions by 5
  pos = torch.arange(6).unsqueeze(0).expand(2, 6)
  sq, sk = l.FPropDefaultTheta(q, pos + 5), l.FPropDefaultTheta(k, pos + 5)
  a = torch.einsum('btnh,bsnh->bnts', rq, rk)
  b = torch.einsum('btnh,bsnh->bnts', sq, sk)
  torch.testing.assert_close(a, b, atol=1e-4, rtol=1e-4)


# ------------------------------------------------------------------------- softmaxes --
def _XentOracle(logits, ids, weights, label_smoothing=0.0):
  lp = F.log_softmax(logits.float(), -1)
  nll = -lp.gather(-1, ids.unsqueeze(-1)).squeeze(-1)
  if label_smoothing:
    nll = (1 - label_smoothing) * nll - label_smoothing * lp.mean(-1)
  return nll, (nll * weights).sum(), weights.sum()


@pytest.mark.parametrize('cls,kw', [
    (layers.SimpleFullSoftmax, {}),
    (layers.SimpleFullSoftmax, dict(num_shards=3)),
    (layers.SimpleFullSoftmax, dict(use_num_classes_major_weight=True)),
    (layers.SingleShardFullSoftmax, {}),
    (layers.EinsumSoftmax, {}),
])
def test_full_softmax_xent_matches_oracle(cls, kw):
  l = _Make(cls, input_dim=6, num_classes=9, **kw)
  x = torch.randn(4, 5, 6)
  ids = torch.randint(0, 9, (4, 5))
  w = torch.ones(4, 5)
  w[0, 3:] = 0
  out = l.FPropDefaultTheta(x, w, class_ids=ids)
  logits = l.Logits(l.theta, x)
  assert logits.shape == (4, 5, 9)
  nll, tot, wsum = _XentOracle(logits.detach(), ids, w)
  torch.testing.assert_close(out.per_example_xent.reshape(4, 5), nll, atol=1e-4, rtol=1e-4)
  assert float(out.total_xent) == pytest.approx(float(tot), rel=1e-4)
  assert float(out.total_weight) == pytest.approx(float(wsum))
  assert float(out.avg_xent) == pytest.approx(float(tot / wsum), rel=1e-4)
  torch.testing.assert_close(out.per_example_argmax.reshape(4, 5), logits.argmax(-1))
  _GradsFlow(l, out.total_xent)
  # class_probabilities path ≡ class_ids path for one-hot targets
  probs = F.one_hot(ids, 9).float()
  out2 = l.FPropDefaultTheta(x, w, class_probabilities=probs)
  assert float(out2.total_xent) == pytest.approx(float(tot), rel=1e-4)


def test_softmax_logit_caps_label_smoothing_and_chunks():
  l = _Make(layers.SimpleFullSoftmax, input_dim=4, num_classes=7, logits_abs_max=0.5)
  x = torch.randn(8, 4) * 10
  assert float(l.Logits(l.theta, x).abs().max()) <= 0.5 + 1e-6
  s = _Make(layers.SimpleFullSoftmax, input_dim=4, num_classes=7, logits_soft_max=2.0)
  assert float(s.Logits(s.theta, x).abs().max()) <= 2.0
  ls = _Make(layers.SimpleFullSoftmax, input_dim=4, num_classes=7, label_smoothing=0.1)
  ids = torch.randint(0, 7, (8,))
  out = ls.FPropDefaultTheta(x, torch.ones(8), class_ids=ids)
  _, tot, _ = _XentOracle(ls.Logits(ls.theta, x).detach(), ids, torch.ones(8), 0.1)
  assert float(out.total_xent) == pytest.approx(float(tot), rel=1e-4)
  ch = _Make(layers.SimpleFullSoftmax, input_dim=4, num_classes=7, chunk_size=3)
  plain = _Make(layers.SimpleFullSoftmax, name='plain', input_dim=4, num_classes=7)
  with torch.no_grad():
    for a, b in zip(plain.vars.Flatten(), ch.vars.Flatten()):
      a.copy_(b)
  oc = ch.FPropDefaultTheta(x, torch.ones(8), class_ids=ids)
  op = plain.FPropDefaultTheta(x, torch.ones(8), class_ids=ids)
  assert float(oc.total_xent) == pytest.approx(float(op.total_xent), rel=1e-5)


def test_sampled_softmax_trains_and_full_softmax_evaluates():
  l = _Make(layers.SimpleFullSoftmax, input_dim=4, num_classes=50, num_sampled=8)
  x = torch.randn(6, 4)
  ids = torch.randint(0, 50, (6,))
  out = l.FPropDefaultTheta(x, torch.ones(6), class_ids=ids)
  assert torch.isfinite(out.total_xent)
  _GradsFlow(l, out.total_xent)
  with cluster_factory.SetEval(True):
    le = _Make(layers.SimpleFullSoftmax, name='e', input_dim=4, num_classes=50, num_sampled=8)
    oe = le.FPropDefaultTheta(x, torch.ones(6), class_ids=ids)
    _, tot, _ = _XentOracle(le.Logits(le.theta, x).detach(), ids, torch.ones(6))
    assert float(oe.total_xent) == pytest.approx(float(tot), rel=1e-4)


def test_sigmoid_focal_and_scones_losses():
  x = torch.randn(5, 4)
  probs = torch.zeros(5, 3)
  probs[torch.arange(5), torch.tensor([0, 2, 1, 1, 0])] = 1.0
  sg = _Make(layers.SimpleFullSigmoidCrossEntropy, input_dim=4, num_classes=3)
  out = sg.FPropDefaultTheta(x, torch.ones(5), class_probabilities=probs)
  logits = sg.Logits(sg.theta, x).detach()
  want = F.binary_cross_entropy_with_logits(logits, probs, reduction='none').sum(-1)
  torch.testing.assert_close(out.per_example_xent, want, atol=1e-5, rtol=1e-5)
  fo = _Make(layers.FocalFullSoftmax, input_dim=4, num_classes=3, focal_loss_gamma=2.0)
  ids = probs.argmax(-1)
  of = fo.FPropDefaultTheta(x, torch.ones(5), class_ids=ids)
  lp = F.log_softmax(fo.Logits(fo.theta, x).detach(), -1).gather(-1, ids[:, None])[:, 0]
  torch.testing.assert_close(of.per_example_xent, -((1 - lp.exp()) ** 2) * lp, atol=1e-5,
                             rtol=1e-4)
  g0 = _Make(layers.FocalFullSoftmax, name='g0', input_dim=4, num_classes=3, focal_loss_gamma=0.)
  o0 = g0.FPropDefaultTheta(x, torch.ones(5), class_ids=ids)
  lp0 = F.log_softmax(g0.Logits(g0.theta, x).detach(), -1).gather(-1, ids[:, None])[:, 0]
  torch.testing.assert_close(o0.per_example_xent, -lp0, atol=1e-5, rtol=1e-5)
  sc = _Make(layers.Scones, input_dim=4, num_classes=3, pos_weight=2.0)
  os_ = sc.FPropDefaultTheta(x, torch.ones(5), class_probabilities=probs)
  assert os_.per_example_xent.shape == (5,) and torch.isfinite(os_.total_xent)
  _GradsFlow(sc, os_.total_xent)


def test_shared_embedding_softmax_ties_weights():
  for cls, kw in [(layers.SingleShardSharedEmbeddingSoftmax,
                   dict(input_dim=6, num_classes=11, vocab_size=11, embedding_dim=6)),
                  (layers.SharedSoftmaxLayer, dict(vocab_size=11, embedding_dim=6,
                                                   softmax=layers.SimpleFullSoftmax.Params().Set(
                                                       input_dim=6, num_classes=11)))]:
    l = _Make(cls, **kw)
    ids = torch.tensor([[1, 5, 10]])
    emb = l.EmbLookup(l.theta, ids)
    assert emb.shape == (1, 3, 6)
    logits = l.Logits(l.theta, emb)
    # tied: logit of class c for input e_c is ‖e_c‖² (+ bias 0)
    torch.testing.assert_close(logits[0, 0, 1], (emb[0, 0] ** 2).sum(), atol=1e-4, rtol=1e-4)
    n_mats = [v for v in l.vars.Flatten() if v.dim() == 2]
    assert len(n_mats) == 1                                        # one shared matrix


def test_conv_softmax_and_bias_layer():
  l = _Make(layers.ConvSoftmax, input_dim=4, hidden_dim=8, num_classes=5)
  x = torch.randn(2, 7, 4)
  logits = l.Logits(l.theta, x)
  assert logits.shape == (2, 7, 5)
  b = _Make(layers.BiasLayerSimple, dims=3)
  with torch.no_grad():
    b.vars.Flatten()[0].copy_(torch.tensor([1.0, 2.0, 3.0]))
  torch.testing.assert_close(b.FPropDefaultTheta(torch.zeros(2, 3)),
                             torch.tensor([[1.0, 2, 3], [1, 2, 3]]))


# --------------------------------------------------------------------------- dropout --
def test_dropout_layers():
  d = _Make(layers.DropoutLayer, keep_prob=0.8)
  x = torch.ones(200, 50)
  y = d.FPropDefaultTheta(x)
  kept = (y != 0).float().mean()
  assert 0.75 < float(kept) < 0.85
  torch.testing.assert_close(y[y != 0], torch.full_like(y[y != 0], 1.25))
  with cluster_factory.SetEval(True):
    de = _Make(layers.DropoutLayer, keep_prob=0.5)
    assert torch.equal(de.FPropDefaultTheta(x), x)
  nb = _Make(layers.DropoutLayer, keep_prob=0.5, noise_shape_broadcast_dims=[1])
  yb = nb.FPropDefaultTheta(x)
  assert ((yb != 0).all(1) | (yb == 0).all(1)).all()               # whole rows kept/dropped
  det = _Make(layers.DeterministicDropoutLayer, keep_prob=0.5)
  with py_utils.StepSeedContext(1234) if hasattr(py_utils, 'StepSeedContext') else \
      torch.random.fork_rng():
    a = det.FPropDefaultTheta(x)
  assert 0.4 < float((a != 0).float().mean()) < 0.6


# ----------------------------------------------------------------------------- norms --
@pytest.mark.parametrize('kw', [{}, dict(direct_scale=True), dict(center=False, bias=False),
                                dict(use_fused_layernorm=True)])
def test_layer_norm_matches_oracle(kw):
  l = _Make(layers.LayerNorm, input_dim=6, **kw)
  with torch.no_grad():
    l.vars.scale.copy_(torch.linspace(-0.2, 0.3, 6))
    if 'bias' in l.vars:
      l.vars.bias.copy_(torch.linspace(0.1, 0.6, 6))
  x = torch.randn(3, 4, 6) * 2 + 1
  y = l.FPropDefaultTheta(x)
  center = kw.get('center', True)
  mean = x.mean(-1, keepdim=True) if center else 0.0
  var = ((x - mean) ** 2).mean(-1, keepdim=True) if center else (x ** 2).mean(-1, keepdim=True)
  n = (x - mean) * torch.rsqrt(var + 1e-6)
  scale = l.vars.scale.detach() if kw.get('direct_scale') else 1.0 + l.vars.scale.detach()
  want = n * scale + (l.vars.bias.detach() if 'bias' in l.vars else 0.0)
  torch.testing.assert_close(y, want, atol=1e-5, rtol=1e-5)
  _GradsFlow(l, y)


def test_reshaped_and_categorical_layer_norm():
  r = _Make(layers.ReshapedLayerNorm, input_dim=6)
  x = torch.randn(2, 5, 2, 3)                                      # [..., N, H] with N·H = 6
  y = r.FPropDefaultTheta(x)
  assert y.shape == x.shape
  flat = F.layer_norm(x.reshape(2, 5, 6), (6,), eps=1e-6).reshape(2, 5, 2, 3)
  torch.testing.assert_close(y, flat, atol=1e-5, rtol=1e-5)
  c = _Make(layers.CategoricalLayerNorm, input_dim=4, num_classes=3)
  with torch.no_grad():
    for v in c.vars.Flatten():
      v.copy_(torch.randn_like(v) * 0.1)
  xin = torch.randn(2, 4)
  c.SetClassIndex(torch.tensor(0))
  y0 = c.FPropDefaultTheta(xin)
  c.SetClassIndex(torch.tensor(2))
  y2 = c.FPropDefaultTheta(xin)
  assert not torch.allclose(y0, y2)
  c.SetClassIndex(torch.tensor(0))
  torch.testing.assert_close(c.FPropDefaultTheta(xin), y0)


# ------------------------------------------------------------------- label smoothing --
def test_uniform_label_smoother_reference_golden():
  l = _Make(layers.UniformLabelSmoother, num_classes=5, uncertainty=0.1)
  labels = torch.tensor([[0, 1, 2, 3, 3, 3, 4]])
  ids = torch.tensor([[0, 0, 1, 2, 3, 3, 3]])
  out = l.FPropDefaultTheta(torch.zeros(1, 7), labels, ids)
  want = np.full((1, 7, 5), 0.025, np.float32)
  for t, c in enumerate([0, 1, 2, 3, 3, 3, 4]):
    want[0, t, c] = 0.9
  np.testing.assert_allclose(out.numpy(), want, atol=1e-6)           # ref :5828
  big = _Make(layers.UniformLabelSmoother, name='big', num_classes=5, uncertainty=0.1,
              uncertainty_larger=0.2, token_id_uncertainty_larger=4)
  o2 = big.FPropDefaultTheta(torch.zeros(1, 7), torch.tensor([[0, 1, 2, 3, 3, 3, 3]]),
                             torch.tensor([[0, 0, 1, 2, 4, 4, 4]]))
  np.testing.assert_allclose(o2[0, 3].numpy(), [0.025, 0.025, 0.025, 0.9, 0.025], atol=1e-6)
  np.testing.assert_allclose(o2[0, 4].numpy(), [0.05, 0.05, 0.05, 0.8, 0.05], atol=1e-6)
  np.testing.assert_allclose(o2.sum(-1).numpy(), 1.0, atol=1e-6)


def test_localized_label_smoother_spreads_mass_over_time():
  l = _Make(layers.LocalizedLabelSmoother, num_classes=6, offsets=[-1, 1], weights=[0.1, 0.2])
  labels = torch.tensor([[1, 2, 3, 4, 0]])
  pad = torch.tensor([[0., 0, 0, 0, 1]])
  out = l.FPropDefaultTheta(pad, labels, labels)
  np.testing.assert_allclose(out.sum(-1).numpy(), 1.0, atol=1e-6)
  # position 1 (label 2): neighbours' labels 1 (offset −1, w .1) and 3 (offset +1, w .2)
  np.testing.assert_allclose(out[0, 1].numpy(), np.array([0, .1, 1, .2, 0, 0]) / 1.3, atol=1e-6)
  # position 0 has no left neighbour; position 3 is the last valid label: its right neighbour
  # is padding, and it does not leak into position 2 (EOS is not made more probable)
  np.testing.assert_allclose(out[0, 0].numpy(), np.array([0, 1, .2, 0, 0, 0]) / 1.2, atol=1e-6)
  np.testing.assert_allclose(out[0, 2].numpy(), np.array([0, 0, .1, 1, 0, 0]) / 1.1, atol=1e-6)
  np.testing.assert_allclose(out[0, 3].numpy(), np.array([0, 0, 0, .1, 1, 0]) / 1.1, atol=1e-6)


# --------------------------------------------------------------------- gates & merge --
def test_highway_skip_and_gating_layers():
  h = _Make(layers.HighwaySkipLayer, input_dim=4, carry_bias_init=1.0)
  x, t = torch.randn(3, 4), torch.randn(3, 4)
  y = h.FPropDefaultTheta(x, t)
  assert y.shape == (3, 4)
  hc = _Make(layers.HighwaySkipLayer, name='hc', input_dim=4, couple_carry_transform_gates=True)
  yc = hc.FPropDefaultTheta(x, t)
  # coupled gates: y is a convex combination of x and t elementwise
  lo, hi = torch.minimum(x, t), torch.maximum(x, t)
  assert ((yc >= lo - 1e-5) & (yc <= hi + 1e-5)).all()
  _GradsFlow(hc, yc)
  g = _Make(layers.GatingLayer, input_dim=4, carry_bias_init=0.0)
  yg = g.FPropDefaultTheta(x, t)
  assert ((yg >= lo - 1e-5) & (yg <= hi + 1e-5)).all()
  sat = _Make(layers.GatingLayer, name='sat', input_dim=4, has_bias=True, carry_bias_init=1e4)
  torch.testing.assert_close(sat.FPropDefaultTheta(x, t), x, atol=1e-4, rtol=1e-4)


def test_weighted_sum_and_gated_average():
  w = _Make(layers.WeightedSumLayer, num_sources=3, weighted_merger_dropout_prob=0.0)
  xs = [torch.randn(2, 4) for _ in range(3)]
  y = w.FPropDefaultTheta(xs)
  torch.testing.assert_close(y, sum(xs) / 3.0, atol=1e-5, rtol=1e-5)   # zero-init logits
  with torch.no_grad():
    w.vars.Flatten()[0].copy_(torch.tensor([10.0, -10.0, -10.0]))
  torch.testing.assert_close(w.FPropDefaultTheta(xs), xs[0], atol=1e-4, rtol=1e-4)
  ws = _Make(layers.WeightedSumLayer, name='ws', num_sources=2, weighted_merger_softmax=False,
             weighted_merger_dropout_prob=0.0, global_weight_scale=2.0)
  v = ws.vars.Flatten()[0].detach()
  torch.testing.assert_close(ws.FPropDefaultTheta(xs[:2]),
                             2.0 * (v[0] * xs[0] + v[1] * xs[1]), atol=1e-5, rtol=1e-5)
  g = _Make(layers.GatedAverageLayer, num_nodes=4, num_inputs=3)
  yg = g.FPropDefaultTheta(xs)
  stack = torch.stack(xs, -1)
  assert ((yg >= stack.min(-1).values - 1e-5) & (yg <= stack.max(-1).values + 1e-5)).all()
  _GradsFlow(g, yg)


def test_lhuc_residual_adapter_and_glu():
  l = _Make(layers.LHUCLayer, input_dim=4)
  x = torch.randn(3, 4)
  torch.testing.assert_close(l.FPropDefaultTheta(x), x)             # 2·sigmoid(0) = 1
  with torch.no_grad():
    l.vars.Flatten()[0].fill_(100.0)
  torch.testing.assert_close(l.FPropDefaultTheta(x), 2.0 * x, atol=1e-5, rtol=1e-5)
  r = _Make(layers.ResidualAdapterLayer, input_dim=4, bottleneck_dim=2)
  xr = torch.randn(2, 5, 4)
  yr = r.FPropDefaultTheta(xr)
  assert yr.shape == xr.shape
  _GradsFlow(r, yr)
  g = _Make(layers.GluLayer, input_dim=4, output_dim=4)
  yg = g.FPropDefaultTheta(xr, torch.zeros(2, 5, 1))
  assert yg.shape == xr.shape
  g2 = _Make(layers.GluLayer, name='g2', input_dim=4, output_dim=6, apply_residual=False)
  assert g2.FPropDefaultTheta(xr, torch.zeros(2, 5, 1)).shape == (2, 5, 6)


def test_grad_norm_tracker_rejects_outliers():
  """ref layers_test.py:5642: ~16% of e^{N(7,1)} outliers among e^{N(5,1)} norms are rejected."""
  t = _Make(layers.GradNormTracker, clip_threshold=3.0)
  rng = np.random.RandomState(12345)
  normal = np.exp(rng.normal(5.0, 1.0, size=10000))
  outliers = np.exp(rng.normal(7.0, 1.0, size=100))
  rejected = 0
  for i in range(100):
    for j in range(100):
      t.FPropDefaultTheta(torch.tensor(float(normal[i * 100 + j])))
    if float(t.FPropDefaultTheta(torch.tensor(float(outliers[i])))) == 0.0:
      rejected += 1
  assert 5 < rejected < 60
  # NaN steps are rejected and do not pollute the statistics
  before = [v.detach().clone() for v in t.vars.Flatten()]
  assert float(t.FPropDefaultTheta(torch.tensor(100.0), has_nan=torch.tensor(True))) == 0.0
  for a, b in zip(before, t.vars.Flatten()):
    torch.testing.assert_close(a, b.detach())
  capped = _Make(layers.GradNormTracker, name='cap', clip_threshold=3.0,
                 grad_norm_clip_cap_min=math.exp(10.0))
  for v in normal[:200]:
    capped.FPropDefaultTheta(torch.tensor(float(v)))
  assert float(capped.FPropDefaultTheta(torch.tensor(float(np.exp(8.0))))) == 1.0


def test_fetch_layer_and_cct_gating_network():
  f = _Make(layers.FetchLayer)
  x = torch.randn(2, 3)
  out = f.FPropDefaultTheta(x)
  assert out is x or torch.equal(out, x)
  c = _Make(layers.CCTGatingNetwork, input_dim=4, hidden_layer_dim=6, num_outputs=3,
            noise_std=1.0, noise_warmup_steps=10)
  g = c.FPropDefaultTheta(torch.randn(5, 4))
  assert g.shape == (5, 3) and float(g.min()) >= 0 and float(g.max()) <= 1
  with cluster_factory.SetEval(True):
    ce = _Make(layers.CCTGatingNetwork, name='e', input_dim=4, hidden_layer_dim=6, num_outputs=3)
    ge = ce.FPropDefaultTheta(torch.randn(5, 4))
    assert set(ge.unique().tolist()) <= {0.0, 1.0}                  # hard gates at inference


def test_cond_scale_shift_ffn():
  l = _Make(layers.CondScaleShiftFFNLayer, input_dim=4, output_dim=3,
            ffn=layers.FeedForwardNet.Params().Set(hidden_layer_dims=[5], activation='RELU'),
            scale_fn='SIGMOID', shift_fn='TANH')
  scale, shift = l.FPropDefaultTheta(torch.randn(6, 4))
  assert scale.shape == shift.shape == (6, 3)
  assert 0 <= float(scale.min()) and float(scale.max()) <= 1 and float(shift.abs().max()) <= 1


# ------------------------------------------------------------------ multitask adapters --
@pytest.mark.parametrize('cls', [layers.MultitaskAdapterLayer, layers.MultitaskAdapterEinsumLayer])
def test_multitask_adapters(cls):
  l = _Make(cls, num_tasks=3, input_dim=4, bottleneck_dim=2, data_format='BTC')
  x = torch.randn(2, 5, 4)
  tasks = torch.tensor([0, 2])
  y = l.FPropDefaultTheta(x, tasks)
  assert y.shape == x.shape
  # residual adapter: zero up-projection ⇒ identity
  with torch.no_grad():
    for v in l.vars.Flatten():
      if 'up' in v.var_name:
        v.zero_()
  torch.testing.assert_close(l.FPropDefaultTheta(x, tasks), x, atol=1e-6, rtol=1e-6)
  per_step = torch.tensor([[0, 1, 2, 0, 1], [2, 2, 2, 2, 2]])
  assert l.FPropDefaultTheta(x, per_step).shape == x.shape
  tbc = _Make(cls, name='tbc', num_tasks=3, input_dim=4, bottleneck_dim=2, data_format='TBC')
  assert tbc.FPropDefaultTheta(x.transpose(0, 1), tasks).shape == (5, 2, 4)


# ------------------------------------------------------------------ statistical pooling --
def test_statistical_pooling_layers():
  l = _Make(layers.StatisticalPoolingLayer, has_stddev=True)
  x = torch.randn(2, 9, 3)
  pad = torch.zeros(2, 9)
  pad[1, 5:] = 1
  y = l.FProp(x, pad)
  assert y.shape == (2, 6)
  torch.testing.assert_close(y[0, :3], x[0].mean(0), atol=1e-5, rtol=1e-5)
  torch.testing.assert_close(y[1, :3], x[1, :5].mean(0), atol=1e-5, rtol=1e-5)
  torch.testing.assert_close(y[1, 3:], x[1, :5].var(0, unbiased=False).clamp_min(1e-6).sqrt(),
                             atol=1e-4, rtol=1e-4)
  pf = _Make(layers.PerFrameStatisticalPoolingLayer, has_stddev=False, left_context=-1,
             right_context=0)
  yf = pf.FProp(x, pad)
  assert yf.shape == (2, 9, 3)
  torch.testing.assert_close(yf[0, 3], x[0, :4].mean(0), atol=1e-5, rtol=1e-5)  # causal mean
  win = _Make(layers.PerFrameStatisticalPoolingLayer, name='w', has_stddev=False, left_context=1,
              right_context=1)
  yw = win.FProp(x, pad)
  torch.testing.assert_close(yw[0, 4], x[0, 3:6].mean(0), atol=1e-5, rtol=1e-5)


# ----------------------------------------------------------------------- LSH layers --
@pytest.mark.parametrize('cls', [layers.LSHMemoryRankKOneHotTaskLayer,
                                 layers.LSHTaskWithMultiplierLayer])
def test_lsh_task_layers(cls):
  l = _Make(cls, input_dim=6, output_dim=4, num_tasks=2, num_hash_bits=4, rank=2)
  x = torch.randn(7, 6)
  y0 = l.FPropDefaultTheta(x, torch.zeros(7, dtype=torch.long))
  y1 = l.FPropDefaultTheta(x, torch.ones(7, dtype=torch.long))
  assert y0.shape == (7, 4) and not torch.allclose(y0, y1)
  _GradsFlow(l, y0)


# --------------------------------------------------------------------- MLM augmenter --
def test_masked_lm_data_augmenter():
  l = _Make(layers.MaskedLmDataAugmenter, vocab_size=100, mask_prob=0.3, random_prob=0.1,
            same_prob=0.1, mask_token_id=99)
  ids = torch.randint(0, 90, (64, 32))
  pad = torch.zeros(64, 32)
  pad[:, 28:] = 1
  new_ids, mask = l.FPropDefaultTheta(ids, pad)
  assert new_ids.shape == ids.shape and mask.shape == ids.shape
  assert float(mask[:, 28:].sum()) == 0                              # never on paddings
  frac = float(mask[:, :28].float().mean())
  assert 0.4 < frac < 0.6                                            # 0.3 + 0.1 + 0.1
  changed = (new_ids != ids)
  assert (changed <= (mask > 0)).all()                               # only selected tokens change
  assert 0.2 < float((new_ids[:, :28] == 99).float().mean()) < 0.4
