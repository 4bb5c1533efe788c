"""Small reference helpers added in round 2, each against a straightforward oracle."""
import queue

import numpy as np
import pytest
import torch

from lingvo_b200.core import activations
from lingvo_b200.core import attention_util
from lingvo_b200.core import batch_major_attention as bma
from lingvo_b200.core import batch_utils
from lingvo_b200.core import bn_layers
from lingvo_b200.core import entmax
from lingvo_b200.core import flat_beam_search_helper as fbs
from lingvo_b200.core import gshard_builder
from lingvo_b200.core import gshard_decode
from lingvo_b200.core import gshard_utils
from lingvo_b200.core import hyperparams
from lingvo_b200.core import layers
from lingvo_b200.core import lstm_frnn_layer
from lingvo_b200.core.nested_map import NestedMap


def test_glu_variants():
  x = torch.randn(3, 8)
  a, b = x.chunk(2, -1)
  torch.testing.assert_close(activations.GetFn('GLU')(x), a * torch.sigmoid(b))
  torch.testing.assert_close(activations.GetFn('BILINEAR_GLU')(x), a * b)
  torch.testing.assert_close(activations.GetFn('SWISH_GLU')(x), a * torch.nn.functional.silu(b))
  torch.testing.assert_close(activations.GLUVariants(x, 'RELU'), a * torch.relu(b))
  assert activations.DimMultiplier('GELU_GLU') == 2 and activations.DimMultiplier('GELU') == 1
  assert activations.GetFlops('GLU') == 5


def test_extract_block_context_v2_matches_slices():
  b, t, d, w, l, r = 2, 7, 3, 3, 3, 2
  x = torch.randn(b, t, d)
  pad = torch.zeros(b, t); pad[1, 5:] = 1
  patches, ppad = attention_util.ExtractBlockContextV2(x, w, l, r, paddings=pad)
  u, c = 3, l - 1 + w + r
  assert patches.shape == (b, u, c, d) and ppad.shape == (b, u, c)
  for ui in range(u):
    for ci in range(c):
      src = ui * w - (l - 1) + ci
      if 0 <= src < t:
        torch.testing.assert_close(patches[:, ui, ci], x[:, src])
        assert torch.equal(ppad[:, ui, ci], pad[:, src])
      else:
        assert patches[:, ui, ci].abs().sum() == 0 and bool((ppad[:, ui, ci] == 1).all())
  assert attention_util.ExtractBlockContextV2(x, w, l, r)[1] is None


def test_entmax_general_alpha_and_loss():
  torch.manual_seed(0)
  x = torch.randn(4, 9, dtype=torch.float64, requires_grad=True)
  p15 = entmax.entmax_support(x, alpha=1.5)
  torch.testing.assert_close(p15, entmax.entmax15(x), atol=1e-6, rtol=1e-6)
  p2 = entmax.entmax_support(x, alpha=2.0)
  torch.testing.assert_close(p2, entmax.sparsemax(x), atol=1e-6, rtol=1e-6)
  p11 = entmax.entmax_support(x, alpha=1.05)
  assert (p11 > 0).float().mean() > (p2 > 0).float().mean()          # less sparse near softmax
  assert torch.allclose(p11.sum(-1), torch.ones(4, dtype=torch.float64))
  # gradient of the bisection version = gradient of the exact sort-based one
  w = torch.randn(4, 9, dtype=torch.float64)
  g_b, = torch.autograd.grad((p15 * w).sum(), x, retain_graph=True)
  g_e, = torch.autograd.grad((entmax.entmax15(x) * w).sum(), x)
  torch.testing.assert_close(g_b, g_e, atol=1e-5, rtol=1e-5)
  # loss: ≥ 0, zero iff the prediction is the (one-hot) label; gradient = p − y
  labels = torch.nn.functional.one_hot(torch.tensor([1, 3, 0, 8]), 9).double()
  loss = entmax.entmax_loss(labels, x)
  assert loss.shape == (4,) and bool((loss >= -1e-9).all())
  g, = torch.autograd.grad(loss.sum(), x)
  torch.testing.assert_close(g, p15.detach() - labels, atol=1e-6, rtol=1e-6)
  sure = entmax.entmax_loss(labels, labels * 50.0)
  assert float(sure.abs().max()) < 1e-6


def test_batch_utils_global_scaling():
  assert batch_utils.scale_global_to_infeed(64, False) == 64
  assert batch_utils.scale_global_to_worker(8) * 1 >= 1
  from lingvo_b200.core import cluster_factory
  n = max(int(cluster_factory.Current().total_worker_devices), 1)
  assert batch_utils.scale_global_to_worker(8 * n) == 8
  if n > 1:
    with pytest.raises(ValueError):
      batch_utils.scale_global_to_worker(8 * n + 1)


def test_adding_accumulator():
  acc = bn_layers.AddingAccumulator([2], torch.float32)
  assert acc.GetValue().tolist() == [0, 0]
  acc.Update(torch.tensor([1.0, 2.0]))
  acc.Update(torch.tensor([0.5, 0.5], dtype=torch.float64))
  assert acc.GetValue().tolist() == [1.5, 2.5]
  acc.Disable()
  assert acc.GetValue().tolist() == [0, 0]
  acc.Enable()
  assert acc.GetValue().tolist() == [1.5, 2.5]


def test_einsum_i32_and_infinite_repeat():
  a = torch.tensor([[1.0, 0.0], [1.0, 1.0]])
  out = fbs.einsum_i32('ij,jk->ik', a, a)
  assert out.dtype == torch.int32 and out.tolist() == [[1, 0], [2, 1]]
  q = queue.Queue()
  for i in range(4):
    q.put((i,))
  q.put(None)
  seen = []

  def Body(*args):
    total = (args[0] if len(args) == 2 else 0) + args[-1]
    seen.append(total)
    return [total]

  assert gshard_decode.infinite_repeat(Body, q) == [6] and seen == [0, 1, 3, 6]
  n = [0]

  def Count():
    n[0] += 1
    if n[0] == 5:
      raise StopIteration

  assert gshard_decode.infinite_repeat(Count) == [] and n[0] == 5


def test_kl_div_and_reshape_dim_and_sharding_spec():
  p = torch.softmax(torch.randn(2, 3, 5), -1)
  q = torch.softmax(torch.randn(2, 3, 5), -1)
  want = torch.nn.functional.kl_div(q.log(), p, reduction='sum') / 2
  torch.testing.assert_close(gshard_builder.KLDiv(p, q), want, atol=1e-5, rtol=1e-5)
  assert float(gshard_builder.KLDiv(p, p)) == pytest.approx(0.0, abs=1e-6)
  x = torch.arange(24).reshape(2, 12)
  assert gshard_utils.ReshapeDim(x, 1, 3).shape == (2, 3, 4)
  assert gshard_utils.ReshapeDim(x, -1, 4).shape == (2, 4, 3)
  assert gshard_utils.ReshapeDim(x, 1) is x
  mesh = np.arange(4).reshape(2, 2)
  with gshard_utils.MeshSplitDimPrefixContext(1):
    spec = gshard_utils.GetMeshSplitSharding(mesh, [0, -1])
  assert list(spec.split_dims_mapping) == [1, 0, -1]


def test_copy_fields_subset_and_conv_flops():
  a = hyperparams.Params(); b = hyperparams.Params()
  for p in (a, b):
    p.Define('x', 1, ''); p.Define('y', 2, ''); p.Define('sub', hyperparams.Params(), '')
  a.Set(x=10, y=20)
  a.sub.Define('z', 5, '')
  hyperparams.CopyFieldsSubsetTo(a, b, ['x', 'sub'])
  assert (b.x, b.y) == (10, 2) and b.sub.z == 5 and b.sub is not a.sub
  hyperparams.CopyFieldsSubsetTo(a, b, 'y')
  assert b.y == 20
  assert layers.Conv2DFlops([2, 8, 8, 3], [3, 3, 3, 16], (2, 2), 'SAME') == 2 * 4 * 4 * 27 * 16 * 2
  assert layers.Conv2DFlops([2, 8, 8, 3], [3, 3, 3, 16], (1, 1), 'VALID') == 2 * 6 * 6 * 27 * 16 * 2


def test_reshaped_multi_headed_projection_matches_plain():
  mesh = np.arange(4).reshape(2, 2)
  common = dict(input_dim=8, num_heads=2, dim_per_head=3, random_seed=5)
  plain = bma.MultiHeadedProjectionLayer.Params().Set(name='p', **common).Instantiate()
  resh = bma.ReshapedMultiHeadedProjectionLayer.Params().Set(
      name='p', device_mesh=mesh, **common).Instantiate()
  with torch.no_grad():
    resh.vars.w.copy_(plain.vars.w); resh.vars.b.copy_(torch.randn_like(plain.vars.b))
    plain.vars.b.copy_(resh.vars.b)
  x = torch.randn(2, 5, 8)
  torch.testing.assert_close(resh.FPropDefaultTheta(x.reshape(2, 5, 2, 4)),
                             plain.FPropDefaultTheta(x), atol=1e-5, rtol=1e-5)
  out_plain = bma.MultiHeadedProjectionLayer.Params().Set(
      name='o', is_output_projection=True, **common).Instantiate()
  out_resh = bma.ReshapedMultiHeadedProjectionLayer.Params().Set(
      name='o', is_output_projection=True, device_mesh=mesh, **common).Instantiate()
  with torch.no_grad():
    out_resh.vars.w.copy_(out_plain.vars.w)
    out_plain.vars.b.copy_(torch.randn(8)); out_resh.vars.b.copy_(out_plain.vars.b)
  h = torch.randn(2, 5, 2, 3)
  torch.testing.assert_close(out_resh.FPropDefaultTheta(h).reshape(2, 5, 8),
                             out_plain.FPropDefaultTheta(h), atol=1e-5, rtol=1e-5)


def test_lstm_cell_ext_projected_inputs_match_fprop():
  p = lstm_frnn_layer.LSTMCellSimpleExt.Params().Set(
      name='c', num_input_nodes=6, num_output_nodes=4, random_seed=3)
  cell = p.Instantiate()
  t, b = 5, 3
  x1, x2 = torch.randn(t, b, 2), torch.randn(t, b, 4)
  pad = torch.zeros(b, 1)
  proj = cell.ProjectInputSequence(cell.theta, NestedMap(act=[x1, x2]))
  assert proj.shape == (t, b, 16)
  sa = sb = cell.zero_state(cell.theta, b)
  for i in range(t):
    sa, _ = cell.FProp(cell.theta, sa, NestedMap(act=[x1[i], x2[i]], padding=pad))
    sb, _ = cell.FPropWithProjectedInput(cell.theta, sb,
                                         NestedMap(proj_inputs=proj[i], padding=pad))
  torch.testing.assert_close(sa.m, sb.m, atol=1e-5, rtol=1e-5)
  torch.testing.assert_close(sa.c, sb.c, atol=1e-5, rtol=1e-5)
  mixed = cell._MixWithProjectedInput(cell.theta, sa, proj[0])
  assert mixed.shape == (b, 16)


def test_cluster_device_strings_and_infeed_context():
  from lingvo_b200.core import cluster
  s = cluster.MakeDeviceString('/job:trainer', 1, 2, 'GPU', 3)
  assert s == '/job:trainer/replica:1/task:2/device:GPU:3'
  d = cluster.ParseDeviceString(s)
  assert (d.job, d.replica, d.task, d.device) == ('trainer', 1, 2, 'GPU')
  assert 'task' not in cluster.ParseDeviceString('/job:a/replica:0')
  base = cluster.GetInfeedContext()
  assert base.num_infeed_hosts >= 1
  with cluster.InfeedContextScope(3, 8):
    assert cluster.GetInfeedContext() == (3, 8)
    with cluster.InfeedContextScope(1, 2):
      assert cluster.GetInfeedContext().infeed_host_index == 1
    assert cluster.GetInfeedContext().num_infeed_hosts == 8
  assert cluster.GetInfeedContext() == base


def test_set_cluster_swaps_implementation():
  from lingvo_b200.core import cluster_factory
  orig = cluster_factory.Cluster

  class MyCluster(orig):
    marker = True

  try:
    cluster_factory.SetCluster(MyCluster)
    assert getattr(cluster_factory.Current(), 'marker', False) or \
        cluster_factory.Cluster is MyCluster
  finally:
    cluster_factory.SetCluster(orig)
  assert cluster_factory.Cluster is orig


def test_nested_map_assertions():
  import unittest
  from lingvo_b200.core import compare

  class T(compare.NestedMapAssertions):
    def runTest(self):
      pass

  t = T()
  a = NestedMap(x=torch.zeros(2, 3), y=NestedMap(z=torch.ones(4)))
  t.assertNestedMapEqual(a, a.DeepCopy())
  t.assertNestedMapEqual({'x': torch.zeros(2, 3), 'y': NestedMap(z=torch.ones(4))}, a)
  b = a.DeepCopy(); b.y.z = torch.ones(5)
  with pytest.raises(AssertionError) as e:
    t.assertNestedMapEqual(a, b)
  assert 'y.z' in str(e.value)
  with pytest.raises(AssertionError):
    compare.assertNestedMapEqual(None, a, b)
  del unittest


def test_standalone_adafactor_optimizer_matches_layer():
  from lingvo_b200.core import optimizer
  from lingvo_b200.core import py_utils
  assert optimizer.GetLrValue(0.5) == 0.5 and optimizer.GetLrValue(lambda: 0.25) == 0.25
  torch.manual_seed(0)
  w_a = torch.nn.Parameter(torch.randn(160, 130))
  b_a = torch.nn.Parameter(torch.randn(130))
  w_b, b_b = [torch.nn.Parameter(t.detach().clone()) for t in (w_a, b_a)]
  lr = [0.1]
  opt = optimizer.XLAShardingAdafactorOptimizer(
      [w_a, b_a], learning_rate=lambda: lr[0], decay_rate=lambda: 0.8, clipping_threshold=1.0)
  layer = optimizer.XLAShardingAdafactor.Params().Set(
      name='af', clipping_threshold=1.0, decay_exponent_pow=None).Instantiate()
  layer.DecayRate = lambda step=None: 0.8
  x = torch.randn(16, 160)
  for step in range(3):
    lr[0] = 0.1 / (step + 1)
    for w, b in ((w_a, b_a), (w_b, b_b)):
      w.grad = b.grad = None
      ((x @ w + b) ** 2).mean().backward()
    opt.step()
    layer.Apply(lr[0], [py_utils.VarGrad(w_b, w_b.grad), py_utils.VarGrad(b_b, b_b.grad)])
  torch.testing.assert_close(w_a, w_b)
  torch.testing.assert_close(b_a, b_b)
  # the 2-D variable is factored: row / column accumulators, no full second moment
  kinds = {k for slots in opt.slots().values() for k in slots}
  assert {'vr', 'vc', 'v'} <= kinds


def test_nested_map_and_params_small_methods():
  import dataclasses
  m = NestedMap(a=1, b=NestedMap(c=[NestedMap(d=5), 7]))
  assert m.Keys() == ['a', 'b.c[0].d', 'b.c[1]']
  assert m.GetSlice(['a', 'b.c[0].d']) == NestedMap(a=1, b=NestedMap(c=[NestedMap(d=5)]))
  u = NestedMap(a=1, x=NestedMap(y=2)).Union(NestedMap(a=9, x=NestedMap(z=3)))
  assert u == NestedMap(a=9, x=NestedMap(y=2, z=3))
  base = NestedMap(a=1)
  assert base.Update(NestedMap(b=NestedMap(c=2))) is base and base.b.c == 2
  assert NestedMap.SquareBracketIndex('k[12]') == ('k', 12)
  assert NestedMap.SquareBracketIndex('k') == ('k', None)

  @dataclasses.dataclass
  class Inner:
    v: int = 3

  @dataclasses.dataclass
  class Outer:
    name: str = 'n'
    inner: Inner = dataclasses.field(default_factory=Inner)

  nm = NestedMap.FromNestedDataclass(Outer())
  assert nm.name == 'n' and nm.inner.v == 3 and isinstance(nm.inner, NestedMap)
  with pytest.raises(ValueError):
    NestedMap.FromNestedDataclass({'a': 1})

  p = hyperparams.Params()
  p.Define('x', 5, 'x'); p.Define('y', None, 'y'); p.Define('sub', hyperparams.Params(), '')
  p.sub.Define('z', 0, '')
  p.x = 6
  assert p.ParamIsSet('x') and not p.ParamIsSet('y') and p.ParamIsSet('sub.z')
  with pytest.raises(AttributeError):
    p.ParamIsSet('nope')
  assert p._slots['x'].GetDefault() == 5 and p.Copy()._slots['x'].GetDefault() == 5
  assert p._slots['x'].ToString(1) == '  x: 6'
  q = hyperparams.Params()
  q.Define('x', 0, ''); q.Define('other', 1, '')
  q.MergeCommonKeysFrom(p)
  assert q.x == 6 and q.other == 1
  with pytest.raises(AttributeError):
    hyperparams.CopyFieldsTo(p, q)                       # unknown keys are an error by default
  assert hyperparams.CopyFieldsTo(p, q, skip='y', ignore_unknown_keys=True).x == 6


def test_single_shard_softmax_chunked_xent_matches_dense():
  common = dict(name='sm', input_dim=6, num_classes=9, random_seed=4)
  dense = layers.SingleShardFullSoftmax.Params().Set(**common).Instantiate()
  chunked = layers.SingleShardFullSoftmax.Params().Set(chunk_size=4, **common).Instantiate()
  x = torch.randn(12, 6, requires_grad=True)
  ids = torch.randint(0, 9, (12, 1))
  w = torch.rand(12, 1)
  a = dense.FProp(dense.theta, x, w, class_ids=ids)
  b = chunked.FProp(chunked.theta, x, w, class_ids=ids)
  torch.testing.assert_close(a.per_example_xent, b.per_example_xent)
  torch.testing.assert_close(a.avg_xent, b.avg_xent)
  assert torch.equal(a.per_example_argmax, b.per_example_argmax) and b.logits is None
  ga, = torch.autograd.grad(a.total_xent, x, retain_graph=True)
  gb, = torch.autograd.grad(b.total_xent, x)
  torch.testing.assert_close(ga, gb)
  probs = torch.softmax(torch.randn(12, 9), -1)
  xe, _ = chunked.XentLossByChunk(chunked.theta, x, None, probs)
  want = -(probs * torch.log_softmax(dense.Logits(dense.theta, x), -1)).sum(-1)
  torch.testing.assert_close(xe, want, atol=1e-5, rtol=1e-5)
  dw = dense.DenseWeights(dense.theta)
  assert dw.wm.shape == (6, 9) and dw.b.shape == (9,)
  with pytest.raises(AssertionError):
    chunked.XentLossByChunk(chunked.theta, x[:10], ids[:10].reshape(-1))


def test_bleu_free_functions_and_spectrum_augmenter_einsum_hooks():
  import numpy as np
  import torch
  from lingvo_b200.core import ml_perf_bleu_metric as bleu
  from lingvo_b200.core import spectrum_augmenter
  assert bleu.native_to_unicode(b'caf\xc3\xa9') == 'café' and bleu.is_unicode('x')
  assert bleu.native_to_unicode(b'\xff-ok') == '-ok'
  ids = np.array([[5, 6, 7, 8, 9, 0, 0], [3, 4, 5, 6, 7, 8, 0]])
  score, weight = bleu.bleu_score(torch.from_numpy(ids), ids)
  assert abs(score - 1.0) < 1e-6 and weight == 1.0
  worse, _ = bleu.bleu_score(np.array([[5, 6, 1, 8, 9, 0, 0], [3, 4, 5, 6, 2, 8, 0]]), ids)
  assert 0.0 <= worse < 1.0
  aug = spectrum_augmenter.SpectrumAugmenter.Params().Set(name='aug').Instantiate()
  a, b = torch.randn(2, 3, 4, 1), torch.randn(2, 5, 3)
  torch.testing.assert_close(aug.EinsumBxycBzxBzyc(a, b), torch.einsum('bxyc,bzx->bzyc', a, b))
  torch.testing.assert_close(aug.EinsumBBmBm(torch.ones(2), torch.ones(2, 3)), torch.ones(2, 3))
