"""Oracle checks for the py_utils helpers added in round 2 (ref lingvo/core/py_utils_test.py)."""
import numpy as np
import pytest
import torch

from lingvo_b200.core import hyperparams
from lingvo_b200.core import py_utils
from lingvo_b200.core.nested_map import NestedMap


def test_create_ids_and_labels():
  ids = torch.tensor([[3, 4, 5, 9], [6, 7, 9, 9]])
  pad = torch.tensor([[0., 0, 0, 1], [0, 0, 1, 1]])
  t = py_utils.CreateIdsAndLabels(ids, pad)
  assert t.ids.tolist() == [[1, 3, 4, 5, 2], [1, 6, 7, 2, 2]]
  assert t.labels.tolist() == [[3, 4, 5, 2, 2], [6, 7, 2, 2, 2]]
  assert t.paddings.tolist() == [[0, 0, 0, 0, 1], [0, 0, 0, 1, 1]]
  assert torch.equal(t.weights, 1 - t.paddings)
  t2 = py_utils.CreateIdsAndLabels(ids, pad, trim=True)
  assert t2.ids.shape == (2, 4) and t2.labels.tolist()[0] == [3, 4, 5, 2]


def test_merge_duplicate_ids():
  ids = torch.tensor([[4, 4, 5, 6, 6, 5, 0, 0], [1, 1, 1, 1, 2, 2, 3, 3]])
  pad = torch.tensor([[0., 0, 0, 0, 0, 0, 1, 1], [0] * 8])
  extra = {'x': torch.arange(16, dtype=torch.float32).reshape(2, 8, 1) + 1}
  rid, rpad, rx = py_utils.MergeDuplicateIds(ids, pad, extra)
  assert rid.tolist() == [[4, 5, 6, 5, 0, 0, 0, 0], [1, 2, 3, 0, 0, 0, 0, 0]]
  assert rpad.tolist() == [[0, 0, 0, 0, 1, 1, 1, 1], [0, 0, 0, 1, 1, 1, 1, 1]]
  assert rx.x[0, :, 0].tolist() == [1, 3, 4, 6, 0, 0, 0, 0]
  assert rx.x[1, :, 0].tolist() == [9, 13, 15, 0, 0, 0, 0, 0]


def test_mix_by_weight_only_advances_chosen():
  calls = [0, 0]

  def Make(i):
    def Fn():
      calls[i] += 1
      return torch.tensor(float(i))
    return Fn

  picks = []
  for s in range(200):
    v, onehot = py_utils.MixByWeight([Make(0), Make(1)], [0.25, 0.75], seed=s)
    assert int(onehot.argmax()) == int(v)
    picks.append(int(v))
  assert sum(calls) == 200
  assert 0.6 < np.mean(picks) < 0.9


def test_pad_sequence_to_and_expand():
  x = torch.ones(2, 3, 4)
  pad = torch.zeros(2, 3)
  y, p = py_utils.PadSequenceTo(x, pad, 5, 7.0)
  assert y.shape == (2, 5, 4) and float(y[0, 4, 0]) == 7.0
  assert p.tolist() == [[0, 0, 0, 1, 1]] * 2
  (a, b), p2 = py_utils.PadSequenceTo([x, x[..., 0]], pad, 4, 0)
  assert a.shape == (2, 4, 4) and b.shape == (2, 4) and p2.shape == (2, 4)
  e = py_utils.ExpandAndPadOrTrimTo(torch.ones(2, 3), [2, 5, 7])
  assert e.shape == (2, 5, 1) and e[:, 3:].abs().sum() == 0
  assert py_utils.ExpandTo(torch.ones(2), 3).shape == (2, 1, 1)
  assert py_utils.AppendDims(torch.ones(2), 2).shape == (2, 1, 1)


def test_causal_padding_and_gathers():
  c = py_utils.CausalSelfAttenPadding(3)
  assert c.tolist() == [[0, 1, 1], [0, 0, 1], [0, 0, 0]]
  t = torch.randn(2, 3, 5)
  idx = torch.randint(0, 5, (2, 3))
  g = py_utils.GatherTensorValuesBySeqIndices(t, idx)
  for b in range(2):
    for s in range(3):
      assert g[b, s] == t[b, s, idx[b, s]]
  pr = py_utils.GetSoftmaxProbsBySeqIndices(t, idx, keepdims=True)
  assert pr.shape == (2, 3, 1)
  np.testing.assert_allclose(pr[..., 0].numpy(),
                             torch.softmax(t, -1).gather(-1, idx[..., None])[..., 0].numpy(),
                             rtol=1e-6)


def test_numeric_helpers():
  x = torch.tensor([1.0, 2.0, 3.0])
  y = torch.tensor([2.0, 0.0, 4.0])
  assert py_utils.DivideNoNan(x, y).tolist() == [0.5, 0.0, 0.75]
  assert abs(float(py_utils.ReduceRms(x)) - np.sqrt(14 / 3)) < 1e-6
  assert float(py_utils.SumAbs([x, None, -y])) == 12.0
  assert not bool(py_utils.HasNanOrInf(x))
  assert bool(py_utils.HasNanOrInf([x, torch.tensor([float('nan')])]))
  assert bool(py_utils.HasNanOrInf(NestedMap(a=torch.tensor([float('inf')]))))
  capped = py_utils.MaybeSoftCapLogits(torch.tensor([100.0, -100.0, 0.1]), 5.0)
  assert float(capped.abs().max()) <= 5.0 and abs(float(capped[2]) - 0.1) < 1e-3
  assert py_utils.MaybeSoftCapLogits(x, 0.0) is x
  assert py_utils.clip_by_value(x, 1.5, 2.5).tolist() == [1.5, 2.0, 2.5]
  v, i = py_utils.TopK(torch.tensor([[1., 5, 3]]), 2)
  assert v.tolist() == [[5, 3]] and i.tolist() == [[1, 2]]
  assert py_utils.ArgMax(torch.tensor([[1., 5, 3]])).tolist() == [1]
  for step, want in [(0, 1.0), (9, 1.0), (10, 0.5), (25, 0.1)]:
    assert abs(float(py_utils.PiecewiseConstant(step, [10, 20], [1.0, 0.5, 0.1])) - want) < 1e-6


def test_asserts():
  assert py_utils.assert_greater(torch.tensor([2, 3]), 1)
  with pytest.raises(AssertionError):
    py_utils.assert_less(torch.tensor([2, 3]), 3)
  assert py_utils.assert_less_equal(torch.tensor([2, 3]), 3)
  assert py_utils.assert_greater_equal(3, 3)
  with pytest.raises(AssertionError):
    py_utils.Assert(torch.tensor(False), ['boom'])
  assert py_utils.AssertIdShape([2, None], [2, 7], [2, 7])
  with pytest.raises(AssertionError):
    py_utils.AssertIdShape([2, None], [3, 7])
  with pytest.raises(AssertionError):
    py_utils.AssertIdShape([2, None], [2, 7], [2, 8])
  assert py_utils.with_dependencies([1], 'x') == 'x'


def test_structure_helpers():
  a = NestedMap(x=1, y=NestedMap(z=[1, 2]))
  b = NestedMap(x='q', y=NestedMap(z=[3, 4]))
  assert py_utils.IsCompatible(a, b)
  assert not py_utils.IsCompatible(a, NestedMap(x=1))
  with pytest.raises(ValueError):
    py_utils.AssertIsCompatible(a, NestedMap(x=1, y=NestedMap(z=[1])))
  assert py_utils.Chunked([1, 2, 3, 4]) == [(1, 2), (3, 4)]
  t = torch.ones(1)
  assert len(py_utils.ToUniqueList(NestedMap(a=t, b=t, c=torch.ones(1)))) == 2
  d = py_utils.MergeDictsWithValueCheck({'a': t}, {'a': t, 'b': 2})
  assert set(d) == {'a', 'b'}
  with pytest.raises(RuntimeError):
    py_utils.MergeDictsWithValueCheck({'a': t}, {'a': torch.ones(1)})
  view = py_utils.ReadOnlyAttrDictView({'k': 3})
  assert view.k == 3 and view['k'] == 3 and 'k' in view and len(view) == 1
  with pytest.raises(AttributeError):
    view.k = 4
  with pytest.raises(AttributeError):
    view['k'] = 4
  with pytest.raises(AttributeError):
    _ = view.missing
  assert py_utils.Pack(NestedMap(a=0, b=0), [1, 2]) == NestedMap(a=1, b=2)
  assert py_utils.HasSameShape(torch.ones(2, 3), torch.zeros(2, 3)) is not None


def test_file_pattern_helpers():
  assert py_utils.ShardedFilePatternToGlob('/x/y@8') == '/x/y-?????-of-00008'
  assert py_utils.ShardedFilePatternToGlob('/x/y@*') == '/x/y-?????-of-*'
  assert py_utils.ShardedFilePatternToGlob('/x/y') == '/x/y'
  with pytest.raises(ValueError):
    py_utils.ShardedFilePatternToGlob('a@2,b@3')
  assert py_utils.RecordFormatFromFilePattern('tfrecord:/a/b*') == ('tfrecord', '/a/b*')
  assert py_utils.RecordFormatFromFilePattern('/a/b*') == ('sstable', '/a/b*')
  assert py_utils.SanitizeScopeKey('_a[0]') == 'a_0'
  s = py_utils.GenerateSeedFromId(17)
  assert s == py_utils.GenerateSeedFromId(17) and s != py_utils.GenerateSeedFromId(18)
  assert 0 <= s < 2**31 - 1


def test_gradient_helpers():
  v1, v2 = torch.ones(2), torch.ones(3)
  v1.var_name, v2.var_name = 'a', 'b'
  vg = NestedMap(a=py_utils.VarGrad(v1, torch.full((2,), 2.0)), b=py_utils.VarGrad(v2, None))
  kept = py_utils.SkipNoneGradients(vg)
  assert list(kept.keys()) == ['a']
  masked = py_utils.MaskGradients(kept, {'a': torch.tensor([1.0, 0.0])})
  assert masked.a.grad.tolist() == [2.0, 0.0]
  xs = NestedMap(a=torch.ones(2), b=torch.ones(3))
  dxs = NestedMap(a=None, b=torch.full((3,), 5.0))
  out = py_utils.ConvertNoneGradientToZeros(xs, [None, dxs.b])
  assert out.a.tolist() == [0, 0] and out.b.tolist() == [5, 5, 5]


def test_nce_and_auc():
  torch.manual_seed(0)
  targets = (torch.rand(4, 50) > 0.5).float()
  mask = torch.ones(4, 50)
  good = targets * 0.9 + 0.05
  nce, auc = py_utils.ComputeNceAndAuc(good, targets, mask)
  assert float(nce) > 0.6 and float(auc) > 0.95
  rnd = torch.rand(4, 50)
  nce2, auc2 = py_utils.ComputeNceAndAuc(rnd, targets, mask)
  assert float(nce2) < 0.1 and 0.3 < float(auc2) < 0.7


def test_uniform_sampler_is_uniform():
  counts = np.zeros(20)
  for seed in range(300):
    s = py_utils.UniformSampler(5, seed=seed)
    for i in range(20):
      s.Add(i)
    assert len(s.samples) == 5
    for i in s.samples:
      counts[i] += 1
  assert counts.min() > 40 and counts.max() < 110   # expectation 75


def test_scopes_and_control_flow():
  assert py_utils.GetTaskCallScope() is None
  with py_utils.TaskCallScope('t1'):
    assert py_utils.GetTaskCallScope() == 't1'
    with py_utils.TaskCallScope('t2'):
      assert py_utils.GetTaskCallScope() == 't2'
    assert py_utils.GetTaskCallScope() == 't1'
  assert py_utils.GetTaskCallScope() is None
  with py_utils.SampleStep(3) as s:
    assert s == 3
  assert py_utils.ForLoop(lambda i, st: st + i, 0, 5, 1, 0) == 10
  assert py_utils.WhileLoop(lambda st: st < torch.tensor(7), lambda st: st + 2, 1) == 7
  assert py_utils.If(torch.tensor(True), 3, lambda x: x + 1, lambda x: x - 1) == 4
  assert py_utils.IsEagerMode() and not py_utils.IsTpuTraining()
  with py_utils.VariableScope(['a', 'b']):
    assert py_utils.GetVariableName('w').endswith('b/w')


def test_update_dtype_walks_nested_params():
  from lingvo_b200.core import layers
  p = layers.FeedForwardNet.Params().Set(name='ffn', input_dim=4, hidden_layer_dims=[4, 4])
  py_utils.UpdateDtype(p, torch.float64)
  py_utils.UpdateFpropDtype(p, torch.bfloat16)
  assert p.dtype == torch.float64 and p.fprop_dtype == torch.bfloat16
  assert p.projection.dtype == torch.float64 and p.projection.fprop_dtype == torch.bfloat16
  assert isinstance(p, hyperparams.Params)


def test_rnn_cell_state_init():
  from lingvo_b200.core import rnn_cell
  z = py_utils.InitRNNCellState([2, 3])
  assert z.abs().sum() == 0
  init = py_utils.RNNCellStateInit.RandomNormal(seed=7)
  a = py_utils.InitRNNCellState([2, 3], init=init, name='s')
  b = py_utils.InitRNNCellState([2, 3], init=init, name='s')
  assert torch.equal(a, b) and a.abs().sum() > 0
  assert py_utils.InitRNNCellState([2, 3], init=init, is_eval=True).abs().sum() == 0
  p = rnn_cell.LSTMCellSimple.Params().Set(name='c', num_input_nodes=4, num_output_nodes=5,
                                           zero_state_init_params=init)
  cell = p.Instantiate()
  st = cell.zero_state(cell.theta, 3)
  assert st.m.shape == (3, 5) and st.m.abs().sum() > 0 and not torch.equal(st.m, st.c)
  cell_eval = p.Copy().Set(is_eval=True).Instantiate() if 'is_eval' in p else None
  if cell_eval is not None:
    assert cell_eval.zero_state(cell_eval.theta, 3).m.abs().sum() == 0


def test_call_defun_custom_backward():
  calls = {'bak': 0}

  def Fwd(xs):
    return NestedMap(y=xs.a * xs.b, n=xs.k + 1, aux=[xs.a.sum()])

  def Bak(xs, ys, dys):
    calls['bak'] += 1
    # deliberately NOT the true gradient of a (×2) so we can tell it was used
    return NestedMap(a=2.0 * dys.y * xs.b + dys.aux[0], b=dys.y * xs.a, k=None)

  a = torch.randn(3, requires_grad=True)
  b = torch.randn(3, requires_grad=True)
  ys = py_utils.CallDefun(Fwd, NestedMap(a=a, b=b, k=torch.tensor(4)), bak=Bak)
  assert int(ys.n) == 5 and isinstance(ys.aux, list)
  (ys.y.sum() + 3.0 * ys.aux[0]).backward()
  assert calls['bak'] == 1
  np.testing.assert_allclose(a.grad.numpy(), (2.0 * b + 3.0).detach().numpy(), rtol=1e-6)
  np.testing.assert_allclose(b.grad.numpy(), a.detach().numpy(), rtol=1e-6)
  # no bak: plain call
  out = py_utils.CallDefun(lambda xs: xs * 2, torch.ones(2))
  assert out.tolist() == [2, 2]


def test_function_decorator_and_simple_gradients():
  @py_utils.Function(bak=lambda xs, ys, dys: dys * 10.0)
  def Ident(x):
    return x.clone()

  x = torch.ones(2, requires_grad=True)
  Ident(x).sum().backward()
  assert x.grad.tolist() == [10.0, 10.0]
  w = torch.tensor([1.0, 2.0], requires_grad=True)
  u = torch.tensor([1.0], requires_grad=True)
  g = py_utils.ComputeGradientsSimple((w * w).sum(), [w, u])
  assert g[0].tolist() == [2.0, 4.0] and g[1] is None
  with py_utils.GradientTape():
    assert py_utils.CurrentGradientTape() is None
  vn = py_utils.DisableVN()
  assert not vn.global_vn and not vn.per_step_vn


def test_misc_runtime_helpers(tmp_path):
  with py_utils.VariableListDtypeRegexScope([('.*bias', torch.float16)]):
    assert py_utils.FindDataType('a/bias') == torch.float16
    assert py_utils.FindDataType('a/w') is None
  x = torch.arange(3.0)
  assert py_utils.Save(x, str(tmp_path / 'dbg'), x=x) is x
  files = list(tmp_path.glob('dbg.*.x.npy'))
  assert len(files) == 1 and np.load(files[0]).tolist() == [0, 1, 2]
  with py_utils.RemoveAssertContext():
    py_utils.HasShape(torch.ones(2, 3), [5, 5])      # asserts are off
  with pytest.raises(Exception):
    py_utils.HasShape(torch.ones(2, 3), [5, 5])
  py_utils.SetShapes(NestedMap(a=torch.ones(2)), NestedMap(a=torch.zeros(2)))
  with py_utils.outside_all_rewrites():
    assert py_utils.RunOnTpuHost(lambda v: v + 1, 1) == 2
  n = {'c': 0}

  @py_utils.RetryOnTransientTfError(initial_delay_sec=0.001, max_retries=5)
  def Flaky():
    n['c'] += 1
    if n['c'] < 3:
      raise OSError('transient')
    return 'ok'

  assert Flaky() == 'ok' and n['c'] == 3


def test_override_vars_from_checkpoints(tmp_path):
  from lingvo_b200.core import layers, saver as saver_lib
  p = layers.ProjectionLayer.Params().Set(name='proj', input_dim=3, output_dim=2, has_bias=True)
  src = p.Instantiate()
  with torch.no_grad():
    src.vars.w.fill_(0.5)
    src.vars.b.fill_(-1.0)
  sv = saver_lib.Saver(str(tmp_path), lambda: {v.var_name: v for v in src.vars.Flatten()})
  sv.Save(7)
  dst = p.Copy().Set(name='other').Instantiate()
  claimed = py_utils.OverrideVarsFromCheckpoints(
      dst.vars.Flatten(), {str(tmp_path): ([('other/(.*)', 'proj/%s')], ['.*/b$'])})
  assert list(claimed) == ['other/w']
  assert float(dst.vars.w.mean()) == 0.5 and float(dst.vars.b.abs().sum()) == 0.0
