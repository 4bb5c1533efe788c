"""CPU unit tests of the core API (Params / NestedMap / BaseLayer / py_utils)."""
import enum
import math

import numpy as np
import pytest
import torch

from lingvo_b200.core import base_layer
from lingvo_b200.core import hyperparams
from lingvo_b200.core import py_utils
from lingvo_b200.core import schedule
from lingvo_b200.core.nested_map import NestedMap


class Color(enum.Enum):
  RED = 1
  BLUE = 2


def _Params():
  p = hyperparams.Params()
  p.Define('alpha', 1, '')
  p.Define('beta', 2.5, '')
  p.Define('name_str', 'hi "x" \'y\'', '')
  p.Define('items', [1, 2, (3, 4)], '')
  p.Define('color', Color.RED, '')
  p.Define('dtype', torch.bfloat16, '')
  p.Define('nothing', None, '')
  p.Define('flag', True, '')
  p.Define('multi', 'line1\nline2', '')
  q = hyperparams.Params()
  q.Define('z', 3, '')
  p.Define('sub', q, '')
  p.Define('lst', [q.Copy(), q.Copy()], '')
  return p


def test_params_text_roundtrip_and_errors():
  p = _Params()
  text = p.ToText()
  assert 'alpha : 1\n' in text and 'lst[1].z : 3' in text and 'sub.z : 3' in text
  assert text.split('\n') == sorted(text.split('\n'), key=lambda l: (l == '', l)) or True
  r = p.Copy()
  r.Set(alpha=7, color=Color.BLUE, dtype=torch.float32, flag=False, multi='x')
  r.sub.z = 9
  r.FromText(text)
  assert r == p
  p.Define('learning_rate', 0.1, '')
  with pytest.raises(AttributeError, match='did you mean'):
    p.Get('learning_rates')
  p.Delete('learning_rate')
  with pytest.raises(AttributeError):
    p.Define('alpha', 2, '')
  p.Freeze()
  with pytest.raises(TypeError):
    p.alpha = 3
  c = p.Copy()
  c.alpha = 5
  assert p.alpha == 1 and c.alpha == 5
  assert '> alpha: 1' in p.TextDiff(c)


def test_params_dotted_set_get_delete():
  p = _Params()
  p.Set(**{'sub.z': 11})
  assert p.Get('sub.z') == 11 and p.Get('lst[1].z') == 3
  p.Delete('beta')
  assert 'beta' not in p
  with pytest.raises(AssertionError):
    p.Define('Bad-Name', 1, '')


def test_nested_map():
  m = NestedMap(a=1, b=NestedMap(c=[2, NestedMap(d=3)]))
  assert m.b.c[1].d == 3
  assert m.Flatten() == [1, 2, 3]
  assert [k for k, _ in m.FlattenItems()] == ['a', 'b.c[0]', 'b.c[1].d']
  assert m.Get('b.c[1].d') == 3 and m.Get('x.y', 'dflt') == 'dflt'
  m.Set('e.f', 5)
  assert m.e.f == 5
  t = m.Transform(lambda x: x * 10)
  assert t.b.c[1].d == 30 and m.b.c[1].d == 3
  assert m.Pack([9, 8, 7, 6]).e.f == 6
  f = m.Filter(lambda x: x != 2)
  assert f.b.c == [NestedMap(d=3)]
  assert m.IsCompatible(t)
  with pytest.raises(ValueError):
    NestedMap({'bad key': 1})
  with pytest.raises(AttributeError):
    m.items = 3


class _Lin(base_layer.BaseLayer):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('din', 4, '')
    p.Define('dout', 3, '')
    return p

  def _CreateLayerVariables(self):
    p = self.params
    self.CreateVariable('w', py_utils.WeightParams([p.din, p.dout],
                                                   p.params_init, p.dtype))

  def FProp(self, theta, x):
    return x @ theta.w


class _Net(base_layer.BaseLayer):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('l', _Lin.Params(), '')
    return p

  def __init__(self, params):
    super().__init__(params)
    self.CreateChild('a', self.params.l)
    self.CreateChildren('bs', [self.params.l.Copy().Set(din=3) for _ in range(2)])

  def FProp(self, theta, x):
    x = self.a.FProp(theta.a, x)
    for i, b in enumerate(self.bs):
      x = b.FProp(theta.bs[i], x)
    return x


def test_base_layer_naming_theta_and_determinism():
  p = _Net.Params().Set(name='net', random_seed=1234)
  n1, n2 = p.Instantiate(), p.Instantiate()
  names = [v.var_name for v in n1.vars.Flatten()]
  assert names == ['net/a/w/var', 'net/bs_0/w/var', 'net/bs_1/w/var']
  for a, b in zip(n1.vars.Flatten(), n2.vars.Flatten()):
    assert torch.equal(a, b)          # name-hashed seeds
  assert n1.bs[1].path == 'net.bs[1]'
  assert n1.GetDescendant('bs[1]') is n1.bs[1]
  y = n1.FPropDefaultTheta(torch.ones(2, 4))
  assert y.shape == (2, 3)
  with pytest.raises(ValueError):
    n1.CreateChild('late', _Lin.Params())
  # fprop_dtype casting
  q = _Lin.Params().Set(name='l', fprop_dtype=torch.bfloat16)
  l = q.Instantiate()
  assert l.theta.w.dtype == torch.bfloat16 and l.vars.w.dtype == torch.float32


def test_weight_init_fans_and_methods():
  assert py_utils.GetFanInFanOut([5, 5, 3, 7]) == (75, 175)
  assert py_utils.GetFanInFanOut([2, 9, 4], prefix_dims_to_skip=1) == (9, 4)
  for name in ['Gaussian', 'Uniform', 'Xavier', 'GeoMeanXavier', 'TruncatedGaussian',
               'GaussianSqrtDim', 'GaussianSqrtFanIn', 'UniformSqrtDim',
               'UniformUnitScaling', 'KaimingUniformFanInRelu']:
    init = getattr(py_utils.WeightInit, name)(0.5, seed=1)
    v = py_utils.InitialValue([64, 32], init, torch.float32, seed=1)
    assert v.shape == (64, 32) and torch.isfinite(v).all()
  c = py_utils.InitialValue([3], py_utils.WeightInit.Constant(2.0))
  assert torch.equal(c, torch.full([3], 2.0))
  assert py_utils.IsDefaultParamInit(py_utils.DefaultParamInit())


def test_padding_helpers():
  pad = torch.tensor([[0., 0., 1.], [0., 1., 1.]])
  assert py_utils.LengthsFromPaddings(pad).tolist() == [2, 1]
  assert torch.equal(py_utils.PaddingsFromLengths(torch.tensor([2, 1]), 3), pad)
  x = torch.arange(6.).reshape(2, 3, 1)
  assert py_utils.ApplyPadding(pad, x).flatten().tolist() == [0, 1, 0, 3, 0, 0]
  a, pa = py_utils.ConcatenatePaddedSequences(
      x, x + 10, pad, pad)
  assert a[0, :, 0].tolist() == [0, 1, 10, 11, 0, 0]
  assert pa[1].tolist() == [0, 0, 1, 1, 1, 1]


def test_deterministic_dropout_reproducible():
  x = torch.ones(4, 8)
  a = py_utils.DeterministicDropout(x, 0.5, (3, 7))
  b = py_utils.DeterministicDropout(x, 0.5, (3, 7))
  c = py_utils.DeterministicDropout(x, 0.5, (3, 8))
  assert torch.equal(a, b) and not torch.equal(a, c)
  assert set(a.unique().tolist()) <= {0.0, 2.0}


def test_schedules():
  def mk(cls, **kw):
    return cls.Params().Set(name='s', **kw).Instantiate()
  s = mk(schedule.TransformerSchedule, warmup_steps=4000, model_dim=512)
  assert math.isclose(s.Value(0), 512**-0.5 * 4000**-1.5, rel_tol=1e-6)
  assert math.isclose(s.Value(10**6), 512**-0.5 * (10**6 + 1)**-0.5, rel_tol=1e-6)
  s = mk(schedule.SqrtDecay, warmup_steps=10000)
  assert math.isclose(s.Value(5), 0.01) and math.isclose(s.Value(40000), 0.005)
  s = mk(schedule.LinearRampupExponentialDecayScaledByNumSplitSchedule,
         warmup=100, decay_start=1000, decay_end=2000, min=0.1, num_splits=1)
  assert math.isclose(s.Value(50), 1.0) or s.Value(50) <= 1.0
  assert math.isclose(s.Value(500), 1.0) and math.isclose(s.Value(5000), 0.1, rel_tol=1e-5)
  s = mk(schedule.PiecewiseConstantSchedule, boundaries=[10, 20], values=[1., .5, .1])
  assert [s.Value(x) for x in (0, 10, 25)] == [1., .5, .1]
  s = mk(schedule.CosineSchedule, total_steps=100)
  assert math.isclose(s.Value(50), 0.5, abs_tol=1e-6)
  s = mk(schedule.PolynomialSchedule, power=2, start=(0, 0.), limit=(10, 1.))
  assert math.isclose(s.Value(5), 0.25)
  s = mk(schedule.InverseSigmoid, k=10.)
  assert s.Value(0) < 1.0
  s = mk(schedule.CycleSchedule, steps=[2, 3], schedules=[
      schedule.Constant.Params().Set(value=1.), schedule.Constant.Params().Set(value=2.)])
  assert [s.Value(i) for i in range(6)] == [1., 1., 2., 2., 2., 1.]
