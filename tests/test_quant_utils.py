"""Quantization domains, clipping-cap schedules and QuantizableLayer hooks
(ref `lingvo/core/quant_utils_test.py`)."""

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from lingvo_b200.core import cluster_factory
from lingvo_b200.core import py_utils
from lingvo_b200.core import quant_test_lib
from lingvo_b200.core import quant_utils
from lingvo_b200.core.quant_utils import QDistribution


@pytest.fixture(autouse=True)
def _ResetStep():
  py_utils.SetGlobalStep(0)
  yield
  py_utils.SetGlobalStep(0)


def _Layer(qdomain=None, **kw):
  p = quant_test_lib.SampleQuantizedProjectionLayer.Params().Set(name='proj', **kw)
  if qdomain is not None:
    p.qdomain.default = qdomain
  layer = p.Instantiate()
  layer.InstantiateVariables()
  return layer


# ---------------------------------------------------------------- fake quant primitive --
def test_fake_quant_grid_contains_zero_and_has_256_levels():
  x = torch.linspace(-2.0, 2.0, 20001)
  y = quant_utils.FakeQuantWithMinMax(x, -1.0, 1.0, num_bits=8)
  levels = torch.unique(y)
  assert levels.numel() == 256
  assert (levels == 0).any()                       # zero is exactly representable
  step = (levels[1:] - levels[:-1])
  np.testing.assert_allclose(step.numpy(), 2.0 / 255, rtol=1e-4)
  # nudging: [-1, 1] has zero point 127.5 → the grid is shifted by half a quantum so that
  # zero lands on a level; the span stays 255 quanta
  np.testing.assert_allclose(float(levels[-1] - levels[0]), 2.0, rtol=1e-5)
  assert abs(abs(float(levels[0])) - 1.0) == pytest.approx(1 / 255, rel=1e-3)


def test_fake_quant_matches_integer_reference_grid():
  g = torch.Generator().manual_seed(0)
  x = torch.randn(1000, generator=g) * 3
  lo, hi = -2.5, 5.0
  y = quant_utils.FakeQuantWithMinMax(x, lo, hi, num_bits=4)
  scale = (hi - lo) / 15
  zp = round(-lo / scale)
  ref = (np.clip(np.floor(x.numpy() / scale + 0.5) + zp, 0, 15) - zp) * scale
  np.testing.assert_allclose(y.numpy(), ref, atol=1e-5)


def test_fake_quant_straight_through_gradient_masks_outside_range():
  x = torch.tensor([-3.0, -0.5, 0.2, 0.9, 4.0], requires_grad=True)
  quant_utils.FakeQuantWithMinMax(x, -1.0, 1.0).sum().backward()
  np.testing.assert_array_equal(x.grad.numpy(), [0, 1, 1, 1, 0])


def test_fake_quant_tensor_ranges_and_degenerate_range():
  x = torch.tensor([0.3, -0.7])
  y = quant_utils.FakeQuantWithMinMax(x, torch.tensor(-1.0), torch.tensor(1.0), num_bits=8)
  assert torch.allclose(y, x, atol=2 / 255)
  z = quant_utils.FakeQuantWithMinMax(x, 0.0, 0.0)
  assert torch.all(z == 0)


# ------------------------------------------------------------------------- schedules --
def test_linear_clipping_cap_schedule():
  p = quant_utils.LinearClippingCapSchedule.Params().Set(
      start_step=100, end_step=200, start_cap=6.0, end_cap=1.0)
  s = p.Instantiate()
  assert s.Value(0) == pytest.approx(6.0)
  assert s.Value(100) == pytest.approx(6.0)
  assert s.Value(150) == pytest.approx(3.5)
  assert s.Value(200) == pytest.approx(1.0)
  assert s.Value(10_000) == pytest.approx(1.0)
  assert s.GetEndRange() == (-1.0, 1.0)
  assert not s.is_quantized
  py_utils.SetGlobalStep(150)
  x = torch.tensor([-10.0, -1.0, 2.0, 5.0])
  np.testing.assert_allclose(s.ApplyClipping(s.theta, x).numpy(), [-3.5, -1.0, 2.0, 3.5])
  np.testing.assert_allclose(s.ApplyConstantClip(x, 0.0, 1.0).numpy(), [0, 0, 1, 1])


def test_identity_clipping_cap_schedule():
  s = quant_utils.IdentityClippingCapSchedule.Params().Set(name='id').Instantiate()
  x = torch.randn(5) * 100
  assert torch.equal(s.ApplyClipping(s.theta, x), x)
  assert torch.equal(s.ApplyConstantClip(x, 0, 1), x)
  lo, hi = s.GetEndRange()
  assert lo < -1e30 and hi > 1e30


def test_fake_quantization_schedule_phases():
  p = quant_utils.FakeQuantizationSchedule.Params().Set(
      clip_start_step=10, clip_end_step=20, quant_start_step=30, start_cap=4.0, end_cap=1.0)
  s = p.Instantiate()
  assert s.is_quantized and s.bits == 8
  assert s.GetEndRange() == (-1.0, 1.0)
  assert s.GetQuantizedEndRange() == (-1.0, 127 / 128)
  assert s.GetQuantizedEndRange(end_cap=2.0, bits=16) == (-2.0, 2.0 * 32767 / 32768)
  x = torch.tensor([-9.0, -0.3337, 0.25, 0.61803, 9.0])

  py_utils.SetGlobalStep(5)            # before the ramp: identity
  assert torch.equal(s.ApplyClipping(s.theta, x), x)
  py_utils.SetGlobalStep(15)           # half-way: cap 2.5, clip only
  y = s.ApplyClipping(s.theta, x)
  np.testing.assert_allclose(y.numpy(), [-2.5, -0.3337, 0.25, 0.61803, 2.5 * 127 / 128],
                             rtol=1e-6)
  py_utils.SetGlobalStep(25)           # ramp done, still clip only
  y = s.ApplyClipping(s.theta, x)
  np.testing.assert_allclose(y.numpy(), [-1.0, -0.3337, 0.25, 0.61803, 127 / 128], rtol=1e-6)
  py_utils.SetGlobalStep(30)           # quantizing: multiples of 1/128
  y = s.ApplyClipping(s.theta, x)
  np.testing.assert_allclose(y.numpy() * 128, np.round(y.numpy() * 128), atol=1e-4)
  np.testing.assert_allclose(y.numpy(), [-1.0, -43 / 128, 0.25, 79 / 128, 127 / 128], atol=1e-6)
  # per-call overrides of caps/bits
  y16 = s.ApplyClipping(s.theta, x, end_cap=2.0, bits=16)
  assert float(y16[0]) == pytest.approx(-2.0)
  assert abs(float(y16[3]) - 0.61803) < 2.0 / 32768
  assert s.Value(15) == (pytest.approx(2.5), False)
  assert s.Value(31) == (pytest.approx(1.0), True)


def test_fake_quantization_schedule_requires_quant_after_clip():
  p = quant_utils.FakeQuantizationSchedule.Params().Set(clip_end_step=10, quant_start_step=5)
  with pytest.raises(AssertionError):
    p.Instantiate()


def test_fake_quantization_schedule_inference_is_end_state():
  p = quant_utils.FakeQuantizationSchedule.Params().Set(
      clip_start_step=10, clip_end_step=20, quant_start_step=30, is_inference=True)
  s = p.Instantiate()
  y = s.ApplyClipping(s.theta, torch.tensor([-3.0, 0.3, 3.0]))
  np.testing.assert_allclose(y.numpy(), [-1.0, 38 / 128, 127 / 128], atol=1e-6)


# --------------------------------------------------------------------- QuantizableLayer --
def test_unquantized_layer_hooks_are_identities():
  layer = _Layer()
  x = torch.randn(3, 4)
  assert layer.QWeight(x) is x
  assert layer.QAct('inputs', x) is x
  assert layer.QRAct(x, QDistribution.TANH) is x
  assert torch.equal(layer.fns.qtanh(x), torch.tanh(x))
  assert torch.equal(layer.fns.qadd(x, x, qout_name='inputs'), x + x)
  assert torch.equal(layer.QMatmul(x, x.t()), x @ x.t())
  assert torch.allclose(layer.QEinsum('ab,cb->ac', x, x), x @ x.t(), atol=1e-5)
  a, b = layer.ToAqtActActInputs(x, x)
  assert a is x and b is x and layer.FromAqtActActMatmul(x) is x


def test_tracking_validation():
  layer = _Layer()
  with pytest.raises(ValueError, match='already tracked'):
    layer.TrackQActs('inputs')
  with pytest.raises(ValueError, match='must first be tracked'):
    layer.QAct('nope', torch.zeros(1))
  with pytest.raises(ValueError, match='requires qout_name'):
    layer.fns.qadd(torch.zeros(1), torch.zeros(1))
  with pytest.raises(ValueError, match='to be None or one of'):
    layer.QMatmul(torch.zeros(1, 1), torch.zeros(1, 1), lhs_name='untracked')
  with pytest.raises(ValueError, match='TrackQWeight'):
    layer.ToAqtWeight('w', torch.zeros(2, 2), feature_axis=-1)
  layer.TrackQWeight('w', [2, 3], feature_axis=-1)
  w = torch.randn(2, 3)
  assert layer.ToAqtWeight('w', w, feature_axis=-1) is w
  assert layer.FromAqtWeight('w', w) is w
  a, ww = layer.ToAqtInputs('w', torch.zeros(1, 2), w, w_feature_axis=-1)
  assert ww is w and layer.FromAqtMatmul('w', a) is a
  with pytest.raises(ValueError, match='already tracked'):
    layer.TrackQWeight('w', [2, 3], feature_axis=-1)


def test_qdomain_type_is_checked():
  p = quant_test_lib.SampleQuantizedProjectionLayer.Params().Set(name='proj')
  p.qdomain.default = quant_utils.LinearClippingCapSchedule.Params()
  with pytest.raises(TypeError):
    p.Instantiate()


def test_named_domain_falls_back_to_default():
  p = quant_test_lib.SampleQuantizedProjectionLayer.Params().Set(name='proj')
  p.qdomain.Define('softmax', None, '')
  p.qdomain.default = quant_utils.PassiveAsymQDomain.Params()
  layer = p.Instantiate()
  assert layer._GetQDomain('softmax') is layer.qdomain_default
  assert layer.GetQDomainParams('softmax') is layer.params.qdomain.default
  p2 = p.Copy()
  p2.qdomain.softmax = quant_utils.PassiveAsymQDomain.Params().Set(bits=16)
  layer2 = p2.Instantiate()
  assert layer2._GetQDomain('softmax').bits == 16
  assert layer2._GetQDomain('default').bits == 8


def test_qconv_matches_torch_conv():
  layer = _Layer()
  x = torch.randn(2, 9, 3)
  w = torch.randn(4, 3, 5)
  y = layer.QConv1D(x, w, 2, 'SAME')
  assert y.shape == (2, 5, 5)
  ref = F.conv1d(F.pad(x.transpose(1, 2), (1, 2)), w.permute(2, 1, 0), stride=2).transpose(1, 2)
  np.testing.assert_allclose(y.numpy(), ref.numpy(), atol=1e-5)
  x2 = torch.randn(2, 6, 6, 3)
  w2 = torch.randn(3, 3, 3, 4)
  y2 = layer.QConv2D(x2, w2, (1, 1), 'SAME')
  ref2 = F.conv2d(x2.permute(0, 3, 1, 2), w2.permute(3, 2, 0, 1), padding=1).permute(0, 2, 3, 1)
  np.testing.assert_allclose(y2.numpy(), ref2.numpy(), atol=1e-5)
  wd = torch.randn(3, 3, 3, 2)
  yd = layer.QConv2D(x2, wd, (1, 1), 'VALID', is_depthwise=True)
  assert yd.shape == (2, 4, 4, 6)
  # channel c*mult+m of the output only sees input channel c
  ref_c1m0 = F.conv2d(x2[..., 1][:, None], wd[:, :, 1, 0][None, None])[:, 0]
  np.testing.assert_allclose(yd[..., 2].numpy(), ref_c1m0.numpy(), atol=1e-5)


# ----------------------------------------------------------- SymmetricScheduledClip --
def _FqDomain(**kw):
  return quant_utils.SymmetricScheduledClipQDomain.Params().Set(
      cc_schedule=quant_utils.FakeQuantizationSchedule.Params().Set(
          clip_start_step=0, clip_end_step=10, quant_start_step=10, start_cap=8.0, end_cap=1.0,
          **kw))


def test_symmetric_scheduled_clip_domain():
  layer = _Layer(_FqDomain())
  qd = layer.qdomain_default
  assert qd.bits == 8
  x = torch.tensor([-20.0, -0.4, 0.7, 20.0])
  py_utils.SetGlobalStep(0)                           # cap 8, clip only
  np.testing.assert_allclose(layer.QWeight(x).numpy(), [-8, -0.4, 0.7, 8 * 127 / 128], rtol=1e-6)
  py_utils.SetGlobalStep(10)                          # cap 1, quantized
  y = layer.QAct('inputs', x)
  np.testing.assert_allclose(y.numpy(), [-1, -51 / 128, 90 / 128, 127 / 128], atol=1e-6)
  # eval_only acts are untouched while training, quantized in eval
  assert layer.QAct('inputs', x, eval_only=True) is x
  with cluster_factory.SetEval(True):
    elayer = _Layer(_FqDomain())
    np.testing.assert_allclose(elayer.QAct('inputs', x, eval_only=True).numpy(), y.numpy())
  # natural ranges go through the schedule; constant ranges are only clipped
  np.testing.assert_allclose(layer.QRAct(torch.tensor([0.3, 5.0]), QDistribution.TANH).numpy(),
                             [38 / 128, 127 / 128], atol=1e-6)
  np.testing.assert_allclose(layer.QRAct(torch.tensor([0.3, 5.0]), QDistribution.PADDING).numpy(),
                             [0.3, 1.0])
  with pytest.raises(ValueError, match='log_softmax_range'):
    layer.fns.qlogsoftmax(torch.zeros(2, 3))


def test_symmetric_domain_changes_layer_output_and_has_gradients():
  plain = _Layer()
  quant = _Layer(_FqDomain())
  py_utils.SetGlobalStep(20)
  x = torch.from_numpy(quant_test_lib.QuantUtilsBaseTest.INPUTS)
  pad = torch.from_numpy(quant_test_lib.QuantUtilsBaseTest.PADDINGS)
  a = plain.FProp(plain.theta, x, pad)
  b = quant.FProp(quant.theta, x, pad)
  assert not torch.allclose(a, b)
  assert float((b * 128 - torch.round(b * 128)).abs().max()) < 1e-3
  b.sum().backward()
  assert quant.vars.w.grad is not None and float(quant.vars.w.grad.abs().sum()) > 0


# ------------------------------------------------------------------- PassiveAsym --
def test_passive_asym_tracks_ranges_and_uses_them_in_eval():
  qp = quant_utils.PassiveAsymQDomain.Params().Set(ema_decay=0.5)
  layer = _Layer(qp)
  qd = layer.qdomain_default
  assert sorted(qd.act_names) == ['inputs', 'transformed']
  assert float(qd.vars.inputs_min) == -1.0 and float(qd.vars.inputs_max) == 1.0
  assert not qd.vars.inputs_min.requires_grad
  x = torch.tensor([-4.0, 0.5, 2.0])
  y = layer.QAct('inputs', x)                          # batch range [-4, 2]
  assert torch.allclose(y, x, atol=6 / 255)
  assert not torch.equal(y, x)
  layer.QAct('inputs', torch.tensor([0.1, 6.0]))       # widens the accumulated max to 6
  acc = qd.accumulators.qact_inputs.GetValue()
  np.testing.assert_allclose(acc.numpy(), [2, -4, 6])
  layer.PostTrainingStepUpdate()
  # EMA toward (-4, 6) with decay .5 from the defaults (-1, 1)
  assert float(qd.vars.inputs_min) == pytest.approx(-2.5)
  assert float(qd.vars.inputs_max) == pytest.approx(3.5)
  np.testing.assert_allclose(qd.accumulators.qact_inputs.GetValue().numpy(), [0, 0, 0])
  layer.PostTrainingStepUpdate()                       # nothing seen → unchanged
  assert float(qd.vars.inputs_min) == pytest.approx(-2.5)
  # untouched activation keeps its defaults
  assert float(qd.vars.transformed_max) == 1.0

  with cluster_factory.SetEval(True):
    ep = quant_test_lib.SampleQuantizedProjectionLayer.Params().Set(name='proj')
    ep.qdomain.default = qp
    elayer = ep.Instantiate()
    elayer.InstantiateVariables()
    eqd = elayer.qdomain_default
    eqd.vars.inputs_min.data.fill_(-2.5)
    eqd.vars.inputs_max.data.fill_(3.5)
    out = elayer.QAct('inputs', torch.tensor([-10.0, 0.0, 1.234, 10.0]))
    scale = 6.0 / 255
    zp = round(2.5 / scale)
    np.testing.assert_allclose(out.numpy(), [-zp * scale, 0.0, round(1.234 / scale) * scale,
                                             (255 - zp) * scale], atol=1e-5)
    # eval never records
    np.testing.assert_allclose(eqd.accumulators.qact_inputs.GetValue().numpy(), [0, 0, 0])


def test_passive_asym_weight_epsilon_delay_and_freeze():
  layer = _Layer(quant_utils.PassiveAsymQDomain.Params().Set(quantize_weight_epsilon=0.5))
  w = torch.tensor([0.01, 0.02])
  qw = layer.QWeight(w)                                # range forced to [-.5, .5]
  np.testing.assert_allclose(qw.numpy(), np.round(w.numpy() * 255) / 255, atol=1e-6)

  delayed = _Layer(quant_utils.PassiveAsymQDomain.Params().Set(delay_start_steps=100))
  x = torch.tensor([-1.0, 0.123456, 1.0])
  py_utils.SetGlobalStep(10)
  assert torch.equal(delayed.QAct('inputs', x), x)
  py_utils.SetGlobalStep(100)
  assert not torch.equal(delayed.QAct('inputs', x), x)
  never = _Layer(quant_utils.PassiveAsymQDomain.Params().Set(delay_start_steps=-1))
  py_utils.SetGlobalStep(10**6)
  assert torch.equal(never.QAct('inputs', x), x)

  frozen = _Layer(quant_utils.PassiveAsymQDomain.Params().Set(freeze=True))
  frozen.QAct('inputs', torch.tensor([-50.0, 50.0]))
  frozen.PostTrainingStepUpdate()
  assert float(frozen.qdomain_default.vars.inputs_max) == 1.0


def test_passive_asym_natural_ranges():
  layer = _Layer(quant_utils.PassiveAsymQDomain.Params().Set(log_softmax_range=(-16.0, 0.0)))
  t = layer.fns.qtanh(torch.tensor([10.0, 0.0, -10.0]))
  # tanh range narrowed to [-1, 1 - 2/256]: max representable is 127/128
  np.testing.assert_allclose(t.numpy(), [127 / 128, 0.0, -1.0], atol=1e-6)
  s = layer.fns.qsoftmax(torch.tensor([[0.0, 100.0]]))
  np.testing.assert_allclose(s.numpy(), [[0.0, 255 / 256]], atol=1e-6)
  r6 = layer.fns.qrelu6(torch.tensor([7.0, 3.0]))
  np.testing.assert_allclose(r6.numpy(), [6.0, round(3.0 / (6 / 255)) * 6 / 255], atol=1e-5)
  ls = layer.fns.qlogsoftmax(torch.tensor([[0.0, -30.0]]))
  assert float(ls[0, 1]) == pytest.approx(-16.0, abs=1e-4)
  loose = _Layer(quant_utils.PassiveAsymQDomain.Params().Set(narrow_to_asym_bit_depth=False))
  top = float(loose.fns.qtanh(torch.tensor([10.0])))
  assert min(abs(top - 127 * 2 / 255), abs(top - 128 * 2 / 255)) < 1e-6


def test_qdistribution_ispositive():
  assert QDistribution.IsPositive(QDistribution.RELU)
  assert not QDistribution.IsPositive(QDistribution.SYMMETRIC)
  assert QDistribution('softmax') is QDistribution.SOFTMAX


def test_quantized_layer_trains_through_a_base_task_step():
  """PostTrainingStepUpdate reaches the domain through the layer tree."""
  layer = _Layer(quant_utils.PassiveAsymQDomain.Params().Set(ema_decay=0.0))
  x = torch.from_numpy(quant_test_lib.QuantUtilsBaseTest.INPUTS)
  pad = torch.from_numpy(quant_test_lib.QuantUtilsBaseTest.PADDINGS)
  layer.FProp(layer.theta, x, pad).sum().backward()
  layer.PostTrainingStepUpdate()
  assert float(layer.qdomain_default.vars.inputs_min) == pytest.approx(-3.0)
  assert float(layer.qdomain_default.vars.inputs_max) == pytest.approx(2.0)
