"""The test-support libraries themselves (quant / stream-step / trainer fixtures)."""

import unittest

import torch

from lingvo_b200 import model_registry
from lingvo_b200 import models_test_helper
from lingvo_b200.core import cluster_factory
from lingvo_b200.core import conv_layers_with_time_padding as conv_tp
from lingvo_b200.core import quant_test_lib
from lingvo_b200.core import quant_utils
from lingvo_b200.core import stream_step_test_base
from lingvo_b200.core import test_trainer_utils
from lingvo_b200.core import trainer_test_utils


class QuantLibTest(quant_test_lib.QuantUtilsBaseTest):

  def testPlainAndQuantizedDiffer(self):
    p = quant_test_lib.SampleQuantizedProjectionLayer.Params().Set(name='proj')
    plain = self._testLayerHelper('plain', p)
    q = p.Copy()
    q.qdomain.default = quant_utils.SymmetricScheduledClipQDomain.Params().Set(
        cc_schedule=quant_utils.FakeQuantizationSchedule.Params().Set(
            clip_start_step=0, clip_end_step=1, quant_start_step=1, start_cap=1.0, end_cap=1.0))
    with cluster_factory.SetEval(True):
      self._testLayerHelper('quant', q, not_expected=plain, global_step=10)


class CausalConvStreamTest(stream_step_test_base.StreamStepTestBase):

  def _GetParams(self, input_dim=8, stride=1, right_context=0, kernel=3, **kwargs):
    del stride, right_context, kwargs
    return conv_tp.CausalDepthwiseConv2DLayer.Params().Set(
        name='conv', filter_shape=[kernel, 1, input_dim, 1], filter_stride=[1, 1])

  @property
  def input_rank(self):
    return 4

  def _GetInputs(self, batch_size, max_seqlen, input_dim, full_seq=False):
    x, pad = super()._GetInputs(batch_size, max_seqlen, input_dim, full_seq)
    return x.reshape(batch_size, max_seqlen, 1, input_dim), pad

  def testStreamEqualsFProp(self):
    self._TestStreamStepHelper(batch_size=2, max_seqlen=12, input_dim=4, stride=2)


def test_identity_regression_model_trains():
  cls = trainer_test_utils.RegisterIdentityRegressionModel(
      name='IdentityRegressionForTest', learning_rate=0.05, max_train_steps=4)
  mp = cls().Model()
  mp.input = cls().Train()
  with cluster_factory.ForTestingWorker(mode='sync', job='trainer_client'):
    model = mp.Instantiate()
  task = model.tasks[0]
  m0, b0 = float(task.vars.m), float(task.vars.b)
  losses = [float(task.TrainStep()[0]['loss'][0]) for _ in range(4)]
  # batch 0 is all zeros: loss = b0², gradient only on b
  assert abs(losses[0] - b0 ** 2) < 1e-6
  assert float(task.vars.b) != b0 and float(task.vars.m) != m0
  gen = cls().Train().Instantiate()
  assert float(gen.GetPreprocessedInputBatch().value[0, 0, 0]) == 0.0
  assert float(gen.GetPreprocessedInputBatch().value[0, 0, 0]) == 1.0


def test_model_validator_and_models_helper():
  import lingvo_b200.models.lm.params.params  # noqa: F401
  import lingvo_b200.models.image.params.mnist  # noqa: F401
  case = test_trainer_utils.MakeModelValidatorTestCase(
      ['lm.synthetic_packed_input.DenseLm8B2x2', 'image.mnist.LeNet5'])
  res = unittest.TextTestRunner(verbosity=0).run(
      unittest.defaultTestLoader.loadTestsFromTestCase(case))
  assert res.wasSuccessful(), res.failures + res.errors

  class _Models(models_test_helper.BaseModelsTest):
    pass
  _Models.CreateTestMethodsForAllRegisteredModels(
      model_registry, task_regexes=[r'^lm\.synthetic_packed_input\.(DenseLmTiny|MoELm8ETiny)$'])
  names = [n for n in dir(_Models) if n.startswith('testModelParams_')]
  assert len(names) == 2
  res = unittest.TextTestRunner(verbosity=0).run(
      unittest.defaultTestLoader.loadTestsFromTestCase(_Models))
  assert res.wasSuccessful(), res.failures + res.errors


def test_profiling_roofline_and_nvtx_wrappers():
  import torch
  from lingvo_b200.utils import profiling
  r = profiling.Roofline({'hbm_gbs': 6000.0, 'bf16_tflops': 1400.0})
  g = r.Gemm(8192, 2048, 2048)
  assert abs(g['flops'] - 2 * 8192 * 2048 * 2048) < 1
  rep = r.Report('gemm', 60.0, **g)
  assert rep['limiter'] == 'compute' and 0.7 < rep['fraction_of_roofline'] < 0.9
  rep = r.Report('adafactor', 400.0, **r.AdafactorFactored(8 * 2048 * 8192))
  assert rep['limiter'] == 'memory' and rep['achieved_gbs'] > 5000
  assert r.BoundUs(**r.NormBwd(8192, 2048)) > r.BoundUs(**r.NormFwd(8192, 2048))
  peaks = profiling.MeasuredPeaks('/nonexistent.json')
  assert 'fallback' in peaks['source']
  # NVTX instrumentation is a transparent wrapper (no CUDA needed when disabled)
  from lingvo_b200.core import layers
  p = layers.ProjectionLayer.Params().Set(name='proj', input_dim=4, output_dim=3)
  layer = p.Instantiate()
  x = torch.randn(2, 4)
  want = layer.FPropDefaultTheta(x)
  assert profiling.InstrumentLayers(layer) >= 1
  assert profiling.InstrumentLayers(layer) == 0          # idempotent
  torch.testing.assert_close(layer.FPropDefaultTheta(x), want)
  with profiling.Range('noop'):
    pass


def test_test_utils_numeric_gradient_tape_and_summary_reader(tmp_path):
  import numpy as np
  import torch
  from lingvo_b200.core import test_utils
  from lingvo_b200.utils import tfevents
  x = torch.tensor([[1.0, 2.0], [3.0, 4.0]], dtype=torch.float64)
  g = test_utils.ComputeNumericGradientEager(lambda v: (v ** 2).sum(), x, step=2)
  np.testing.assert_allclose(g.numpy(), [[2.0, 0.0], [6.0, 0.0]], atol=1e-6)
  assert test_utils.PickEveryN(np.arange(6).reshape(2, 3), 2).tolist() == [0, 2, 4]
  with test_utils.TapeIfEager() as tape:
    w = torch.tensor([1.0, 2.0])
    tape.watch(w)
    y = (w ** 3).sum()
    dw = tape.gradient(y, w)
  torch.testing.assert_close(dw, torch.tensor([3.0, 12.0]))
  got = test_utils.DefineAndTrace(((2, 3), torch.float32), torch.ones(3))(
      lambda a, b: (a + b).shape)
  assert tuple(got) == (2, 3)
  wr = tfevents.EventFileWriter(str(tmp_path))
  wr.add_scalar('loss', 1.5, 1)
  wr.add_scalar('loss', 0.5, 2)
  wr.add_scalar('acc', 0.9, 2)
  wr.close()
  tc = test_utils.TestCase()
  vals = tc.GetScalarSummaryValues(str(tmp_path), ['loss'])
  assert vals == {'loss': {1: 1.5, 2: 0.5}}

  class _T(test_utils.TestCase):
    @test_utils.SkipIfEager
    def testGraphOnly(self):
      raise AssertionError('must be skipped')
  import unittest
  res = unittest.TestResult()
  _T('testGraphOnly').run(res)
  assert len(res.skipped) == 1 and not res.failures and not res.errors
  line = '    test_utils.CompareToGoldenSingleFloat(self, 0.25, v)\n'
  assert '0.500000' in test_utils.ReplaceGoldenSingleFloat(line, 0.5)
