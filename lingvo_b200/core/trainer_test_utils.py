"""Tiny deterministic model for trainer / executor tests (ref
`lingvo/core/trainer_test_utils.py`).

`CountingInputGenerator` emits batch i with every element = i; `IdentityRegressionTask`
learns `y = m·x + b` towards the identity with plain SGD, so every step's loss and the
variable trajectory can be predicted in closed form. `RegisterIdentityRegressionModel`
registers a one-task model around them.
"""

from __future__ import annotations

import torch

from lingvo_b200 import model_registry
from lingvo_b200.core import base_input_generator
from lingvo_b200.core import base_model
from lingvo_b200.core import base_model_params
from lingvo_b200.core import metrics as metrics_lib
from lingvo_b200.core import optimizer
from lingvo_b200.core import program
from lingvo_b200.core import py_utils
from lingvo_b200.core.nested_map import NestedMap


class CountingInputGenerator(base_input_generator.BaseInputGenerator):
  """Batch n holds `value = n` repeated `batch_size` times (ref :30)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('shape', [2, 2], 'Shape of one example.')
    p.batch_size = 2
    return p

  def __init__(self, params):
    super().__init__(params)
    self._count = 0

  def _InputBatch(self):
    p = self.params
    v = float(self._count)
    self._count += 1
    return NestedMap(value=torch.full([p.batch_size] + list(p.shape), v),
                     counter=torch.tensor(int(v)))

  def Reset(self, sess=None):
    self._count = 0


class IdentityRegressionTask(base_model.BaseTask):
  """ref :70."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('weight_init_value', 0.8, 'Initial m.')
    p.Define('bias_init_value', 0.4, 'Initial b.')
    p.name = 'identity_regression_task'
    p.train.optimizer = optimizer.SGD.Params()
    p.train.learning_rate = 0.01
    p.train.max_steps = 10
    return p

  def __init__(self, params):
    super().__init__(params)
    self.global_steps, self.metrics, self.result_per_example_tensors = [], [], []

  def _CreateLayerVariables(self):
    super()._CreateLayerVariables()
    p = self.params
    self.CreateVariable('m', py_utils.WeightParams(
        [], py_utils.WeightInit.Constant(p.weight_init_value), torch.float32))
    self.CreateVariable('b', py_utils.WeightParams(
        [], py_utils.WeightInit.Constant(p.bias_init_value), torch.float32))

  def ComputePredictions(self, theta, input_batch):
    return theta.m * input_batch.value + theta.b

  def ComputeLoss(self, theta, predicted, input_batch):
    diff = predicted - input_batch.value
    loss = diff.square().mean()
    n = float(input_batch.value.shape[0])
    rep = lambda x: x.detach().reshape(1).expand(int(n))
    metrics = {'loss': (loss, n), 'm': (theta.m.detach(), n), 'b': (theta.b.detach(), n),
               'num_samples_in_batch': (torch.tensor(n), 1.0)}
    per_example = {'input': input_batch.value, 'loss': diff.square().flatten(1).mean(1),
                   'diff': diff, 'm': rep(theta.m), 'b': rep(theta.b)}
    return metrics, per_example

  def FilterPerExampleTensors(self, per_example):
    return per_example

  def ProcessFPropResults(self, sess, global_step, metrics, per_example):
    self.global_steps.append(global_step)
    self.metrics.append(metrics)
    self.result_per_example_tensors.append(per_example)

  def CreateDecoderMetrics(self):
    return {'num_samples_in_batch': metrics_lib.AverageMetric(),
            'diff': metrics_lib.AverageMetric()}

  def DecodeWithTheta(self, theta, input_batch):
    pred = self.ComputePredictions(theta, input_batch)
    return NestedMap(diff=(pred - input_batch.value).abs().flatten(1).mean(1))

  def Decode(self, input_batch):
    return self.DecodeWithTheta(self.theta, input_batch)

  def PostProcessDecodeOut(self, dec_out_dict, dec_metrics_dict):
    d = dec_out_dict['diff']
    dec_metrics_dict['num_samples_in_batch'].Update(len(d))
    for v in (d.tolist() if hasattr(d, 'tolist') else list(d)):
      dec_metrics_dict['diff'].Update(float(v))
    return []


class ModelTrackingFPropResults(base_model.SingleTaskModel):
  """Records what the runner passes to `ProcessFPropResults` (ref :147)."""

  def __init__(self, params, **kwargs):
    super().__init__(params, **kwargs)
    self.global_steps, self.metrics, self.result_per_example_tensors = [], [], []

  def ProcessFPropResults(self, sess, global_step, metrics, per_example):
    self.global_steps.append(global_step)
    self.metrics.append(metrics)
    self.result_per_example_tensors.append(per_example)


def RegisterIdentityRegressionModel(name='IdentityRegressionModel',  # pylint: disable=invalid-name
                                    weight_init_value=0.8, bias_init_value=0.4,
                                    learning_rate=0.01, max_train_steps=10,
                                    train_batch_size=2, eval_batch_size=2, train_steps_per_loop=2,
                                    eval_decode_steps_per_loop=2, eval_decode_samples=10):
  """Registers `test.<name>` and returns the params class (ref :163)."""

  class _Model(base_model_params.SingleTaskModelParams):

    def Train(self):
      return CountingInputGenerator.Params().Set(batch_size=train_batch_size)

    def Test(self):
      return CountingInputGenerator.Params().Set(batch_size=eval_batch_size,
                                                 num_samples=eval_decode_samples)

    def Task(self):
      p = IdentityRegressionTask.Params().Set(weight_init_value=weight_init_value,
                                              bias_init_value=bias_init_value)
      p.train.Set(learning_rate=learning_rate, max_steps=max_train_steps)
      p.eval.samples_per_summary = eval_decode_samples
      return p

    def Model(self):
      return ModelTrackingFPropResults.Params(self.Task())

    def ProgramSchedule(self):
      return program.SimpleProgramScheduleForTask(
          train_dataset_name='Train', train_steps_per_loop=train_steps_per_loop,
          eval_dataset_names=['Test'], eval_steps_per_loop=eval_decode_steps_per_loop,
          decode_steps_per_loop=eval_decode_steps_per_loop)

  _Model.__name__ = name
  _Model.__module__ = 'lingvo_b200.models.test.params.test'
  model_registry.RegisterSingleTaskModel(_Model)
  return _Model
