"""Builder-style base class for Milan experiments (ref
`lingvo/tasks/milan/params/dual_encoder_recipe.py`)."""

import functools

from lingvo_b200.core import base_model_params
from lingvo_b200.core import layers as lingvo_layers
from lingvo_b200.core import optimizer
from lingvo_b200.core import schedule
from lingvo_b200.models.milan import constants
from lingvo_b200.models.milan import dataset_spec
from lingvo_b200.models.milan import dual_encoder
from lingvo_b200.models.milan import input_generator


class RecipeError(Exception):
  pass


class DualEncoderRecipe(base_model_params.SingleTaskModelParams):
  """Subclasses call `AddModality` / `AddPreprocessor` in `__init__` and provide
  `default_dataset` (ref :28)."""

  def __init__(self):
    self.dataset = self._ChooseDatasetSpec()
    self.input_params = input_generator.MilanInputGenerator.Params().Set(
        batch_size=64, use_per_host_infeed=True)
    self.task_params = dual_encoder.MilanTask.Params()
    self.task_params.train.Set(
        clip_gradient_norm_to_value=1.0,
        grad_norm_tracker=lingvo_layers.GradNormTracker.Params().Set(
            name='grad_norm_tracker', grad_norm_clip_cap_min=0.1),
        save_max_to_keep=2000, save_keep_checkpoint_every_n_hours=0.1667,
        optimizer=optimizer.Adam.Params().Set(beta1=0.9, beta2=0.999, epsilon=1e-8),
        learning_rate=0.0001,
        lr_schedule=schedule.StepwiseExponentialSchedule.Params().Set(
            decay=0.999, num_steps_per_decay=1000),
        max_steps=40000)

  def _ChooseDatasetSpec(self):
    return self.default_dataset

  @property
  def default_dataset(self) -> dataset_spec.DatasetSpec:
    raise NotImplementedError()

  @property
  def encoder_configs(self):
    return self.task_params.dual_encoder.encoder_configs

  def AddModality(self, name, **kwargs):
    config = dual_encoder.EncoderConfig().Set(**kwargs)
    self.encoder_configs[name] = config
    return config

  def AddPreprocessor(self, input_feature, preprocessor):
    self.input_params.preprocessors[input_feature] = preprocessor.Copy()

  def StartFromCheckpoint(self, checkpoint_path):
    """Fine-tune from a checkpoint: everything but the global step and the grad-norm
    tracker statistics is restored (ref :115)."""
    self.task_params.train.init_from_checkpoint_rules = {
        checkpoint_path: ([('(.*)', '%s')], ['.*grad_norm_tracker/.*', 'global_step'])}

  def GetAllDatasetParams(self):
    def _Split(name, split, shuffle):
      return self.input_params.Copy().Set(
          name=name, dataset_fn=functools.partial(self.dataset.Read, split=split,
                                                  shuffle=shuffle))
    return {'Train': _Split('Train', constants.Split.TRAIN, True),
            'Dev': _Split('Dev', constants.Split.DEV, False),
            'Test': _Split('Test', constants.Split.TEST, False)}

  def Train(self):
    return self.GetAllDatasetParams()['Train']

  def Dev(self):
    return self.GetAllDatasetParams()['Dev']

  def Test(self):
    return self.GetAllDatasetParams()['Test']

  def Task(self):
    task_params = self.task_params.Copy()
    if not task_params.dual_encoder.encoder_configs:
      raise RecipeError('Must configure at least one encoder.')
    assert task_params.dual_encoder.label_fn is None
    task_params.dual_encoder.label_fn = self.dataset.Label
    return task_params
