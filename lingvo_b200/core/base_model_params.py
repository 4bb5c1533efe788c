"""Experiment-params base classes (reference `core/base_model_params.py:23-148`)."""

from lingvo_b200.core import base_model
from lingvo_b200.core import hyperparams


class DatasetError(Exception):
  pass


class GetAllDatasetParamsNotImplementedError(NotImplementedError):
  pass


class _BaseModelParams:
  """Base class for storing model Params for a single experiment."""

  def GetAllDatasetParams(self):
    """Optionally returns {dataset_name: Params} for all datasets."""
    raise GetAllDatasetParamsNotImplementedError()

  def GetDatasetParams(self, dataset):
    """Input generator params for `dataset` ('Train', 'Dev', 'Test', …)."""
    try:
      all_datasets = self.GetAllDatasetParams()
      if dataset not in all_datasets:
        raise DatasetError(f'Dataset {dataset} not found')
      return all_datasets[dataset]
    except GetAllDatasetParamsNotImplementedError:
      pass
    try:
      f = getattr(self, dataset)
    except AttributeError as e:
      raise DatasetError(str(e)) from e
    return f()

  def Train(self):
    raise NotImplementedError()

  def Dev(self):
    raise NotImplementedError()

  def Test(self):
    raise NotImplementedError()

  def Search(self):
    """Model search params (hyper-parameter tuning)."""
    return None


class SingleTaskModelParams(_BaseModelParams):
  """Model Params for a `.SingleTaskModel`."""

  def Task(self):
    """Returns task params."""
    raise NotImplementedError('Abstract method')

  def Model(self):
    """Wraps Task() params into SingleTaskModel params."""
    return base_model.SingleTaskModel.Params(self.Task())

  def ProgramSchedule(self):
    """Returns a schedule for the Executor."""
    raise NotImplementedError('Abstract method')


class MultiTaskModelParams(_BaseModelParams):
  """Model Params for a `.MultiTaskModel`."""

  def Model(self):
    raise NotImplementedError('Abstract method')

  def ProgramSchedule(self):
    raise NotImplementedError('Abstract method')
