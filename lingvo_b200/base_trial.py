"""Hyper-parameter tuner hook interface (reference `lingvo/base_trial.py:22-136`)."""


class Trial:
  """Base class for tuner trials."""

  @classmethod
  def CreateForTest(cls):
    return cls()

  def Name(self):
    raise NotImplementedError('Abstract method')

  def OverrideModelParams(self, model_params):
    raise NotImplementedError('Abstract method')

  def ShouldStop(self):
    raise NotImplementedError('Abstract method')

  def ReportDone(self, infeasible=False, infeasible_reason=''):
    raise NotImplementedError('Abstract method')

  def ShouldStopAndMaybeReport(self, global_step, metrics_dict):
    raise NotImplementedError('Abstract method')

  def ReportEvalMeasure(self, global_step, metrics_dict, checkpoint_path):
    raise NotImplementedError('Abstract method')


class NoOpTrial(Trial):
  """A Trial implementation that does nothing."""

  def Name(self):
    return ''

  def OverrideModelParams(self, model_params):
    return model_params

  def ShouldStop(self):
    return False

  def ReportDone(self, infeasible=False, infeasible_reason=''):
    return False

  def ShouldStopAndMaybeReport(self, global_step, metrics_dict):
    del global_step, metrics_dict
    return False

  def ReportEvalMeasure(self, global_step, metrics_dict, checkpoint_path):
    del global_step, metrics_dict, checkpoint_path
    return False


class TunerManagedError(BaseException):
  """Raised when the tuner wants to terminate the trial."""
