"""Dataset-name discovery on params classes (reference `lingvo/datasets.py:34-96`)."""

import inspect

from lingvo_b200.core import base_model_params

DatasetFunctionError = type('DatasetFunctionError', (TypeError,), {})
GetAllDatasetParamsNotImplementedError = (
    base_model_params.GetAllDatasetParamsNotImplementedError)


def GetDatasets(cls, warn_on_error=True):
  """Names of the dataset methods (`Train`, `Dev`, …) of a params class."""
  mdl_params = None
  if inspect.isclass(cls):
    try:
      mdl_params = cls()
    except TypeError:
      mdl_params = None
  else:
    mdl_params = cls
    cls = type(cls)
  if mdl_params is not None:
    try:
      return sorted(mdl_params.GetAllDatasetParams().keys())
    except GetAllDatasetParamsNotImplementedError:
      pass
  datasets = []
  skip = {'GetAllDatasetParams', 'GetDatasetParams', 'Model', 'Task',
          'ProgramSchedule', 'Search'}
  for name, fn in inspect.getmembers(cls, inspect.isroutine):
    if name in skip or name.startswith('_') or not name[0].isupper():
      continue
    try:
      sig = inspect.signature(fn)
    except (TypeError, ValueError):
      continue
    params = [p for p in sig.parameters.values() if p.name not in ('self', 'cls')]
    if any(p.default is inspect.Parameter.empty and p.kind in (
        p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD) for p in params):
      if warn_on_error:
        continue
      raise DatasetFunctionError(
          'Found a public function %s in %s with required arguments' %
          (name, cls.__name__))
    datasets.append(name)
  return sorted(datasets)
