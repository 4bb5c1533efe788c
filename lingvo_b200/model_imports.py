"""Lazy import of experiment-params modules by model name.

Reference `lingvo/model_imports.py:76-107`: `--model=image.mnist.LeNet5`
imports `<root>.image.params.mnist`. `ImportAllParams` walks every task's
`params/` package (used by the all-models smoke test).
"""

import importlib
import logging
import pkgutil
import re

_TASK_ROOT = 'lingvo_b200.models'
_TASK_DIRS = ('asr', 'car', 'image', 'lm', 'milan', 'mt', 'punctuator')


def _Import(name, report_error=False):
  try:
    importlib.import_module(name)
    return True
  except ModuleNotFoundError as e:
    missing = getattr(e, 'name', '') or ''
    if report_error and not name.startswith(missing + '.') and name != missing:
      logging.warning('Could not import %s: %s', name, e)
    if not (name == missing or name.startswith(missing + '.')):
      raise
    return False


def ImportAllParams(task_root=_TASK_ROOT, task_dirs=_TASK_DIRS,
                    require_success=False):
  success = False
  for task in task_dirs:
    pkg = '%s.%s.params' % (task_root, task)
    try:
      mod = importlib.import_module(pkg)
    except ModuleNotFoundError:
      continue
    for info in pkgutil.iter_modules(mod.__path__):
      if info.name.endswith('_test'):
        continue
      success = _Import('%s.%s' % (pkg, info.name)) or success
    # Nested sub-packages (e.g. mt.params.xendec).
    for info in pkgutil.walk_packages(mod.__path__, pkg + '.'):
      if info.ispkg:
        continue
      success = _Import(info.name) or success
  if require_success and not success:
    raise LookupError('Could not import any task params from %s' % task_root)
  return success


def ImportParams(model_name, task_root=_TASK_ROOT, task_dirs=_TASK_DIRS,
                 require_success=True):
  """`image.mnist.LeNet5` → import lingvo_b200.models.image.params.mnist."""
  if '.' not in model_name:
    raise ValueError('Invalid model name %s' % model_name)
  model_module = model_name.rpartition('.')[0]
  # Try the name as a fully qualified module first.
  success = _Import(model_module)
  for task in task_dirs:
    if model_module.startswith(task + '.'):
      path = model_module[len(task) + 1:]
      success = _Import('%s.%s.params.%s' % (task_root, task, path)) or success
  if require_success and not success:
    raise LookupError('Could not find any valid import paths for module %s. '
                      'Check the logs above to see if there were errors '
                      'importing the module.' % model_module)
  return success
