"""Registry of experiment params classes.

Reference `lingvo/model_registry.py`: `@RegisterSingleTaskModel` /
`@RegisterMultiTaskModel` (:200-215), class keys `<task>.<file>.<Class>`
(:85-105), `GetClass/GetParams/GetProgramSchedule` (:378-390),
`--model_params_override` / `--model_params_file_override` (:172-197),
`Task_<Dataset>` overrides (:282-294).

Here experiment modules live under `lingvo_b200.models.<task>.params.<file>`.
"""

import inspect
import os

from lingvo_b200 import flags
from lingvo_b200.core import base_model_params

flags.DEFINE_string('model_params_override', '',
                    'Optional text specifying `key : value` overrides '
                    'separated by ";" or newlines.')
flags.DEFINE_string('model_params_file_override', '',
                    'Optional file with newline-separated overrides.')
flags.DEFINE_string('executor_datasets_to_eval', None,
                    'Semicolon-separated datasets for the executor to eval.')
flags.DEFINE_string('executor_oneoff_checkpoint_to_load', None,
                    'Override checkpoint for one-off eval/decode.')
FLAGS = flags.FLAGS

_PREFIXES = ('lingvo_b200.models.', 'lingvo.tasks.')


class _ModelRegistryHelper:
  """Holds the global key → params-class map."""

  _MODEL_PARAMS_ALLOW_REDEF = False
  _MODEL_PARAMS = {}
  _REGISTERED_MODULES = set()

  @classmethod
  def _ModelParamsClassKey(cls, src_cls):
    path = src_cls.__module__
    for pre in _PREFIXES:
      if path.startswith(pre):
        path = path[len(pre):]
        break
    path = path.replace('params.', '')
    try:
      if inspect.getfile(src_cls).endswith('test.py'):
        return 'test.{}'.format(src_cls.__name__)
    except (TypeError, OSError):
      pass
    return '{}.{}'.format(path, src_cls.__name__)

  @classmethod
  def _GetSourceInfo(cls, src_cls):
    try:
      src_file = inspect.getsourcefile(src_cls)
    except (TypeError, OSError):        # classes built at run time have no source file
      src_file = '<dynamic>'
    info = '%s@%s' % (cls._ModelParamsClassKey(src_cls), src_file)
    try:
      return '%s:%d' % (info, inspect.getsourcelines(src_cls)[-1])
    except (TypeError, OSError):
      return info

  @classmethod
  def _RegisterModel(cls, wrapper_cls, src_cls):
    key = cls._ModelParamsClassKey(src_cls)
    if not cls._MODEL_PARAMS_ALLOW_REDEF and key in cls._MODEL_PARAMS:
      existing = cls._MODEL_PARAMS[key]
      if getattr(existing, '_src_cls', None) is not src_cls and (
          getattr(existing, '_src_qualname', None) != (
              src_cls.__module__, src_cls.__qualname__)):
        raise ValueError('Duplicate model registered for key {}: {}.{}'.format(
            key, src_cls.__module__, src_cls.__name__))
    cls._REGISTERED_MODULES.add(src_cls.__module__)
    cls._MODEL_PARAMS[key] = wrapper_cls
    return key

  @classmethod
  def _CreateWrapperClass(cls, src_cls):
    helper = cls

    class Registered(src_cls):
      """Registered model wrapper: annotates Model() with source info."""
      _src_cls = src_cls
      _src_qualname = (src_cls.__module__, src_cls.__qualname__)

      def Model(self):
        p = super().Model()
        p.model = helper._GetSourceInfo(src_cls)
        return p

    Registered.__name__ = src_cls.__name__
    return Registered

  @classmethod
  def MaybeUpdateParamsFromFlags(cls, cfg):
    if FLAGS.model_params_override and FLAGS.model_params_file_override:
      raise ValueError('Only one of --model_params_override and '
                       '--model_params_file_override may be specified.')
    if FLAGS.model_params_override:
      text = FLAGS.model_params_override.replace(';', '\n')
      cfg.FromText(text, type_overrides={
          'task.train.init_from_checkpoint_override': 'str'})
    if FLAGS.model_params_file_override and os.path.exists(
        FLAGS.model_params_file_override):
      with open(FLAGS.model_params_file_override) as f:
        cfg.FromText(f.read())

  @classmethod
  def RegisterSingleTaskModel(cls, src_cls):
    if not issubclass(src_cls, base_model_params.SingleTaskModelParams):
      raise TypeError('src_cls %s is not a SingleTaskModelParams!' %
                      src_cls.__name__)
    cls._RegisterModel(cls._CreateWrapperClass(src_cls), src_cls)
    return src_cls

  @classmethod
  def RegisterMultiTaskModel(cls, src_cls):
    if not issubclass(src_cls, base_model_params.MultiTaskModelParams):
      raise TypeError('src_cls %s is not a MultiTaskModelParams!' %
                      src_cls.__name__)
    cls._RegisterModel(cls._CreateWrapperClass(src_cls), src_cls)
    return src_cls

  @staticmethod
  def GetAllRegisteredClasses():
    return dict(_ModelRegistryHelper._MODEL_PARAMS)

  @classmethod
  def GetClass(cls, class_key):
    all_params = cls._MODEL_PARAMS
    if class_key not in all_params:
      from lingvo_b200 import model_imports
      model_imports.ImportParams(class_key)
    if class_key not in all_params:
      for k in sorted(all_params):
        pass
      raise LookupError('Model %s not found from list of above known models: %s'
                        % (class_key, sorted(all_params)))
    return all_params[class_key]

  @classmethod
  def GetParamsFromModelParamsObject(cls, model_params_obj, dataset_name):
    cfg = model_params_obj.Model()
    cfg.input = model_params_obj.GetDatasetParams(dataset_name)
    # Per-dataset task overrides: Task_<Dataset>().
    override = getattr(model_params_obj, 'Task_' + dataset_name, None)
    if override is not None and 'task' in cfg:
      cfg.task = override()
    cls.MaybeUpdateParamsFromFlags(cfg)
    return cfg

  @classmethod
  def GetParams(cls, class_key, dataset_name):
    model_params_cls = cls.GetClass(class_key)
    return cls.GetParamsFromModelParamsObject(model_params_cls(), dataset_name)

  @classmethod
  def GetProgramSchedule(cls, class_key):
    return cls.GetClass(class_key)().ProgramSchedule()


RegisterSingleTaskModel = _ModelRegistryHelper.RegisterSingleTaskModel
RegisterMultiTaskModel = _ModelRegistryHelper.RegisterMultiTaskModel
GetAllRegisteredClasses = _ModelRegistryHelper.GetAllRegisteredClasses
GetClass = _ModelRegistryHelper.GetClass
GetParams = _ModelRegistryHelper.GetParams
GetParamsFromModelParamsObject = (
    _ModelRegistryHelper.GetParamsFromModelParamsObject)
GetProgramSchedule = _ModelRegistryHelper.GetProgramSchedule
