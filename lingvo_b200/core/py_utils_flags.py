"""Process-wide behaviour switches (reference `core/py_utils_flags.py`).

Each flag has a process default (overridable by env `LINGVO_B200_<NAME>`),
and can be overridden per-cluster through `cluster.params` attributes of the
same name (reference `_FromGlobal`, :94-122).
"""

import os
import threading

_DEFAULTS = {
    'enable_asserts': True,
    'enable_check_numerics': True,
    'print_debug_tensors': False,
    'xla_device': '',
    'tpu_compatible': False,
    'pin_vars_to_cpu': False,
    'stateless_vars_init': False,
    'use_eager_v2_checkpoints': False,
    'disable_py_utils_debug': False,
    'testonly_skip_norm_layers': False,
    'if_use_tf_function': False,
    'use_fused_kernels': True,
}

_VALUES = dict(_DEFAULTS)
_LOCK = threading.Lock()


def _Coerce(default, text):
  if isinstance(default, bool):
    return text not in ('0', 'false', 'False', '')
  return type(default)(text)


for _k, _v in _DEFAULTS.items():
  _env = os.environ.get('LINGVO_B200_' + _k.upper())
  if _env is not None:
    _VALUES[_k] = _Coerce(_v, _env)


def SetFlag(name, value):
  assert name in _VALUES, name
  with _LOCK:
    _VALUES[name] = value


def GetFlag(name):
  """Cluster override (if the current cluster defines `name`) else global."""
  try:
    from lingvo_b200.core import cluster_factory
    cluster = cluster_factory.Current()
    val = getattr(cluster.params, name, None) if name in cluster.params else None
    if val is not None:
      return val
  except Exception:  # pylint: disable=broad-except
    pass
  return _VALUES[name]


def enable_asserts():
  return _VALUES['enable_asserts']


def enable_check_numerics():
  return _VALUES['enable_check_numerics']


def print_debug_tensors():
  return _VALUES['print_debug_tensors']


def use_fused_kernels():
  return _VALUES['use_fused_kernels']


def use_xla():
  return False


def use_tpu():
  return False


def use_gpu():
  """True when the current cluster's job runs on CUDA devices."""
  try:
    from lingvo_b200.core import cluster_factory
    return cluster_factory.Current().RunsOnGpu()
  except Exception:  # pylint: disable=broad-except
    import torch
    return torch.cuda.is_available()


def tpu_compat():
  return _VALUES['tpu_compatible']
