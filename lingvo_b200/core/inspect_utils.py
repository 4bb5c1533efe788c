"""Params ⟷ callable signature helpers (ref `lingvo/core/inspect_utils.py`)."""
import inspect

from lingvo_b200.core import hyperparams


def _IsDefinableParameter(parameter):
  return parameter.kind in (inspect.Parameter.POSITIONAL_OR_KEYWORD,
                            inspect.Parameter.KEYWORD_ONLY)


def _ExtractParameters(func, ignore, bound):
  ignore = set(ignore or [])
  params = list(inspect.signature(func).parameters.values())
  if bound and params and params[0].name in ('self', 'cls'):
    params = params[1:]
  return [p for p in params if _IsDefinableParameter(p) and p.name not in ignore]


def DefineParams(func, params, ignore=None, bound=False):
  """Defines one param per argument of `func` (default = the argument's default)."""
  for p in _ExtractParameters(func, ignore, bound):
    default = None if p.default is inspect.Parameter.empty else p.default
    params.Define(p.name, default, 'Function parameter.')
  return params


def _MakeArgs(func, params, ignore, bound, **kwargs):
  out = {}
  for p in _ExtractParameters(func, ignore, bound):
    if p.name in kwargs:
      out[p.name] = kwargs[p.name]
    elif p.name in params:
      out[p.name] = params.Get(p.name)
  for k, v in kwargs.items():
    out.setdefault(k, v)
  return out


def CallWithParams(func, params, **kwargs):
  """Calls `func` taking its arguments from `params` (kwargs override)."""
  return func(**_MakeArgs(func, params, None, False, **kwargs))


def ConstructWithParams(cls, params, **kwargs):
  return cls(**_MakeArgs(cls.__init__, params, None, True, **kwargs))


def ParamsFromCallable(func, ignore=None, bound=False):
  return DefineParams(func, hyperparams.Params(), ignore, bound)
