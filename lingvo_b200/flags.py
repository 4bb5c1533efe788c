"""A tiny absl-like flags module (`DEFINE_*`, global `FLAGS`).

The reference uses absl flags for process topology and a few global
behaviours (`trainer.py:54-205`, `trainer_utils.py:19-59`). absl is not
available in this image, so this module provides the same surface on top of
argparse: flags are defined at import time and parsed from `sys.argv`
(`--name=value`, `--name value`, `--[no]bool`).
"""

import argparse
import sys
from typing import Any, Dict, List, Optional


class _Flag:

  def __init__(self, name, default, help_text, parser):
    self.name, self.default, self.help, self.parser = name, default, help_text, parser
    self.value = default
    self.present = False


def _ParseBool(v):
  if isinstance(v, bool):
    return v
  return str(v).lower() in ('1', 'true', 't', 'yes', 'y')


class _FlagValues:

  def __init__(self):
    self.__dict__['_flags'] = {}
    self.__dict__['_parsed'] = False

  def _Define(self, name, default, help_text, parser):
    flags = self.__dict__['_flags']
    if name in flags:
      return  # idempotent: modules may be re-imported under test.
    flags[name] = _Flag(name, default, help_text, parser)

  def __getattr__(self, name):
    flags = self.__dict__['_flags']
    if name not in flags:
      raise AttributeError('Unknown flag --%s' % name)
    return flags[name].value

  def __setattr__(self, name, value):
    flags = self.__dict__['_flags']
    if name not in flags:
      raise AttributeError('Unknown flag --%s' % name)
    flags[name].value = value
    flags[name].present = True

  def __contains__(self, name):
    return name in self.__dict__['_flags']

  def __getitem__(self, name):
    return self.__dict__['_flags'][name]

  def is_parsed(self):
    return self.__dict__['_parsed']

  def __call__(self, argv: Optional[List[str]] = None, known_only=True):
    """Parses argv; returns the remaining (unknown) args."""
    argv = list(sys.argv if argv is None else argv)
    prog, args = argv[0], argv[1:]
    flags = self.__dict__['_flags']
    rest = []
    i = 0
    while i < len(args):
      a = args[i]
      if not a.startswith('--') or a == '--':
        rest.append(a)
        i += 1
        continue
      body = a[2:]
      if '=' in body:
        k, v = body.split('=', 1)
      else:
        k, v = body, None
      k = k.replace('-', '_')
      if k not in flags and k.startswith('no') and k[2:] in flags and (
          flags[k[2:]].parser is _ParseBool):
        flags[k[2:]].value = False
        flags[k[2:]].present = True
        i += 1
        continue
      if k not in flags:
        if not known_only:
          raise ValueError('Unknown flag %s' % a)
        rest.append(a)
        i += 1
        continue
      f = flags[k]
      if v is None:
        if f.parser is _ParseBool:
          v = True
        else:
          i += 1
          if i >= len(args):
            raise ValueError('Flag --%s needs a value' % k)
          v = args[i]
      f.value = f.parser(v)
      f.present = True
      i += 1
    self.__dict__['_parsed'] = True
    return [prog] + rest

  def flag_values_dict(self) -> Dict[str, Any]:
    return {k: f.value for k, f in self.__dict__['_flags'].items()}

  def reset(self):
    for f in self.__dict__['_flags'].values():
      f.value = f.default
      f.present = False


FLAGS = _FlagValues()


def DEFINE_string(name, default, help_text=''):  # pylint: disable=invalid-name
  FLAGS._Define(name, default, help_text, str)


def DEFINE_integer(name, default, help_text=''):  # pylint: disable=invalid-name
  FLAGS._Define(name, default, help_text, int)


def DEFINE_float(name, default, help_text=''):  # pylint: disable=invalid-name
  FLAGS._Define(name, default, help_text, float)


def DEFINE_bool(name, default, help_text=''):  # pylint: disable=invalid-name
  FLAGS._Define(name, default, help_text, _ParseBool)


DEFINE_boolean = DEFINE_bool


def DEFINE_list(name, default, help_text=''):  # pylint: disable=invalid-name
  FLAGS._Define(name, default, help_text,
                lambda v: v if isinstance(v, list) else
                [s for s in str(v).split(',') if s])
