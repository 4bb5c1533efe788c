"""Import every module of the package (catches dormant syntax / import / name errors) and
poke the public classes of modules no other test touches."""

import importlib
import os
import pkgutil

import numpy as np
import pytest
import torch

import lingvo_b200


def _AllModules():
  root = os.path.dirname(lingvo_b200.__file__)
  names = []
  for m in pkgutil.walk_packages([root], prefix='lingvo_b200.'):
    names.append(m.name)
  return sorted(names)


@pytest.mark.parametrize('name', _AllModules())
def test_module_imports(name):
  if name.endswith(('.gke_launch',)):
    pytest.importorskip('yaml')
  try:
    importlib.import_module(name)
  except Exception as e:  # pylint: disable=broad-except
    # Command-line tools define absl flags at import time; two tools may reuse a flag name
    # (they are separate binaries), which only clashes when both live in one process.
    if type(e).__name__ == 'DuplicateFlagError' and '.tools.' in name:
      pytest.skip('flag name shared with another tool binary')
    raise


def test_every_registered_layer_params_roundtrip_text():
  """`cls.Params()` of every BaseLayer subclass in core/ serialises to text and back."""
  from lingvo_b200.core import base_layer, hyperparams
  seen = 0
  for name in _AllModules():
    if '.core.' not in name:
      continue
    mod = importlib.import_module(name)
    for attr in dir(mod):
      cls = getattr(mod, attr)
      if not (isinstance(cls, type) and issubclass(cls, base_layer.BaseLayer)) or cls.__module__ != name:
        continue
      try:
        p = cls.Params()
      except (NotImplementedError, TypeError):
        continue
      text = p.ToText()
      q = cls.Params()
      q.FromText(text)
      assert q.ToText() == text, (name, attr)
      seen += 1
  assert seen > 150, seen
