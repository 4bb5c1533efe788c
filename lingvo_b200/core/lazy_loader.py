"""Deferred module import (ref `lingvo/core/lazy_loader.py`)."""
import importlib
import types


class LazyLoader(types.ModuleType):
  """Imports `name` on first attribute access and installs it into `parent_globals`."""

  def __init__(self, local_name, parent_module_globals, name):
    self._local_name = local_name
    self._parent = parent_module_globals
    super().__init__(name)

  def _Load(self):
    module = importlib.import_module(self.__name__)
    self._parent[self._local_name] = module
    self.__dict__.update(module.__dict__)
    return module

  def __getattr__(self, item):
    return getattr(self._Load(), item)

  def __dir__(self):
    return dir(self._Load())
