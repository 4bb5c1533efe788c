"""Registry-wide model validation (ref `lingvo/models_test_helper.py`).

`BaseModelsTest.CreateTestMethodsForAllRegisteredModels(registry, task_regexes=…)` adds one
test method per registered model that (1) fetches the params of every dataset, (2) checks
`ToText`/`FromText` round-trips, (3) instantiates the model with variables on the `meta`
device (no weight memory, shape/dtype/name logic still runs), (4) checks that variable
names are unique and every layer's params are frozen copies.
"""

from __future__ import annotations

import re
import unittest

import torch

from lingvo_b200.core import base_layer
from lingvo_b200.core import cluster_factory
from lingvo_b200.core import py_utils


def TraverseLayer(layer, fn):
  """Calls `fn(layer)` on a layer and, recursively, on all of its children (ref :72)."""
  if isinstance(layer, (list, tuple)):
    for l in layer:
      TraverseLayer(l, fn)
    return
  fn(layer)
  for child in getattr(layer, 'children', {}).values():
    if isinstance(child, (list, tuple, base_layer.BaseLayer)):
      TraverseLayer(child, fn)


class BaseModelsTest(unittest.TestCase):
  """ref :96."""

  DATASETS = ('Train', 'Dev', 'Test')

  def _ValidateEMA(self, name, mdl):
    """If the model asks for EMA, every trainable variable must have a shadow."""
    if not mdl.params.train.ema_decay:
      return
    for task in mdl.tasks:
      shadows = task.EmaShadowTensors()
      n_train = sum(1 for v in task.vars.Flatten() if v.requires_grad)
      self.assertGreaterEqual(len(shadows), n_train, 'EMA shadows missing in %s' % name)

  def _testOneModelParams(self, registry, name):  # pylint: disable=invalid-name
    cls = registry.GetClass(name)
    built = False
    for ds in self.DATASETS:
      try:
        mp = cls().Model()
        mp.input = getattr(cls(), ds)()
      except (NotImplementedError, AttributeError):
        continue
      text = mp.ToText()
      self.assertTrue(text)
      if built:
        continue
      mp.cluster.mode = 'sync'
      mp.cluster.job = 'trainer_client'
      if 'task' in mp:
        mp.task.input = mp.input
      try:
        with cluster_factory.Cluster(mp.cluster), py_utils.StubVariablesScope('zeros'):
          tasks = [mp.task.Instantiate()] if 'task' in mp else []
      except FileNotFoundError as e:
        self.skipTest('dataset / vocab files not present: %s' % e)
      for task in tasks:
        names = [v.var_name for v in task.vars.Flatten()]
        self.assertEqual(len(names), len(set(names)), 'duplicate variable names in %s' % name)
        seen = []
        TraverseLayer(task, seen.append)
        self.assertGreater(len(seen), 0)
      built = True
    self.assertTrue(built, 'no dataset of %s could be built' % name)

  @classmethod
  def CreateTestMethodsForAllRegisteredModels(cls, registry, task_regexes=None,
                                              exclude_regexes=None):
    """Adds `testModelParams_<name>` for every matching registered model (ref :172)."""
    task_regexes = [re.compile(r) for r in (task_regexes or ['.*'])]
    exclude_regexes = [re.compile(r) for r in (exclude_regexes or [])]
    for name in sorted(registry.GetAllRegisteredClasses()):
      if not any(r.search(name) for r in task_regexes):
        continue
      if any(r.search(name) for r in exclude_regexes):
        continue

      def _Test(self, name=name):
        self._testOneModelParams(registry, name)  # pylint: disable=protected-access

      setattr(cls, 'testModelParams_' + re.sub(r'\W', '_', name), cls.TransformTest(_Test))

  @classmethod
  def TransformTest(cls, test_method):
    """Hook for subclasses to wrap every generated test (e.g. to add skips)."""
    return test_method
