"""Generates "does every registered model build?" test cases (ref
`lingvo/core/test_trainer_utils.py`).

`MakeModelValidatorTestCase(['lm.one_billion_wds.X', …])` returns a `unittest.TestCase`
subclass with `testTrain` / `testDecoder`: for each model the task is instantiated under
a trainer (resp. decoder) cluster with its variables created lazily on the `meta` device,
so even 100 B-parameter configs are validated in milliseconds without allocating weights.
"""

from __future__ import annotations

import unittest

import torch

from lingvo_b200 import model_registry
from lingvo_b200.core import cluster_factory
from lingvo_b200.core import py_utils


def _ModelTuples(model_classes):
  return [(str(m).replace('.', '_'), m) for m in model_classes]


def MakeModelValidatorTestCase(model_classes):
  """ref :36."""

  class _ModelValidator(unittest.TestCase):
    """`TrainerBuilds` / `DecoderBuilds` for every model in `model_classes`."""

    def _Params(self, model, dataset):
      mp = model_registry.GetParams(model, dataset)
      mp.cluster.mode = 'sync'
      return mp

    def _Build(self, mp):
      """Instantiates the task (with its input generator) with variables on `meta`."""
      mp.task.input = mp.input
      try:
        with cluster_factory.Cluster(mp.cluster), py_utils.StubVariablesScope('zeros'):
          return mp.task.Instantiate()
      except (FileNotFoundError, RuntimeError) as e:
        msg = str(e)
        if isinstance(e, FileNotFoundError) or 'No such file' in msg or 'no files match' in msg \
            or 'cannot open' in msg:
          self.skipTest('dataset / vocab files not present: %s' % msg)
        raise

    def TrainerBuilds(self, model):
      mp = self._Params(model, 'Train')
      mp.cluster.job = 'trainer_client'
      task = self._Build(mp)
      self.assertGreater(len(task.vars.Flatten()), 0)
      return task

    def DecoderBuilds(self, model):
      try:
        mp = self._Params(model, 'Test')
      except (NotImplementedError, AttributeError):
        mp = self._Params(model, 'Train')
      mp.cluster.job = 'decoder'
      mp.cluster.do_eval = True
      task = self._Build(mp)
      self.assertTrue(hasattr(task, 'Decode'))
      return task

    def testTrain(self):  # pylint: disable=invalid-name
      for label, model in _ModelTuples(model_classes):
        with self.subTest(model=label):
          self.TrainerBuilds(model)

    def testDecoder(self):  # pylint: disable=invalid-name
      for label, model in _ModelTuples(model_classes):
        with self.subTest(model=label):
          self.DecoderBuilds(model)

  return _ModelValidator
