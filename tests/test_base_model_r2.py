"""Round-2 base_model helpers (reference base_model.py: _UpdateVnConfig :1084,
_ComputeGradientMask :837, CopyTaskParams :1392, TaskNames :1512, EMA helpers)."""

import pytest
import torch

from lingvo_b200 import model_registry
from lingvo_b200.core import base_model
from lingvo_b200.core import cluster_factory
from lingvo_b200.core import py_utils
import lingvo_b200.models.lm.params.synthetic_packed_input  # noqa: F401

MODEL = 'lm.synthetic_packed_input.DenseLmTiny'


def _TaskParams():
  return model_registry.GetParams(MODEL, 'Train').task


def test_vn_config_follows_train_params():
  tp = _TaskParams()
  tp.vn = py_utils.VariationalNoiseParams(None, global_vn=True)
  tp.train.vn_std = 0.3
  tp.train.vn_start_step = 7
  task = tp.Instantiate()
  assert task.params.vn.scale == 0.3 and task.params.vn.start_step == 7
  assert task.params.vn.global_vn
  # children inherit the task's noise config
  child = next(iter(task.children.values()))
  leaf = [l for l in task.Flatten() if hasattr(l, 'params') and 'vn' in l.params] \
      if hasattr(task, 'Flatten') else [child]
  assert tp.vn.scale is None                                   # caller's params untouched
  # vn_std == 0 ⇒ off, whatever p.vn says
  off = _TaskParams()
  off.vn = py_utils.VariationalNoiseParams(None, global_vn=True)
  off.train.vn_std = 0.0
  assert not off.Instantiate().params.vn.global_vn
  # eval ⇒ off
  with cluster_factory.SetEval(True):
    ev = _TaskParams()
    ev.vn = py_utils.VariationalNoiseParams(None, global_vn=True)
    ev.train.vn_std = 0.3
    assert ev.Instantiate().params.vn.scale is None
  bad = _TaskParams()
  bad.vn = py_utils.VariationalNoiseParams(0.1, global_vn=True)
  bad.train.vn_std = 0.3
  with pytest.raises(ValueError):
    bad.Instantiate()
  del leaf


def test_gradient_mask_ema_helpers_and_export(tmp_path):
  tp = _TaskParams()
  tp.train.ema_decay = 0.9
  mp = base_model.SingleTaskModel.Params(tp)
  assert mp.name == tp.name and mp.train.ema_decay == 0.9
  model = mp.Instantiate()
  model.InstantiateVariables()
  task = model.GetTask()
  names = [v.var_name for v in task.vars.Flatten()]
  mask = task._ComputeGradientMask(['emb', 'nothing_matches', '.*'])
  assert set(mask.keys()) == set(names)
  emb = [n for n in names if 'emb' in n][0]
  assert mask[emb].tolist() == [1.0, 0.0, 1.0]
  other = [n for n in names if 'emb' not in n][0]
  assert mask[other].tolist() == [0.0, 0.0, 1.0]
  ema_vars = model.variables_for_ema
  assert ema_vars and all(v.requires_grad for v in ema_vars)
  assert model.ema_decay == 0.9
  task.CreateExponentialMovingAverage()
  shadows = model.MakeEMAVariablesDictTF2()
  assert len(shadows) == len(ema_vars)
  assert all(k.endswith('/ExponentialMovingAverage') for k in shadows)
  model.Export(str(tmp_path))                                   # no-op hook, must not raise
  with pytest.raises(AssertionError):
    _ = task.post_training_loop_op
  task.PostTrainingLoop()
  assert len(task.post_training_loop_op) == len(task.learners)
  assert task.InferenceEager is not None
  with pytest.raises(NotImplementedError):
    task.EmailDecodeSummary({}, [], base_model.DecodeEmailOptions('decoder', 1, 0))
  assert base_model.ExecutorEma().ema is None


def test_multi_task_names_are_sorted():
  from lingvo_b200.core import hyperparams
  p = hyperparams.Params()
  p.Define('task_params', hyperparams.Params(), '')
  p.task_params.Define('zeta', None, '')
  p.task_params.Define('alpha', None, '')
  assert base_model.MultiTaskModel.TaskNames(p) == ['alpha', 'zeta']
