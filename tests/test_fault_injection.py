"""Recovery paths driven by injected faults."""

import os

import pytest
import torch

from lingvo_b200.core import fault_injection as fi


def test_parse_spec():
  fs = fi.ParseSpec('transient@5, nan_grad@7,transient@12x2,exit@20:rank=1,stall@3:seconds=0.5')
  assert [(f.kind, f.step, f.remaining) for f in fs] == [
      ('transient', 5, 1), ('nan_grad', 7, 1), ('transient', 12, 2), ('exit', 20, 1), ('stall', 3, 1)]
  assert fs[3].opts == {'rank': 1} and fs[4].opts == {'seconds': 0.5}
  with pytest.raises(ValueError):
    fi.ParseSpec('bogus')
  with pytest.raises(AssertionError):
    fi.ParseSpec('explode@3')


def test_rank_filter_and_repeat():
  inj = fi.Injector(fi.ParseSpec('transient@2x2:rank=1'), rank=0)
  inj.BeforeStep(2)                                   # other rank: nothing
  inj = fi.Injector(fi.ParseSpec('transient@2x2:rank=1'), rank=1)
  for _ in range(2):
    with pytest.raises(ConnectionError):
      inj.BeforeStep(2)
  inj.BeforeStep(2)                                   # exhausted
  assert inj.fired == {'transient': 2}


def test_nan_gradient_is_skipped_by_the_learner():
  from lingvo_b200.core import learner as learner_lib, py_utils
  from lingvo_b200.core.nested_map import NestedMap
  fi.Arm('nan_grad@0')
  try:
    lrn = learner_lib.Learner.Params().Set(name='loss', learning_rate=0.1).Instantiate()
    w = torch.nn.Parameter(torch.ones(4))
    w.var_name = 'w/var'
    py_utils.SetGlobalStep(0)
    for step in range(2):
      with py_utils.GlobalStepContext(step):
        loss = (w * w).sum()
        _, metrics = lrn.Apply({'loss': (loss, torch.tensor(1.0))}, NestedMap(w=w))
      flag = [v for k, v in metrics.items() if k.startswith('has_nan_or_inf')][0]
      if step == 0:
        assert float(flag[0]) == 1.0
        torch.testing.assert_close(w.detach(), torch.ones(4))       # update skipped
      else:
        assert float(flag[0]) == 0.0
        assert (w.detach() < 1.0).all()                             # training goes on
  finally:
    fi.Disarm()


def test_trainer_retries_transient_fault_and_resumes_from_checkpoint(tmp_path, monkeypatch):
  from lingvo_b200 import flags, trainer
  from lingvo_b200.models.image import input_generator
  data = input_generator.FakeMnistData(str(tmp_path), train_size=64, test_size=32)
  monkeypatch.setenv('LINGVO_B200_MNIST', data)
  logdir = str(tmp_path / 'log')
  inj = fi.Arm('transient@3')
  try:
    flags.FLAGS.reset()
    trainer.main(['trainer', '--run_locally=cpu', '--mode=sync', '--model=image.mnist.LeNet5',
                  '--logdir=' + logdir,
                  '--model_params_override=task.train.max_steps:5;input.batch_size:8;'
                  'task.train.save_interval_steps:2;task.train.summary_interval_steps:100'])
  finally:
    fi.Disarm()
    flags.FLAGS.reset()
  assert inj.fired == {'transient': 1}
  assert os.path.exists(os.path.join(logdir, 'train', 'ckpt-00000005.index'))


def test_fatal_fault_fails_fast(tmp_path, monkeypatch):
  from lingvo_b200 import flags, trainer
  from lingvo_b200.models.image import input_generator
  data = input_generator.FakeMnistData(str(tmp_path), train_size=64, test_size=32)
  monkeypatch.setenv('LINGVO_B200_MNIST', data)
  fi.Arm('fatal@1')
  try:
    flags.FLAGS.reset()
    with pytest.raises(ValueError, match='fatal fault'):
      trainer.main(['trainer', '--run_locally=cpu', '--mode=sync', '--model=image.mnist.LeNet5',
                    '--logdir=' + str(tmp_path / 'log2'),
                    '--model_params_override=task.train.max_steps:3;input.batch_size:8'])
  finally:
    fi.Disarm()
    flags.FLAGS.reset()
