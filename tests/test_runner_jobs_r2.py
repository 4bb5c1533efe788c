"""Checkpoint-following jobs (Controller / TrainSummaries), eager runner table, input policy,
RNN param factories."""

import glob
import os

import pytest
import torch

from lingvo_b200 import eager_runners
from lingvo_b200 import runners
from lingvo_b200.core import base_input_generator
from lingvo_b200.core import hyperparams
from lingvo_b200.core import input_policy
from lingvo_b200.core import model_helper
from lingvo_b200.core import rnn_cell
from lingvo_b200.core import rnn_layers
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.utils import tfevents


def _MnistCfg(tmp_path, monkeypatch, max_steps):
  from lingvo_b200 import model_registry
  from lingvo_b200.models.image import input_generator
  import lingvo_b200.models.image.params.mnist  # noqa: F401  pylint: disable=unused-import
  data = input_generator.FakeMnistData(str(tmp_path), train_size=64, test_size=32)
  monkeypatch.setenv('LINGVO_B200_MNIST', data)
  cfg = model_registry.GetParams('image.mnist.LeNet5', 'Train')
  cfg.input.batch_size = 8
  cfg.task.train.max_steps = max_steps
  cfg.train.max_steps = max_steps
  cfg.task.train.summary_interval_steps = 1
  cfg.cluster.mode = 'sync'
  cfg.cluster.job = 'trainer_client'
  return cfg


def test_train_summaries_job_follows_checkpoints(tmp_path, monkeypatch):
  logdir = str(tmp_path / 'log')
  cfg = _MnistCfg(tmp_path, monkeypatch, max_steps=3)
  tr = runners.Trainer(cfg, '', logdir, '', None)
  tr.Start()
  assert os.path.exists(os.path.join(logdir, 'train', 'ckpt-00000003.index'))
  before = {v.var_name: v.detach().clone() for v in tr.task.vars.Flatten()}

  job = eager_runners.TrainSummaries(cfg.Copy(), '', logdir, '', None)
  assert isinstance(job, runners.TrainSummaries)
  job.Start()                      # sees the final checkpoint (step 3 == max_steps), then stops
  assert job.num_summaries_written == 1
  files = glob.glob(os.path.join(logdir, 'train_summaries', 'events.out.tfevents.*'))
  assert files
  got = {}
  for f in files:
    for step, tag, val in tfevents.ReadScalars(f):
      got[tag] = (step, val)
  assert got['global_step'] == (3, 3.0)
  assert got['grad_norm/all'][1] > 0 and got['var_norm/all'][1] > 0
  assert 'log_pplx' in got and 'loss' in got and 'total_num_params' in got
  # summarising never updates the weights: same values as the checkpoint it restored
  for v in job._model.GetTask().vars.Flatten():   # pylint: disable=protected-access
    torch.testing.assert_close(v.detach(), before[v.var_name])
  # an explicit call on an older checkpoint works too
  older = sorted(glob.glob(os.path.join(logdir, 'train', 'ckpt-*.index')))[0][:-len('.index')]
  assert job.SummarizeCheckpoint(older) <= 3 and job.num_summaries_written == 2


def test_eager_runner_table():
  assert eager_runners.GetRunnerClass('trainer') is runners.Trainer
  assert eager_runners.GetRunnerClass('trainer_client') is runners.Trainer
  assert eager_runners.GetRunnerClass('evaler_dev') is runners.Evaler
  assert eager_runners.GetRunnerClass('decoder_test') is runners.Decoder
  assert eager_runners.GetRunnerClass('train_summaries') is runners.TrainSummaries
  with pytest.raises(ValueError):
    eager_runners.GetRunnerClass('controller')      # no controller in the eager family


class _Gen(base_input_generator.BaseInputGenerator):

  def _InputBatch(self):
    return NestedMap(x=torch.zeros(4, 2), ids=torch.arange(4))


def test_input_policy_pins_batches_to_the_input_device():
  p = _Gen.Params().Set(name='g', batch_size=4)
  q = input_policy.Apply(p)
  assert q.cls is not _Gen and issubclass(q.cls, _Gen) and q.cls.__name__ == '_Gen'
  assert input_policy.Apply(q) is q                              # idempotent
  gen = q.Instantiate()
  with torch.device('meta'):                                      # a stray device context …
    batch = gen.GetPreprocessedInputBatch()
    splits = gen.SplitInputBatch(2)
  assert batch.x.device.type == 'cpu' and batch.ids.device.type == 'cpu'   # … is overridden
  assert [s.x.shape[0] for s in splits] == [2, 2] and splits[0].x.device.type == 'cpu'


def test_base_task_applies_the_input_policy(tmp_path, monkeypatch):
  cfg = _MnistCfg(tmp_path, monkeypatch, max_steps=1)
  from lingvo_b200.core import cluster_factory
  task = cfg.task.Copy()
  task.input = cfg.input
  with cluster_factory.Cluster(cfg.cluster):
    t = task.Instantiate()
  assert getattr(type(t.input), '_input_policy_applied', False)


def test_rnn_param_factories():
  cell = rnn_cell.LSTMCellSimple.Params().Set(num_input_nodes=4, num_output_nodes=4)
  lp = hyperparams.Params()
  lp.Define('unidi_rnn_type', 'func', '')
  lp.Define('bidi_rnn_type', 'func', '')
  assert model_helper.CreateUnidirectionalRNNParams(lp, cell).cls is rnn_layers.FRNN
  assert model_helper.CreateBidirectionalRNNParams(lp, cell, cell).cls is \
      rnn_layers.BidirectionalFRNN
  for t in ('quasi_ifo', 'sru'):
    lp.unidi_rnn_type = lp.bidi_rnn_type = t
    assert model_helper.CreateUnidirectionalRNNParams(lp, cell).cls is rnn_layers.FRNN
    bp = model_helper.CreateBidirectionalRNNParams(lp, cell, cell)
    assert bp.cls is rnn_layers.BidirectionalFRNNQuasi and bp.fwd is cell
  lp.unidi_rnn_type = lp.bidi_rnn_type = 'native_cudnn'
  with pytest.raises(ValueError):
    model_helper.CreateUnidirectionalRNNParams(lp, cell)
  with pytest.raises(ValueError):
    model_helper.CreateBidirectionalRNNParams(lp, cell, cell)
