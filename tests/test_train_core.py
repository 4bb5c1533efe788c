"""Optimizers, learner, gating parity, checkpoint & trainer integration (CPU)."""
import glob
import math
import os

import numpy as np
import pytest
import torch

from lingvo_b200.core import gshard_layers
from lingvo_b200.core import optimizer
from lingvo_b200.core import py_utils
from lingvo_b200.core.nested_map import NestedMap


def _Apply(opt_params, steps=3, shape=(8, 4)):
  torch.manual_seed(0)
  w = torch.nn.Parameter(torch.randn(shape))
  w.var_name = 'w/var'
  opt = opt_params.Set(name='opt').Instantiate()
  gs = [torch.randn(shape) * 0.1 for _ in range(steps)]
  for i, g in enumerate(gs):
    with py_utils.GlobalStepContext(i):
      opt.Apply(0.1, [py_utils.VarGrad(w, g)])
  return w.detach(), gs, opt


def test_adam_matches_formula():
  w, gs, opt = _Apply(optimizer.Adam.Params().Set(beta1=0.9, beta2=0.999,
                                                  epsilon=1e-6, fused=False))
  torch.manual_seed(0)
  ref = torch.randn(8, 4)
  m = torch.zeros_like(ref)
  v = torch.zeros_like(ref)
  for t, g in enumerate(gs, 1):
    m = 0.9 * m + 0.1 * g
    v = 0.999 * v + 0.001 * g * g
    lr_t = 0.1 * math.sqrt(1 - 0.999**t) / (1 - 0.9**t)
    ref = ref - lr_t * m / (v.sqrt() + 1e-6)
  assert torch.allclose(w, ref, atol=1e-6)
  slots = opt.GetOptimizerSlots()
  assert 'w/Adam' in slots and 'w/Adam_1' in slots and 'beta1_power' in slots


def test_sgd_momentum_adagrad_run():
  for p in [optimizer.SGD.Params(), optimizer.Momentum.Params(),
            optimizer.RMSProp.Params(), optimizer.Adagrad.Params(),
            optimizer.AdaDelta.Params(),
            optimizer.Accumulator.Params().Set(accum_steps=2),
            optimizer.AdaGraft.Params(),
            optimizer.DistributedShampoo.Params().Set(
                start_preconditioning_steps=1, preconditioning_compute_steps=1)]:
    w, _, _ = _Apply(p)
    assert torch.isfinite(w).all()


def test_adafactor_factored_reference():
  w, gs, _ = _Apply(optimizer.XLAShardingAdafactor.Params().Set(
      beta1=0.0, clipping_threshold=1.0, factored=True, decay_exponent_pow=0.8,
      min_dim_size_to_factor=2, fused=False), steps=2, shape=(6, 4))
  torch.manual_seed(0)
  ref = torch.randn(6, 4)
  vr = torch.zeros(4)   # dims sorted by size: d0=0 (6) → vr = mean over axis 0
  vc = torch.zeros(6)
  for t, g in enumerate(gs):
    decay = 1.0 - (t + 1.0)**-0.8
    g2 = g * g + 1e-30
    vr = vr * decay + g2.mean(0) * (1 - decay)
    vc = vc * decay + g2.mean(1) * (1 - decay)
    x = g * torch.rsqrt(vr / vr.mean()).unsqueeze(0) * torch.rsqrt(vc).unsqueeze(1)
    x = x / max(1.0, float(x.square().mean().sqrt()) / 1.0)
    scale = max(float(ref.square().mean().sqrt()), 1e-3) * 0.1
    ref = ref - x * scale
  assert torch.allclose(w, ref, atol=1e-5)


@pytest.mark.parametrize('legacy', [True, False])
@pytest.mark.parametrize('with_pad', [True, False])
def test_top2_index_gating_matches_dense_oracle(legacy, with_pad):
  torch.manual_seed(3)
  g, s, e = 3, 40, 8
  logits = torch.randn(g, s, e)
  pad = None
  if with_pad:
    pad = torch.zeros(g, s)
    pad[:, -5:] = 1.0
  aux, comb, disp = gshard_layers.Top2GatingOnLogits(
      None, pad, logits, 1, e, 0, torch.float32, False, 'all', 0.0, legacy, 1.0)
  idx = gshard_layers.Top2GatingIndices(logits, pad, e, 0, torch.float32, 'all',
                                        0.0, legacy, 1.0)
  comb2 = gshard_layers.CombineTensorFromIndices(idx, e)
  assert comb.shape == comb2.shape
  assert torch.allclose(comb, comb2, atol=1e-6)
  assert torch.allclose(aux, idx.aux_loss, atol=1e-6)
  assert torch.equal(disp, (comb2 != 0).float())
  # capacity respected & first-choice priority
  assert disp.sum(dim=(0, 1)).max() <= g * s


def test_moe_indexed_equals_dense_apply():
  torch.manual_seed(4)
  g, s, m, h, e = 2, 32, 16, 24, 4
  x = torch.randn(g, s, m)
  gw = torch.randn(m, e)
  wi, wo = torch.randn(e, m, h) * 0.1, torch.randn(e, h, m) * 0.1
  gating = gshard_layers.ComputeGating(gw, x, None, 1, e, 0, True, torch.float32,
                                       capacity_factor=2.0, use_xla_sharding=False)
  dense, _ = gshard_layers.FeedForwardNetworksApplyGating(gating, x, x, wi, wo, 1, g)
  idx = gshard_layers.Top2GatingIndices(torch.einsum('GSM,ME->GSE', x, gw), None,
                                        e, 0, capacity_factor=2.0)
  out = gshard_layers.MoEApplyIndexed(x, idx, wi, wo)
  assert torch.allclose(dense, out, atol=1e-5)


def test_tensor_bundle_and_checkpointer(tmp_path):
  from lingvo_b200 import model_registry
  from lingvo_b200.core import checkpointer
  from lingvo_b200.utils import tensor_bundle
  import lingvo_b200.models.lm.params.synthetic_packed_input  # noqa
  cfg = model_registry.GetParams('lm.synthetic_packed_input.MoELm8ETiny', 'Train')
  cfg.task.fprop_dtype = torch.float32
  cfg.task.builder.fprop_dtype = torch.float32
  cfg.train.async_checkpointing = False
  model = cfg.Instantiate()
  task = model.tasks[0]
  for _ in range(2):
    task.TrainStep()
  ck = checkpointer.Checkpointer(str(tmp_path), model)
  path = ck.Save(gsteps=task.global_step)
  assert os.path.basename(path) == 'ckpt-00000002'
  assert os.path.exists(path + '.index') and os.path.exists(path + '.data-00000-of-00001')
  assert 'model_checkpoint_path: "ckpt-00000002"' in open(tmp_path / 'checkpoint').read()
  keys = tensor_bundle.BundleReader(path).Keys()
  assert 'global_step' in keys
  assert any(k.endswith('/wi/var') for k in keys)
  assert any(k.endswith('/Adafactor_vr') or k.endswith('/Adafactor_v') for k in keys)
  model2 = cfg.Instantiate()
  ck2 = checkpointer.Checkpointer(str(tmp_path), model2)
  assert ck2.Restore() == path
  for a, b in zip(model.vars.Flatten(), model2.vars.Flatten()):
    assert torch.equal(a, b)
  assert model2.tasks[0].global_step == 2
  # identical continuation
  batch = task.GetInputBatch()
  l1 = task.FProp(task.theta, batch)[0]['loss'][0]
  l2 = model2.tasks[0].FProp(model2.tasks[0].theta, batch)[0]['loss'][0]
  assert torch.allclose(l1, l2)


def test_trainer_mnist_logdir_artifacts(tmp_path, monkeypatch):
  """Mirror of reference trainer_test.py:177-250 (artefacts in logdir)."""
  import shutil
  from lingvo_b200 import flags, trainer
  from lingvo_b200.models.image import input_generator
  data = input_generator.FakeMnistData(str(tmp_path), train_size=64, test_size=32)
  monkeypatch.setenv('LINGVO_B200_MNIST', data)
  logdir = str(tmp_path / 'log')
  flags.FLAGS.reset()
  argv = ['trainer', '--run_locally=cpu', '--mode=sync', '--model=image.mnist.LeNet5',
          '--logdir=' + logdir,
          '--model_params_override=task.train.max_steps:2;input.batch_size:8;'
          'task.train.summary_interval_steps:1']
  trainer.main(argv)
  for f in ['control/params.txt', 'control/model_analysis.txt', 'train/checkpoint',
            'train/ckpt-00000002.index', 'train/trainer_params.txt']:
    assert os.path.exists(os.path.join(logdir, f)), f
  assert glob.glob(os.path.join(logdir, 'train', 'events.out.tfevents.*'))
  # the controller followed the trainer: it wrote training summaries (metrics, grad/var
  # norms, total_num_params) for the final checkpoint into control/
  from lingvo_b200.utils import tfevents
  events = glob.glob(os.path.join(logdir, 'control', 'events.out.tfevents.*'))
  assert events
  tags = {}
  for path in events:
    for step, tag, val in tfevents.ReadScalars(path):
      tags.setdefault(tag, []).append((step, val))
  assert 'total_num_params' in tags and tags['total_num_params'][-1][1] > 1000
  assert tags['global_step'][-1] == (2, 2.0)
  assert 'log_pplx' in tags and 'grad_norm/all' in tags and tags['grad_norm/all'][-1][1] > 0
  flags.FLAGS.reset()
  trainer.main(['trainer', '--run_locally=cpu', '--mode=sync',
                '--model=image.mnist.LeNet5', '--logdir=' + logdir,
                '--job=evaler_test,decoder_test',
                '--model_params_override=task.train.max_steps:2;input.batch_size:8'])
  assert os.path.exists(os.path.join(logdir, 'eval_test', 'score-00000002.txt'))
  assert os.path.exists(os.path.join(logdir, 'eval_test', 'processed_ckpts.txt'))
  assert os.path.exists(os.path.join(logdir, 'decoder_test', 'decoder_out_000000002'))
  score = open(os.path.join(logdir, 'eval_test', 'score-00000002.txt')).read()
  assert 'accuracy:' in score and 'log_pplx:' in score
  flags.FLAGS.reset()


def test_executor_job_runs_program_schedule(tmp_path, monkeypatch):
  """`--job=executor_tpu` (reference call stack §3.4): train program → eval program per loop,
  checkpoints and per-program event files / scores in the logdir."""
  from lingvo_b200 import flags, trainer
  from lingvo_b200.models.image import input_generator
  data = input_generator.FakeMnistData(str(tmp_path), train_size=64, test_size=32)
  monkeypatch.setenv('LINGVO_B200_MNIST', data)
  logdir = str(tmp_path / 'log')
  flags.FLAGS.reset()
  trainer.main(['trainer', '--run_locally=cpu', '--mode=sync', '--model=image.mnist.LeNet5',
                '--job=executor_tpu', '--logdir=' + logdir,
                '--model_params_override=task.train.max_steps:6;input.batch_size:8'])
  flags.FLAGS.reset()
  for f in ['control/params.txt', 'train/checkpoint', 'train/ckpt-00000100.index',
            'EvalProgram_test/score-00000100.txt']:
    assert os.path.exists(os.path.join(logdir, f)), f
  assert glob.glob(os.path.join(logdir, 'TrainProgram_train', 'events.out.tfevents.*'))
  score = open(os.path.join(logdir, 'EvalProgram_test', 'score-00000100.txt')).read()
  assert 'accuracy' in score
