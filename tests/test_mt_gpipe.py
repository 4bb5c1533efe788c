"""BASELINE config #5: Transformer MT as a GPipe pipeline (`models/mt/model.py
GPipeTransformerModel` over `layers_with_gpipe.GPipeTransformerStack`), single process and
one rank per cell through `parallel.pp.PipelineEngine`."""

import os

import pytest
import torch

from lingvo_b200.core import test_utils
import torch.distributed as dist
import torch.multiprocessing as mp

from lingvo_b200.core.nested_map import NestedMap


def _Task():
  from lingvo_b200 import model_registry
  from lingvo_b200.core import optimizer
  from lingvo_b200.core import schedule
  import lingvo_b200.models.mt.params.wmt14_en_de  # noqa: F401
  cfg = model_registry.GetParams('mt.wmt14_en_de.WmtEnDeTransformerGPipeTiny', 'Train')
  tp = cfg.task
  tp.input = None
  tp.random_seed = 11                       # name-seeded init: identical in every process
  tp.label_smoothing = 0.0
  tp.stack.emb_tpl.input_dropout_prob = 0.0
  for tpl in (tp.stack.encoder_tpl, tp.stack.decoder_tpl):
    tpl.tr_atten_tpl.residual_dropout_prob = 0.0
    tpl.tr_fflayer_tpl.residual_dropout_prob = 0.0
  tp.train.optimizer = optimizer.Adam.Params()
  tp.train.learning_rate = 3e-3
  tp.train.lr_schedule = schedule.Constant.Params()
  return tp.Instantiate()


def _Batch(step):
  g = torch.Generator().manual_seed(step)
  b, t = 4, 6
  ids = torch.randint(1, 64, (b, t), generator=g)
  pad = torch.zeros(b, t)
  pad[1, 4:] = 1.0
  return NestedMap(src=NestedMap(ids=ids, paddings=pad),
                   tgt=NestedMap(ids=torch.roll(ids, 1, 1), labels=ids, paddings=pad,
                                 weights=1 - pad))


def _Train(task, steps=4):
  losses = []
  for s in range(steps):
    task.FPropDefaultTheta(_Batch(s % 2))
    task.BProp()
    losses.append(float(task._eval_metrics['loss'][0]))
  return losses


def test_gpipe_mt_task_single_process_learns():
  task = _Task()
  assert task.stack.num_stages == 2 and task.engine is None
  losses = _Train(task, 12)
  assert losses[-1] < 0.8 * losses[0], losses


def _Worker(rank, world, port, q):
  import faulthandler
  faulthandler.dump_traceback_later(150, exit=True)
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                    WORLD_SIZE=str(world))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  task = _Task()
  assert task.engine is not None and task.engine.remat
  losses = _Train(task, 4)
  own = {v.var_name: v.data.clone().numpy() for v in task.vars.Flatten()
         if ('cell_%d' % rank) in v.var_name}
  q.put((rank, losses, own))
  dist.barrier()
  dist.destroy_process_group()


def test_gpipe_mt_task_two_ranks_match_single_process():
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = test_utils.FreePort()
  procs = [ctx.Process(target=_Worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  res = {r[0]: r for r in [q.get(timeout=240) for _ in range(2)]}
  for p in procs:
    p.join(timeout=60)
  ref = _Task()
  want = _Train(ref, 4)
  # every rank reports the last stage's loss; pipeline ≡ single process step by step
  for r in range(2):
    assert res[r][1] == pytest.approx(want, rel=2e-4, abs=2e-5), (res[r][1], want)
  ref_vars = {v.var_name: v.data for v in ref.vars.Flatten()}
  checked = 0
  for r in range(2):
    for name, val in res[r][2].items():
      torch.testing.assert_close(torch.from_numpy(val), ref_vars[name], atol=2e-5, rtol=2e-4)
      checked += 1
  assert checked == len(ref_vars)           # each variable trained on exactly one rank
