"""Asynchronous data parallelism (SURVEY row "DP async / parameter server"): local steps
with delayed non-blocking parameter averaging, 2 ranks over gloo."""

import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lingvo_b200.core import test_utils


def _Task(async_mode, sync_every=1):
  from lingvo_b200.core import base_input_generator
  from lingvo_b200.core import base_model
  from lingvo_b200.core import layers
  from lingvo_b200.core import optimizer
  from lingvo_b200.core import schedule
  from lingvo_b200.core.nested_map import NestedMap

  class Inp(base_input_generator.BaseInputGenerator):
    """y = x·w* + noise; every rank sees its own stream."""

    def __init__(self, params):
      super().__init__(params)
      self._g = torch.Generator().manual_seed(100 + int(os.environ.get('RANK', '0')))
      self._w = torch.linspace(-1, 1, 6).reshape(6, 1)

    def _InputBatch(self):
      x = torch.randn(16, 6, generator=self._g)
      return NestedMap(x=x, y=x @ self._w)

  class Reg(base_model.BaseTask):

    def __init__(self, params):
      super().__init__(params)
      self.CreateChild('fc', layers.FCLayer.Params().Set(input_dim=6, output_dim=1,
                                                        activation='NONE'))

    def ComputePredictions(self, theta, batch):
      return self.fc.FProp(theta.fc, batch.x)

    def ComputeLoss(self, theta, pred, batch):
      loss = (pred - batch.y).square().mean()
      return {'loss': (loss, 16.0)}, {}

  p = Reg.Params().Set(name='reg', input=Inp.Params().Set(batch_size=16))
  p.random_seed = 7
  p.train.learning_rate = 0.05
  p.train.lr_schedule = schedule.Constant.Params()
  p.train.optimizer = optimizer.SGD.Params()
  p.train.async_data_parallel = async_mode
  p.train.async_sync_every_n_steps = sync_every
  return p


def _Worker(rank, world, port, sync_every, q):
  import faulthandler
  faulthandler.dump_traceback_later(120, exit=True)
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                    WORLD_SIZE=str(world))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  from lingvo_b200.core import cluster_factory
  from lingvo_b200.core import train_engine
  with cluster_factory.ForTestingWorker(mode='async', job='trainer_client'):
    task = _Task(True, sync_every).Instantiate()
    eng = train_engine.TrainEngine(task)
    assert eng.async_dp is not None and eng.dp is None
    assert all(l.grad_sync is None for l in task.learners)
    flat = lambda: torch.cat([v.detach().reshape(-1) for v in task.vars.Flatten()]).clone()
    start = flat()
    losses, mid = [], None
    for i in range(40):
      m, _ = eng.Step()
      losses.append(float(m['loss'][0]))
      if i == 20:
        mid = flat()
    n_rec = eng.async_dp.num_reconciliations
    eng.PreSave()                                  # Finalize: replicas agree
    q.put(test_utils.ToNumpyTree((rank, start, mid, flat(), losses, n_rec)))
  dist.barrier()
  dist.destroy_process_group()


def _Launch(sync_every):
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = test_utils.FreePort()
  procs = [ctx.Process(target=_Worker, args=(r, 2, port, sync_every, q)) for r in range(2)]
  for pr in procs:
    pr.start()
  res = {r[0]: r for r in [test_utils.ToTorchTree(q.get(timeout=120)) for _ in range(2)]}
  for pr in procs:
    pr.join(timeout=60)
  return res


def test_async_data_parallel_replicas_drift_then_reconcile():
  res = _Launch(sync_every=4)
  (_, s0, m0, f0, l0, n0), (_, s1, m1, f1, l1, n1) = res[0], res[1]
  torch.testing.assert_close(s0, s1)                       # same starting point
  assert not torch.allclose(m0, m1, atol=1e-6)             # replicas step independently …
  assert float((m0 - m1).abs().max()) < 0.5                # … but stay within the staleness bound
  torch.testing.assert_close(f0, f1, atol=1e-6, rtol=1e-6)  # Finalize ⇒ identical
  assert n0 == n1 == 40 // 4 - 1                           # one reconciliation per 4 steps, delayed
  for losses in (l0, l1):
    assert losses[-1] < 0.05 * losses[0]                   # both learn the shared target
  w_star = torch.linspace(-1, 1, 6)
  assert float((f0[1:] - w_star).abs().max()) < 0.1 and abs(float(f0[0])) < 0.1   # [b, w]


def test_async_correction_is_the_delayed_average():
  """sync_every=1: after each step p_r ← p_r + mean(prev snapshots) − own prev snapshot."""
  res = _Launch(sync_every=1)
  assert res[0][5] == res[1][5] == 39
  torch.testing.assert_close(res[0][3], res[1][3], atol=1e-6, rtol=1e-6)


def test_single_process_ignores_async_flag():
  from lingvo_b200.core import cluster_factory
  from lingvo_b200.core import train_engine
  with cluster_factory.ForTestingWorker(mode='async', job='trainer_client'):
    task = _Task(True).Instantiate()
    eng = train_engine.TrainEngine(task)
    assert eng.async_dp is None
    m, _ = eng.Step()
    assert np.isfinite(float(m['loss'][0]))
