"""Row-sharded sparse embedding tables across ranks + the v1 collection / v2 manager APIs
(ref `tpu_embedding_layers{,_v1,_v2}.py`, `tpu_embedding_manager.py`)."""

import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lingvo_b200.core import test_utils
from lingvo_b200.core.nested_map import NestedMap


def _LayerParams(mod, lr=0.5):
  table = mod.TPUEmbeddingTable.Params().Set(
      name='t', vocab_size=21, embedding_dim=4, input_keys=['a', 'b'], combiner='sum')
  p = mod.TPUEmbeddingLayer.Params().Set(
      name='emb', tables=[table], learning_rate=lr,
      optimizer=mod.TPUEmbeddingSGDOptimizer.Params())
  p.random_seed = 11
  return p


def _Ids(rank):
  g = torch.Generator().manual_seed(50 + rank)
  a = torch.randint(0, 21, (3, 4), generator=g)
  a[0, 3] = -1
  b = torch.randint(0, 21, (3, 2), generator=g)
  return NestedMap(a=a, b=b)


def _Worker(rank, world, port, q):
  import faulthandler
  faulthandler.dump_traceback_later(120, exit=True)
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                    WORLD_SIZE=str(world))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  from lingvo_b200.core import tpu_embedding_layers as tel
  layer = _LayerParams(tel).Instantiate()
  layer.InstantiateVariables()
  t = layer.tables[0]
  assert t._local_rows == 11                        # ceil(21 / 2) rows per rank
  # make the logical table = arange so every row is recognisable: row r lives on r % 2
  with torch.no_grad():
    local = torch.arange(t._local_rows).float() * world + rank
    t.table.copy_(local.unsqueeze(1).expand(-1, 4) * 0.1)
  ids = _Ids(rank)
  out = layer.EmbLookup(layer.theta, ids)
  loss = (out.a * (rank + 1)).sum() + out.b.sum() * 2.0
  loss.backward()
  layer.ApplyGradients(global_step=0)
  q.put(test_utils.ToNumpyTree((rank, out.a.detach(), out.b.detach(), t.table.detach().clone())))
  dist.barrier()
  dist.destroy_process_group()


def test_row_sharded_lookup_and_sparse_update_across_two_ranks():
  world = 2
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = test_utils.FreePort()
  procs = [ctx.Process(target=_Worker, args=(r, world, port, q)) for r in range(world)]
  for pr in procs:
    pr.start()
  res = {r[0]: r for r in [test_utils.ToTorchTree(q.get(timeout=120)) for _ in range(world)]}
  for pr in procs:
    pr.join(timeout=60)
  # oracle: one logical table [22, 4] with row r = 0.1·r, SGD lr .5 on Σ of both ranks' losses
  logical = (torch.arange(22).float() * 0.1).unsqueeze(1).expand(-1, 4).clone()
  grad = torch.zeros_like(logical)
  for rank in range(world):
    ids = _Ids(rank)
    for key, scale in (('a', float(rank + 1)), ('b', 2.0)):
      v = ids[key]
      want = torch.where((v >= 0).unsqueeze(-1), logical[v.clamp_min(0)],
                         torch.zeros(1)).sum(1)
      torch.testing.assert_close(res[rank][1 if key == 'a' else 2], want)   # fetched rows
      for r in v[v >= 0].flatten().tolist():
        grad[r] += scale
  after = logical - 0.5 * grad
  for rank in range(world):
    table = res[rank][3]
    rows = torch.arange(table.shape[0]) * world + rank
    torch.testing.assert_close(table, after[rows.clamp(max=21)][:table.shape[0]],
                               atol=1e-6, rtol=1e-6)


def test_v1_collection_protocol():
  from lingvo_b200.core import schedule
  from lingvo_b200.core import tpu_embedding_layers_v1 as v1
  v1.TpuEmbeddingCollection.Reset()
  p = _LayerParams(v1)
  p.gradient_multiplier_schedule = schedule.Constant.Params().Set(value=0.5)
  layer = p.Instantiate()
  layer.InstantiateVariables()
  coll = v1.TpuEmbeddingCollection.Get()
  assert coll is v1.TpuEmbeddingCollection.Get() and coll.layers == [layer]
  assert coll.feature_names == frozenset({'a', 'b'})
  assert list(coll.table_variables.keys()) == ['t']
  before = layer.tables[0].table.detach().clone()
  ids = NestedMap(a=torch.tensor([[1, 2]]), b=torch.tensor([[3, -1]]))
  coll.SetTaskMode('eval_task', 'eval')
  coll.SetTaskMode('train_task', 'train')
  assert coll.ShouldStopGradient('eval_task') and not coll.ShouldStopGradient('train_task')
  out = layer.EmbLookup(layer.theta, ids, task_call_scope='eval_task')
  assert coll.GetActivations('eval_task') is out
  (out.a.sum() + out.b.sum()).backward()
  assert coll.ApplyGradients('eval_task') == 0                      # gradients are dropped
  torch.testing.assert_close(layer.tables[0].table.detach(), before)
  out = layer.EmbLookup(layer.theta, ids, task_call_scope='train_task')
  (out.a.sum() + out.b.sum()).backward()
  assert coll.ApplyGradients('train_task', global_step=3) == 1
  after = layer.tables[0].table.detach()
  # lr .5 × multiplier .5 × grad 1
  torch.testing.assert_close(before[1] - after[1], torch.full((4,), 0.25))
  assert torch.equal(after[4], before[4])
  coll.AddSummaryTensor('x', torch.tensor(1.0))
  assert coll.summary_tensors[0][0] == 'x'
  import pytest
  with pytest.raises(ValueError):
    coll.SetGradientMultiplierSchedule(object())
  with pytest.raises(ValueError):
    coll.ShouldStopGradient('unknown')
  v1.TpuEmbeddingCollection.Reset()


def test_v2_layers_register_with_the_manager():
  from lingvo_b200.core import tpu_embedding_layers_v2 as v2
  from lingvo_b200.core import tpu_embedding_manager as mgr_lib
  mgr = mgr_lib.Default()
  mgr.Reset()
  layer = _LayerParams(v2, lr=1.0).Instantiate()
  layer.InstantiateVariables()
  assert mgr.enabled and mgr.layers == [layer]
  cfg = layer.table_configs[0]
  assert cfg.vocabulary_size == 21 and cfg.dim == 4 and cfg.features == ['a', 'b']
  assert layer.tables[0].OwnerOf(13) == 0                           # single process
  d = mgr.Describe()[0]
  assert d.table == 't' and d.local_rows == 21
  before = layer.tables[0].table.detach().clone()
  out = layer.EmbLookup(layer.theta, NestedMap(a=torch.tensor([[5]]), b=torch.tensor([[-1]])))
  out.a.sum().backward()
  mgr.ApplyGradients(global_step=0)
  assert mgr.steps_applied == 1
  torch.testing.assert_close(before[5] - layer.tables[0].table.detach()[5], torch.ones(4))
  fn = layer.tables[0].optimizer.CreateOptimizerFn() if hasattr(
      layer.tables[0].optimizer, 'CreateOptimizerFn') else None
  opt = v2.TPUEmbeddingSGDOptimizer.Params().Instantiate()
  table = torch.zeros(3, 2)
  opt.CreateOptimizerFn()(table, {}, torch.tensor([1]), torch.ones(1, 2), 0.1)
  torch.testing.assert_close(table[1], torch.full((2,), -0.1))
  del fn
  mgr.Reset()
