"""Tensor parallelism as a model feature of the GShard dense builder:
a 2-rank (gloo) tensor-parallel UniTransformer reproduces the single-process model —
loss, every gradient (shards gathered), three optimizer steps, incremental decoding — and
its checkpoint is an ordinary one that a single process restores."""

import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lingvo_b200.core import cluster_factory
from lingvo_b200.core import gshard_builder as gb
from lingvo_b200.core import optimizer
from lingvo_b200.core import schedule
from lingvo_b200.core import test_utils
from lingvo_b200.core.nested_map import NestedMap


def _Params(gated=False, bias=False):
  b = gb.DenseBuilder.Params().Set(
      model_dim=16, attention_num_heads=4, attention_key_value_dim=4, ff_dim=32,
      relative_attention_type='bias', relative_attention_num_buckets=8,
      relative_attention_max_distance=16, relative_attention_use_universal_1d_position=True,
      ff_use_bias=bias, device_mesh_shape=[1, 2])
  p = gb.UniTransformer.Params().Set(
      name='lm', builder=b, vocab_size=40, num_transformer_layers=2, max_length=16,
      positional_embedding=False, label_smoothing=0.0, z_loss=0.0, gated_gelu=gated,
      decoder_max_steps=3)
  p.random_seed = 1234
  p.train.learning_rate = 0.05
  p.train.lr_schedule = schedule.Constant.Params()
  p.train.optimizer = optimizer.Adam.Params().Set(beta1=0.9, beta2=0.99, epsilon=1e-6)
  p.train.clip_gradient_norm_to_value = 1.0
  return p


def _Batch():
  g = torch.Generator().manual_seed(5)
  ids = torch.randint(2, 40, (3, 8), generator=g)
  seg = torch.ones(3, 8, dtype=torch.long)
  seg[1, 5:] = 2
  pos = torch.arange(8).repeat(3, 1)
  pos[1, 5:] = torch.arange(3)
  return NestedMap(ids=ids, labels=torch.roll(ids, -1, 1), paddings=torch.zeros(3, 8),
                   segment_ids=seg, segment_pos=pos)


def _Run(task, steps):
  """loss + grads of step 0 (logical/gathered), then `steps` train steps → losses."""
  from lingvo_b200.parallel import mesh as mesh_lib
  from lingvo_b200.parallel import tp_layers
  ctx = mesh_lib.TensorParallel()
  batch = _Batch()
  metrics, _ = task.FPropDefaultTheta(batch)
  loss0 = metrics['loss'][0]
  loss0.backward()
  grads = {}
  for v in task.vars.Flatten():
    g = v.grad.detach().clone()
    shard = getattr(v, 'tp_shard', None)
    if shard is not None:
      g = tp_layers.GatherShards(g, ctx, shard[2])
    grads[v.var_name] = g
    v.grad = None
  task._metrics = None   # pylint: disable=protected-access
  losses = []
  for _ in range(steps):
    m, _ = task.TrainStep(batch)
    losses.append(float(m['loss'][0]))
  return float(loss0), grads, losses


def _Worker(rank, world, port, gated, bias, logdir, q):
  import faulthandler
  faulthandler.dump_traceback_later(200, exit=True)
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                    WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  from lingvo_b200.core import checkpointer
  from lingvo_b200.core import base_model
  from lingvo_b200.parallel import dp
  from lingvo_b200.parallel import mesh as mesh_lib
  with cluster_factory.ForTestingWorker(mode='sync', job='trainer_client'):
    model = base_model.SingleTaskModel.Params(_Params(gated, bias)).Instantiate()
    task = model.GetTask()
    ctx = mesh_lib.TensorParallel()
    assert ctx is not None and ctx.tp_size == 2 and ctx.dp_size == 1
    dp.Attach(task)
    sharded = {v.var_name: tuple(v.shape) for v in task.vars.Flatten()
               if getattr(v, 'tp_shard', None) is not None}
    loss0, grads, losses = _Run(task, 3)
    ck = checkpointer.Checkpointer(logdir, model)
    ck.Save(gsteps=3, sync=True)
    ck.Sync()
    # incremental decode under TP (local heads + all-reduced projection)
    b = _Batch()
    with torch.no_grad():
      pred = task.ComputePredictions(task.theta, b)
      full = task._ComputeLogits(task.theta, pred.dec_outputs).float()   # pylint: disable=protected-access
      st = task.InitDecodeState(3, 10, b.ids.device)
      step0 = task.DecodeStep(task.theta, b.ids[:, 0], st, 0)
    q.put(test_utils.ToNumpyTree((rank, loss0, grads, losses, sharded,
                                  {v.var_name: v.detach() for v in task.vars.Flatten()},
                                  full[:, 0], step0)))
  dist.barrier()
  dist.destroy_process_group()


@pytest.mark.parametrize('gated,bias', [(False, False), (True, True)])
def test_two_rank_tensor_parallel_matches_single_process(tmp_path, gated, bias):
  world = 2
  logdir = str(tmp_path / 'train')
  os.makedirs(logdir)
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = test_utils.FreePort()
  procs = [ctx.Process(target=_Worker, args=(r, world, port, gated, bias, logdir, q))
           for r in range(world)]
  for pr in procs:
    pr.start()
  res = {r[0]: r for r in [test_utils.ToTorchTree(q.get(timeout=120)) for _ in range(world)]}
  for pr in procs:
    pr.join(timeout=60)

  # single-process oracle (no process group → no TP; same name-seeded logical weights)
  from lingvo_b200.core import base_model
  from lingvo_b200.core import checkpointer
  with cluster_factory.ForTestingWorker(mode='sync', job='trainer_client'):
    model = base_model.SingleTaskModel.Params(_Params(gated, bias)).Instantiate()
    task = model.GetTask()
    loss0, grads, losses = _Run(task, 3)
    final = {v.var_name: v.detach().clone() for v in task.vars.Flatten()}

  for r in range(world):
    _, l0, g, ls, sharded, weights, full0, step0 = res[r]
    assert l0 == pytest.approx(loss0, rel=1e-5)
    np.testing.assert_allclose(ls, losses, rtol=2e-4)
    assert set(g) == set(grads)
    for name, want in grads.items():
      torch.testing.assert_close(g[name], want, atol=2e-5, rtol=2e-4, msg=name)
    # attention heads and the FFN hidden dim are really sharded (half-size weights)
    assert any('wq' in n for n in sharded) and any('/wi' in n or 'wi_0' in n for n in sharded)
    for name, shape in sharded.items():
      assert int(np.prod(shape)) * 2 == final[name].numel(), name
    torch.testing.assert_close(step0, full0, atol=1e-4, rtol=1e-4)
  assert losses[-1] < losses[0]

  # the TP job wrote an ordinary (logical-shape) checkpoint: restore it without TP
  with cluster_factory.ForTestingWorker(mode='sync', job='trainer_client'):
    fresh = base_model.SingleTaskModel.Params(_Params(gated, bias)).Instantiate()
    ck = checkpointer.Checkpointer(logdir, fresh)
    assert ck.Restore() and fresh.GetTask().global_step == 3
    for v in fresh.GetTask().vars.Flatten():
      torch.testing.assert_close(v.detach(), final[v.var_name], atol=2e-4, rtol=2e-3,
                                 msg=v.var_name)


def test_mesh_topology_helpers():
  from lingvo_b200.parallel import mesh as mesh_lib
  ctx = mesh_lib.Reset()
  assert ctx.tp_size == 1 and ctx.tp_rank == 0 and ctx.dp_size == 1
  assert mesh_lib.TensorParallel() is None
  assert mesh_lib.ConfigureFromMeshShape([1, 2]) is ctx          # no process group: ignored


# ----------------------------------------------------------------- context parallelism --
def _CpParams(cp):
  p = _Params()
  p.builder.device_mesh_shape = None
  p.builder.mhd_w_split = [-1, -1, -1]
  p.builder.mh_wi_split = [-1, -1]
  p.builder.hm_wo_split = [-1, -1]
  p.builder.context_parallel = cp
  return p


def _CpWorker(rank, world, port, q):
  import faulthandler
  faulthandler.dump_traceback_later(200, exit=True)
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                    WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  from lingvo_b200.core import base_model
  from lingvo_b200.parallel import dp
  from lingvo_b200.parallel import mesh as mesh_lib
  mesh_lib.Reset(mode='nccl')
  with cluster_factory.ForTestingWorker(mode='sync', job='trainer_client'):
    task = base_model.SingleTaskModel.Params(_CpParams(True)).Instantiate().GetTask()
    dp.Attach(task)
    batch = _Batch()                                   # every rank sees the full batch …
    probe = task._ComputeInputBatch(_Batch())          # pylint: disable=protected-access
    assert probe.tgt.ids.shape == (3, 4)               # … and keeps its half of each sequence
    metrics, _ = task.FPropDefaultTheta(batch)
    loss = metrics['loss'][0]
    loss.backward()
    # what DP sync would apply: mean over ranks of the local gradients
    grads = {}
    for v in task.vars.Flatten():
      g = v.grad.detach().clone()
      dist.all_reduce(g)
      grads[v.var_name] = g / world
    lt = loss.detach().clone()
    dist.all_reduce(lt)
    q.put(test_utils.ToNumpyTree((rank, float(lt) / world, grads)))
  dist.barrier()
  dist.destroy_process_group()


def test_two_rank_context_parallel_matches_single_process():
  """Sequence dimension sharded over 2 ranks (K/V all-gather, dK/dV reduce-scatter, global
  causal + relative bias + packed segments): mean-of-shard losses and DP-averaged gradients
  equal the unsharded model's."""
  world = 2
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = test_utils.FreePort()
  procs = [ctx.Process(target=_CpWorker, args=(r, world, port, q)) for r in range(world)]
  for pr in procs:
    pr.start()
  res = {r[0]: r for r in [test_utils.ToTorchTree(q.get(timeout=60)) for _ in range(world)]}
  for pr in procs:
    pr.join(timeout=60)
  from lingvo_b200.core import base_model
  with cluster_factory.ForTestingWorker(mode='sync', job='trainer_client'):
    task = base_model.SingleTaskModel.Params(_CpParams(False)).Instantiate().GetTask()
    batch = _Batch()
    metrics, _ = task.FPropDefaultTheta(batch)
    # the CP objective is the mean of the two half-sequence means: build it explicitly
    from lingvo_b200.core.nested_map import NestedMap as NM
    loss = metrics['loss'][0]
    loss.backward()
    want = {v.var_name: v.grad.detach().clone() for v in task.vars.Flatten()}
  for r in range(world):
    _, l, g = res[r]
    assert l == pytest.approx(float(loss), rel=1e-5)
    for name, w in want.items():
      torch.testing.assert_close(g[name], w, atol=2e-5, rtol=2e-4, msg=name)
  del NM
