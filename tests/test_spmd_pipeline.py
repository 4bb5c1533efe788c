"""SPMD shifting-buffer / circular pipeline (`gshard_layers.LayerwiseShardablePipelinedLayer`,
reference :180-1185) and `PipelinedTransformerLayers` (bma :7512)."""

import os

import pytest
import torch

from lingvo_b200.core import test_utils
import torch.distributed as dist
import torch.multiprocessing as mp

from lingvo_b200.core import gshard_layers as G
from lingvo_b200.core import layers
from lingvo_b200.core.nested_map import NestedMap


def _Pipe(stages, rep, nmb, dim=8):
  body = layers.FCLayer.Params().Set(name='fc', input_dim=dim, output_dim=dim,
                                     activation='TANH', random_seed=7)
  return G.LayerwiseShardablePipelinedLayer.Params().Set(
      name='pipe', num_stages=stages, circular_repeat=rep, single_stage_body=body,
      num_microbatches=nmb, random_seed=7)


@pytest.mark.parametrize('stages,rep,nmb', [(3, 1, 4), (4, 1, 2), (2, 2, 4), (3, 2, 6)])
def test_local_pipeline_equals_sequential_layers(stages, rep, nmb):
  torch.manual_seed(0)
  layer = _Pipe(stages, rep, nmb).Instantiate()
  x = torch.randn(12, 8, requires_grad=True)
  y = layer.FPropDefaultTheta(x)
  ref = x
  for k in range(stages * rep):
    ref = getattr(layer, 'body_%03d' % k).FPropDefaultTheta(ref)
  torch.testing.assert_close(y, ref)
  y.sum().backward()
  assert all(v.grad is not None for v in layer.vars.Flatten())
  # the schedule is a proper wave: stage s first works at iteration s
  sched = layer._Schedule(nmb)
  for s in range(stages):
    assert sched[s][s] == (0, 0)
    assert all(sched[t][s] is None for t in range(s))


def test_pipeline_carries_nested_side_inputs():
  class AddPad(layers.FCLayer):
    def FProp(self, theta, inp):
      out = super().FProp(theta, inp.vec) * (1.0 - inp.paddings).unsqueeze(-1)
      return NestedMap(vec=out, paddings=inp.paddings)
  body = AddPad.Params().Set(name='fc', input_dim=4, output_dim=4, activation='NONE')
  p = G.LayerwiseShardablePipelinedLayer.Params().Set(
      name='pipe', num_stages=2, single_stage_body=body, num_microbatches=2)
  layer = p.Instantiate()
  x = NestedMap(vec=torch.randn(4, 3, 4), paddings=(torch.rand(4, 3) > 0.5).float())
  y = layer.FPropDefaultTheta(x)
  assert y.vec.shape == (4, 3, 4)
  torch.testing.assert_close(y.paddings, x.paddings)
  assert float((y.vec * x.paddings.unsqueeze(-1)).abs().max()) == 0.0


def _Worker(rank, world, port, rep, nmb, q):
  import faulthandler
  faulthandler.dump_traceback_later(120, exit=True)
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  p = _Pipe(world, rep, nmb)
  layer = G.LayerwiseShardablePipelinedLayer.ForStageGroup(p)
  torch.manual_seed(0)
  x = torch.randn(12, 8, requires_grad=True)
  y = layer.FPropDefaultTheta(x)
  (y * torch.arange(8.0)).sum().backward()
  grads = {v.var_name: v.grad.clone() for v in layer.vars.Flatten()}
  q.put(test_utils.ToNumpyTree((rank, y.detach(), x.grad.clone() if rank == 0 else None, grads,
                                sorted(layer._owned))))
  dist.barrier()
  dist.destroy_process_group()


@pytest.mark.parametrize('rep,nmb', [(1, 4), (2, 4)])
def test_rank_sharded_pipeline_matches_local(rep, nmb):
  world = 2
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = test_utils.FreePort()
  procs = [ctx.Process(target=_Worker, args=(r, world, port, rep, nmb, q))
           for r in range(world)]
  for pr in procs:
    pr.start()
  res = {r[0]: r for r in [test_utils.ToTorchTree(q.get(timeout=180)) for _ in range(world)]}
  for pr in procs:
    pr.join(timeout=60)
  # oracle: everything in one process (same name-seeded weights)
  local = _Pipe(world, rep, nmb).Instantiate()
  torch.manual_seed(0)
  x = torch.randn(12, 8, requires_grad=True)
  y = local.FPropDefaultTheta(x)
  (y * torch.arange(8.0)).sum().backward()
  for r in range(world):
    torch.testing.assert_close(res[r][1], y.detach(), atol=1e-6, rtol=1e-5)
    assert res[r][4] == [k for k in range(world * rep) if k % world == r]   # weights 1/S each
  torch.testing.assert_close(res[0][2], x.grad, atol=1e-6, rtol=1e-5)
  want = {v.var_name: v.grad for v in local.vars.Flatten()}
  seen = 0
  for r in range(world):
    for name, g in res[r][3].items():
      torch.testing.assert_close(g, want[name], atol=1e-6, rtol=1e-5)
      seen += 1
  assert seen == len(want)


def test_pipelined_transformer_layers_equal_unpipelined_stack():
  from lingvo_b200.core import batch_major_attention as bma
  torch.manual_seed(0)
  stage = bma.StackedTransformerLayers.Params().Set(
      num_layers=1, mdl_dim=16, hidden_dim=32, num_atten_heads=2, random_seed=5)
  p = bma.PipelinedTransformerLayers.Params().Set(
      name='pl', pipeline_stage=stage, num_pipeline_stages=3, num_pipeline_microbatches=2,
      final_layer_norm=True, random_seed=5)
  layer = p.Instantiate()
  x = torch.randn(4, 6, 16, requires_grad=True)
  pad = torch.zeros(4, 6)
  pad[1, 4:] = 1.0
  y, out_pad = layer.FPropDefaultTheta(x, pad)
  assert y.shape == x.shape
  torch.testing.assert_close(out_pad, pad)
  # same weights, no pipelining: run the three stage bodies back to back on the full batch
  ref = x
  for k in range(3):
    body = getattr(layer.pipeline, 'body_%03d' % k)
    ref, _ = body.stage.FPropDefaultTheta(ref, pad)
  ref = layer.final_ln.FPropDefaultTheta(ref)
  torch.testing.assert_close(y, ref, atol=1e-5, rtol=1e-4)
  y.sum().backward()
  assert x.grad is not None
