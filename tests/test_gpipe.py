"""GPipe layers: partitioning, micro-batch equivalence, 2-rank pipeline engine."""

import os
import sys

import pytest
import torch

from lingvo_b200.core import test_utils
import torch.distributed as dist
import torch.multiprocessing as mp

from lingvo_b200.core import gpipe
from lingvo_b200.core import layers
from lingvo_b200.core import py_utils
from lingvo_b200.core import tshape


def _Fc(name, i, o):
  return layers.FCLayer.Params().Set(name=name, input_dim=i, output_dim=o,
                                     activation='TANH')


def _Pipe(num_micro, cells=2):
  subs = [_Fc('fc%d' % i, 8, 8) for i in range(4)]
  per = len(subs) // cells
  cell_tpl = [gpipe.FeatureExtractionLayer.Params().Set(
      name='cell_%d' % c, sub=subs[c * per:(c + 1) * per]) for c in range(cells)]
  p = gpipe.PipeliningLayer.Params().Set(name='pipe', cell_tpl=cell_tpl,
                                         num_micro_batches=num_micro)
  p.params_init = py_utils.WeightInit.Uniform(0.5)
  p.random_seed = 1234
  return p


def test_partition_sequential_layers_balances_flops():
  subs = [_Fc('a', 8, 8), _Fc('b', 8, 32), _Fc('c', 32, 32), _Fc('d', 32, 8)]
  parts = gpipe.PartitionSequentialLayers(subs, 2, tshape.Shape([4, 8]))
  assert len(parts) == 2
  assert sum(len(p.sub) for p in parts) == 4
  assert len(parts[0].sub) >= 1 and len(parts[1].sub) >= 1


def test_micro_batching_is_equivalent():
  x = torch.randn(8, 8)
  a = _Pipe(1).Instantiate()
  b = _Pipe(4).Instantiate()
  ya = a.FPropDefaultTheta(x)
  yb = b.FProp(a.theta, x)
  torch.testing.assert_close(ya, yb)
  ya.sum().backward()
  assert all(v.grad is not None for v in a.vars.Flatten())


def test_feature_extraction_forwarding():
  p = gpipe.FeatureExtractionLayer.Params().Set(
      name='fe', sub=[_Fc('x', 8, 8)], num_act_inputs=1, num_act_outputs=1)
  l = p.Instantiate()
  x, extra = torch.randn(2, 8), torch.randn(2, 3)
  out = l.FPropDefaultTheta(x, extra)
  assert len(out) == 2 and out[1] is extra


def _Worker(rank, world, port, q, remat=False):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  from lingvo_b200.parallel import pp
  torch.manual_seed(0)
  x = torch.randn(8, 8)
  layer = _Pipe(4).Instantiate()          # same seed → same weights on both ranks
  eng = pp.PipelineEngine(remat=remat)
  layer.AttachEngine(eng)
  out = layer.FProp(layer.theta, x)
  loss = out.pow(2).sum() if eng.is_last else None
  eng.Backward(loss)
  grads = {v.var_name: (v.grad.clone() if v.grad is not None else None)
           for v in layer.vars.Flatten()}
  q.put(test_utils.ToNumpyTree((rank, float(loss) if loss is not None else None, grads)))
  dist.barrier()
  dist.destroy_process_group()


import pytest  # noqa: E402


@pytest.mark.parametrize('remat', [False, True])
def test_pipeline_engine_two_ranks_matches_single_process(remat):
  """GPipe over 2 ranks (overlapped isend/irecv links, tensor-header handshake) — with and
  without rematerialisation — reproduces the single-process loss and gradients."""
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = test_utils.FreePort()
  procs = [ctx.Process(target=_Worker, args=(r, 2, port, q, remat)) for r in range(2)]
  for p in procs:
    p.start()
  res = [test_utils.ToTorchTree(q.get(timeout=120)) for _ in range(2)]
  for p in procs:
    p.join(timeout=60)
  res = {r[0]: r for r in res}
  # single-process oracle
  torch.manual_seed(0)
  x = torch.randn(8, 8)
  ref = _Pipe(4).Instantiate()
  loss = ref.FPropDefaultTheta(x).pow(2).sum()
  loss.backward()
  assert abs(res[1][1] - float(loss)) < 1e-4
  for v in ref.vars.Flatten():
    stage = 0 if 'cell_0' in v.var_name else 1
    g = res[stage][2][v.var_name]
    assert g is not None, v.var_name
    torch.testing.assert_close(g, v.grad, atol=1e-5, rtol=1e-4)
    other = res[1 - stage][2][v.var_name]
    assert other is None        # a stage never touches the other stage's weights


def _EpNormWorker(rank, world, port, q):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  from lingvo_b200.core import learner as learner_lib
  from lingvo_b200.core import py_utils
  from lingvo_b200.core.nested_map import NestedMap
  from lingvo_b200.parallel import mesh
  mesh.Reset()
  assert mesh.ExpertParallelFor(2) is not None
  lrn = learner_lib.Learner.Params().Set(name='l', learning_rate=0.1).Instantiate()
  shared = torch.nn.Parameter(torch.ones(4))
  expert = torch.nn.Parameter(torch.ones(3))
  expert.expert_parallel = True
  vgs = NestedMap(
      a=py_utils.VarGrad(shared, torch.full((4,), 2.0)),
      e=py_utils.VarGrad(expert, torch.full((3,), float(rank + 1))))
  out = lrn.ScaleGradients(vgs)
  q.put((rank, float(out.stats['grad_norm/all'][0])))
  dist.barrier()
  dist.destroy_process_group()


def test_global_grad_norm_sums_expert_parallel_grads_over_ranks():
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = test_utils.FreePort()
  procs = [ctx.Process(target=_EpNormWorker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  res = dict(q.get(timeout=60) for _ in range(2))
  for p in procs:
    p.join(timeout=60)
  want = (4 * 4.0 + 3 * 1.0 + 3 * 4.0) ** 0.5
  assert abs(res[0] - want) < 1e-5 and abs(res[1] - want) < 1e-5, res


def _CpWorker(rank, world, port, q):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  from lingvo_b200.parallel import cp
  torch.manual_seed(0)
  b, l, h, d = 2, 16, 2, 8
  full = [torch.randn(b, l, h, d, requires_grad=True) for _ in range(3)]
  seg = torch.tensor([[1] * 6 + [2] * 7 + [0] * 3, [1] * 16])
  bias = torch.randn(h, l, l)
  loc = [cp.ShardSequence(t.detach(), 1).requires_grad_(True) for t in full]
  lq = l // world
  out = cp.Attention(*loc, causal=True, segment_ids=cp.ShardSequence(seg, 1),
                     bias=bias[:, rank * lq:(rank + 1) * lq])
  dy = torch.randn(b, l, h, d)
  valid = cp.ShardSequence((seg != 0).reshape(b, l, 1, 1).float(), 1)
  (out * valid).backward(cp.ShardSequence(dy, 1))
  q.put(test_utils.ToNumpyTree((rank, out.detach(), [t.grad.clone() for t in loc])))
  dist.barrier()
  dist.destroy_process_group()


def test_context_parallel_attention_matches_single_device():
  from lingvo_b200.parallel import cp
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = test_utils.FreePort()
  procs = [ctx.Process(target=_CpWorker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  res = {r[0]: r for r in (test_utils.ToTorchTree(q.get(timeout=120)) for _ in range(2))}
  for p in procs:
    p.join(timeout=60)
  torch.manual_seed(0)
  b, l, h, d = 2, 16, 2, 8
  full = [torch.randn(b, l, h, d, requires_grad=True) for _ in range(3)]
  seg = torch.tensor([[1] * 6 + [2] * 7 + [0] * 3, [1] * 16])
  bias = torch.randn(h, l, l)
  ref = cp.AttentionRef(*full, causal=True, segment_ids=seg, bias=bias)
  dy = torch.randn(b, l, h, d)
  # padding rows (segment 0) attend to nothing: their outputs are arbitrary, exclude them
  valid = (seg != 0).reshape(b, l, 1, 1).float()
  (ref * valid).backward(dy)
  out = torch.cat([res[0][1], res[1][1]], 1)
  torch.testing.assert_close(out * valid, ref.detach() * valid, atol=1e-4, rtol=1e-4)
  # dq stays local; dk / dv come back through the reduce-scatter
  for i in range(3):
    got = torch.cat([res[0][2][i], res[1][2][i]], 1)
    torch.testing.assert_close(got, full[i].grad, atol=1e-4, rtol=1e-3)


def test_gpipe_transformer_stacks_split_invariance():
  """The GPipe transformer stacks give the same output for 1 and 2 pipeline cells and for
  1 and 2 micro-batches (same seeds → same weights)."""
  from lingvo_b200.core import layers_with_gpipe as lg

  def Build(splits, micro):
    torch.manual_seed(0)
    p = lg.GPipeTransformerStack.Params().Set(
        name='stack', model_dim=16, num_encoder_layers=2, num_decoder_layers=0, num_splits=splits,
        splits=splits, num_micro_batches=micro, random_seed=123)
    p.encoder_tpl.tr_atten_tpl.num_attention_heads = 2
    p.encoder_tpl.tr_fflayer_tpl.hidden_dim = 32
    return p.Instantiate()
  x = torch.randn(5, 4, 16)                      # [time, batch, dim]
  pad = torch.zeros(5, 4); pad[3:, 1] = 1.0
  import re

  def Key(v):                                   # variable identity independent of the cell split
    return re.sub(r'cell_\d+/', '', v.var_name)

  def CopyWeights(dst, src):
    table = {Key(v): v for v in src.vars.Flatten()}
    for v in dst.vars.Flatten():
      v.data.copy_(table[Key(v)].data)
  outs = []
  ref_layer = None
  for splits, micro in [(1, 1), (2, 1), (2, 2)]:
    layer = Build(splits, micro)
    if ref_layer is None:
      ref_layer = layer
    else:
      CopyWeights(layer, ref_layer)
    outs.append(layer.FPropDefaultTheta(x, pad))
    assert outs[-1].shape == x.shape
  torch.testing.assert_close(outs[0], outs[1], atol=1e-5, rtol=1e-5)
  torch.testing.assert_close(outs[0], outs[2], atol=1e-5, rtol=1e-5)

  def BuildBm(splits):
    torch.manual_seed(0)
    p = lg.GPipeBatchMajorTransformerStack.Params().Set(
        name='bm', model_dim=16, num_encoder_layers=2, num_splits=splits, random_seed=7)
    p.encoder_tpl.tr_atten_tpl.num_heads = 2
    p.encoder_tpl.tr_fflayer_tpl.hidden_dim = 32
    return p.Instantiate()
  xb = torch.randn(4, 5, 16)                     # [batch, time, dim]
  pb = torch.zeros(4, 5)
  one, two = BuildBm(1), BuildBm(2)
  CopyWeights(two, one)
  torch.testing.assert_close(one.FPropDefaultTheta(xb, pb), two.FPropDefaultTheta(xb, pb),
                             atol=1e-5, rtol=1e-5)
