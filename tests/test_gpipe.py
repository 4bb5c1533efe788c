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


def _SmallStack(cls=None, **kw):
  from lingvo_b200.core import layers_with_gpipe as lg
  cls = cls or lg.GPipeTransformerStack
  p = cls.Params().Set(name='stack', model_dim=16, random_seed=321, **kw)
  for tpl in (p.encoder_tpl, p.decoder_tpl):
    t = tpl.transformer_tpl if 'transformer_tpl' in tpl else tpl
    t.tr_atten_tpl.num_attention_heads = 2
    t.tr_fflayer_tpl.hidden_dim = 32
    if 'tr_double_heads_atten_tpl' in tpl:
      tpl.tr_atten_tpl.num_attention_heads = 2
      tpl.tr_double_heads_atten_tpl.num_attention_heads = 4
  return p


def test_deterministic_weights_layer():
  from lingvo_b200.core import layers_with_gpipe as lg
  p = lg.DeterministicWeightsLayer.Params().Set(name='w', num_sources=4, minimal_prob=0.05)
  layer = p.Instantiate()
  w = layer.FPropDefaultTheta()
  torch.testing.assert_close(w, torch.full((4,), 0.25))
  with torch.no_grad():
    layer.vars.sum_weight.copy_(torch.tensor([2.0, 0.0, 0.0, -50.0]))
  w = layer.FPropDefaultTheta()
  assert abs(float(w.sum()) - 1.0) < 1e-6 and float(w.min()) >= 0.05 - 1e-7
  p2 = p.Copy().Set(weighted_merger_softmax=False, global_weight_scale=2.0)
  l2 = p2.Instantiate()
  with torch.no_grad():
    l2.vars.sum_weight.copy_(torch.tensor([1.0, 0.0, 0.0, 0.0]))
  torch.testing.assert_close(l2.FPropDefaultTheta(), torch.tensor([2.25, 0.25, 0.25, 0.25]))


def test_transparent_gpipe_stack_merges_all_encoder_layers():
  from lingvo_b200.core import layers_with_gpipe as lg
  p = _SmallStack(num_encoder_layers=3, splits=[1, 3], is_transparent=True,
                  transparent_merger_dropout_prob=0.0, normalize_encoder=True)
  stack = p.Instantiate()
  encs = stack.GetEncoders()
  assert len(encs) == 3 and hasattr(encs[0], 'transparent_merger')
  assert not hasattr(encs[1], 'transparent_merger') and encs[2].params.final_enc_layer
  assert hasattr(encs[2], 'layer_norm')
  # several cells ⇒ every dropout is deterministic
  assert type(encs[0].self_atten.residual_dropout).__name__ == 'DeterministicDropoutLayer'
  assert type(encs[0].fflayer.fflayer.dropout[0]).__name__ == 'DeterministicDropoutLayer'
  x = torch.randn(5, 2, 16)
  pad = torch.zeros(5, 2)
  out = stack.FPropDefaultTheta(x, pad)
  assert out.shape == x.shape
  # manual: acc = Σ w_i · input_i + w_3 · h_3, then LN
  with torch.no_grad():
    encs[0].transparent_merger.vars.sum_weight.copy_(torch.tensor([0.3, -0.2, 0.1, 0.5]))
  out = stack.FPropDefaultTheta(x, pad)
  w = encs[0].transparent_merger.FPropDefaultTheta()
  from lingvo_b200.core import layers_with_attention as lwa
  h, acc = x, torch.zeros_like(x)
  for i, e in enumerate(encs):
    acc = acc + w[i] * h
    h, _ = lwa.TransformerLayer.FProp(e, e.theta, h, pad)
  want = encs[2].layer_norm.FPropDefaultTheta(acc + w[3] * h)
  torch.testing.assert_close(out, want, atol=1e-5, rtol=1e-5)
  # FPropMeta threads the shrinking helper shape through the cells
  from lingvo_b200.core import tshape
  shapes = stack._CalculateOutputShapes(
      [tshape.Shape([5, 2, 16]), tshape.Shape([5, 2])] + [None] * 8)
  assert list(shapes[0][7]) == [3] and shapes[1][7] is None
  (out.sum()).backward()
  assert encs[0].transparent_merger.vars.sum_weight.grad.abs().sum() > 0


def test_evolved_transformer_gpipe_stack_and_pipelined_embeddings():
  from lingvo_b200.core import layers_with_gpipe as lg
  p = _SmallStack(lg.GPipeEvolvedTransformerStack, num_encoder_layers=1, num_decoder_layers=1,
                  splits=2, num_splits=2, use_pipelined_embeddings=True, num_micro_batches=2)
  p.emb_tpl.Set(vocab_size=11, model_dim=16, max_seq_len=32)
  p.softmax_tpl.Set(num_classes=11, input_dim=16)
  from lingvo_b200.core import layers
  p.label_smoothing = layers.UniformLabelSmoother.Params().Set(
      name='smooth', num_classes=11, uncertainty=0.1)
  stack = p.Instantiate()
  assert type(stack.GetEncoders()[0]).__name__ == 'GPipeEvolvedTransformerEncoderLayer'
  assert type(stack.GetDecoders()[0]).__name__ == 'GPipeEvolvedTransformerDecoderLayer'
  src = torch.randint(0, 11, (6, 4)); tgt = torch.randint(0, 11, (5, 4))
  sp, tp = torch.zeros(6, 4), torch.zeros(5, 4)
  logits = stack.FPropDefaultTheta(src, sp, tgt, tp)
  assert logits.shape == (5, 4, 11)
  labels = torch.randint(0, 11, (5, 4))
  xent, logits2 = stack.FPropDefaultTheta(src, sp, tgt, tp, labels=labels,
                                          label_weights=torch.ones(5, 4))
  assert xent.shape == (5, 4) and torch.isfinite(xent).all()
  torch.testing.assert_close(logits, logits2)
  # embedding / softmax helpers used by decoders
  emb = stack.EncoderEmbedFPropDefaultTheta(src)
  assert emb.shape == (6, 4, 16)
  step = stack.DecoderEmbedFPropDefaultTheta(tgt[2:3], t=2)
  full = stack.DecoderEmbedFPropDefaultTheta(tgt)
  torch.testing.assert_close(step[0], full[2])
  lg2 = stack.Logits(stack.theta, torch.randn(3, 16))
  assert lg2.shape == (3, 11)
  enc = stack.EncoderFPropDefaultTheta(emb, sp)
  assert enc.shape == (6, 4, 16)
  xent.sum().backward()
  assert all(v.grad is not None for v in stack.vars.Flatten())


def test_batch_major_gpipe_stack_with_embeddings_and_softmax():
  from lingvo_b200.core import layers_with_gpipe as lg
  p = lg.GPipeBatchMajorTransformerStack.Params().Set(
      name='bm', model_dim=16, num_encoder_layers=2, num_decoder_layers=1, num_splits=2,
      packed_input=True, random_seed=5,
      emb_tpl=lg.GPipeBatchMajorTransformerEmbeddingLayer.Params().Set(vocab_size=13),
      softmax_tpl=lg.GPipeBatchMajorTransformerSoftmaxLayer.Params().Set(num_classes=13))
  for tpl in (p.encoder_tpl, p.decoder_tpl):
    tpl.tr_atten_tpl.num_heads = 2
    tpl.tr_fflayer_tpl.hidden_dim = 32
  stack = p.Instantiate()
  src = torch.randint(0, 13, (2, 6)); tgt = torch.randint(0, 13, (2, 4))
  sseg = torch.tensor([[1, 1, 1, 2, 2, 2]] * 2); tseg = torch.tensor([[1, 1, 2, 2]] * 2)
  spos = torch.tensor([[0, 1, 2, 0, 1, 2]] * 2); tpos = torch.tensor([[0, 1, 0, 1]] * 2)
  out = stack.FPropDefaultTheta(src, torch.zeros(2, 6), tgt, torch.zeros(2, 4), sseg, tseg,
                                source_segment_pos=spos, target_segment_pos=tpos)
  assert out.shape == (2, 4, 13)
  # packed: the second target segment must not see the first source segment
  src2 = src.clone(); src2[:, :3] = (src2[:, :3] + 1) % 13
  out2 = stack.FPropDefaultTheta(src2, torch.zeros(2, 6), tgt, torch.zeros(2, 4), sseg, tseg,
                                 source_segment_pos=spos, target_segment_pos=tpos)
  torch.testing.assert_close(out[:, 2:], out2[:, 2:], atol=1e-5, rtol=1e-5)
  assert (out[:, :2] - out2[:, :2]).abs().max() > 1e-4
