"""Behavioural tests of core modules that have no dedicated suite elsewhere."""

import math
import os

import numpy as np
import pytest
import torch

from lingvo_b200.core import layers, py_utils
from lingvo_b200.core.nested_map import NestedMap


def test_entmax_and_sparsemax():
  from lingvo_b200.core import entmax
  x = torch.tensor([[2.0, 1.0, 0.1, -3.0], [0.0, 0.0, 0.0, 0.0]], requires_grad=True)
  for fn in (entmax.entmax15, entmax.sparsemax):
    p = fn(x)
    torch.testing.assert_close(p.sum(-1), torch.ones(2))
    assert (p >= 0).all()
    assert p[0, 3] == 0                                   # sparse: far-away logit gets exactly 0
    torch.testing.assert_close(p[1], torch.full((4,), 0.25))
  g, = torch.autograd.grad(entmax.entmax15(x)[0, 0], x)
  assert abs(float(g[0].sum())) < 1e-5                    # gradients live on the simplex tangent
  assert entmax.entmax_support(x)[0].tolist()[-1] in (0, False)


def test_favor_attention_approximates_softmax_attention():
  from lingvo_b200.core import favor_attention as fa
  torch.manual_seed(0)
  b, l, h, d = 2, 12, 2, 16
  q, k, v = (torch.randn(b, l, h, d) * 0.3 for _ in range(3))
  proj = fa.create_projection_matrix(512, d, seed=1)
  out = fa.favor_attention(q, k, v, None, fa.softmax_kernel_transformation, causal=False,
                           projection_matrix=proj)
  ref = torch.einsum('blhm,bmhd->blhd', torch.softmax(
      torch.einsum('blhd,bmhd->blhm', q, k) / math.sqrt(d), -1), v)
  assert (out - ref).abs().mean() < 0.1 * ref.abs().mean() + 0.05
  # causal variant: position t must not depend on the future
  v2 = v.clone(); v2[:, -1] += 10.0
  o1 = fa.favor_attention(q, k, v, None, fa.relu_kernel_transformation, causal=True, projection_matrix=proj)
  o2 = fa.favor_attention(q, k, v2, None, fa.relu_kernel_transformation, causal=True, projection_matrix=proj)
  torch.testing.assert_close(o1[:, :-1], o2[:, :-1])


def _Fc(name, i, o):
  return layers.FCLayer.Params().Set(name=name, input_dim=i, output_dim=o, activation='TANH')


def test_revnet_matches_plain_autograd():
  from lingvo_b200.core import reversible_layers as rev
  torch.manual_seed(0)
  p = rev.StackedRevNetLayer.Params().Set(name='rev', sub_layer_params=[
      rev.RevNetLayer.Params().Set(name='r%d' % i, f_params=_Fc('f', 6, 6), g_params=_Fc('g', 6, 6))
      for i in range(3)])
  layer = p.Instantiate()
  x1, x2 = torch.randn(4, 6, requires_grad=True), torch.randn(4, 6, requires_grad=True)
  out = layer.FPropDefaultTheta(NestedMap(split1=x1, split2=x2))
  loss = (out.split1 * out.split2).sum()
  grads = torch.autograd.grad(loss, [x1, x2] + list(layer.vars.Flatten()))
  # oracle: the same computation written out without the memory-saving Function
  def Plain(a, b):
    for blk in layer.sub_layers:
      a2 = a + blk.f_block.FPropDefaultTheta(b)
      b2 = b + blk.g_block.FPropDefaultTheta(a2)
      a, b = a2, b2
    return a, b
  a, b = Plain(x1, x2)
  ref = torch.autograd.grad((a * b).sum(), [x1, x2] + list(layer.vars.Flatten()))
  for g, r in zip(grads, ref):
    torch.testing.assert_close(g, r, atol=1e-5, rtol=1e-4)


def test_generic_repeat_layer_shares_or_separates_variables():
  from lingvo_b200.core import repeat_layer
  p = repeat_layer.GenericRepeatLayer.Params().Set(name='rep', body=_Fc('b', 5, 5), repeat=3)
  layer = p.Instantiate()
  x = torch.randn(2, 5)
  y = layer.FPropDefaultTheta(x)
  z = x
  for i in range(3):
    z = layer.body_iter[i].FPropDefaultTheta(z)
  torch.testing.assert_close(y, z)
  assert len(layer.vars.Flatten()) == 6                     # w, b per iteration


def test_graddrop_masks_conflicting_gradient_signs():
  from lingvo_b200.core import graddrop
  layer = graddrop.GradDrop.Params().Set(name='gd', keep_gradnorm_constant=False,
                                         marginalize_batch_dim=False, use_input_sign_only=False,
                                         random_seed=1).Instantiate()
  x = torch.ones(3, 4)
  layer.FPropDefaultTheta(x)
  agree = layer.CombineLossGrads([torch.ones(3, 4), 2 * torch.ones(3, 4)])
  torch.testing.assert_close(agree, 3 * torch.ones(3, 4))   # same sign: nothing dropped
  conflict = layer.CombineLossGrads([torch.ones(3, 4), -torch.ones(3, 4)])
  assert set(conflict.unique().tolist()) <= {-1.0, 1.0}     # exactly one side survives per element


def test_pcgrad_projects_conflicting_task_gradients():
  from lingvo_b200.core import gradient_combiner as gc
  w = torch.nn.Parameter(torch.zeros(2)); w.var_name = 'w/var'
  vmap = NestedMap(w=w)
  def Entry(g):
    return NestedMap(loss_metric=(torch.tensor(1.0), torch.tensor(1.0)),
                     grads=NestedMap(w=py_utils.VarGrad(w, torch.tensor(g))))
  comb = gc.PCGradCombiner.Params().Set(name='pc').Instantiate()
  out, _ = comb.Combine(vmap, {'a': Entry([1.0, 0.0]), 'b': Entry([-1.0, 1.0])})
  g = out.w.grad
  # each task gradient is projected onto the normal plane of the other: no component of the sum
  # points against either original gradient
  assert float(g @ torch.tensor([1.0, 0.0])) >= -1e-6 and float(g @ torch.tensor([-1.0, 1.0])) >= -1e-6
  summed, _ = gc.SumCombiner.Params().Set(name='s').Instantiate().Combine(
      vmap, {'a': Entry([1.0, 0.0]), 'b': Entry([-1.0, 1.0])})
  torch.testing.assert_close(summed.w.grad, torch.tensor([0.0, 1.0]))


def test_bleu_scorers_agree_on_perfect_and_imperfect_hyps():
  from lingvo_b200.core import ml_perf_bleu_metric as mlb, scorers
  refs = ['the quick brown fox jumps over the lazy dog', 'hello there general kenobi you are bold']
  s = scorers.BleuScorer()
  for r in refs:
    s.AddSentence(r, r)
  assert abs(s.ComputeOverallScore() - 1.0) < 1e-6
  assert abs(mlb.bleu_wrapper(refs, refs) - 1.0) < 1e-6       # fraction, like the reference
  hyps = ['the quick brown cat jumps over the lazy dog', 'hello there general kenobi you are old']
  m = mlb.MlPerfBleuMetric()
  for r, h in zip(refs, hyps):
    m.Update(r, h)
  assert 0.40 < m.value < 0.95
  assert mlb.bleu_tokenize('Hello, world!') == ['Hello', ',', 'world', '!']


def test_early_stop_and_metric_history(tmp_path):
  from lingvo_b200.core import early_stop
  mh = early_stop.MetricHistory.Params().Set(logdir=str(tmp_path), jobname='eval',
                                             metric='loss').Instantiate()
  for step, v in [(10, 3.0), (20, 2.0), (30, 2.5), (40, 2.4), (50, 2.3)]:
    mh.ConditionalAppend('eval', 'loss', step, v)
  best, last = early_stop.BestStep(mh.hist_file, 0.0, True)
  assert (best, last) == (20, 50)
  es = early_stop.EarlyStop.Params().Set(name='es', window=20, tolerance=0.0,
                                         metric_history=mh.params).Instantiate()
  assert es.Stop()                                          # 30 steps since the best > window
  es2 = early_stop.EarlyStop.Params().Set(name='es', window=100, metric_history=mh.params).Instantiate()
  assert not es2.Stop()


def test_differentiable_assignment_respects_marginals():
  from lingvo_b200.core import differentiable_assignment as da
  torch.manual_seed(0)
  score = torch.randn(2, 4, 6, requires_grad=True)
  rows, cols = torch.full((2, 4), 1.5), torch.full((2, 6), 1.0)
  a, iters, eps, delta = da.max_assignment(
      score, elementwise_upper_bound=torch.ones(2, 4, 6), row_sums=rows, col_sums=cols,
      epsilon=0.05, num_iterations=200)
  assert a.shape == score.shape and (a >= -1e-6).all() and (a <= 1 + 1e-4).all()
  torch.testing.assert_close(a.sum(-1), rows, atol=2e-2, rtol=2e-2)
  torch.testing.assert_close(a.sum(-2), cols, atol=2e-2, rtol=2e-2)
  assert iters == 200 and abs(eps - 0.05) < 1e-9 and float(delta) < 0.05
  # low temperature → near-integral, and it prefers high scores
  assert float(((a > 0.9) | (a < 0.1)).float().mean()) > 0.7
  assert float((a * score).sum()) > float((score.detach().mean() * a.sum()))
  g, = torch.autograd.grad((a * torch.randn_like(a)).sum(), score)
  assert torch.isfinite(g).all() and g.abs().sum() > 0


def test_egdd_optimizer_descends():
  from lingvo_b200.core import egdd
  torch.manual_seed(0)
  w = torch.nn.Parameter(torch.randn(6)); w.var_name = 'w/var'
  opt = egdd.EGDD.Params().Set(name='egdd').Instantiate()
  first = None
  for _ in range(60):
    loss = (w ** 2).sum()
    g, = torch.autograd.grad(loss, w)
    opt.Apply(0.05, [py_utils.VarGrad(w, g)])
    first = first or float(loss)
  assert float((w ** 2).sum()) < 0.5 * first


def test_task_schedulers():
  from lingvo_b200.core import task_scheduler as ts
  c = ts.ConstantScheduler.Params().Set(name='c', task_probs=[('a', 0.8), ('b', 0.2)], random_seed=3).Instantiate()
  picks = [c.Sample(s) for s in range(2000)]
  assert 0.74 < picks.count('a') / 2000 < 0.86
  assert [c.Sample(7) for _ in range(3)] == [c.Sample(7)] * 3          # seeded: a function of the step
  e = ts.ExponentialScheduler.Params().Set(name='e', alpha=0.01, random_seed=1,
                                           task_probs=[('a', (1.0, 0.0)), ('b', (0.0, 1.0))]).Instantiate()
  e.Sample(0); p0 = e.cur_probs
  e.Sample(2000); p1 = e.cur_probs
  assert p0[0] > 0.99 and p1[1] > 0.99                                   # drifts from a to b
  rr = ts.RoundRobinScheduler.Params().Set(name='r', tasks=['x', 'y', 'z']).Instantiate()
  assert [rr.Sample(i) for i in range(5)] == ['x', 'y', 'z', 'x', 'y']
  sq = ts.SequentialScheduler.Params().Set(name='s', task_steps=[('p', 2), ('q', 3)]).Instantiate()
  assert [sq.Sample(i) for i in range(6)] == ['p', 'p', 'q', 'q', 'q', 'q']


def test_wpm_encoder_roundtrip(tmp_path):
  from lingvo_b200.core import wpm_encoder
  bow = wpm_encoder.BOW_STR
  vocab = ['<unk>', '<s>', '</s>', bow + 'hel', 'lo', bow + 'wor', 'ld', bow, 'h', 'e', 'l', 'o']
  f = tmp_path / 'wpm.txt'
  f.write_text('\n'.join(vocab) + '\n', encoding='utf-8')
  enc = wpm_encoder.WpmEncoder(str(f))
  ids, pieces = enc.Encode('hello world')
  assert pieces == [bow + 'hel', 'lo', bow + 'wor', 'ld']
  assert enc.Decode(ids) == 'hello world'
  assert enc.EncodeWord('zzz')[0] == bow and '<unk>' in enc.EncodeWord('zzz')
  assert (enc.sentence_start_id, enc.sentence_end_id, enc.unk_id) == (1, 2, 0)


def test_flat_beam_search_finds_the_most_likely_sequence():
  from lingvo_b200.core import flat_beam_search_helper as fbs
  # A fixed first-order model: after token t the next-token log-probs are row t of `table`.
  v, eos = 5, 2
  table = torch.full((v, v), -4.0)
  table[1, 3] = 0.0      # <s> → 3
  table[3, 4] = 0.0      # 3 → 4
  table[4, eos] = 0.0    # 4 → </s>
  table[1, 4] = -0.5     # a tempting but worse branch: <s> → 4 → </s>

  def Callback(ids, pos, seg, mask, state, t):
    del pos, seg, mask, t
    return table[ids], state

  (out_ids, lens, scores), _ = fbs.flat_beam_search(2, 3, 6, Callback, None, bos_id=1, eos_id=eos,
                                                    beam_gap=None)
  assert out_ids.shape[:2] == (2, 3)
  best = out_ids[0, 0, :int(lens[0, 0])].tolist()
  assert best == [3, 4, eos], (best, scores[0])
  assert scores[0, 0] >= scores[0, 1] >= scores[0, 2]
  second = out_ids[0, 1, :int(lens[0, 1])].tolist()
  assert second == [4, eos]
  m, s = fbs.update_nbest((torch.zeros(1, 2, 4, dtype=torch.bool), torch.tensor([[0.5, 0.1]])),
                          (torch.ones(1, 2, 4, dtype=torch.bool), torch.tensor([[0.3, 0.9]])))
  torch.testing.assert_close(s, torch.tensor([[0.9, 0.5]]))
  assert m[0, 0].all() and not m[0, 1].any()


def test_activations_table_and_inspect_utils():
  from lingvo_b200.core import activations, hyperparams, inspect_utils
  x = torch.linspace(-2, 2, 9)
  torch.testing.assert_close(activations.GetFn('RELU')(x), torch.relu(x))
  torch.testing.assert_close(activations.GetFn('SWISH')(x), x * torch.sigmoid(x))
  assert activations.IsSupported('GELU') and not activations.IsSupported('NOPE')
  assert activations.DimMultiplier('GATED_GELU') == 2 and activations.DimMultiplier('RELU') == 1
  assert activations.GetFlops('NONE') == 0 and activations.GetFlops('TANH') > 0

  def Fn(a, b=2, *, c='x'):
    return (a, b, c)
  p = hyperparams.Params()
  inspect_utils.DefineParams(Fn, p)
  assert (p.a, p.b, p.c) == (None, 2, 'x')
  p.a = 5
  assert inspect_utils.CallWithParams(Fn, p, c='y') == (5, 2, 'y')

  class Thing:
    def __init__(self, size, name='t'):
      self.size, self.name = size, name
  q = inspect_utils.ParamsFromCallable(Thing.__init__, ignore=['self'])
  q.size = 3
  t = inspect_utils.ConstructWithParams(Thing, q)
  assert (t.size, t.name) == (3, 't')


def _Lin(name, i, o):
  from lingvo_b200.core import builder_layers as bl
  return bl.LinearLayer.Params().Set(name=name, input_dims=i, output_dims=o)


def test_builder_layers_combinators():
  from lingvo_b200.core import builder_layers as bl
  torch.manual_seed(0)
  x = torch.randn(3, 4)
  seq = bl.SequentialLayer.Params().Set(name='seq', sub=[
      _Lin('a', 4, 6), bl.MapLayer.Params().Set(name='act', fn=torch.tanh), _Lin('b', 6, 2)]).Instantiate()
  y = seq.FPropDefaultTheta(x)
  want = torch.tanh(x @ seq.a.vars.w) @ seq.b.vars.w
  torch.testing.assert_close(y, want)
  # stacked-variable repeat: one [repeat, …] variable, applied slice by slice
  rep = bl.RepeatLayer.Params().Set(name='rep', body=_Lin('l', 4, 4), repeat=3).Instantiate()
  w = rep.body.vars.w
  assert tuple(w.shape) == (3, 4, 4)
  z = x
  for i in range(3):
    z = z @ w[i]
  torch.testing.assert_close(rep.FPropDefaultTheta(x), z, atol=1e-5, rtol=1e-5)
  par = bl.ParallelLayer.Params().Set(
      name='par', sub=[_Lin('left', 4, 2), _Lin('right', 4, 2)],
      merge=lambda outs: tuple(sum(o[0] for o in outs) for _ in range(1))).Instantiate()
  torch.testing.assert_close(par.FPropDefaultTheta(x), x @ par.left.vars.w + x @ par.right.vars.w)
  g = bl.GraphLayer.Params().Set(name='g', input_endpoints=['x'], output_endpoints=['y', 'h'], sub=[
      ('x->h', _Lin('first', 4, 5)),
      ('h->t', bl.MapLayer.Params().Set(name='sq', fn=torch.square)),
      ('t->y', _Lin('second', 5, 1))]).Instantiate()
  gy, gh = g.FPropDefaultTheta(x)
  torch.testing.assert_close(gh, x @ g.first.vars.w)
  torch.testing.assert_close(gy, torch.square(gh) @ g.second.vars.w)
  first = bl.FirstNLayer.Params().Set(name='f', n=2).Instantiate()
  assert first.FPropDefaultTheta(1, 2, 3) == (1, 2)
  remat = bl.RematerializationLayer.Params().Set(name='rm', body=_Lin('inner', 4, 4)).Instantiate()
  xr = x.clone().requires_grad_(True)
  out = remat.FPropDefaultTheta(xr)
  gx, = torch.autograd.grad(out.sum(), xr)
  torch.testing.assert_close(gx, (torch.ones(3, 4) @ remat.body.vars.w.t()))
  sig = bl.GraphSignature('a,b.c,[d,e],(k=f)->g,h.i')
  assert sig.outputs == ['g', 'h.i'] and len(sig.inputs) == 4
  with pytest.raises(ValueError):
    bl.GraphSignature('a->1bad')


def test_pruning_schedule_and_masks():
  from lingvo_b200.core import pruning_utils as pu
  pu.PruningOp.Reset()
  hp = dict(begin_pruning_step=10, end_pruning_step=110, initial_sparsity=0.0, target_sparsity=0.75,
            sparsity_function_exponent=3.0)

  class Obj:
    pass
  torch.manual_seed(0)
  w = torch.nn.Parameter(torch.randn(16, 8)); w.var_name = 'lstm/wm/var'
  obj = Obj(); obj.vars = NestedMap(wm=w)
  mask = pu.PruningOp.ApplyPruning(hp, obj, 'wm', None, torch.float32)
  assert mask.all()
  assert pu.PruningOp.Sparsity(0) == 0.0 and abs(pu.PruningOp.Sparsity(110) - 0.75) < 1e-9
  assert 0.0 < pu.PruningOp.Sparsity(40) < pu.PruningOp.Sparsity(80) < 0.75      # cubic ramp
  s = pu.PruningOp.UpdateMasks(110)
  kept = float(mask.mean())
  assert abs(s - 0.75) < 1e-9 and abs(kept - 0.25) < 0.02
  # the smallest-magnitude weights are the ones removed
  assert float(w.detach().abs()[mask == 0].max()) <= float(w.detach().abs()[mask == 1].min())
  mw = pu.PruningOp.MaskedWeight(w)
  assert float((mw == 0).float().mean()) >= 0.74
  pu.PruningOp.Reset()


def test_saver_sanity_checks_block_bad_checkpoints(tmp_path):
  from lingvo_b200.core import saver as saver_lib
  state = {'w/var': torch.ones(3), 'global_step': torch.tensor(5)}
  sv = saver_lib.Saver(str(tmp_path), lambda: state,
                       sanity_checks=[(r'w/', [saver_lib.IsFinite()])])
  path = sv.Save(5)
  assert saver_lib.LatestCheckpoint(str(tmp_path)) == path
  assert saver_lib.ReadCheckpointState(str(tmp_path))['model_checkpoint_path']
  state['w/var'] = torch.tensor([1.0, float('nan'), 0.0])
  with pytest.raises(saver_lib.SanityCheckFailed):
    sv.Save(6)
  assert saver_lib.LatestCheckpoint(str(tmp_path)) == path                       # nothing committed
  rng = saver_lib.InRange(-1.0, 1.0)
  assert rng.Check('v', torch.tensor([0.5])) and not rng.Check('v', torch.tensor([2.0]))


def test_tpu_embedding_layer_sparse_updates_touch_only_looked_up_rows():
  from lingvo_b200.core import tpu_embedding_layers as tel
  torch.manual_seed(0)
  table = tel.TPUEmbeddingTable.Params().Set(
      name='t', vocab_size=20, embedding_dim=4, input_keys=['a', 'b'], combiner='mean')
  layer = tel.TPUEmbeddingLayer.Params().Set(
      name='emb', tables=[table], learning_rate=0.5,
      optimizer=tel.TPUEmbeddingSGDOptimizer.Params()).Instantiate()
  before = layer.tables[0].table.detach().clone()
  ids = NestedMap(a=torch.tensor([[1, 3, -1], [3, 3, 5]]), b=torch.tensor([[7, -1, -1], [-1, -1, -1]]))
  out = layer.EmbLookup(layer.theta, ids)
  assert out.a.shape == (2, 4) and out.b.shape == (2, 4)
  torch.testing.assert_close(out.a[0], (before[1] + before[3]) / 2)            # mean over valid ids
  torch.testing.assert_close(out.b[1], torch.zeros(4))                          # all-missing row
  (out.a.sum() + out.b.sum()).backward()
  layer.ApplyGradients(global_step=0)
  after = layer.tables[0].table.detach()
  changed = (after != before).any(-1).nonzero().flatten().tolist()
  assert changed == [1, 3, 5, 7]
  # row 3 was hit with weights 1/2 (example 0) and 2·1/3 (example 1): SGD step = lr · Σ
  torch.testing.assert_close(before[3] - after[3], torch.full((4,), 0.5 * (0.5 + 2.0 / 3.0)))


def test_attention_util_blocks_masks_and_sparse_attention():
  from lingvo_b200.core import attention_util as au
  x = torch.arange(2 * 7 * 3, dtype=torch.float32).reshape(2, 7, 3)
  blocks = au.ConvertToBlocks(x, 3)
  assert blocks.shape == (2, 3, 3, 3) and (blocks[:, 2, 1:] == 0).all()
  ctx = au.ExtractBlockContext(x, block_size=3, left_context=2, right_context=1)
  assert ctx.shape == (2, 3, 5, 3)
  torch.testing.assert_close(ctx[:, 1, 1:4], x[:, 3:6])                        # the block itself
  torch.testing.assert_close(ctx[:, 1, 0], x[:, 2])                            # one step of left context
  mask = au.MakeLocalMask(7, 3, 2, 1)
  assert mask.shape == (3, 3, 5)
  # query t=4 (block 1, w=1) sees keys 3..5: context positions k = block*3 - 1 + c
  assert mask[1, 1].tolist() == [0.0, 1.0, 1.0, 1.0, 0.0]
  t = 4
  rel = torch.randn(1, 1, t, 2 * t - 1)
  shifted = au.RelShift(rel)
  for i in range(t):
    for j in range(t):
      assert shifted[0, 0, i, j] == rel[0, 0, i, j - i + t - 1]
  torch.manual_seed(0)
  q, k, v = torch.randn(1, 2, 5, 8), torch.randn(1, 2, 6, 8), torch.randn(1, 2, 6, 8)
  idx = torch.tensor([0, 2, -1]).expand(1, 2, 5, 3)
  out, probs = au.ComputeSparseAttention(q, k, v, idx)
  logits = torch.einsum('bnth,bnsh->bnts', q, k)[..., [0, 2]] / math.sqrt(8)
  want = torch.einsum('bntw,bnwh->bnth', torch.softmax(logits, -1), v[:, :, [0, 2]])
  torch.testing.assert_close(out, want, atol=1e-5, rtol=1e-5)
  assert (probs[..., 2] == 0).all()


def test_lstm_frnn_matches_generic_frnn():
  from lingvo_b200.core import lstm_frnn_layer, rnn_cell, rnn_layers
  torch.manual_seed(0)
  cell = lstm_frnn_layer.LSTMCellSimpleExt.Params().Set(name='cell', num_input_nodes=5, num_output_nodes=7)
  fast = lstm_frnn_layer.LstmFRNN.Params().Set(name='fast', cell=cell, remat_steps=3).Instantiate()
  slow = rnn_layers.FRNN.Params().Set(
      name='slow', cell=rnn_cell.LSTMCellSimple.Params().Set(name='cell', num_input_nodes=5,
                                                             num_output_nodes=7)).Instantiate()
  for a, b in zip(slow.vars.Flatten(), fast.vars.Flatten()):
    a.data.copy_(b.data)
  x = torch.randn(9, 2, 5, requires_grad=True)
  pad = torch.zeros(9, 2, 1); pad[6:, 1] = 1.0
  y1, s1 = fast.FPropDefaultTheta(x, pad)
  y2, s2 = slow.FPropDefaultTheta(x, pad)
  torch.testing.assert_close(y1, y2, atol=1e-5, rtol=1e-5)
  torch.testing.assert_close(s1.c, s2.c, atol=1e-5, rtol=1e-5)
  g1, = torch.autograd.grad(y1.sum(), x, retain_graph=True)
  g2, = torch.autograd.grad(y2.sum(), x)
  torch.testing.assert_close(g1, g2, atol=1e-5, rtol=1e-4)
  rev = lstm_frnn_layer.LstmFRNN.Params().Set(name='rev', cell=cell, reverse=True).Instantiate()
  for a, b in zip(rev.vars.Flatten(), fast.vars.Flatten()):
    a.data.copy_(b.data)
  yr, _ = rev.FPropDefaultTheta(x, torch.zeros(9, 2, 1))
  yf, _ = fast.FPropDefaultTheta(torch.flip(x, [0]), torch.zeros(9, 2, 1))
  torch.testing.assert_close(yr, torch.flip(yf, [0]), atol=1e-5, rtol=1e-5)


def test_gshard_utils_sharding_specs():
  from lingvo_b200.core import gshard_utils as gu
  mesh = np.arange(8).reshape(2, 4)
  spec = gu.TensorShardingSpec.FromFullShape([10, 16, 3], [1, 0, -1], mesh)
  assert not spec.is_replicated and spec.uneven_padding == [2, 0, 0]
  assert spec.ShardShape([10, 16, 3]) == [3, 8, 3] and spec.NumShards(0) == 4
  # shards tile the tensor exactly once (last shard of an uneven dim is shorter)
  cover = np.zeros((10, 16, 3), int)
  for a in range(2):
    for b in range(4):
      cover[spec.ShardSlices([10, 16, 3], (a, b))] += 1
  assert (cover == 1).all()
  assert spec.AddLeadingDims(2).split_dims_mapping == [-1, -1, 1, 0, -1]
  assert spec.RemoveDim(1).split_dims_mapping == [1, -1]
  assert gu.TensorShardingSpec.ReplicatedSpec().is_replicated
  x = torch.zeros(4, 6)
  y = gu.Split(x, 1, 2)
  assert gu.GetSharding(y).split_dims_mapping == [-1, 0]
  with gu.MeshSplitDimPrefixContext(1):
    z = gu.MeshSplit(torch.zeros(3, 4, 6), mesh, [0, -1])
    assert gu.GetMeshSplitDimPrefixContext() == [1]
  assert gu.GetSharding(z).split_dims_mapping == [1, 0, -1]
  assert gu.GetMeshSplitDimPrefixContext() == []
  zz = gu.ZigzagOrderOnDeviceMesh(np.arange(8), 0)
  assert zz.tolist() == [0, 2, 4, 6, 7, 5, 3, 1]
  v = torch.nn.Parameter(torch.zeros(10, 16)); v.device_mesh = mesh; v.tensor_split_dims_mapping = [0, 1]
  assert gu.GetVarSharding(v).ShardShape([10, 16]) == [5, 4]
  assert gu.GetVarSharding(torch.nn.Parameter(torch.zeros(2))).is_replicated


def test_datasources_mixing_curriculum_and_iterators():
  from lingvo_b200.core import datasource as ds
  mk = lambda tag: ds.IteratorDataSource.Params().Set(
      name='it_' + tag, iter_fn=lambda: iter([NestedMap(src=tag, i=k) for k in range(3)]))
  it = mk('a').Instantiate()
  assert [it.GetNext().i for _ in range(5)] == [0, 1, 2, 0, 1]           # repeats
  once = ds.IteratorDataSource.Params().Set(name='once', repeat=False,
                                            iter_fn=lambda: iter([NestedMap(i=0)])).Instantiate()
  once.GetNext()
  with pytest.raises(StopIteration):
    once.GetNext()
  mix = ds.CrossBatchMixingDataSource.Params().Set(
      name='mix', sub=[mk('a'), mk('b')], weights=[0.9, 0.1], random_seed=1).Instantiate()
  picks = [mix.GetNext() for _ in range(400)]
  frac_a = sum(b.src == 'a' for b in picks) / 400
  assert 0.84 < frac_a < 0.96 and int(picks[0].source_selected[0]) in (0, 1)
  cur = ds.CurriculumDataSource.Params().Set(name='cur', sub=[mk('a'), mk('b')], boundaries=[10]).Instantiate()
  with py_utils.GlobalStepContext(3):
    assert cur.GetNext().src == 'a'
  with py_utils.GlobalStepContext(10):
    assert cur.GetNext().src == 'b'


def test_summary_collector_step_rate_and_model_analysis(tmp_path):
  from lingvo_b200.core import plot, summary_utils
  from lingvo_b200.utils import tfevents
  with summary_utils.SummaryCollector() as col:
    summary_utils.scalar('a/loss', torch.tensor(1.5))
    summary_utils.scalar('b/const', 2)
    summary_utils.histogram('h', torch.arange(10.0))
    summary_utils.text('t', 'hello')
  summary_utils.scalar('ignored', 3.0)                       # no active collector: dropped
  vals = col.Resolve()
  assert vals == {'a/loss': 1.5, 'b/const': 2.0}
  w = tfevents.EventFileWriter(str(tmp_path))
  col.WriteTo(w, step=7)
  w.close()
  scalars = list(tfevents.ReadScalars(w.path))
  assert (7, 'a/loss', 1.5) in scalars and (7, 'b/const', 2.0) in scalars
  tr = summary_utils.StepRateTracker()
  import time
  r0 = tr.ComputeStepRate(0, 0)
  time.sleep(0.05)
  rate, ex_rate, total = tr.ComputeStepRate(10, 320)
  assert r0[0] == 0.0 and 50 < rate < 400 and abs(ex_rate / rate - 32.0) < 1e-6 and total == 320
  layer = layers.FCLayer.Params().Set(name='fc', input_dim=4, output_dim=3).Instantiate()
  table, n = summary_utils.ModelAnalysis(layer)
  assert n == 4 * 3 + 3 and 'fc/w/var' in table and '(4, 3)' in table
  # plotting degrades to None without matplotlib instead of failing
  fig = plot.MatplotlibFigureSummary('fig')
  fig.AddSubplot([np.zeros((2, 3, 3))])
  out = fig.Finalize()
  assert out is None or isinstance(out, list)
  assert plot.ToUnicode(b'abc') == 'abc'


def test_target_sequence_sampler_filters_and_stops():
  from lingvo_b200.core import target_sequence_sampler as tss
  v, eos = 6, 2
  table = torch.full((v, v), -8.0)
  table[1, 3] = 0.0; table[1, 4] = -0.1        # after <s>: 3 or 4 about equally likely
  table[3, eos] = 0.0; table[4, eos] = 0.0      # then always </s>
  table[eos, eos] = 0.0

  def Init(theta, enc, k):
    return NestedMap(log_probs=torch.zeros(4 * k, v)), NestedMap(step=torch.zeros(1))

  def Pre(theta, enc, ids, state, k, t):
    return NestedMap(log_probs=table[ids.squeeze(1)]), state

  s = tss.TargetSequenceSampler.Params().Set(target_seq_len=5, top_k=2, temperature=1.0).Instantiate()
  out = s.Sample(None, None, 7, Init, Pre, None)
  assert out.ids.shape == (4, 5) and out.logits.shape == (4, 5, v)
  assert set(out.ids[:, 0].tolist()) <= {3, 4}                       # top-2 filter
  assert (out.ids[:, 1:] == eos).all()
  torch.testing.assert_close(out.paddings[:, 2:], torch.ones(4, 3))  # padded after </s>
  assert out.paddings[:, :2].sum() == 0
  again = s.Sample(None, None, 7, Init, Pre, None)
  assert torch.equal(out.ids, again.ids)                             # seeded
  greedy = tss.TargetSequenceSampler.Params().Set(target_seq_len=3, top_k=1).Instantiate()
  assert (greedy.Sample(None, None, 0, Init, Pre, None).ids[:, 0] == 3).all()
  nucleus = tss.TargetSequenceSampler.Params().Set(target_seq_len=1, nucleus_p=0.3).Instantiate()
  f = nucleus._Filter(table[1:2])
  assert int((f > -1e29).sum()) == 1                                 # only the head of the nucleus survives
  eps = tss.TargetSequenceSampler.Params().Set(target_seq_len=1, epsilon=0.9).Instantiate()
  assert int((eps._Filter(table[1:2]) > -1e29).sum()) == 1           # fail-safe keeps the arg-max


def test_conv_layers_builder_blocks_respect_paddings():
  from lingvo_b200.core import conv_layers_builder as clb
  torch.manual_seed(0)
  b = clb.Builder.Params().Set(norm_layer_tpl=None).Instantiate()
  x = torch.randn(2, 10, 8, 3)                       # [batch, time, freq, channels]
  pad = torch.zeros(2, 10); pad[1, 6:] = 1.0
  conv = b.Conv2D('c', (3, 3, 3, 5), stride=(2, 2)).Instantiate()
  y, yp = conv.FPropDefaultTheta(x, pad)
  assert y.shape == (2, 5, 4, 5) and yp.shape == (2, 5)
  # reference (non-v2) rule: an output frame is padding if ANY input frame of its window is
  assert yp[1].tolist() == [0, 0, 1, 1, 1]
  assert (y[1, 2:] == 0).all()                       # padded frames stay zero
  sep = b.SeparableConv2D('s', (3, 3, 3, 7), depth_multiplier=2).Instantiate()
  y2, _ = sep.FPropDefaultTheta(x, pad)
  assert y2.shape == (2, 10, 8, 7)
  causal = b.Conv2D('cc', (3, 1, 3, 4), is_causal=True, activation='NONE').Instantiate()
  xa = x.clone(); xa[:, 7:] += 5.0                   # change the future …
  za, _ = causal.FPropDefaultTheta(x, torch.zeros(2, 10))
  zb, _ = causal.FPropDefaultTheta(xa, torch.zeros(2, 10))
  torch.testing.assert_close(za[:, :7], zb[:, :7])   # … the past does not move
  pool = b.CausalPooling('p', 'AVG', left_context=2).Instantiate()
  ones = torch.ones(1, 4, 1, 1)
  out, _ = pool.FPropDefaultTheta(ones * torch.arange(4.0).reshape(1, 4, 1, 1), torch.zeros(1, 4))
  assert out.flatten().tolist() == [0.0, 0.5, 1.5, 2.5]


def test_stacked_transformer_encoder_layers_wrapper():
  from lingvo_b200.core import self_attention_layer as sal
  torch.manual_seed(0)
  p = sal.StackedTransformerEncoderLayers.Params().Set(
      name='enc', num_layers=2, mdl_dim=16, hidden_dim=32, num_atten_heads=4)
  layer = p.Instantiate()
  x = torch.randn(2, 7, 16)
  pad = torch.zeros(2, 7); pad[0, 5:] = 1.0
  y, yp = layer.FPropDefaultTheta(x, pad)
  assert y.shape == x.shape and torch.equal(yp, pad)
  x2 = x.clone(); x2[0, 5:] += 3.0                    # padded positions must not influence the rest
  y2, _ = layer.FPropDefaultTheta(x2, pad)
  torch.testing.assert_close(y[0, :5], y2[0, :5], atol=1e-5, rtol=1e-5)
  cast = sal.StackedTransformerEncoderLayers.Cast(
      p.Copy().Set(name='again'))
  assert cast.num_layers == 2 and cast.mdl_dim == 16


def test_tpu_summary_context_and_predictor_runner(tmp_path):
  from lingvo_b200.core import predictor_runner_base as prb, saver as saver_lib, tpu_summary
  tpu_summary.scalar('outside', torch.tensor(1.0))               # no context: ignored
  with tpu_summary.context() as ctx:
    tpu_summary.scalar('loss', torch.tensor(2.5))
    tpu_summary.tensor('vec', torch.arange(3.0))
    tpu_summary.pw_tensor('pw', torch.ones(2))
    merged = tpu_summary.merge_all()
    assert merged['loss'] == 2.5 and merged['vec'].tolist() == [0.0, 1.0, 2.0]
    assert list(tpu_summary.merge_all_pw_tensor()) == ['pw']
    assert len(ctx.summary_tensors) == 2
  assert tpu_summary.merge_all() == {}

  # predictor runner: picks the newest checkpoint, runs every batch once, marks the step DONE
  ckpt_dir, out_dir = tmp_path / 'train', tmp_path / 'out'
  state = {'w/var': torch.ones(2), 'global_step': torch.tensor(12)}
  saver_lib.Saver(str(ckpt_dir), lambda: state).Save(12)

  class FakePredictor:
    def __init__(self):
      self.loaded, self.calls = [], 0
    def Load(self, path):
      self.loaded.append(path)
    def Run(self, fetch, subgraph_name='default', **feeds):
      self.calls += 1
      return {'y': feeds['x'] * 2}

  class Runner(prb.PredictorRunnerBase):
    def InputGenerator(self):
      for i in range(3):
        yield {'x': np.full(2, i)}
    def OutputWriter(self, output_dir, outputs):
      with open(os.path.join(output_dir, 'out.txt'), 'a') as f:
        f.write('%s\n' % outputs['y'].tolist())

  pred = FakePredictor()
  r = Runner(str(ckpt_dir), str(out_dir), 'inference_graph.pbtxt', pred, batch_size=2)
  r.Run()
  assert len(pred.loaded) == 1 and pred.loaded[0].endswith('ckpt-00000012') and pred.calls == 3
  step_dir = out_dir / 'step_00000012'
  assert (step_dir / 'DONE').read_text().startswith('3 batches')
  assert (step_dir / 'out.txt').read_text().splitlines() == ['[0, 0]', '[2, 2]', '[4, 4]']
  r.Run()                                                         # already DONE: nothing re-runs
  assert pred.calls == 3


def test_layer_variable_dict_fns_global_vn_and_optimizer_hooks(tmp_path):
  import pytest
  import torch
  from lingvo_b200.core import early_stop, layers, learner, optimizer, py_utils, rnn_cell
  from lingvo_b200.core.nested_map import NestedMap
  p = layers.FeedForwardNet.Params().Set(name='ffn', input_dim=4, hidden_layer_dims=[8, 2],
                                         activation=['RELU', 'NONE'])
  p.vn = py_utils.VariationalNoiseParams(0.5, global_vn=True)
  net = p.Instantiate()
  vd = net.GetVariablesDict()
  assert len(vd) == len(net.vars.Flatten()) and all(k.endswith('/var') for k in vd)
  some = next(iter(vd))
  leaf = [l for l in net.children.Flatten()][0] if hasattr(net.children, 'Flatten') else None
  net.AddFunction('double', lambda x: 2 * x)
  assert net.fns.double(3) == 6 and net.fns['double'](1) == 2
  with pytest.raises(AttributeError):
    net.AddFunction('double', lambda x: x)
  net.AddFunction('double', lambda x: 3 * x, replace=True)
  assert net.fns.double(1) == 3
  assert net.ema is None
  clean = net.vars.Transform(lambda v: v.detach().clone())
  noisy = net.AddGlobalVN(clean)
  diffs = [float((a - b).abs().max()) for a, b in zip(noisy.Flatten(), clean.Flatten())]
  assert max(diffs) > 0                                  # noise was added on a raw theta
  from lingvo_b200.core import cluster_factory
  with cluster_factory.SetEval(True):
    same = net.AddGlobalVN(clean)
  assert all(torch.equal(a, b) for a, b in zip(same.Flatten(), clean.Flatten()))
  del some, leaf
  # optimizer / learner hooks
  lp = learner.Learner.Params().Set(name='loss', optimizer=optimizer.SGD.Params(),
                                    learning_rate=0.5)
  lrn = lp.Instantiate()
  w = torch.nn.Parameter(torch.tensor([1.0, 2.0]))
  w.var_name = 'w/var'
  bound = lrn.optimizer.GetOptimizer(0.5)
  bound.apply_gradients(NestedMap(w=py_utils.VarGrad(w, torch.tensor([1.0, 1.0]))))
  torch.testing.assert_close(w.detach(), torch.tensor([0.5, 1.5]))
  assert lrn.ApplyPostTrainingLoop() is None and lrn.optimizer.GetLrScheduleValue() == 1.0
  assert [float(x) for x in lrn.ComputeLosses({'loss': (torch.tensor(2.0), 1.0)})] == [2.0]
  with pytest.raises(ValueError):
    lrn.ComputeLosses({'other': (torch.tensor(2.0), 1.0)})
  # early stop helpers
  mh = early_stop.MetricHistory.Params().Set(jobname='eval', metric='loss',
                                             logdir=str(tmp_path)).Instantiate()
  mh.Append(10, 1.0)
  mh.Append(20, 0.5)
  assert mh.metric == 'loss' and open(mh.hist_file).read().split() == \
      ['10', '1.000000', '20', '0.500000']
  es = early_stop.EarlyStop.Params()
  early_stop.MetricHistory.SetLogdirInMetricHistories(es, '/some/dir')
  assert es.metric_history.logdir == '/some/dir'
  # rnn cell sizes
  cell = rnn_cell.LSTMCellSimple.Params().Set(name='c', num_input_nodes=3,
                                               num_output_nodes=5).Instantiate()
  assert cell.output_size == 5
  y = rnn_cell.RNNCell.LayerNorm(torch.randn(2, 6))
  torch.testing.assert_close(y.mean(-1), torch.zeros(2), atol=1e-5, rtol=0)
