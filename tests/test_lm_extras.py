"""LM breadth: conditional / mixture RNN LMs, BERT tokenizer + input + task, gshard
incremental decoding, lifelong / tunable transformers, decode tool."""

import numpy as np
import torch

from lingvo_b200 import model_registry
from lingvo_b200 import ops
from lingvo_b200.core import cluster_factory
from lingvo_b200.core import gshard_builder as gb
from lingvo_b200.core import gshard_decode
from lingvo_b200.core import rnn_cell
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.lm import input_generator
from lingvo_b200.models.lm import layers as lm_layers
from lingvo_b200.models.lm import tokenizer as lm_tokenizer
from lingvo_b200.utils import tf_example


def _Labels(t, b, v):
  return NestedMap(class_ids=torch.randint(0, v, (t, b)), class_weights=torch.ones(t, b))


def test_conditional_rnn_lm():
  p = lm_layers.ConditionalRnnLm.Params().Set(name='clm', vocab_size=20, condition_dim=4)
  p.emb.Set(vocab_size=20, embedding_dim=6)
  p.rnns.Set(num_layers=1, num_input_nodes=10, num_output_nodes=8,
             cell_tpl=rnn_cell.LSTMCellSimple.Params())
  p.softmax.Set(input_dim=8, num_classes=20)
  lm = p.Instantiate()
  t, b = 5, 3
  ids, pad = torch.randint(0, 20, (t, b)), torch.zeros(t, b)
  c1, c2 = torch.randn(b, 4), torch.randn(b, 4)
  o1, _ = lm.FProp(lm.theta, ids, pad, lm.zero_state(lm.theta, b), c1, _Labels(t, b, 20))
  o2, _ = lm.FProp(lm.theta, ids, pad, lm.zero_state(lm.theta, b), c2, None)
  assert o1.logits.shape == (t, b, 20) and torch.isfinite(o1.avg_xent)
  assert not torch.allclose(o1.logits, o2.logits)     # the condition matters


def test_moe_lm():
  v = 16
  p = lm_layers.MoeLm.Params().Set(name='moelm', vocab_size=v, number_of_experts=3)
  p.emb.Set(vocab_size=v, embedding_dim=6)
  p.rnns.Set(num_layers=1, num_input_nodes=6, num_output_nodes=8,
             cell_tpl=rnn_cell.LSTMCellSimple.Params())
  p.merge.Set(vocab_size=v)
  p.merge.rnns.Set(num_layers=1, num_input_nodes=8, num_output_nodes=8,
                   cell_tpl=rnn_cell.LSTMCellSimple.Params())
  p.merge.softmax.Set(input_dim=8, num_classes=v)
  lm = p.Instantiate()
  t, b = 4, 2
  ids, pad = torch.randint(0, v, (t, b)), torch.zeros(t, b)
  out, st = lm.FProp(lm.theta, ids, pad, lm.zero_state(lm.theta, b), _Labels(t, b, v))
  assert out.gating.shape == (t, b, 3)
  torch.testing.assert_close(out.gating.sum(-1), torch.ones(t, b))
  assert len(st.rnns) == 4 and 'merge' in st
  out.avg_xent.backward()
  assert lm.domain_predictor_softmax.vars.Flatten()[0].grad is not None


def test_bert_tokenizer(tmp_path):
  pieces = ['[PAD]', '[UNK]', '[CLS]', '[SEP]', '[MASK]', 'un', '##aff', '##able', 'hello',
            ',', 'wor', '##ld', '!']
  vf = tmp_path / 'vocab.txt'
  vf.write_text('\n'.join(pieces) + '\n')
  p = lm_tokenizer.BertTokenizer.Params().Set(
      vocab_filepath=str(vf), target_sos_id=2, target_eos_id=3, target_unk_id=1)
  tok = p.Instantiate()
  assert tok.Encode('Unaffable, HELLO world!') == [5, 6, 7, 9, 8, 10, 11, 12]
  assert tok.Encode('xyz') == [1]
  ids, labels, pads = tok.StringsToIds(['hello world'], 8)
  assert ids[0, :4].tolist() == [2, 8, 10, 11] and labels[0, :4].tolist() == [8, 10, 11, 3]
  assert tok.IdsToStrings(labels, torch.tensor([3])) == ['hello world']


def _BertRecords(path, n=24, t=16, k=4, vocab=128):
  rng = np.random.RandomState(0)
  w = ops.host().TFRecordWriter(str(path))
  for _ in range(n):
    la, lb = rng.randint(2, 5), rng.randint(2, 5)
    toks = np.concatenate([[101], rng.randint(5, vocab, la), [102], rng.randint(5, vocab, lb),
                           [102]])
    ids = np.zeros(t, np.int64); ids[:len(toks)] = toks
    mask = np.zeros(t, np.int64); mask[:len(toks)] = 1
    pos = np.sort(rng.choice(np.arange(1, len(toks) - 1), 2, replace=False))
    pos = pos[toks[pos] != 102]
    mpos = np.zeros(k, np.int64); mpos[:len(pos)] = pos
    mids = np.zeros(k, np.int64); mids[:len(pos)] = toks[pos]
    mw = np.zeros(k, np.float32); mw[:len(pos)] = 1
    ids[pos] = 103
    w.write(tf_example.MakeExample({'input_ids': ids, 'input_mask': mask,
                                    'masked_lm_positions': mpos, 'masked_lm_ids': mids,
                                    'masked_lm_weights': mw}))
  w.close()


def test_tfrecord_bert_input_and_task(tmp_path):
  f = tmp_path / 'bert.tfrecord'
  _BertRecords(f)
  for packing in (False, True):
    p = input_generator.TFRecordBertInput.Params().Set(
        name='inp', input_file=str(f), max_sequence_length=16, max_predictions_per_seq=4,
        batch_size=4, enable_packing=packing, prepacking_batch_size=12, shuffle=packing)
    b = p.Instantiate().GetPreprocessedInputBatch()
    assert b.ids.shape == (4, 16) and b.masked_ids.shape == (4, 16)
    m = b.masked_pos > 0
    assert (b.masked_ids[m] == 103).all() and (b.ids[m] != 103).all()
    assert (b.ids[~m] == b.masked_ids[~m]).all()
    real = b.segment_ids > 0
    assert (b.ids[real] == 102).sum() == (b.segment_ids.max(1).values.sum() if packing else 4)
    if packing:
      assert float(b.segment_ids.max()) >= 2
  import lingvo_b200.models.lm.params.wiki_bert  # noqa: F401
  mp = model_registry.GetParams('lm.wiki_bert.BertDenseTiny', 'Train')
  mp.task.fprop_dtype = torch.float32
  mp.task.builder.fprop_dtype = torch.float32
  with cluster_factory.ForTestingWorker(mode='sync', job='trainer_client'):
    task = mp.task.Instantiate()
  b16 = b.Transform(lambda x: x)
  pred = task.ComputePredictions(task.theta, b16)
  metrics, _ = task.ComputeLoss(task.theta, pred, b16)
  assert float(metrics['num_masked'][0]) == float(m.sum())
  metrics['loss'][0].backward()
  ps = model_registry.GetProgramSchedule('lm.wiki_bert.MLPerfTrainBertDense2B')
  assert ps.ml_perf.benchmark_name == 'bert' and ps.ml_perf.decoder_metric_name == 'acc1'


def _Builder(cls=gb.DenseBuilder, **kw):
  return cls.Params().Set(
      model_dim=16, attention_num_heads=2, attention_key_value_dim=8, ff_dim=32, e_dim=4,
      c_dim=0, capacity_factor=2.0, moe_hidden_dim=32, relative_attention_type='bias',
      relative_attention_num_buckets=8, relative_attention_max_distance=16,
      relative_attention_use_universal_1d_position=True, **kw)


def _LmBatch(b=3, t=9, v=50):
  ids = torch.randint(2, v, (b, t))
  return NestedMap(ids=ids, labels=ids, paddings=torch.zeros(b, t),
                   segment_ids=torch.ones(b, t, dtype=torch.long),
                   segment_pos=torch.arange(t).repeat(b, 1))


def test_gshard_incremental_decode_matches_full():
  p = gb.UniTransformer.Params().Set(
      name='lm', builder=_Builder(), vocab_size=50, num_transformer_layers=2, max_length=32,
      positional_embedding=False, label_smoothing=0.0, z_loss=0.0, decoder_max_steps=5)
  with cluster_factory.ForTestingWorker(mode='sync', job='trainer_client', do_eval=True):
    task = p.Instantiate()
  batch = _LmBatch()
  with torch.no_grad():
    pred = task.ComputePredictions(task.theta, batch)
    full = task._ComputeLogits(task.theta, pred.dec_outputs).float()
    st = task.InitDecodeState(3, 12, batch.ids.device)
    for i in range(9):
      lg = task.DecodeStep(task.theta, batch.ids[:, i], st, i)
      torch.testing.assert_close(lg, full[:, i], atol=1e-4, rtol=1e-4)
  out = gshard_decode.DecodeIds(task, task.theta, batch)
  assert out.ids.shape == (3, 14) and (out.ids[:, :9] == batch.ids).all()
  assert (out.lens >= out.prefix_lens).all() and (out.scores <= 0).all()
  # greedy continuation = argmax of the full forward on the extended sequence
  ext = NestedMap(ids=out.ids[:, :10], labels=out.ids[:, :10], paddings=torch.zeros(3, 10),
                  segment_ids=torch.ones(3, 10, dtype=torch.long),
                  segment_pos=torch.arange(10).repeat(3, 1))
  with torch.no_grad():
    lg = task._ComputeLogits(task.theta, task.ComputePredictions(task.theta, ext).dec_outputs)
  assert (lg[:, 8].argmax(-1) == out.ids[:, 9]).all()


def test_tunable_and_lifelong_transformers():
  p = gb.TunableUniTransformer.Params().Set(
      name='tun', builder=_Builder(), vocab_size=50, num_transformer_layers=3, max_length=32,
      positional_embedding=False, top_layer_types=['attn', 'ffw'],
      sub_layer_types=['attn', 'moe'], bottom_layer_types=['attn', 'ffw'])
  with cluster_factory.ForTestingWorker(mode='sync', job='trainer_client'):
    task = p.Instantiate()
  kinds = [type(blk.layer).__name__ for blk in task.dec.layers]
  assert kinds == ['SelfAttentionLayer', 'DenseReluDenseLayer', 'SelfAttentionLayer', 'MoELayer',
                   'SelfAttentionLayer', 'DenseReluDenseLayer']
  lb = _Builder(gb.DenseLifelongBuilder, e_dim_old=2, lwf_scale=0.5)
  p = gb.LifelongUniTransformer.Params().Set(
      name='life', builder=lb, vocab_size=50, num_transformer_layers=2, max_length=32,
      positional_embedding=False, moe=True)
  with cluster_factory.ForTestingWorker(mode='sync', job='trainer_client'):
    task = p.Instantiate()
  assert task.old.dec.layers[1].layer.vars.gw.shape[1] == 2
  assert task.dec.layers[1].layer.vars.gw.shape[1] == 4
  batch = _LmBatch()
  metrics, _ = task.ComputeLoss(task.theta, task.ComputePredictions(task.theta, batch), batch)
  assert 'lwf_loss' in metrics and float(metrics['lwf_loss'][0]) >= 0
  metrics['loss'][0].backward()
  assert all(v.grad is None for v in task.old.vars.Flatten())
  g = gb.DenseLifelongBuilder.ExpandGate(torch.randn(16, 2), 4)
  w = gb.DenseLifelongBuilder.ExpandExperts(torch.randn(2, 16, 32), 4)
  assert g.shape == (16, 4) and w.shape == (4, 16, 32) and torch.equal(w[2], w[0])


def test_gshard_lm_decode_tool(tmp_path):
  from lingvo_b200.models.lm.tools import gshard_lm_decode
  import lingvo_b200.models.lm.params.synthetic_packed_input  # noqa: F401
  f = tmp_path / 'prompts.tsv'
  f.write_text('5 6 7\n8 9\n10 11 12 13\n')
  dec = gshard_lm_decode.GShardLMDecodeBatch(
      'lm.synthetic_packed_input.DenseLmTiny', prefix_max_len=8, batch_size=2,
      max_decode_steps=4, device='cpu')
  outs = dec.DecodeFiles([str(f)], str(tmp_path / 'out'))
  dec.stop()
  lines = open(outs[0]).read().strip().split('\n')
  assert len(lines) == 3 and lines[0].split('\t')[0] == '5 6 7'
  assert len(lines[0].split('\t')[1].split()) >= 1
  assert dec.DecodeFiles([str(f)], str(tmp_path / 'out')) == []      # restart-safe skip


def test_lm_layer_state_combination_and_xent_output():
  import torch
  from lingvo_b200.core import layers as core_layers
  from lingvo_b200.core.nested_map import NestedMap
  from lingvo_b200.models.lm import layers as lm_layers
  p = lm_layers.RnnLm.CommonParams(vocab_size=11, emb_dim=4, num_layers=2, rnn_dims=6)
  p.name = 'lm'
  lm_layers.RnnLm.UpdateTargetVocabSize(p, 13)
  assert p.vocab_size == 13 and p.softmax.num_classes == 13 and p.emb.vocab_size == 13
  lm = p.Instantiate()
  s0 = lm.zero_state(lm.theta, 3)
  s1 = s0.Transform(lambda x: x + 1.0)
  mixed = lm.CombineStates(s0, s1, torch.tensor([True, False, True]))
  for a in mixed.Flatten():
    assert a[0].abs().sum() == 0 and a[2].abs().sum() == 0 and bool((a[1] == 1).all())
  assert lm.GetFeedDict() == {}
  sm = core_layers.SimpleFullSoftmax.Params().Set(name='sm', input_dim=5, num_classes=7).Instantiate()
  acts = torch.randn(4, 6, 5)                         # batch 3 × 2 samples
  only_logits = lm_layers.ComputeXentOutput(sm, sm.theta, acts, None)
  assert only_logits.logits.shape == (4, 6, 7)
  labels = NestedMap(class_ids=torch.randint(0, 7, (4, 3)), class_weights=torch.ones(4, 3))
  out = lm_layers.ComputeXentOutput(sm, sm.theta, acts, labels, num_samples=2)
  want = torch.nn.functional.cross_entropy(
      sm.Logits(sm.theta, acts.reshape(24, 5)).float(),
      labels.class_ids.repeat(1, 2).reshape(-1), reduction='mean')
  torch.testing.assert_close(out.avg_xent, want, atol=1e-5, rtol=1e-5)
  probs = torch.softmax(torch.randn(4, 3, 7), -1)
  out2 = lm_layers.ComputeXentOutput(
      sm, sm.theta, acts[:, :3], NestedMap(class_probabilities=probs,
                                           class_weights=torch.ones(4, 3)))
  assert out2.per_example_xent.shape[0] == 12


def test_new_dense_lm_configs_and_sharded_adam_micro_batches():
  import numpy as np
  import torch
  from lingvo_b200 import model_registry
  from lingvo_b200.core import py_utils
  from lingvo_b200.models.lm.params import synthetic_packed_input as sp
  for name, mesh in [('DenseLm175B1K', [64, 16]), ('DenseLm175B8x8Decode2D', [8, 16]),
                     ('DenseLm12kWide162BAdamBS25616x16', None),
                     ('DenseLm12kWide162BAdam32x32', [64, 32])]:
    cls = getattr(sp, name)
    mp = model_registry.GetParams('lm.synthetic_packed_input.' + name, 'Train')
    if mesh is not None:
      assert list(cls.DEVICE_MESH.shape) == mesh
    assert mp.task.builder.model_dim == cls.MODEL_DIM
  dec = model_registry.GetParams('lm.synthetic_packed_input.DenseLm175B8x8Decode2D', 'Train')
  assert dec.task.builder.relative_attention_use_universal_1d_position is False
  assert dec.task.builder.model_dim_reshape_segments == 8 and dec.task.builder.emb_w_split == [1, 0]
  bs = model_registry.GetParams('lm.synthetic_packed_input.DenseLm12kWide162BAdamBS25616x16',
                                'Train')
  assert bs.task.train.optimizer.num_micro_batches == 4
  big = sp.DenseLm12kWide162BAdam32x32
  assert big.DEVICE_MESH[1, 0] == 1 and big.DEVICE_MESH[0, 1] == 64     # transposed device order
  assert sp.ShardedAdam is sp.ShardedAdamOptimizer
  # 2 micro-batches: the update happens on the second Apply, with the averaged gradient
  opt = sp.ShardedAdam.Params().Set(name='a', num_micro_batches=2, beta1=0.0, beta2=0.0,
                                    epsilon=1e-8).Instantiate()
  ref = sp.ShardedAdam.Params().Set(name='b', beta1=0.0, beta2=0.0, epsilon=1e-8).Instantiate()
  w = torch.nn.Parameter(torch.ones(3)); w2 = torch.nn.Parameter(torch.ones(3))
  g1, g2 = torch.tensor([1.0, 2.0, 3.0]), torch.tensor([3.0, 2.0, -1.0])
  opt.Apply(0.1, [py_utils.VarGrad(w, g1)])
  assert torch.equal(w.detach(), torch.ones(3))
  opt.Apply(0.1, [py_utils.VarGrad(w, g2)])
  ref.Apply(0.1, [py_utils.VarGrad(w2, (g1 + g2) / 2)])
  torch.testing.assert_close(w.detach(), w2.detach())
  assert float(opt._Slot(w, 'grad_accum').abs().sum()) == 0.0
  del np
