"""MT breadth: extra encoders/decoders, XEnDec, insertion model, input variants."""

import numpy as np
import pytest
import torch

from lingvo_b200 import model_registry
from lingvo_b200 import ops
from lingvo_b200.core import cluster_factory
from lingvo_b200.core import layers
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.mt import base_config
from lingvo_b200.models.mt import data_augmenter
from lingvo_b200.models.mt import decoder as mt_decoder
from lingvo_b200.models.mt import encoder as mt_encoder
from lingvo_b200.models.mt import input_generator
from lingvo_b200.models.mt import layers as mt_layers
from lingvo_b200.models.mt import model as mt_model
from lingvo_b200.utils import tf_example

V = 20


def _Batch(b=4, s=7, t=6, seed=0):
  g = torch.Generator().manual_seed(seed)
  src = torch.randint(3, V, (b, s), generator=g)
  lab = torch.randint(3, V, (b, t), generator=g)
  ids = torch.cat([torch.ones(b, 1, dtype=torch.long), lab[:, :-1]], 1)
  sp = torch.zeros(b, s); sp[0, -2:] = 1
  tp = torch.zeros(b, t); tp[1, -1:] = 1
  return NestedMap(src=NestedMap(ids=src, paddings=sp),
                   tgt=NestedMap(ids=ids, labels=lab, paddings=tp, weights=1 - tp))


def test_transformer_stack_transparent():
  p = mt_layers.TransformerStack.Params().Set(
      name='st', model_dim=16, num_transformer_layers=2, is_transparent=True,
      num_transparent_outputs=3, ln_output=True)
  p.transformer_tpl.tr_atten_tpl.num_attention_heads = 2
  p.transformer_tpl.tr_fflayer_tpl.hidden_dim = 32
  st = p.Instantiate()
  x, pad = torch.randn(5, 3, 16), torch.zeros(5, 3)
  outs, _, seg = st.FProp(st.theta, x, pad)
  assert isinstance(outs, list) and len(outs) == 3 and outs[0].shape == (5, 3, 16)
  assert seg is None


@pytest.mark.parametrize('cls', [mt_encoder.MTEncoderV1, mt_encoder.MTEncoderUniRNN,
                                 mt_encoder.MTEncoderBiRNNPrecomputedEmbedding])
def test_rnn_encoders(cls):
  p = cls.Params().Set(name='enc', lstm_cell_size=8, num_lstm_layers=3)
  p.emb.Set(vocab_size=V, embedding_dim=8)
  if 'encoder_out_dim' in p:
    p.encoder_out_dim = 12
  enc = p.Instantiate()
  b = _Batch().src
  if cls is mt_encoder.MTEncoderBiRNNPrecomputedEmbedding:
    b.embs = torch.randn(4, 7, 8)
  out = enc.FProp(enc.theta, b)
  assert out.encoded.shape[:2] == (7, 4) and out.padding.shape == (7, 4)
  assert float(out.encoded[-1, 0].abs().sum()) == 0.0          # padded frame zeroed
  out.encoded.sum().backward()


def test_batch_major_encoder_decoder_formats():
  ep = base_config.SetupTransformerEncoder(16, V, 1, 2, 32)
  bp = mt_encoder.TransformerBatchMajorEncoder.Params()
  for k, v in ep.IterParams():
    if k in bp and k not in ('cls',):
      bp.Set(**{k: v})
  bp.Set(name='enc', final_layer_norm=True, output_data_format='BTC')
  enc = bp.Instantiate()
  out = enc.FProp(enc.theta, _Batch().src)
  assert out.encoded.shape == (4, 7, 16) and out.padding.shape == (4, 7)


def _XEnDecParams():
  p = base_config.SetupXEnDecTransformerParams(
      mt_model.TransformerXEnDecModel.Params(), name='xendec', vocab_size=V, model_dim=16,
      hidden_dim=32, num_heads=2, num_layers=1, learning_rate=1e-3, warmup_steps=10)
  return p


def test_xendec_losses_and_grads():
  p = _XEnDecParams()
  with cluster_factory.ForTestingWorker(mode='sync', job='trainer_client'):
    task = p.Instantiate()
  batch = _Batch()
  batch.src.source_mask = (torch.rand(4, 7) < 0.4).float()
  pred = task.ComputePredictions(task.theta, batch)
  assert pred.attention.probs.shape == (4, 6, 7)
  metrics, per = task.ComputeLoss(task.theta, pred, batch)
  for k in ('clean_loss', 'other_loss', 'mix_loss', 'loss'):
    assert k in metrics and torch.isfinite(metrics[k][0])
  total = metrics.clean_loss[0] + metrics.other_loss[0] + metrics.mix_loss[0]
  torch.testing.assert_close(metrics.loss[0], total)
  metrics.loss[0].backward()
  g = task.enc.token_emb.vars.wm.grad
  assert g is not None and float(g.abs().sum()) > 0


def test_xendec_target_lambdas():
  att = [torch.full((2, 3, 4), 0.25), torch.full((2, 3, 4), 0.25)]
  src_lam = [torch.tensor([[1., 1, 0, 0]] * 2), torch.tensor([[0., 0, 1, 1]] * 2)]
  zs, zt = torch.zeros(2, 4), torch.zeros(2, 3)
  _, inp, lab = mt_model.TransformerXEnDecModel._CreateTargetLambdas(
      att, src_lam, [zs, zs], [zt, zt])
  torch.testing.assert_close(lab[0], torch.full((2, 3), 0.5))
  assert inp[0][:, 0].tolist() == [1.0, 1.0]          # first decoder input = own SOS


def test_insertion_model_trains_one_step():
  p = mt_model.InsertionModel.Params().Set(name='ins')
  p.decoder.Set(model_dim=16, num_trans_layers=1)
  p.decoder.token_emb.Set(vocab_size=2 * V)
  p.decoder.softmax.num_classes = V
  p.decoder.trans_tpl.tr_atten_tpl.num_attention_heads = 2
  p.decoder.trans_tpl.tr_fflayer_tpl.hidden_dim = 32
  with cluster_factory.ForTestingWorker(mode='sync', job='trainer_client'):
    task = p.Instantiate()
  batch = _Batch()
  pred = task.ComputePredictions(task.theta, batch)
  n = pred.tgt.target_indices.shape[0]
  assert pred.tgt.target_weights.shape == (n,)
  assert int(pred.tgt.target_indices[:, 1].max()) < pred.outputs.shape[1]
  metrics, aux = task.ComputeLoss(task.theta, pred, batch)
  assert float(metrics['loss'][0]) > 0
  metrics['loss'][0].backward()


def test_mass_layer():
  m = data_augmenter.MASS.Params().Set(mask_id=3, mask_ratio=0.5, vocab_size=V).Instantiate()
  ids = torch.randint(4, V, (3, 10))
  out = m.Mask(ids, torch.ones(3, 10), torch.tensor([10, 8, 6]))
  assert out.src.ids.shape == (3, 10) and out.tgt.labels.shape == (3, 10)
  n_masked = (out.tgt.weights > 0).sum(1)
  assert n_masked.tolist() == [5, 4, 3]
  assert (out.tgt.labels[out.tgt.weights > 0] == ids[out.tgt.weights > 0]).all()


def test_mlperf_and_double_input(tmp_path):
  rng = np.random.RandomState(0)
  w = ops.host().TFRecordWriter(str(tmp_path / 'ml.tfrecords'))
  w2 = ops.host().TFRecordWriter(str(tmp_path / 'nmt.tfrecords'))
  for _ in range(64):
    n = rng.randint(3, 8)
    src = rng.randint(3, V, n)
    w.write(tf_example.MakeExample({'inputs': np.append(src, 2), 'targets': np.append(src, 2)}))
    w2.write(tf_example.MakeExample({
        'source_id': src, 'source_padding': np.zeros(n, np.float32),
        'target_id': np.concatenate([[1], src]), 'target_padding': np.zeros(n + 1, np.float32),
        'target_label': np.concatenate([src, [2]]), 'target_weight': np.ones(n + 1, np.float32)}))
  w.close(); w2.close()
  vocab = tmp_path / 'vocab.txt'
  vocab.write_text('\n'.join(['<unk>', '<s>', '</s>'] + ['w%d' % i for i in range(3, V)]))
  def _Common(p, f):
    p.Set(name='inp', file_pattern='tfrecord:' + str(tmp_path / f), bucket_upper_bound=[10],
          bucket_batch_limit=[8], file_buffer_size=16, file_parallelism=1,
          num_batcher_threads=1)
    p.tokenizer.token_vocab_filepath = str(vocab)
    p.tokenizer.load_token_ids_from_vocab = False
    p.tokenizer.vocab_size = V
    return p
  ml = _Common(input_generator.MlPerfInput.Params(), 'ml.tfrecords').Instantiate()
  b = ml.GetPreprocessedInputBatch()
  assert b.tgt.ids.shape == b.tgt.labels.shape
  assert (b.tgt.ids[:, 0] == 0).all()
  real = b.tgt.paddings[:, 1:] < 0.5
  assert torch.equal(b.tgt.ids[:, 1:][real], b.tgt.labels[:, :-1][real])
  dp = _Common(input_generator.NmtDoubleInput.Params(), 'nmt.tfrecords').Set(
      source_mask_ratio=0.5, permutation_distance=2, mask_words_ratio=0.3)
  b = dp.Instantiate().GetPreprocessedInputBatch()
  assert b.src.source_mask.shape == b.src.ids.shape
  assert float((b.src.source_mask * b.src.paddings).sum()) == 0.0
  assert 'other_src' in b and b.other_src.ids.shape == b.src.ids.shape
  assert (b.other_src.ids == 3).any()


def test_text_packed_input(tmp_path):
  lines = ['w%d w%d w%d\tw%d w%d' % (3 + i % 5, 4 + i % 7, 5 + i % 3, 6 + i % 4, 7 + i % 6)
           for i in range(64)]
  f = tmp_path / 'pairs.tsv'
  f.write_text('\n'.join(lines) + '\n')
  vocab = tmp_path / 'vocab.txt'
  vocab.write_text('\n'.join(['<unk>', '<s>', '</s>'] + ['w%d' % i for i in range(3, V)]))
  p = input_generator.TextPackedInput.Params().Set(
      name='inp', file_pattern='text:' + str(f), bucket_upper_bound=[10],
      bucket_batch_limit=[16], file_buffer_size=16, file_parallelism=1, num_batcher_threads=1,
      packing_factor=4.0, source_max_length=24, target_max_length=24,
      file_pattern_task_ids=[0], task_to_src_lang_map=[1], task_to_tgt_lang_map=[2])
  p.tokenizer.token_vocab_filepath = str(vocab)
  p.tokenizer.load_token_ids_from_vocab = False
  p.tokenizer.vocab_size = V
  b = p.Instantiate().GetPreprocessedInputBatch()
  assert b.src.ids.shape == (4, 24) and b.tgt.ids.shape == (4, 24)
  assert int(b.src.segment_ids.max()) >= 3          # several sentences per row
  row = b.tgt.segment_ids[0]
  first = (b.tgt.segment_pos[0] == 0) & (row > 0)
  assert (b.tgt.ids[0][first] == 1).all()            # every packed target starts with <s>
  assert (b.src.source_ids[b.src.segment_ids > 0] == 1).all()


def test_registered_mt_params():
  import lingvo_b200.models.mt.params.params  # noqa: F401
  for name in ('mt.wmtm16_en_de.WmtCaptionEnDeTransformer',
               'mt.wmtm16_en_de.WmtCaptionEnDeTransformerCloudTpu',
               'mt.xendec.wmt14_en_de.WmtEnDeXEnDec', 'mt.xendec.wmt14_en_de.WmtDeEnXEnDec'):
    mp = model_registry.GetParams(name, 'Train')
    assert mp.task.name and mp.input.file_pattern
