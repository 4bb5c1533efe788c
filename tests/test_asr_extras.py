"""ASR satellites: edit distance, EOS normalisation, fusion, decoder metrics, WER tools."""

import numpy as np
import torch

from lingvo_b200.core import tokenizers
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.asr import contextualizer_base
from lingvo_b200.models.asr import decoder_metrics
from lingvo_b200.models.asr import eos_normalization
from lingvo_b200.models.asr import fusion
from lingvo_b200.models.asr import levenshtein_distance as lev
from lingvo_b200.models.asr import model_test_input_generator
from lingvo_b200.models.asr.tools import custom_html_handlers
from lingvo_b200.models.asr.tools import simple_wer
from lingvo_b200.models.asr.tools import simple_wer_v2


def test_levenshtein_split():
  st = lev.LevenshteinDistance('a b c d'.split(), 'a x c d e'.split())
  assert (st.subs, st.insertions, st.deletions, st.total) == (1, 1, 0, 2)
  st = lev.LevenshteinDistance('a b c'.split(), [])
  assert st.deletions == 3 and st.total == 3
  assert lev.CostTable('kitten', 'sitting')[-1][-1] == 3


def test_eos_normalization():
  ids = torch.tensor([[5, 6, 2, 2, 0, 0], [7, 2, 0, 0, 0, 0], [8, 9, 10, 0, 0, 0]])
  lens = torch.tensor([4, 2, 3])
  out, n = eos_normalization.NormalizeTrailingEos(ids, lens, need_trailing_eos=True, eos_id=2)
  assert n.tolist() == [3, 2, 4]
  assert out[0].tolist() == [5, 6, 2, 2, 2, 2]
  assert out[2].tolist()[:4] == [8, 9, 10, 2]
  out, n = eos_normalization.NormalizeTrailingEos(ids, lens, need_trailing_eos=False, eos_id=2)
  assert n.tolist() == [2, 1, 3]
  o2, n2 = eos_normalization.NumpyNormalizeTrailingEos(ids.numpy(), lens.numpy(), False, 2)
  assert np.array_equal(o2, out.numpy()) and np.array_equal(n2, n.numpy())
  filled = eos_normalization.FillPaddingPos(ids, lens, -1)
  assert filled[1].tolist() == [7, 2, -1, -1, -1, -1]


def test_null_contextualizer_and_fusion():
  c = contextualizer_base.NullContextualizer.Params().Set(name='ctx').Instantiate()
  assert c.GetContextDim() == 0
  ctx = torch.ones(2, 4)
  assert c.QueryAttention(None, None, None, ctx, None) is ctx
  f = fusion.NullFusion.Params().Set(name='fuse').Instantiate()
  logits = torch.randn(2, 7)
  assert f.ComputeLogitsWithLM(NestedMap(), logits) is logits
  out, st = f.FProp(f.theta, NestedMap(), logits, None, None)
  assert out is logits


def test_shallow_fusion_steps_lm():
  from lingvo_b200.models.lm import layers as lm_layers
  lm = lm_layers.RnnLm.CommonParams(vocab_size=11, emb_dim=8, num_layers=1, rnn_dims=8,
                                    rnn_hidden_dims=0)
  p = fusion.ShallowFusion.Params().Set(name='fuse', lm=lm, lm_weight=0.5)
  f = p.Instantiate()
  st0 = f.zero_state(f.theta, 3)
  ids = torch.randint(0, 11, (3, 1))
  am = torch.randn(3, 11)
  _, st1 = f.FProp(f.theta, st0, am, ids, torch.zeros(3, 1))
  fused = f.ComputeLogitsWithLM(st1, am)
  assert fused.shape == (3, 11)
  ref = torch.log_softmax(am, -1) + 0.5 * st1.lm_output
  torch.testing.assert_close(fused, ref)
  assert not st1.lm_output.requires_grad


def test_decoder_metrics_end_to_end():
  tok = tokenizers.AsciiTokenizer.Params().Instantiate()
  def ids_to_strings(ids, lens):
    return tok.IdsToStrings(ids, lens)
  refs = ['hello world', 'good morning all']
  hyps = [['hello world', 'hello word'], ['good morning', 'good morning all']]
  def enc(strs, t=24):
    ids, labels, pads = tok.StringsToIds(strs, t)
    return ids, labels, pads
  _, tl, tp = enc(refs)
  flat = [h for row in hyps for h in row]
  _, hl, hp = enc(flat)
  hlens = (1 - hp).sum(1).long()
  outs = NestedMap(topk_hyps=None, topk_ids=hl, topk_lens=hlens,
                   topk_scores=torch.tensor([[-1.0, -2.0], [-1.5, -1.7]]), topk_decoded=None)
  batch = NestedMap(tgt=NestedMap(ids=tl, labels=tl, paddings=tp, weights=1 - tp))
  dm = decoder_metrics.DecoderMetrics.Params().Instantiate()
  d = dm.ComputeMetrics(outs, batch, ids_to_strings)
  assert d.topk_decoded[0][0] == 'hello world'
  assert d.norm_wer_errors.tolist() == [[0.0, 1.0], [1.0, 0.0]]
  metrics = dm.CreateMetrics()
  kv = dm.PostProcess(d, metrics)
  assert len(kv) == 2
  assert abs(metrics['wer'].value - 1.0 / 5.0) < 1e-6
  assert abs(metrics['oracle_norm_wer'].value) < 1e-6
  assert abs(metrics['sacc'].value - 0.5) < 1e-6
  assert abs(metrics['error_rates/del'].value - 0.2) < 1e-6
  assert metrics['ter'].value > 0


def test_test_input_generator_shapes():
  p = model_test_input_generator.TestInputGenerator.Params().Set(
      source_shape=[4, 12, 8, 1], target_shape=[4, 6])
  p.tokenizer = tokenizers.AsciiTokenizer.Params()
  g = p.Instantiate()
  b1, b2 = g.GetPreprocessedInputBatch(), g.GetPreprocessedInputBatch()
  assert b1.src.src_inputs.shape == (4, 12, 8, 1) and b1.tgt.ids.shape == (4, 6)
  assert not torch.equal(b1.src.src_inputs, b2.src.src_inputs)
  assert (b1.tgt.ids[:, 0] == p.tokenizer.target_sos_id).all()
  assert torch.equal(b1.tgt.ids[:, 1:], b1.tgt.labels[:, :-1])
  pk = p.Copy().Set(target_key='k')
  assert 'k' in pk.Instantiate().GetPreprocessedInputBatch().tgt


def test_simple_wer_tools():
  info, html = simple_wer.ComputeWER('a x c d e', 'a b c d', diagnosis=True)
  assert (info['sub'], info['ins'], info['del'], info['nw']) == (1, 1, 0, 4)
  assert 'yellow' in html and 'green' in html
  kp = simple_wer.AnalyzeKeyPhrases('play some jazz', 'play some jazz now', ['some jazz', 'rock'])
  assert kp['ref_nkp'] == 1 and kp['hyp_nkp'] == 1
  assert simple_wer.RemoveCommentTxtPreprocess('Hello, [noise] World!') == 'hello world'
  w = simple_wer_v2.SimpleWER(key_phrases=['big apple'],
                              html_handler=custom_html_handlers.ChainOfHtmlHandlers(
                                  custom_html_handlers.NewlineHtmlHandler(),
                                  simple_wer_v2.HighlightAlignedHtmlHandler()))
  w.AddHypRef('the big apple is nice', 'the big apple was nice')
  w.AddHypRef('hello <eol> there', 'hello <eol> there')
  wer, parts = w.GetWER()
  assert abs(wer - 100.0 / 8) < 1e-6 and parts['ins'] == 0
  j, f1, p_, r_ = w.GetKeyPhraseStats()
  assert (j, f1, p_, r_) == (1.0, 1.0, 1.0, 1.0)
  assert w.GetMostFrequentErrPatterns()['sub'][0][0] == ('was', 'is')
  assert '<br>' in w.aligned_htmls[1] or 'eol' in w.aligned_htmls[1]


def test_metrics_calculator_aggregates_error_rates():
  import collections
  import numpy as np
  from lingvo_b200.core import metrics
  from lingvo_b200.models.asr import metrics_calculator as mc
  d = collections.defaultdict(metrics.AverageMetric)
  pi = mc.PostProcessInputs(
      transcripts=['the cat sat', 'Hello World'],
      topk_decoded=[['the cat sat', 'the cat'], ['hello word', 'Hello World']],
      filtered_transcripts=['the cat sat', 'Hello World'],
      filtered_top_hyps=['the cat sat', 'hello word'],
      topk_scores=[[-0.1, -0.5], [-0.2, -0.3]], utt_id=['u0', 'u1'],
      norm_wer_errors=[[0, 1], [2, 0]],                       # per-hyp word errors
      target_labels=np.array([[5, 6, 7, 2], [8, 9, 2, 0]]),
      target_paddings=np.array([[0, 0, 0, 0], [0, 0, 0, 1]]),
      topk_ids=np.array([[5, 6, 7, 2], [5, 6, 2, 0], [8, 4, 2, 0], [8, 9, 2, 0]]),
      topk_lens=np.array([4, 3, 3, 3]))
  mc.CalculateMetrics(pi, d)
  # cased: utt0 exact, utt1 "hello word" vs "Hello World" = 2 substitutions over 5 ref words
  assert abs(d['wer'].value - 2 / 5) < 1e-9
  assert abs(d['error_rates/sub'].value - 2 / 5) < 1e-9 and d['error_rates/ins'].value == 0
  # case-insensitive: only "word" vs "world" is wrong
  assert abs(d['case_insensitive_error_rates/wer'].value - 1 / 5) < 1e-9
  # oracle picks the best hypothesis of each list: 0 + 0 errors
  assert d['oracle_norm_wer'].value == 0.0
  assert abs(d['sacc'].value - 0.5) < 1e-9                     # top hyp exact for 1 of 2 utterances
  # token error rate: utt0 exact (4 tokens), utt1 one substitution over 3 tokens
  assert abs(d['ter'].value - 1 / 7) < 1e-9
  assert mc.GetRefIds([1, 2, 3], [0, 1, 0]) == [1, 3]


def test_simple_wer_v1_entry_points(capsys):
  from lingvo_b200.models.asr.tools import simple_wer
  d = simple_wer.ComputeEditDistanceMatrix('a b c d'.split(), 'a x c'.split())
  assert d.shape == (4, 5) and int(d[-1, -1]) == 2 and d[0].tolist() == [0, 1, 2, 3, 4]
  assert simple_wer.PreprocessTxtBeforeWER('Hello [noise]  World - again\n') == 'hello world again'
  errs, nref, html = simple_wer.AverageWERs(['the cat sat', 'a b'], ['the cat sat down', 'a c'],
                                            diagnosis=True)
  assert errs == {'sub': 1, 'ins': 0, 'del': 1} and nref == 6 and len(html) == 2
  out = capsys.readouterr().out
  assert 'total error = 2, total word = 6, wer = 33.33%' in out and 'Error breakdown' in out
  s, det = simple_wer.GenerateSummaryFromErrs(10, {'sub': 1, 'ins': 2, 'del': 0})
  assert s.endswith('wer = 30.00%') and 'ins=20.00%' in det
