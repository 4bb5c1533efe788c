"""Speech decoder metrics layer (ref `lingvo/tasks/asr/decoder_metrics.py`).

`DecoderMetrics` turns a beam-search output into the decode dictionary
(`ComputeMetrics`: top-k strings, filtered strings, per-hypothesis normalised word
errors) and scores it on the host (`PostProcess` → `metrics_calculator`).
Tokens stay on the device until `ComputeMetrics` does one D2H copy of ids/lens/scores.
"""

from __future__ import annotations

import collections

import numpy as np
import torch

from lingvo_b200.core import base_layer
from lingvo_b200.core import metrics as metrics_lib
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.asr import decoder_utils
from lingvo_b200.models.asr import metrics_calculator

DecoderTopK = collections.namedtuple(
    'DecoderTopK', ['hyps', 'ids', 'lens', 'scores', 'decoded', 'alignment'])


def BeamSearchDecodeOutputToDecoderTopK(decoder_outs, *, ids_to_strings_fn,
                                        feed_encoder_outs=False, encoder_outs=None, tag=''):
  """Detokenises the top-k ids (dropping each hypothesis' final EOS) (ref :41)."""
  ids, lens, scores = decoder_outs.topk_ids, decoder_outs.topk_lens, decoder_outs.topk_scores
  decoded = decoder_outs.topk_decoded
  if ids is not None:
    body = (lens - 1).clamp_min(0)
    if feed_encoder_outs:
      flat = ids_to_strings_fn(ids, body, encoder_outs=encoder_outs)
    else:
      flat = ids_to_strings_fn(ids, body)
    k = scores.shape[-1] if scores is not None and scores.dim() == 2 else 1
    decoded = [list(flat[i:i + k]) for i in range(0, len(flat), k)]
  return DecoderTopK(decoder_outs.topk_hyps, ids, lens, scores, decoded,
                     getattr(decoder_outs, 'topk_alignment', {}))


class DecoderMetrics(base_layer.BaseLayer):
  """ref :87."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('include_auxiliary_metrics', True,
             'Also compute oracle WER, SACC, TER (slower decode post-processing).')
    p.Define('log_utf8', False, 'Decode bytes to UTF-8 when logging.')
    p.Define('only_output_tpu_tensors', False,
             'Keep ComputeMetrics device-only; PostProcess detokenises.')
    p.Define('pass_through_transcript_field', None,
             'Read reference transcripts from this input field instead of detokenising tgt.')
    p.name = 'decoder_metrics'
    return p

  def __init__(self, params):
    if not params.name:
      raise ValueError('params.name not set.')
    super().__init__(params)

  def GetTopK(self, decoder_outs, ids_to_strings_fn, feed_encoder_outs=False,
              encoder_outs=None, tag=''):
    return BeamSearchDecodeOutputToDecoderTopK(
        decoder_outs, ids_to_strings_fn=ids_to_strings_fn,
        feed_encoder_outs=feed_encoder_outs, encoder_outs=encoder_outs, tag=tag)

  def ComputeNormalizedWER(self, hyps, refs, num_hyps_per_beam):
    """hyps: [B][K] strings, refs: [B] → float array [B, K] of word errors (ref :130)."""
    out = np.zeros((len(refs), num_hyps_per_beam), np.float32)
    for i, ref in enumerate(refs):
      for n, h in enumerate(hyps[i][:num_hyps_per_beam]):
        out[i, n] = decoder_utils.EditDistance(ref, h)[3]
    return out

  def AddAdditionalDecoderMetricsToGraph(self, topk_hyps, filtered_hyps, filtered_refs,
                                         input_batch, decoder_outs):
    return {}

  def ComputeMetrics(self, decoder_outs, input_batch, ids_to_strings_fn):
    """→ decode dictionary (ref :151)."""
    p = self.params
    tgt = input_batch.tgt
    topk = self.GetTopK(decoder_outs, ids_to_strings_fn)
    k = len(topk.decoded[0]) if topk.decoded else 1
    tgt_lens = (1.0 - tgt.paddings.float()).sum(1).long()
    if p.pass_through_transcript_field:
      refs = list(input_batch.Get(p.pass_through_transcript_field))
    else:
      refs = list(ids_to_strings_fn(tgt.labels, (tgt_lens - 1).clamp_min(0)))
    clean = lambda s: decoder_utils.FilterNoise(decoder_utils.FilterEpsilon(s))
    f_refs = [clean(r) for r in refs]
    f_hyps = [[clean(h) for h in row] for row in topk.decoded]
    ret = NestedMap(
        target_ids=tgt.ids, target_labels=tgt.labels, target_weights=tgt.weights,
        target_paddings=tgt.paddings, transcripts=refs, topk_decoded=topk.decoded,
        topk_ids=topk.ids, topk_lens=topk.lens, topk_scores=topk.scores,
        filtered_transcripts=f_refs, filtered_top_hyps=[row[0] for row in f_hyps],
        norm_wer_errors=self.ComputeNormalizedWER(f_hyps, f_refs, k),
        utt_id=input_batch.get('sample_ids'))
    ret.update(self.AddAdditionalDecoderMetricsToGraph(topk, f_hyps, f_refs, input_batch,
                                                       decoder_outs))
    return ret

  def CreateMetrics(self):
    """ref :246."""
    names = ['num_samples_in_batch', 'wer', 'norm_wer', 'oracle_norm_wer', 'sacc', 'ter',
             'error_rates/ins', 'error_rates/sub', 'error_rates/del', 'error_rates/wer',
             'case_insensitive_error_rates/ins', 'case_insensitive_error_rates/sub',
             'case_insensitive_error_rates/del', 'case_insensitive_error_rates/wer']
    m = {n: metrics_lib.AverageMetric() for n in names}
    m['corpus_bleu'] = metrics_lib.CorpusBleuMetric(separator_type='')
    return m

  def FilterRealExamples(self, dec_out_dict):
    """Drops synthetic padding examples flagged by `is_real` (ref :279)."""
    real = dec_out_dict.get('is_real')
    if real is None:
      return
    keep = [i for i, r in enumerate(np.asarray(torch.as_tensor(real).cpu())) if r]
    for key, v in list(dec_out_dict.items()):
      if isinstance(v, list) and len(v) == len(real):
        dec_out_dict[key] = [v[i] for i in keep]
      elif isinstance(v, (torch.Tensor, np.ndarray)) and len(v) == len(real):
        dec_out_dict[key] = v[keep]

  def PreparePostProcess(self, dec_out_dict, dec_metrics_dict):
    """Host copies + field checks → `PostProcessInputs` (ref :320)."""
    def _np(x):
      return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else x
    self.FilterRealExamples(dec_out_dict)
    d = dec_out_dict
    b = len(d['transcripts'])
    scores = _np(d['topk_scores'])
    scores = np.asarray(scores).reshape(b, -1) if scores is not None else np.zeros((b, 1))
    return metrics_calculator.PostProcessInputs(
        transcripts=d['transcripts'], topk_decoded=d['topk_decoded'],
        filtered_transcripts=d['filtered_transcripts'],
        filtered_top_hyps=d['filtered_top_hyps'], topk_scores=scores,
        utt_id=_np(d.get('utt_id')), norm_wer_errors=_np(d['norm_wer_errors']),
        target_labels=_np(d['target_labels']), target_paddings=_np(d['target_paddings']),
        topk_ids=_np(d['topk_ids']), topk_lens=_np(d['topk_lens']))

  def PostProcess(self, dec_out_dict, dec_metrics_dict, tokenizer=None):
    """Scores one decoded batch; returns [(key, text)] for the decode dump (ref :403)."""
    p = self.params
    pi = self.PreparePostProcess(dec_out_dict, dec_metrics_dict)
    n = len(pi.transcripts)
    dec_metrics_dict['num_samples_in_batch'].Update(n)
    if n == 0:
      return []
    k = len(pi.topk_decoded[0])
    first_errs = float(np.sum(pi.norm_wer_errors[:, 0]))
    words = sum(len(decoder_utils.Tokenize(r)) for r in pi.filtered_transcripts)
    dec_metrics_dict['norm_wer'].Update(first_errs / max(1, words), words)
    if p.include_auxiliary_metrics:
      metrics_calculator.CalculateMetrics(pi, dec_metrics_dict, add_summary=False,
                                          use_tpu=False, log_utf8=p.log_utf8)
    else:
      dec_metrics_dict['wer'].Update(first_errs / max(1, words), words)
    if 'corpus_bleu' in dec_metrics_dict:
      for r, h in zip(pi.filtered_transcripts, pi.filtered_top_hyps):
        dec_metrics_dict['corpus_bleu'].Update(r, h)
    kv = []
    for i in range(n):
      key = str(pi.utt_id[i]) if pi.utt_id is not None else str(i)
      lines = ['ref: %s' % pi.transcripts[i]]
      lines += ['hyp[%d] %.4f: %s' % (j, float(pi.topk_scores[i][j]) if j < pi.topk_scores.shape[1] else 0.0,
                                     pi.topk_decoded[i][j]) for j in range(k)]
      kv.append((key, '\n'.join(lines)))
    return kv
