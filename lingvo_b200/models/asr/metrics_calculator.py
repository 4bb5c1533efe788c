"""Aggregation of per-utterance ASR scores into decoder metrics (ref
`lingvo/tasks/asr/metrics_calculator.py`).

`CalculateMetrics` consumes one post-processed decode batch (`PostProcessInputs`) and
updates the metric dictionary created by `DecoderMetrics.CreateMetrics`:
word error rates (case-sensitive and -insensitive, split into ins/sub/del), oracle WER
over the top-k list, sentence accuracy and token error rate.
"""

from __future__ import annotations

import collections
import logging

from lingvo_b200.models.asr import decoder_utils

PostProcessInputs = collections.namedtuple('PostProcessInputs', [
    'transcripts', 'topk_decoded', 'filtered_transcripts', 'filtered_top_hyps',
    'topk_scores', 'utt_id', 'norm_wer_errors', 'target_labels', 'target_paddings',
    'topk_ids', 'topk_lens'])


def GetRefIds(ref_ids, ref_paddings):
  """Labels at non-padded positions (ref :53)."""
  assert len(ref_ids) == len(ref_paddings)
  return [int(t) for t, p in zip(ref_ids, ref_paddings) if p == 0]


class _Tally:
  """Running ins/sub/del/total counts."""

  def __init__(self):
    self.ins = self.sub = self.dele = self.err = 0

  def Add(self, ref, hyp):
    i, s, d, e = decoder_utils.EditDistance(ref, hyp)
    self.ins += i
    self.sub += s
    self.dele += d
    self.err += e
    return i, s, d, e

  def Emit(self, metrics, prefix, denom, weight):
    metrics[prefix + '/ins'].Update(self.ins / denom, weight)
    metrics[prefix + '/sub'].Update(self.sub / denom, weight)
    metrics[prefix + '/del'].Update(self.dele / denom, weight)
    metrics[prefix + '/wer'].Update(self.err / denom, weight)


def _Str(s, utf8):
  if utf8 and isinstance(s, bytes):
    return s.decode('utf-8', 'replace')
  return s


def CalculateMetrics(postprocess_inputs, dec_metrics_dict, add_summary=False, use_tpu=False,
                     log_utf8=False):
  """Updates `dec_metrics_dict` with this batch's statistics (ref :62)."""
  pi = postprocess_inputs
  n_utts = len(pi.transcripts)
  if n_utts == 0:
    return
  cased, uncased = _Tally(), _Tally()
  ref_words = ref_tokens = token_errs = oracle_errs = exact = 0
  for i in range(n_utts):
    hyps = pi.topk_decoded[i]
    k = len(hyps)
    if add_summary:
      logging.info('utt_id: %s', pi.utt_id[i] if pi.utt_id is not None else i)
      logging.info('  ref_str: %s', _Str(pi.transcripts[i], log_utf8))
      for score, h in zip(pi.topk_scores[i], hyps):
        logging.info('  %f: %s', float(score), _Str(h, log_utf8))
    # token error rate: top hypothesis ids vs. un-padded reference ids
    rid = GetRefIds(pi.target_labels[i], pi.target_paddings[i])
    top = i * k
    hid = [int(t) for t in pi.topk_ids[top][:int(pi.topk_lens[top])]]
    ref_tokens += len(rid)
    token_errs += decoder_utils.EditDistanceInIds(rid, hid)[3]
    # word error rates on the filtered strings (top hypothesis only)
    ref, hyp = pi.filtered_transcripts[i], pi.filtered_top_hyps[i]
    _, _, _, errs = cased.Add(ref, hyp)
    uncased.Add(ref.lower(), hyp.lower())
    nw = len(decoder_utils.Tokenize(ref))
    ref_words += nw
    per_hyp = [float(e) for e in pi.norm_wer_errors[i]]
    oracle_errs += min(per_hyp) if per_hyp else errs
    if per_hyp and per_hyp[0] == 0:
      exact += 1
  denom = max(1.0, float(ref_words))
  dec_metrics_dict['wer'].Update(cased.err / denom, ref_words)
  cased.Emit(dec_metrics_dict, 'error_rates', denom, ref_words)
  uncased.Emit(dec_metrics_dict, 'case_insensitive_error_rates', denom, ref_words)
  dec_metrics_dict['oracle_norm_wer'].Update(oracle_errs / denom, ref_words)
  dec_metrics_dict['sacc'].Update(exact / n_utts, n_utts)
  dec_metrics_dict['ter'].Update(token_errs / max(1.0, float(ref_tokens)), ref_tokens)
