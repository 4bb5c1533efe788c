"""MT tasks (ref `lingvo/tasks/mt/model.py`).

`MTBaseModel` (ref :26): encoder → decoder loss for training; beam search +
corpus BLEU for decoding. `TransformerModel` (:176) and `RNMTModel` (:196) only
differ in default encoder/decoder.
"""

from __future__ import annotations

import collections
import math

import torch

from lingvo_b200.core import base_model
from lingvo_b200.core import metrics as metrics_lib
from lingvo_b200.core import py_utils
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.mt import decoder as mt_decoder
from lingvo_b200.models.mt import encoder as mt_encoder


def _NGrams(tokens, n):
  return collections.Counter(tuple(tokens[i:i + n]) for i in range(len(tokens) - n + 1))


class CorpusBleuMetric(metrics_lib.BaseMetric if hasattr(metrics_lib, 'BaseMetric') else object):
  """Corpus-level BLEU-4 with brevity penalty (ref `core/metrics.py` CorpusBleuMetric)."""

  def __init__(self, separator_type='wpm'):
    self._sep = separator_type
    self._match = [0] * 4
    self._total = [0] * 4
    self._hyp_len = 0
    self._ref_len = 0

  def _Tok(self, s):
    return s.split()

  def Update(self, ref_str, hyp_str):
    ref, hyp = self._Tok(ref_str), self._Tok(hyp_str)
    self._hyp_len += len(hyp)
    self._ref_len += len(ref)
    for n in range(1, 5):
      h, r = _NGrams(hyp, n), _NGrams(ref, n)
      self._match[n - 1] += sum(min(c, r[g]) for g, c in h.items())
      self._total[n - 1] += max(len(hyp) - n + 1, 0)

  @property
  def value(self):
    if not self._hyp_len or min(self._total) == 0 or min(self._match) == 0:
      return 0.0
    logp = sum(math.log(m / t) for m, t in zip(self._match, self._total)) / 4
    bp = 1.0 if self._hyp_len > self._ref_len else math.exp(1 - self._ref_len / self._hyp_len)
    return bp * math.exp(logp)

  def Summary(self, name):
    return (name, self.value)


class MTBaseModel(base_model.BaseTask):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.encoder = mt_encoder.TransformerEncoder.Params()
    p.decoder = mt_decoder.TransformerDecoder.Params()
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    self.CreateChild('enc', p.encoder)
    self.CreateChild('dec', p.decoder)

  def ComputePredictions(self, theta, batch):
    enc = self.enc.FProp(theta.enc, batch.src)
    return self.dec.ComputePredictions(theta.dec, enc, batch.tgt)

  def ComputeLoss(self, theta, predictions, batch):
    return self.dec.ComputeLoss(theta.dec, predictions, batch.tgt)

  def _GetTokenizerKeyToUse(self, key):
    return key if key in self.input_generator.tokenizer_dict else None

  def DecodeWithTheta(self, theta, batch):
    with torch.no_grad():
      enc = self.enc.FProp(theta.enc, batch.src)
      out = self.dec.BeamSearchDecodeWithTheta(theta.dec, enc)
    k = out.topk_hyps.ids.shape[1]
    return NestedMap(topk_ids=out.topk_ids, topk_lens=out.topk_lens,
                     topk_scores=out.topk_scores, target_labels=batch.tgt.labels,
                     target_paddings=batch.tgt.paddings, source_ids=batch.src.ids,
                     source_paddings=batch.src.paddings, num_hyps_per_beam=k)

  def Decode(self, batch):
    return self.DecodeWithTheta(self.theta, batch)

  def CreateDecoderMetrics(self):
    return {'num_samples_in_batch': metrics_lib.AverageMetric(),
            'corpus_bleu': CorpusBleuMetric()}

  def PostProcessDecodeOut(self, dec_out, dec_metrics):
    gen = self.input_generator
    tgt_lens = (1.0 - dec_out.target_paddings.float()).sum(1).long()
    refs = gen.IdsToStrings(dec_out.target_labels, (tgt_lens - 1).clamp_min(0))
    k = int(dec_out.num_hyps_per_beam)
    b = len(refs)
    ids = dec_out.topk_ids.reshape(b, k, -1)[:, 0]
    lens = (dec_out.topk_lens.reshape(b, k)[:, 0] - 1).clamp_min(0)   # drop </s>
    hyps = gen.IdsToStrings(ids, lens)
    dec_metrics['num_samples_in_batch'].Update(b)
    kv = []
    for i, (r, h) in enumerate(zip(refs, hyps)):
      dec_metrics['corpus_bleu'].Update(r, h)
      kv.append(('%d' % i, 'ref: %s\nhyp: %s' % (r, h)))
    return kv


class TransformerModel(MTBaseModel):
  """Transformer NMT (ref :176)."""


class RNMTModel(MTBaseModel):
  """RNMT+ (ref :196)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.encoder = mt_encoder.MTEncoderBiRNN.Params()
    p.decoder = mt_decoder.MTDecoderV1.Params()
    return p


class HybridModel(MTBaseModel):
  """Transformer encoder + RNMT decoder (ref :211)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.encoder = mt_encoder.TransformerEncoder.Params()
    p.decoder = mt_decoder.MTDecoderV1.Params()
    return p
