"""ASR task (ref `lingvo/tasks/asr/model.py:30`): frontend → encoder → LAS decoder;
decode = beam search + WER / normalised edit distance metrics
(`decoder_metrics.py:87`)."""

from __future__ import annotations

import torch

from lingvo_b200.core import base_model
from lingvo_b200.core import metrics as metrics_lib
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.asr import decoder as asr_decoder
from lingvo_b200.models.asr import decoder_utils
from lingvo_b200.models.asr import encoder as asr_encoder
from lingvo_b200.models.asr import frontend as asr_frontend


class AsrModel(base_model.BaseTask):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.encoder = asr_encoder.AsrEncoder.Params()
    p.decoder = asr_decoder.AsrDecoder.Params()
    p.Define('frontend', None, 'Optional frontend (e.g. MelAsrFrontend).')
    p.Define('include_auxiliary_metrics', True, 'Emit per-batch auxiliary metrics.')
    p.Define('target_key', '', 'Key of the target in multi-target batches.')
    tp = p.train
    tp.lr_schedule = tp.lr_schedule
    tp.vn_std = 0.075
    tp.l2_regularizer_weight = 1e-6
    tp.clip_gradient_norm_to_value = 1.0
    tp.grad_norm_to_clip_to_zero = 100.0
    tp.learning_rate = 2.5e-4
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    if p.frontend is not None:
      self.CreateChild('frontend', p.frontend)
    self.CreateChild('encoder', p.encoder)
    self.CreateChild('decoder', p.decoder)

  def _Targets(self, batch):
    p = self.params
    return batch.tgt[p.target_key] if p.target_key else batch.tgt

  def _Encode(self, theta, batch):
    src = batch.src
    if self.params.frontend is not None:
      src = self.frontend.FProp(theta.frontend, src)
    return self.encoder.FProp(theta.encoder, src)

  def ComputePredictions(self, theta, batch):
    enc = self._Encode(theta, batch)
    return self.decoder.ComputePredictions(theta.decoder, enc, self._Targets(batch))

  def ComputeLoss(self, theta, predictions, batch):
    return self.decoder.ComputeLoss(theta.decoder, predictions, self._Targets(batch))

  def DecodeWithTheta(self, theta, batch):
    with torch.no_grad():
      enc = self._Encode(theta, batch)
      out = self.decoder.BeamSearchDecodeWithTheta(theta.decoder, enc)
    tgt = self._Targets(batch)
    return NestedMap(topk_ids=out.topk_ids, topk_lens=out.topk_lens,
                     topk_scores=out.topk_scores, num_hyps_per_beam=out.topk_hyps.ids.shape[1],
                     target_labels=tgt.labels, target_paddings=tgt.paddings,
                     utt_id=batch.get('sample_ids'))

  def Decode(self, batch):
    return self.DecodeWithTheta(self.theta, batch)

  def CreateDecoderMetrics(self):
    return {'num_samples_in_batch': metrics_lib.AverageMetric(),
            'wer': metrics_lib.AverageMetric(), 'norm_wer': metrics_lib.AverageMetric(),
            'sacc': metrics_lib.AverageMetric(), 'ter': metrics_lib.AverageMetric(),
            'oracle_norm_wer': metrics_lib.AverageMetric()}

  def PostProcessDecodeOut(self, dec_out, dec_metrics):
    dec_out = base_model.DecodeOutAsTensors(dec_out)
    gen = self.input_generator
    tgt_lens = (1.0 - dec_out.target_paddings.float()).sum(1).long()
    refs = gen.IdsToStrings(dec_out.target_labels, (tgt_lens - 1).clamp_min(0))
    k = int(dec_out.num_hyps_per_beam)
    b = len(refs)
    ids = dec_out.topk_ids.reshape(b, k, -1)
    lens = (dec_out.topk_lens.reshape(b, k) - 1).clamp_min(0)
    dec_metrics['num_samples_in_batch'].Update(b)
    kv = []
    for i, ref in enumerate(refs):
      hyps = gen.IdsToStrings(ids[i], lens[i])
      ref_n = decoder_utils.FilterNoise(decoder_utils.FilterEpsilon(ref))
      hyps = [decoder_utils.FilterNoise(decoder_utils.FilterEpsilon(h)) for h in hyps]
      errs = [decoder_utils.EditDistance(ref_n, h)[3] for h in hyps]
      nref = max(len(decoder_utils.Tokenize(ref_n)), 1)
      dec_metrics['wer'].Update(errs[0] / nref, nref)
      dec_metrics['norm_wer'].Update(errs[0] / nref, nref)
      dec_metrics['oracle_norm_wer'].Update(min(errs) / nref, nref)
      dec_metrics['sacc'].Update(1.0 if errs[0] == 0 else 0.0)
      ref_ids = dec_out.target_labels[i, :int(tgt_lens[i])].tolist()
      hyp_ids = ids[i, 0, :int(lens[i, 0]) + 1].tolist()
      ter = decoder_utils.EditDistanceInIds(ref_ids, hyp_ids)[3]
      dec_metrics['ter'].Update(ter / max(len(ref_ids), 1), max(len(ref_ids), 1))
      kv.append(('%d' % i, 'ref: %s\nhyp: %s' % (ref_n, hyps[0])))
    return kv
