"""ASR task (ref `lingvo/tasks/asr/model.py:30`): frontend → encoder → LAS decoder.

Training: `ComputePredictions` / `ComputeLoss` delegate to the decoder with a (possibly
augmented) decoder theta and targets — the two hooks `_MakeDecoderTheta` /
`_GetDecoderTargets` are what multi-source / multi-target subclasses override.
Decoding: beam search on the device, then the `DecoderMetrics` layer turns ids into strings
and the host scores WER / oracle WER / SACC / TER (`decoder_metrics.py`).
Serving: `Inference()['default'](wav_bytes)` → hypotheses, scores, feature and encoder frames.
"""

from __future__ import annotations

import torch

from lingvo_b200.core import base_model
from lingvo_b200.core import program
from lingvo_b200.core import py_utils
from lingvo_b200.core import schedule
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.asr import decoder as asr_decoder
from lingvo_b200.models.asr import decoder_metrics
from lingvo_b200.models.asr import encoder as asr_encoder
from lingvo_b200.models.asr import frontend as asr_frontend


class AsrModel(base_model.BaseTask):
  """Speech model."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.encoder = asr_encoder.AsrEncoder.Params()
    p.decoder = asr_decoder.AsrDecoder.Params()
    p.Define('frontend', None, 'Feature frontend (e.g. MelAsrFrontend); None: the input '
             'already carries features.')
    p.Define('decoder_metrics', decoder_metrics.DecoderMetrics.Params(),
             'The decoder metrics layer.')
    p.Define('include_auxiliary_metrics', True,
             'Besides WER also compute oracle WER, SACC, TER, … (slower decode job).')
    p.Define('target_key', '', 'Key of the target in multi-target batches.')
    tp = p.train
    tp.lr_schedule = schedule.PiecewiseConstantSchedule.Params().Set(
        boundaries=[350000, 500000, 600000], values=[1.0, 0.1, 0.01, 0.001])
    tp.vn_start_step = 20000
    tp.vn_std = 0.075
    tp.l2_regularizer_weight = 1e-6
    tp.learning_rate = 0.001
    tp.clip_gradient_norm_to_value = 1.0
    tp.grad_norm_to_clip_to_zero = 100.0
    tp.tpu_steps_per_loop = 20
    return p

  def __init__(self, params):
    if not params.name:
      raise ValueError('params.name not set.')
    super().__init__(params)
    p = self.params
    if p.encoder:
      self.CreateChild('encoder', p.encoder.Copy().Set(name=p.encoder.name or 'enc'))
    if p.decoder:
      self.CreateChild('decoder', p.decoder.Copy().Set(name=p.decoder.name or 'dec'))
    if getattr(p.input, 'skip_frontend', False):
      p.frontend = None
    if p.frontend is not None:
      self.CreateChild('frontend', p.frontend)
    self.CreateChild('decoder_metrics', self._DecoderMetricsParams())

  def _DecoderMetricsParams(self):
    p = self.params
    return p.decoder_metrics.Copy().Set(include_auxiliary_metrics=p.include_auxiliary_metrics)

  # -- hooks for multi-target / multi-source subclasses ----------------------------------------
  def _GetDecoderTargets(self, input_batch):
    """The targets forwarded to the decoder (ref :93)."""
    p = self.params
    return input_batch.tgt[p.target_key] if p.target_key else input_batch.tgt

  def _MakeDecoderTheta(self, theta, input_batch):
    """The theta the decoder computes loss / metrics with (ref :108): a copy, so subclasses
    can attach per-batch values (e.g. `input_batch.source_selected`) without touching the
    model's theta."""
    del input_batch
    return theta.decoder.copy()

  def _GetTargetForDecoderMetrics(self, input_batch):
    return self._GetDecoderTargets(input_batch)

  # -- training ------------------------------------------------------------------------------
  def FrontendAndEncoderFProp(self, theta, input_batch_src, initial_state=None):
    """Frontend then encoder; auxiliary losses raised inside the encoder (e.g. MoE load
    balancing) are summed into `encoder_outputs.aux_loss` (ref :154)."""
    p = self.params
    if p.frontend is not None:
      input_batch_src = self.frontend.FProp(theta.frontend, input_batch_src)
    with py_utils.AuxLossContext(reentrant=True) as aux_loss_ctx:
      if initial_state:
        encoder_outputs = self.encoder.FProp(theta.encoder, input_batch_src,
                                             state0=initial_state)
      else:
        encoder_outputs = self.encoder.FProp(theta.encoder, input_batch_src)
      if aux_loss_ctx.aux_losses:
        encoder_outputs.aux_loss = torch.stack(
            [l.float().reshape(()) for l in aux_loss_ctx.aux_losses]).sum()
    return encoder_outputs

  def ComputePredictions(self, theta, input_batch):
    encoder_outputs = self.FrontendAndEncoderFProp(theta, input_batch.src)
    tgt = self._GetDecoderTargets(input_batch)
    decoder_theta = self._MakeDecoderTheta(theta, input_batch)
    predictions = self.decoder.ComputePredictions(decoder_theta, encoder_outputs, tgt)
    predictions.encoder_outputs = encoder_outputs
    return predictions

  def ComputeLoss(self, theta, predictions, input_batch):
    tgt = self._GetDecoderTargets(input_batch)
    decoder_theta = self._MakeDecoderTheta(theta, input_batch)
    metrics, per_sequence = self.decoder.ComputeLoss(decoder_theta, predictions, tgt)
    aux = predictions.get('encoder_outputs', NestedMap()).get('aux_loss')
    if aux is not None:
      loss, weight = metrics['loss']
      metrics['encoder_aux_loss'] = (aux, weight)
      metrics['loss'] = (loss + aux.to(loss.dtype), weight)
    return metrics, per_sequence

  # -- decoding ------------------------------------------------------------------------------
  def _GetTopK(self, decoder_outs, encoder_outs=None, tag=''):
    gen = self.input_generator
    return self.decoder_metrics.GetTopK(
        decoder_outs, ids_to_strings_fn=gen.IdsToStrings,
        feed_encoder_outs=bool(getattr(gen, 'feed_encoder_outs', False)),
        encoder_outs=encoder_outs, tag=tag)

  def _ComputeNormalizedWER(self, hyps, refs):
    return self.decoder_metrics.ComputeNormalizedWER(
        hyps, refs, self.params.decoder.beam_search.num_hyps_per_beam)

  def _ComputeDecoderMetrics(self, decoder_outs, input_batch):
    batch = input_batch.copy()
    if not getattr(self.params.decoder_metrics, 'pass_through_transcript_field', None):
      batch.tgt = self._GetTargetForDecoderMetrics(input_batch)
    return self.decoder_metrics.ComputeMetrics(
        decoder_outs, batch, ids_to_strings_fn=self.input_generator.IdsToStrings)

  def DecodeWithTheta(self, theta, input_batch):
    with torch.no_grad():
      encoder_outputs = self.FrontendAndEncoderFProp(theta, input_batch.src)
      decoder_outs = self.decoder.BeamSearchDecodeWithTheta(theta.decoder, encoder_outputs)
    return self._ComputeDecoderMetrics(decoder_outs, input_batch)

  def Decode(self, input_batch):
    return self.DecodeWithTheta(self.theta, input_batch)

  def CreateDecoderMetrics(self):
    return self.decoder_metrics.CreateMetrics()

  def PostProcessDecodeOut(self, dec_out_dict, dec_metrics_dict):
    return self.decoder_metrics.PostProcess(
        dec_out_dict, dec_metrics_dict, getattr(self.input_generator, 'tokenizer', None))

  # -- serving -------------------------------------------------------------------------------
  def Inference(self):
    """{'default': fn(wav_bytes) → NestedMap(hypotheses, scores, src_frames, encoder_frames)}."""
    return {'default': self._InferenceSubgraph_Default}

  def _InferenceSubgraph_Default(self, wav):   # pylint: disable=invalid-name
    """Offline recognition of one 16-bit PCM WAV file given as bytes (ref :264): decode the
    audio, run the model's frontend (a default `MelAsrFrontend` if the model has none), the
    encoder and beam search; return the top-k strings with their scores."""
    from lingvo_b200.tools import audio_lib  # pylint: disable=g-import-not-at-top
    p = self.params
    frontend = self.frontend if p.frontend is not None else getattr(self, '_default_frontend',
                                                                    None)
    if frontend is None:
      frontend = asr_frontend.MelAsrFrontend.Params().Set(name='frontend').Instantiate()
      self.__dict__['_default_frontend'] = frontend
    _, pcm = audio_lib.DecodeWav(wav)
    audio = torch.from_numpy(pcm).unsqueeze(0).to(self.Device())
    with torch.no_grad():
      src = frontend.FPropDefaultTheta(NestedMap(src_inputs=audio,
                                                 paddings=torch.zeros_like(audio)))
      encoder_outputs = self.encoder.FPropDefaultTheta(src)
      decoder_outputs = self.decoder.BeamSearchDecode(encoder_outputs)
    topk = self._GetTopK(decoder_outputs)
    return NestedMap(hypotheses=topk.decoded, scores=topk.scores, src_frames=src.src_inputs,
                     encoder_frames=encoder_outputs.encoded)

  def ProgramSchedule(self):
    """Train-only executor schedule (decoding runs as a separate job) (ref :310)."""
    ps = program.SimpleProgramScheduleForTask(
        train_dataset_name='Train', train_steps_per_loop=1000, eval_dataset_names=[],
        eval_steps_per_loop=0, decode_steps_per_loop=0)
    ps.train_executions_per_eval = 0
    return ps
