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


class CorpusBleuMetric(metrics_lib.BaseMetric):
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
    if p.encoder is not None:
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
    dec_out = base_model.DecodeOutAsTensors(dec_out)
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


class InsertionModel(MTBaseModel):
  """Insertion-based translation (KERMIT-style, ref :391): source and target are
  concatenated into one canvas; training samples a partial canvas of each side and
  learns to insert the missing tokens."""

  @classmethod
  def Params(cls):
    from lingvo_b200.core import insertion
    from lingvo_b200.models.mt import decoder as mt_decoder
    p = super().Params()
    p.encoder = None
    p.decoder = mt_decoder.InsertionDecoder.Params()
    p.Define('insertion', insertion.SymbolInsertionLayer.Params(), 'Roll-in / oracle policy.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self.CreateChild('insertion', self.params.insertion)

  def _SampleCanvasAndTargets(self, x, x_paddings):
    return self.insertion.FProp(None, x, x_paddings, self.params.decoder.target_eos_id, True)

  def _CreateCanvasAndTargets(self, batch):
    """→ NestedMap(canvas, canvas_paddings[, target_indices, target_weights])."""
    from lingvo_b200.core import insertion
    p = self.params
    vocab = p.decoder.softmax.num_classes
    if self.do_eval:
      # eval: the canvas is the full source followed by an empty target side
      canvas = torch.where(batch.src.paddings < 0.5, batch.src.ids + vocab, batch.src.ids)
      return NestedMap(canvas=canvas, canvas_paddings=batch.src.paddings.float())
    src = self._SampleCanvasAndTargets(batch.src.ids, batch.src.paddings.float())
    tgt = self._SampleCanvasAndTargets(batch.tgt.ids, batch.tgt.paddings.float())
    src.canvas = torch.where(src.canvas_paddings < 0.5, src.canvas + vocab, src.canvas)
    src_lens = (1.0 - src.canvas_paddings).sum(1).long()
    # target-side slots sit after the kept source tokens of the same row
    tgt_idx = tgt.target_indices.clone()
    tgt_idx[:, 1] += src_lens[tgt_idx[:, 0]]
    canvas, canvas_pad = insertion.SequenceConcat(
        src.canvas, src.canvas_paddings, tgt.canvas, tgt.canvas_paddings)
    return NestedMap(
        canvas=canvas, canvas_paddings=canvas_pad,
        target_indices=torch.cat([src.target_indices, tgt_idx], 0),
        target_weights=torch.cat([src.target_weights, tgt.target_weights], 0))

  def ComputePredictions(self, theta, batch):
    desc = self._CreateCanvasAndTargets(batch)
    pred = self.dec.ComputePredictions(
        theta.dec, None, NestedMap(ids=desc.canvas, paddings=desc.canvas_paddings))
    pred.tgt = desc
    return pred

  def ComputeLoss(self, theta, predictions, batch):
    return self.dec.ComputeLoss(theta.dec, predictions)

  def Decode(self, batch):
    raise NotImplementedError('InsertionModel: parallel insertion decoding is not exposed '
                              '(the reference has none either).')


class TransformerXEnDecModel(TransformerModel):
  """XEnDec (crossover encoder-decoder, arXiv 2106.04060; ref :459).

  Three losses per step: the *clean* loss on the batch, the *mono* loss on a partner batch
  (the batch rolled by one unless `other_src/other_tgt` are given), and the *mix* loss on
  the crossover of the two — source embeddings mixed by `src.source_mask`, target input
  embeddings and soft labels mixed by ratios derived from the (stop-gradient)
  cross-attention of both parents."""

  @classmethod
  def Params(cls):
    from lingvo_b200.models.mt import decoder as mt_decoder
    from lingvo_b200.models.mt import encoder as mt_encoder
    p = super().Params()
    p.encoder = mt_encoder.TransformerXEncoder.Params()
    p.decoder = mt_decoder.TransformerXDecoder.Params()
    p.Define('loss_mix_weight', 1.0, 'Weight of the crossover loss.')
    p.Define('loss_clean_weight', 1.0, 'Weight of the clean loss.')
    p.Define('loss_mono_weight', 1.0, 'Weight of the partner-batch loss.')
    p.Define('use_prob_cl', False, 'Curriculum from hard labels to model probabilities.')
    p.Define('use_prob_drop', False, 'Kept for parity.')
    p.Define('atten_drop', 0.0, 'Dropout on the attention used for mixing ratios.')
    return p

  @staticmethod
  def _CreateTargetLambdas(atten_probs, source_lambdas_pair, source_paddings_pair,
                           target_paddings_pair, smooth=0.0):
    """Target mixing ratios from attention mass on each parent's kept source tokens."""
    mass = []
    for k in range(2):
      src_w = source_lambdas_pair[k] * (1.0 - source_paddings_pair[k])          # [B,S]
      m = (atten_probs[k].detach() * src_w.unsqueeze(1)).sum(-1)                # [B,T]
      mass.append((m + smooth) * (1.0 - target_paddings_pair[k]))
    lab0 = mass[0] / (mass[0] + mass[1] + 1e-9)
    label_lambdas = [lab0, 1.0 - lab0]
    inp0 = torch.nn.functional.pad(lab0, (1, 0), value=1.0)[:, :-1]
    input_lambdas = [inp0 * (1.0 - target_paddings_pair[0]),
                     (1.0 - inp0) * (1.0 - target_paddings_pair[1])]
    return source_lambdas_pair, input_lambdas, label_lambdas

  def ComputePredictions(self, theta, batch, other_batch=None, source_lambdas=None,
                         target_lambdas=None):
    enc = self.enc.FProp(theta.enc, batch.src,
                             other_batch.src if other_batch is not None else None,
                             source_lambdas)
    pred = self.dec.ComputePredictions(
        theta.dec, enc, batch.tgt,
        other_batch.tgt if other_batch is not None else None, target_lambdas)
    pred.encoder_outputs = enc
    return pred

  def ComputeLoss(self, theta, predictions, batch):
    p = self.params
    clean = self.dec.ComputeLoss(theta.dec, predictions, batch.tgt)
    if self.do_eval:
      return clean
    roll = lambda x: torch.roll(x, 1, 0)
    if 'other_src' in batch and 'other_tgt' in batch:
      other = NestedMap(src=batch.other_src.DeepCopy(), tgt=batch.other_tgt.DeepCopy())
    else:
      other = NestedMap(src=batch.src.DeepCopy(), tgt=batch.tgt.DeepCopy()).Transform(roll)
    if p.loss_mono_weight > 0 or 'other_src' in batch:
      other_pred = self.ComputePredictions(theta, other)
      mono = self.dec.ComputeLoss(theta.dec, other_pred, other.tgt)
      other_att, other_aux = other_pred.attention.probs, mono[1]
      other_src_embs, other_tgt_embs = other_pred.source_embs, other_pred.target_embs
    else:
      mono = None
      other_att = roll(predictions.attention.probs)
      other_aux = NestedMap(reshape_probs=roll(clean[1].reshape_probs),
                            target_hard_probs=roll(clean[1].target_hard_probs))
      other_src_embs, other_tgt_embs = roll(predictions.source_embs), roll(predictions.target_embs)
    terms = []
    if p.loss_clean_weight > 0:
      terms.append(('clean_loss', clean, p.loss_clean_weight))
    if p.loss_mono_weight > 0:
      terms.append(('other_loss', mono, p.loss_mono_weight))
    if p.loss_mix_weight > 0:
      att, oatt = predictions.attention.probs, other_att
      if p.atten_drop > 0:
        att = torch.nn.functional.dropout(att, p.atten_drop)
        oatt = torch.nn.functional.dropout(oatt, p.atten_drop)
      if p.use_prob_cl:
        ratio = min(float(self.global_step) / 20000.0, 1.0)
        mixp = lambda aux, w: aux.target_hard_probs * (1 - w.unsqueeze(-1) * ratio) + \
            aux.reshape_probs * (w.unsqueeze(-1) * ratio)
        probs, oprobs = mixp(clean[1], batch.tgt.weights), mixp(other_aux, other.tgt.weights)
      else:
        probs, oprobs = clean[1].target_hard_probs, other_aux.target_hard_probs
      src_pads = [batch.src.paddings.float(), other.src.paddings.float()]
      tgt_pads = [batch.tgt.paddings.float(), other.tgt.paddings.float()]
      mask = batch.src.source_mask.float()
      other_lam = mask * (1.0 - src_pads[1])
      src_lam = [(1.0 - other_lam) * (1.0 - src_pads[0]), other_lam]
      src_lam, inp_lam, lab_lam = self._CreateTargetLambdas(
          [att, oatt], src_lam, src_pads, tgt_pads)
      mix_batch = NestedMap(src=batch.src.DeepCopy(), tgt=batch.tgt.DeepCopy())
      mix_batch.tgt.weights = (batch.tgt.weights + other.tgt.weights).clamp(0.0, 1.0)
      mix_batch.src.embs, mix_batch.tgt.embs = predictions.source_embs, predictions.target_embs
      other.src.embs, other.tgt.embs = other_src_embs, other_tgt_embs
      mix_pred = self.ComputePredictions(theta, mix_batch, other, src_lam, inp_lam)
      tp = probs * lab_lam[0].unsqueeze(-1) + oprobs * lab_lam[1].unsqueeze(-1) + 1e-9
      tp = tp / tp.sum(-1, keepdim=True)
      mix = self.dec.ComputeLoss(theta.dec, mix_pred, mix_batch.tgt, tp)
      terms.append(('mix_loss', mix, p.loss_mix_weight))
    metrics = NestedMap()
    total, npred = 0.0, torch.tensor(1.0)
    for name, (m, _), w in terms:
      total = total + m.loss[0] * w
      metrics[name] = (m.loss[0] * w, m.loss[1])
      if name == 'clean_loss':
        npred = m.loss[1]
    metrics.loss = (total, npred)
    for k in ('log_pplx', 'fraction_of_correct_next_step_preds', 'num_predictions'):
      metrics[k] = clean[0][k]
    return metrics, clean[1]


class GPipeTransformerModel(base_model.BaseTask):
  """Transformer NMT whose embeddings, encoder, decoder and softmax form one
  `layers_with_gpipe.GPipeTransformerStack` pipeline (reference building block
  `core/layers_with_gpipe.py:576`; BASELINE.json config #5 "TransformerBig GPipe").

  Single process: the micro-batched stack runs in place. Under torchrun with
  `pipeline_parallel=True` and world == `num_splits`, rank r executes cell r only:
  `parallel.pp.PipelineEngine` moves micro-batch activations / gradients between ranks
  (NCCL p2p over NVLink, overlapped with compute, optional rematerialisation) and every
  rank updates the variables of its own cell.
  """

  @classmethod
  def Params(cls):
    from lingvo_b200.core import layers_with_gpipe  # pylint: disable=g-import-not-at-top
    p = super().Params()
    p.Define('stack', layers_with_gpipe.GPipeTransformerStack.Params(), 'Pipeline stack.')
    p.Define('label_smoothing', 0.1, 'Label smoothing uncertainty.')
    p.Define('pipeline_parallel', True, 'One rank per cell when launched distributed.')
    p.Define('remat', True, 'Re-run stage forwards in backward (O(1) activations).')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    self.CreateChild('stack', p.stack)
    self._engine = None
    import torch.distributed as dist  # pylint: disable=g-import-not-at-top
    if (p.pipeline_parallel and dist.is_available() and dist.is_initialized() and
        dist.get_world_size() > 1):
      from lingvo_b200.parallel import pp  # pylint: disable=g-import-not-at-top
      assert dist.get_world_size() == self.stack.num_stages, (
          'GPipeTransformerModel: %d ranks but %d pipeline cells' %
          (dist.get_world_size(), self.stack.num_stages))
      self._engine = pp.PipelineEngine(remat=p.remat)
      self.stack.AttachEngine(self._engine)

  @property
  def engine(self):
    return self._engine

  def ComputePredictions(self, theta, batch):
    src, tgt = batch.src, batch.tgt
    # GPipe stacks are time-major: ids [T, B]
    logits = self.stack.FProp(
        theta.stack, src.ids.t(), src.paddings.t().float(), tgt.ids.t(),
        tgt.paddings.t().float())
    return NestedMap(logits=logits)

  def ComputeLoss(self, theta, predictions, batch):
    p = self.params
    tgt = batch.tgt
    if self._engine is not None and not self._engine.is_last:
      # placeholder: only the last stage sees logits; metrics are broadcast after BProp
      z = torch.zeros((), device=tgt.ids.device)
      one = torch.ones((), device=tgt.ids.device)
      return {'loss': (z, one), 'log_pplx': (z, one)}, {}
    logits = predictions.logits.float().transpose(0, 1)        # [B, T, V]
    v = logits.shape[-1]
    w = (1.0 - tgt.paddings.float()) * tgt.get('weights', torch.ones_like(tgt.paddings)).float()
    logp = torch.log_softmax(logits, -1)
    nll = -logp.gather(-1, tgt.labels.long().unsqueeze(-1)).squeeze(-1)
    if p.label_smoothing:
      smooth = -logp.mean(-1)
      xent = (1.0 - p.label_smoothing) * nll + p.label_smoothing * smooth
    else:
      xent = nll
    tot = w.sum().clamp_min(1.0)
    loss = (xent * w).sum() / tot
    del v
    return {'loss': (loss, tot), 'log_pplx': ((nll * w).sum() / tot, tot)}, {}

  def BProp(self):
    """With a pipeline engine every rank runs the backward phase of the GPipe schedule and
    applies the optimizer to its own cell; otherwise the standard learner path."""
    if self._engine is None:
      return super().BProp()
    from lingvo_b200.core import py_utils  # pylint: disable=g-import-not-at-top
    import torch.distributed as dist  # pylint: disable=g-import-not-at-top
    eng = self._engine
    for v in self.vars.Flatten():
      v.grad = None
    loss = self._metrics['loss'][0] if eng.is_last else None
    eng.Backward(loss)
    lrn = self.learners[0]
    with py_utils.GlobalStepContext(self._global_step):
      pairs = [py_utils.VarGrad(v, v.grad) for v in self.vars.Flatten()
               if v.requires_grad and v.grad is not None]
      # global gradient norm over all stages (each variable lives on exactly one rank)
      sq = torch.stack([g.grad.float().square().sum() for g in pairs]).sum() if pairs else (
          torch.zeros((), device=self.Device()))
      dist.all_reduce(sq)
      gnorm = sq.sqrt()
      tp = lrn.params
      scale = torch.ones_like(gnorm)
      if tp.clip_gradient_norm_to_value:
        scale = torch.clamp(tp.clip_gradient_norm_to_value / gnorm, max=1.0)
      scale = torch.where(torch.isfinite(gnorm), scale, torch.zeros_like(scale))
      if pairs:
        with torch.no_grad():
          for vg in pairs:
            vg.grad.mul_(scale.to(vg.grad.dtype))
        lrn.optimizer.Apply(lrn.LearningRate(), pairs)
      # share the last stage's metrics so every rank logs / stops identically
      vals = torch.stack([self._metrics['loss'][0].detach().float(),
                          self._metrics['log_pplx'][0].detach().float(),
                          self._metrics['loss'][1].detach().float()])
      dist.broadcast(vals, src=eng.world - 1)
      one = torch.ones((), device=vals.device)
      self._eval_metrics = {'loss': (vals[0], vals[2]), 'log_pplx': (vals[1], vals[2]),
                            'grad_norm/all': (gnorm, one)}
      self.ApplyExponentialMovingAverage()
    self._global_step += 1
    py_utils.SetGlobalStep(self._global_step)
    self._metrics = None
    return None
