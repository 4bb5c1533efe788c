"""Dual-encoder (image↔text, …) retrieval model (ref `lingvo/tasks/milan/dual_encoder.py`).

`DualEncoder` (ref :66): per-modality encoder → optional projection to a joint
space → L2 normalise → all-pairs scores; symmetric multi-label contrastive loss
with a learnable temperature. With data parallelism the result-side embeddings
are concatenated across replicas (the reference's `tpu_utils.CrossReplicaConcat`
with CollectivePermute) via an NCCL all-gather that keeps gradients for the
local shard.
"""

from __future__ import annotations

import torch
import torch.distributed as dist

from lingvo_b200.core import base_layer
from lingvo_b200.core import base_model
from lingvo_b200.core import hyperparams
from lingvo_b200.core import layers
from lingvo_b200.core import py_utils
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.models.milan import labels as label_lib
from lingvo_b200.models.milan import score_functions


def EncoderConfig() -> hyperparams.Params:
  """One modality's config (ref :41)."""
  p = hyperparams.Params()
  p.Define('input_features', '', 'Key of the raw features in the input batch.')
  p.Define('id_feature', '', 'Key of per-example ids (for label functions).')
  p.Define('encoder', None, 'Encoder layer params: FProp(features) → [B, D].')
  p.Define('output_dim', None, 'Encoder output dim.')
  p.Define('encoder_output_dim', None, 'Alias of output_dim.')
  return p


from lingvo_b200.models.milan.tpu_utils import CrossReplicaConcat  # noqa: E402,F401


class DualEncoder(base_layer.BaseLayer):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.name = 'dual_encoder'
    p.Define('encoder_configs', {}, 'modality name → EncoderConfig().')
    p.Define('score_function', score_functions.DotProductScoreFunction.Params(), 'Scores.')
    p.Define('joint_embedding_dim', 0, 'Project every modality to this dim (0: none).')
    p.Define('regularization_loss_weight', 1.0, 'Kept for parity.')
    p.Define('loss_weights', {}, '(query, result) → weight.')
    p.Define('label_fn', None, 'ExamplePairLabeler or callable.')
    p.Define('learnable_temperature', True, 'Learn the softmax temperature.')
    p.Define('initial_temperature', 0.07, 'Initial temperature.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    assert p.encoder_configs and p.loss_weights
    self._modalities = sorted(p.encoder_configs)
    for m in self._modalities:
      cfg = p.encoder_configs[m]
      self.CreateChild('encoder_%s' % m, cfg.encoder)
      if p.joint_embedding_dim:
        self.CreateChild('projection_%s' % m, layers.ProjectionLayer.Params().Set(
            input_dim=cfg.output_dim or cfg.encoder_output_dim,
            output_dim=p.joint_embedding_dim, activation='NONE', batch_norm=False,
            has_bias=True))
    self.CreateChild('score_function', p.score_function)

  def _CreateLayerVariables(self):
    p = self.params
    import math
    self.CreateVariable('log_temperature', py_utils.WeightParams(
        [], py_utils.WeightInit.Constant(math.log(p.initial_temperature)), p.dtype),
        trainable=p.learnable_temperature)

  def EncodeModality(self, theta, modality, features):
    """`features`: one tensor, or a tuple of tensors passed to the encoder positionally
    (e.g. (token_features, lengths))."""
    p = self.params
    enc, th = self.children['encoder_%s' % modality], theta['encoder_%s' % modality]
    emb = enc.FProp(th, *features) if isinstance(features, (tuple, list)) else enc.FProp(th, features)
    if p.joint_embedding_dim:
      emb = self.children['projection_%s' % modality].FProp(
          theta['projection_%s' % modality], emb)
    return torch.nn.functional.normalize(emb.float(), dim=-1)

  def _Features(self, batch, modality):
    key = self.params.encoder_configs[modality].input_features
    if isinstance(key, (tuple, list)):
      return tuple(batch[k] for k in key)
    return batch[key]

  def EncodeBatch(self, theta, batch):
    return NestedMap({m: self.EncodeModality(theta, m, self._Features(batch, m))
                      for m in self._modalities})

  def ComputePredictions(self, theta, input_batch):
    """→ NestedMap({modality: NestedMap(encodings `[batch, …, D]`, ids)}) (ref :191); `ids` is
    the modality's `id_feature` of the batch when the encoder config names one."""
    out = NestedMap()
    for m in self._modalities:
      cfg = self.params.encoder_configs[m]
      entry = NestedMap(encodings=self.EncodeModality(theta, m, self._Features(input_batch, m)))
      id_key = cfg.Get('id_feature') if hasattr(cfg, 'Get') and 'id_feature' in cfg else None
      if id_key and id_key in input_batch:          # feature names may contain '/'
        entry.ids = input_batch[id_key]
      out[m] = entry
    return out

  def ComputeLoss(self, theta, predictions, input_batch):
    """Contrastive retrieval losses between the local queries and the results of ALL ranks
    (ref :223) → ({name: (value, weight)}, {})."""
    p = self.params
    emb = NestedMap({m: predictions[m].encodings.reshape(-1, predictions[m].encodings.shape[-1])
                     for m in predictions})
    temp = torch.exp(theta.log_temperature.float())
    total = 0.0
    metrics = {}
    for (q, r), w in sorted(p.loss_weights.items()):
      if not w:
        continue
      results = CrossReplicaConcat(emb[r])
      scores = self.score_function.FProp(theta.score_function, emb[q], results) / temp
      n, m = scores.shape
      offset = 0
      if m != n and dist.is_available() and dist.is_initialized():
        offset = dist.get_rank() * n
      labels = torch.zeros(n, m, device=scores.device)
      labels[torch.arange(n), torch.arange(n) + offset] = 1.0
      if p.label_fn is not None:
        ids = NestedMap({k: v for k, v in input_batch.items()
                         if isinstance(v, torch.Tensor) and v.dim() == 1 and v.shape[0] == n})
        make = (label_lib.ExamplePairs.WithinBatch if m == n
                else label_lib.ExamplePairs.BetweenLocalAndGlobalBatches)
        labels = p.label_fn(make(ids, query_modality=q, result_modality=r))
      loss = label_lib.MultiLabelContrastiveLoss(labels, scores).mean()
      total = total + w * loss
      acc = (scores.argmax(-1) == torch.arange(n, device=scores.device) + offset).float().mean()
      metrics['loss_%s_to_%s' % (q, r)] = (loss, 1)
      metrics['recall_at_1_%s_to_%s' % (q, r)] = (acc, 1)
    metrics['loss'] = (total, 1)
    return metrics, {}

  def FProp(self, theta, batch):
    """→ (loss, metrics NestedMap of plain values)."""
    metrics, _ = self.ComputeLoss(theta, self.ComputePredictions(theta, batch), batch)
    return metrics['loss'][0], NestedMap({k: v[0] for k, v in metrics.items() if k != 'loss'})


class MilanTask(base_model.BaseTask):
  """Task wrapper (ref :299)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.name = 'milan'
    p.Define('dual_encoder', DualEncoder.Params(), 'Dual encoder.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self.CreateChild('dual_encoder', self.params.dual_encoder)

  def ComputePredictions(self, theta, input_batch):
    return self.dual_encoder.ComputePredictions(theta.dual_encoder, input_batch)

  def ComputeLoss(self, theta, predictions, input_batch):
    metrics, per_example = self.dual_encoder.ComputeLoss(theta.dual_encoder, predictions,
                                                         input_batch)
    n = float(next(iter(input_batch.Flatten())).shape[0])
    return NestedMap({k: (v[0], n) for k, v in metrics.items()}), NestedMap(per_example)

  def Decode(self, input_batch):
    """The encodings of every modality (what an offline retrieval index is built from)."""
    with torch.no_grad():
      preds = self.ComputePredictions(self.theta, input_batch)
      self.ComputeLoss(self.theta, preds, input_batch)
    return preds

  def CreateDecoderMetrics(self):
    from lingvo_b200.core import metrics as metrics_lib   # pylint: disable=g-import-not-at-top
    return {'num_samples_in_batch': metrics_lib.AverageMetric()}

  def PostProcessDecodeOut(self, dec_out_dict, dec_metrics_dict):
    first = next(iter(dec_out_dict.values()))
    enc = first.encodings if hasattr(first, 'encodings') else first['encodings']
    dec_metrics_dict['num_samples_in_batch'].Update(int(enc.shape[0]))
    return []

  def Inference(self):
    return {m: (lambda feats, m=m: self.dual_encoder.EncodeModality(
        self.theta.dual_encoder, m, feats)) for m in self.dual_encoder._modalities}  # pylint: disable=protected-access
