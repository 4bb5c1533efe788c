"""GPipe Transformer stacks (ref `lingvo/core/layers_with_gpipe.py`).

`GPipeTransformerLayer` (ref :165) is a time-major TransformerLayer with the
"all tensors in / all tensors out" convention so layers chain inside a
`FeatureExtractionLayer`; `GPipeTransformerEmbeddingLayer` (ref :397) and
`GPipeTransformerSoftmaxLayer` (ref :355) bracket the stack;
`GPipeTransformerStack` (ref :576) builds encoder/decoder layer lists, splits them
into `num_splits` cells and runs them as a `PipeliningLayer`.
"""

from __future__ import annotations

import torch

from lingvo_b200.core import base_layer
from lingvo_b200.core import batch_major_attention as bma
from lingvo_b200.core import gpipe
from lingvo_b200.core import layers
from lingvo_b200.core import layers_with_attention
from lingvo_b200.core import py_utils
from lingvo_b200.core.gpipe import PipeliningLayer
from lingvo_b200.core.nested_map import NestedMap


class GPipeTransformerLayer(layers_with_attention.TransformerLayer):
  """FProp(source_vecs, source_paddings, target_vecs, target_paddings, source_segment_id,
  target_segment_id, transparent_acc, transparent_acc_helper, source_task_id,
  target_task_id) → the same tuple with the processed stream replaced."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('is_transparent', False, 'Kept for parity.')
    p.Define('num_transparent_outputs', 0, 'Kept for parity.')
    p.Define('transparent_merger_tpl', None, 'Kept for parity.')
    p.Define('normalize_output', False, 'LN on the output.')
    return p

  def FProp(self, theta, source_vecs, source_paddings, target_vecs=None, target_paddings=None,
            source_segment_id=None, target_segment_id=None, transparent_acc=None,
            transparent_acc_helper=None, source_task_id=None, target_task_id=None):
    p = self.params
    if p.has_aux_atten:       # decoder layer
      out, _ = super().FProp(theta, target_vecs, target_paddings, source_vecs, source_paddings,
                             target_segment_id, source_segment_id)
      target_vecs = out
    else:
      out, _ = super().FProp(theta, source_vecs, source_paddings,
                             source_segment_id=source_segment_id)
      source_vecs = out
    return (source_vecs, source_paddings, target_vecs, target_paddings, source_segment_id,
            target_segment_id, transparent_acc, transparent_acc_helper, source_task_id,
            target_task_id)

  @classmethod
  def FPropMeta(cls, p, inputs, *args):
    t, b, d = inputs[0], inputs[1], inputs[2]
    ff = p.tr_fflayer_tpl.hidden_dim
    flops = b * t * (8 * d * d + 4 * t * d + 4 * d * ff)
    return NestedMap(flops=flops, out_shapes=(inputs,) + tuple(args))


class GPipeTransformerSoftmaxLayer(layers.SimpleFullSoftmax):
  """Softmax as the last pipeline stage: consumes the decoder stream (ref :355)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('inputs_from_decoder', False, 'Read target_vecs instead of source_vecs.')
    return p

  def FProp(self, theta, source_vecs, source_paddings, target_vecs=None, *args):
    x = target_vecs if self.params.inputs_from_decoder and target_vecs is not None \
        else source_vecs
    shp = x.shape
    return self.Logits(theta, x.reshape(-1, shp[-1])).reshape(*shp[:-1], -1)

  @classmethod
  def FPropMeta(cls, p, inputs, *args):
    t, b, d = inputs[0], inputs[1], inputs[2]
    return NestedMap(flops=2 * t * b * d * p.num_classes,
                     out_shapes=(type(inputs)([t, b, p.num_classes]),))


class GPipeTransformerEmbeddingLayer(base_layer.BaseLayer):
  """Token + position embeddings for source (and target) ids (ref :397)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('is_transparent', False, 'Kept for parity.')
    p.Define('src', None, 'NestedMap(token_emb, position_emb, input_dropout_prob).')
    p.Define('tgt', None, 'Same for the target side (None: encoder only).')
    p.Define('packed_input', False, 'Packed inputs.')
    p.Define('add_tgt_embedding_layer', False, 'Separate target embeddings.')
    p.Define('batch_dim', 1, 'Batch dimension of the ids ([T, B] → 1).')
    p.Define('vocab_size', 0, 'Vocab size.')
    p.Define('model_dim', 0, 'Model dim.')
    p.Define('max_seq_len', 1024, 'Positional table length.')
    p.Define('input_dropout_prob', 0.0, 'Dropout.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    emb = layers.SimpleEmbeddingLayer.Params().Set(vocab_size=p.vocab_size,
                                                   embedding_dim=p.model_dim)
    pos = layers.PositionalEmbeddingLayer.Params().Set(embedding_dim=p.model_dim)
    self.CreateChild('src_token_emb', emb.Copy())
    self.CreateChild('src_pos_emb', pos.Copy())
    self.CreateChild('src_dropout', layers.DropoutLayer.Params().Set(
        keep_prob=1.0 - p.input_dropout_prob))
    if p.add_tgt_embedding_layer:
      self.CreateChild('tgt_token_emb', emb.Copy())
      self.CreateChild('tgt_pos_emb', pos.Copy())

  def _Embed(self, theta, side, ids):
    p = self.params
    t = ids.shape[0]
    x = self.children[side + '_token_emb'].EmbLookup(theta[side + '_token_emb'], ids.long())
    pos = self.children[side + '_pos_emb'].FProp(theta[side + '_pos_emb'], t).unsqueeze(1)
    return self.src_dropout.FProp(theta.src_dropout,
                                  x * (p.model_dim ** 0.5) + pos.to(device=x.device, dtype=x.dtype))

  def FProp(self, theta, source_id, source_paddings, target_id=None, target_paddings=None,
            source_segment_id=None, target_segment_id=None, *args):
    p = self.params
    src = self._Embed(theta, 'src', source_id)
    tgt = None
    if target_id is not None:
      tgt = self._Embed(theta, 'tgt' if p.add_tgt_embedding_layer else 'src', target_id)
    return (src, source_paddings, tgt, target_paddings, source_segment_id, target_segment_id,
            None, None, None, None)

  @classmethod
  def FPropMeta(cls, p, inputs, *args):
    t, b = inputs[0], inputs[1]
    shape = type(inputs)([t, b, p.model_dim])
    return NestedMap(flops=t * b * p.model_dim * 3, out_shapes=(shape,) + tuple(args))


class GPipeTransformerStack(PipeliningLayer):
  """Encoder (+decoder) Transformer stack pipelined over `num_splits` cells (ref :576)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('model_dim', 1024, 'Model dim.')
    p.Define('num_encoder_layers', 0, 'Encoder layers.')
    p.Define('num_decoder_layers', 0, 'Decoder layers.')
    p.Define('use_pipelined_embeddings', False, 'Embeddings inside the pipeline.')
    p.Define('emb_tpl', GPipeTransformerEmbeddingLayer.Params(), 'Embedding tpl.')
    p.Define('softmax_tpl', GPipeTransformerSoftmaxLayer.Params(), 'Softmax tpl.')
    p.Define('encoder_tpl', GPipeTransformerLayer.Params(), 'Encoder layer tpl.')
    p.Define('decoder_tpl', GPipeTransformerLayer.Params(), 'Decoder layer tpl.')
    p.Define('transparent_merger_dropout_prob', 0.1, 'Kept for parity.')
    p.Define('is_transparent', False, 'Kept for parity.')
    p.Define('num_transparent_outputs', 0, 'Kept for parity.')
    p.Define('packed_input', False, 'Packed inputs.')
    p.Define('normalize_encoder', False, 'Kept for parity.')
    p.Define('normalize_output', False, 'Kept for parity.')
    p.Define('num_splits', 1, 'Pipeline stages.')
    p.Define('splits', 1, 'int or list of last-layer indices per stage.')
    p.encoder_tpl.has_aux_atten = False
    p.decoder_tpl.has_aux_atten = True
    p.decoder_tpl.mask_self_atten = True
    p.batch_dim = 1
    return p

  def __init__(self, params):
    p = params
    layer_ps = []
    if p.use_pipelined_embeddings:
      layer_ps.append(p.emb_tpl.Copy().Set(name='emb', model_dim=p.model_dim))
    for i in range(p.num_encoder_layers):
      layer_ps.append(p.encoder_tpl.Copy().Set(name='encoder_%d' % i, source_dim=p.model_dim,
                                               packed_input=p.packed_input))
    for i in range(p.num_decoder_layers):
      layer_ps.append(p.decoder_tpl.Copy().Set(name='decoder_%d' % i, source_dim=p.model_dim,
                                               packed_input=p.packed_input))
    if p.use_pipelined_embeddings and p.softmax_tpl is not None and p.softmax_tpl.num_classes:
      layer_ps.append(p.softmax_tpl.Copy().Set(name='softmax', input_dim=p.model_dim,
                                               inputs_from_decoder=p.num_decoder_layers > 0))
    n = len(layer_ps)
    splits = p.splits
    if isinstance(splits, int):
      k = max(splits, p.num_splits, 1)
      per = -(-n // k)
      splits = [min((i + 1) * per, n) for i in range(k)]
    cells, start = [], 0
    for si, end in enumerate(splits):
      cells.append(gpipe.FeatureExtractionLayer.Params().Set(
          name='cell_%d' % si, sub=layer_ps[start:end]))
      start = end
    p.cell_tpl = cells
    super().__init__(p)

  def FProp(self, theta, source_input, source_paddings, target_input=None, target_paddings=None,
            source_segment_id=None, target_segment_id=None, labels=None, label_weights=None,
            source_task_id=None, target_task_id=None):
    p = self.params
    args = (source_input, source_paddings, target_input, target_paddings, source_segment_id,
            target_segment_id) + ((None, None, source_task_id, target_task_id)
                                  if not p.use_pipelined_embeddings else ())
    out = super().FProp(theta, *args)
    if isinstance(out, tuple):
      return out[2] if p.num_decoder_layers > 0 and out[2] is not None else out[0]
    return out


class GPipeBatchMajorTransformerStack(PipeliningLayer):
  """Batch-major variant built from `bma.GPipeBatchMajorTransformerLayer` (ref :1147)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('model_dim', 1024, 'Model dim.')
    p.Define('num_encoder_layers', 0, 'Encoder layers.')
    p.Define('num_decoder_layers', 0, 'Decoder layers.')
    p.Define('encoder_tpl', bma.GPipeBatchMajorTransformerLayer.Params(), 'Encoder tpl.')
    p.Define('decoder_tpl', bma.GPipeBatchMajorTransformerLayer.Params(), 'Decoder tpl.')
    p.Define('packed_input', False, 'Packed inputs.')
    p.Define('num_splits', 1, 'Pipeline stages.')
    p.decoder_tpl.has_aux_atten = True
    p.decoder_tpl.mask_self_atten = True
    return p

  def __init__(self, params):
    p = params
    layer_ps = [p.encoder_tpl.Copy().Set(name='encoder_%d' % i, input_dim=p.model_dim,
                                         packed_input=p.packed_input)
                for i in range(p.num_encoder_layers)]
    layer_ps += [p.decoder_tpl.Copy().Set(name='decoder_%d' % i, input_dim=p.model_dim,
                                          packed_input=p.packed_input)
                 for i in range(p.num_decoder_layers)]
    k = max(p.num_splits, 1)
    per = -(-len(layer_ps) // k)
    p.cell_tpl = [gpipe.FeatureExtractionLayer.Params().Set(
        name='cell_%d' % i, sub=layer_ps[i * per:(i + 1) * per]) for i in range(k)]
    super().__init__(p)

  def FProp(self, theta, source_vecs, source_paddings, target_vecs=None, target_paddings=None,
            encoder_self_atten_segment_mask=None, decoder_self_atten_segment_mask=None,
            decoder_cross_atten_segment_mask=None):
    out = super().FProp(theta, source_vecs, source_paddings, target_vecs, target_paddings,
                        encoder_self_atten_segment_mask, decoder_self_atten_segment_mask,
                        decoder_cross_atten_segment_mask)
    if isinstance(out, tuple):
      return out[2] if self.params.num_decoder_layers > 0 else out[0]
    return out
