"""GPipe Transformer stacks (ref `lingvo/core/layers_with_gpipe.py`).

`GPipeTransformerLayer` (ref :165) is a time-major TransformerLayer with the
"all tensors in / all tensors out" convention so layers chain inside a
`FeatureExtractionLayer`; `GPipeTransformerEmbeddingLayer` (ref :397) and
`GPipeTransformerSoftmaxLayer` (ref :355) bracket the stack;
`GPipeTransformerStack` (ref :576) builds encoder/decoder layer lists (plain or Evolved
Transformer, optionally transparent), splits them
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


def _CommonGPipeTransformerParams(p):
  """Params shared by every GPipe transformer layer flavour (ref :28)."""
  p.Define('is_transparent', False,
           'Encoder layers accumulate a learned weighted sum of all layer inputs '
           '(transparent attention); the final encoder layer emits the merged vector.')
  p.Define('transparent_merger_tpl', None,
           'DeterministicWeightsLayer params; set on the FIRST encoder layer only, which '
           'creates the mixing weights that travel down the pipeline.')
  p.Define('final_enc_layer', False, 'Last encoder layer: closes the transparent merge.')
  p.Define('normalize_output', False, 'Layer-normalise the layer output.')
  p.Define('ln_tpl', layers.LayerNorm.Params(), 'Layer norm params.')
  p.Define('num_transparent_outputs', 0, 'Deprecated; unused.')
  return p


def _CommonGPipeTransformerInit(layer):
  p = layer.params
  assert p.name
  if p.normalize_output:
    layer.CreateChild('layer_norm', p.ln_tpl.Copy().Set(name='encoder_ln',
                                                        input_dim=p.source_dim))
  if p.is_transparent and p.transparent_merger_tpl is not None:
    layer.CreateChild('transparent_merger',
                      p.transparent_merger_tpl.Copy().Set(name='transparent_0'))


class _GPipeLayerMixin:
  """FProp(source_vecs, source_paddings, target_vecs, target_paddings, source_segment_id,
  target_segment_id, transparent_acc, transparent_acc_helper[, source_task_id,
  target_task_id]) → the same tuple with the processed stream replaced (ref :64-145).

  Transparent encoders: `transparent_acc_helper` is the vector of remaining mixing weights
  (one per encoder layer input plus one for the final output); every layer adds
  `helper[0] * its input` to `transparent_acc` and pops that weight; the final layer returns
  `acc + helper[-1] * h`."""

  def _InnerFProp(self, theta, *args, **kwargs):
    raise NotImplementedError

  def FProp(self, theta, source_vecs, source_paddings, target_vecs=None, target_paddings=None,
            source_segment_id=None, target_segment_id=None, transparent_acc=None,
            transparent_acc_helper=None, source_task_id=None, target_task_id=None):
    p = self.params
    has_task = source_task_id is not None or target_task_id is not None
    if p.has_aux_atten:       # decoder layer
      assert target_vecs is not None and target_paddings is not None
      h, _ = self._InnerFProp(theta, target_vecs, target_paddings, source_vecs,
                              source_paddings, target_segment_id, source_segment_id)
      target_vecs = h
    else:
      h, _ = self._InnerFProp(theta, source_vecs, source_paddings,
                              source_segment_id=source_segment_id)
      if p.is_transparent:
        if p.transparent_merger_tpl is not None:
          transparent_acc_helper = self.transparent_merger.FProp(theta.transparent_merger)
          transparent_acc = torch.zeros_like(source_vecs)
        w = transparent_acc_helper.to(source_vecs.dtype)
        transparent_acc = transparent_acc + w[0] * source_vecs
        if p.final_enc_layer:
          h = transparent_acc + h * w[-1]
          transparent_acc, transparent_acc_helper = None, None
        else:
          transparent_acc_helper = transparent_acc_helper[1:]
      if p.normalize_output:
        h = self.layer_norm.FProp(theta.layer_norm, h)
      source_vecs = h
    out = (source_vecs, source_paddings, target_vecs, target_paddings, source_segment_id,
           target_segment_id, transparent_acc, transparent_acc_helper)
    return out + ((source_task_id, target_task_id) if has_task else (None, None))

  @classmethod
  def FPropMeta(cls, p, inputs, *args):
    py_utils.CheckShapes((inputs,))
    t, b, d = inputs[0], inputs[1], inputs[2]
    ff = cls._HiddenDim(p)
    flops = b * t * (8 * d * d + 4 * t * d + 4 * d * ff)
    args = tuple(args)
    if not p.has_aux_atten and p.is_transparent and len(args) >= 7:   # transparent encoder
      if p.transparent_merger_tpl is not None:
        args = args[:5] + (inputs, type(inputs)([p.transparent_merger_tpl.num_sources])) \
            + args[7:]
      args = args[:6] + (type(inputs)([args[6][0] - 1]),) + args[7:]
      if p.final_enc_layer:
        args = args[:5] + (None, None) + args[7:]
    return NestedMap(flops=flops, out_shapes=(inputs,) + args)

  @classmethod
  def _SetupAttentionDeterministicDropout(cls, tr_atten_tpl):
    tr_atten_tpl.residual_dropout_tpl = layers.DeterministicDropoutLayer.Params()
    atten = tr_atten_tpl.atten_tpl
    if 'atten_dropout_deterministic' in atten:
      atten.atten_dropout_deterministic = True
    inner = atten.Get('inner_atten_params') if 'inner_atten_params' in atten else None
    if inner is not None and 'atten_dropout_deterministic' in inner:
      inner.atten_dropout_deterministic = True

  @classmethod
  def _SetupTransformerDeterministicDropout(cls, tpl):
    cls._SetupAttentionDeterministicDropout(tpl.tr_atten_tpl)
    if tpl.Get('tr_aux_atten_tpl') is not None:
      cls._SetupAttentionDeterministicDropout(tpl.tr_aux_atten_tpl)
    tpl.tr_fflayer_tpl.residual_dropout_tpl = layers.DeterministicDropoutLayer.Params()
    tpl.tr_fflayer_tpl.fflayer_tpl.dropout = layers.DeterministicDropoutLayer.Params()


class GPipeTransformerLayer(_GPipeLayerMixin, layers_with_attention.TransformerLayer):
  """Time-major TransformerLayer with the pipeline tuple convention (ref :165)."""

  @classmethod
  def Params(cls):
    return _CommonGPipeTransformerParams(super().Params())

  def __init__(self, params):
    super().__init__(params)
    _CommonGPipeTransformerInit(self)

  def _InnerFProp(self, theta, *args, **kwargs):
    return layers_with_attention.TransformerLayer.FProp(self, theta, *args, **kwargs)

  @classmethod
  def _HiddenDim(cls, p):
    return p.tr_fflayer_tpl.hidden_dim

  @classmethod
  def SetupDeterministicDropout(cls, params):
    """Every dropout in the layer becomes keyed by (global step, step seed) so a re-run of
    the forward (micro-batch rematerialisation) draws the same mask (ref :207)."""
    cls._SetupTransformerDeterministicDropout(params)
    return params


class GPipeEvolvedTransformerEncoderLayer(_GPipeLayerMixin,
                                          layers_with_attention.EvolvedTransformerEncoderLayer):
  """Evolved Transformer encoder layer for pipelines (ref :224)."""

  @classmethod
  def Params(cls):
    return _CommonGPipeTransformerParams(super().Params())

  def __init__(self, params):
    super().__init__(params)
    _CommonGPipeTransformerInit(self)

  def _InnerFProp(self, theta, *args, **kwargs):
    return layers_with_attention.EvolvedTransformerEncoderLayer.FProp(self, theta, *args,
                                                                      **kwargs)

  @classmethod
  def _HiddenDim(cls, p):
    return p.transformer_tpl.tr_fflayer_tpl.hidden_dim

  @classmethod
  def SetupDeterministicDropout(cls, params):
    cls._SetupTransformerDeterministicDropout(params.transformer_tpl)
    for name in ('branched_convs_tpl', 'glu_tpl'):
      if name in params and 'dropout_tpl' in params.Get(name):
        params.Get(name).dropout_tpl = layers.DeterministicDropoutLayer.Params()
    return params


class GPipeEvolvedTransformerDecoderLayer(_GPipeLayerMixin,
                                          layers_with_attention.EvolvedTransformerDecoderLayer):
  """Evolved Transformer decoder layer for pipelines (ref :289)."""

  @classmethod
  def Params(cls):
    return _CommonGPipeTransformerParams(super().Params())

  def __init__(self, params):
    super().__init__(params)
    _CommonGPipeTransformerInit(self)

  def _InnerFProp(self, theta, *args, **kwargs):
    return layers_with_attention.EvolvedTransformerDecoderLayer.FProp(self, theta, *args,
                                                                      **kwargs)

  @classmethod
  def _HiddenDim(cls, p):
    return p.transformer_tpl.tr_fflayer_tpl.hidden_dim

  @classmethod
  def SetupDeterministicDropout(cls, params):
    cls._SetupTransformerDeterministicDropout(params.transformer_tpl)
    if 'dropout_tpl' in params.branched_convs_tpl:
      params.branched_convs_tpl.dropout_tpl = layers.DeterministicDropoutLayer.Params()
    cls._SetupAttentionDeterministicDropout(params.tr_atten_tpl)
    cls._SetupAttentionDeterministicDropout(params.tr_double_heads_atten_tpl)
    return params


class DeterministicWeightsLayer(base_layer.BaseLayer):
  """Learned mixing weights `[num_sources]` with (deterministic) dropout and an optional
  softmax floor — the transparent-attention merger of a pipelined encoder (ref :909)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('num_sources', 0, 'Number of inputs to combine.')
    p.Define('weighted_merger_dropout_prob', 0.0, 'Dropout on the weights.')
    p.Define('weighted_merger_softmax', True, 'Softmax-normalise the weights.')
    p.Define('global_weight_scale', 1.0, 'Scale on the learned weights.')
    p.Define('minimal_prob', 0.0, 'Lower bound of every normalised weight.')
    p.Define('dropout_tpl', layers.DeterministicDropoutLayer.Params(), 'Dropout layer.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    assert p.name and p.num_sources > 0
    self.CreateChild('weighted_merger_dropout', p.dropout_tpl.Copy().Set(name='dropout'))
    self.CreateVariable('sum_weight', py_utils.WeightParams(
        [p.num_sources], py_utils.WeightInit.Constant(0.0), p.dtype,
        [self.__class__.__name__ + '_vars']))

  def FProp(self, theta):
    p = self.params
    # the 1/num_sources offset only matters without softmax (it cancels under it)
    w = theta.sum_weight * p.global_weight_scale + 1.0 / p.num_sources
    w = self.weighted_merger_dropout.FProp(theta.weighted_merger_dropout, w)
    if p.weighted_merger_softmax:
      residual = p.minimal_prob * p.num_sources
      assert 0.0 <= residual < 1.0
      w = torch.softmax(w, 0) * (1.0 - residual) + p.minimal_prob
    return w


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
  """Token + position (+ task) embeddings for source and target ids, first layer of the
  pipeline (ref :397)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('token_emb', layers.SimpleEmbeddingLayer.Params(), 'Token embedding params.')
    p.Define('position_emb', layers.PositionalEmbeddingLayer.Params(),
             'Position embedding params.')
    p.Define('input_dropout_prob', 0.0, 'Input dropout.')
    p.Define('dropout_tpl', layers.DropoutLayer.Params(),
             'Dropout flavour (deterministic when the stack is split / micro-batched).')
    p.Define('add_tgt_embedding_layer', False, 'Separate target-side embeddings.')
    p.Define('packed_input', False, 'Position embeddings from segment positions.')
    p.Define('is_transparent', False, 'Transparent encoder downstream.')
    p.Define('max_seq_len', 1024, 'Positional table length used while decoding.')
    p.Define('target_vocab_size', 0, 'Target vocab size if different.')
    p.Define('dec_task_emb', None, 'Task embedding added to every decoder position.')
    p.Define('enc_task_emb', None, 'Task embedding added to every encoder position.')
    p.Define('batch_dim', 1, 'Batch dimension of the ids ([T, B] → 1).')
    p.Define('ret_task_ids', False, 'Forward the task ids down the pipeline.')
    p.Define('scale_sqrt_depth', True, 'Multiply token embeddings by sqrt(model_dim).')
    # shorthand accepted in place of token_emb / position_emb settings
    p.Define('vocab_size', 0, 'Sets token_emb.vocab_size.')
    p.Define('model_dim', 0, 'Sets the embedding dims.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    tok = p.token_emb.Copy()
    pos = p.position_emb.Copy()
    if p.vocab_size:
      tok.vocab_size = p.vocab_size
    if p.model_dim:
      tok.embedding_dim = p.model_dim
      pos.embedding_dim = p.model_dim
    self._dim = tok.embedding_dim
    drop = p.dropout_tpl.Copy().Set(keep_prob=1.0 - p.input_dropout_prob)
    self.CreateChild('src_token_emb', tok.Copy())
    self.CreateChild('src_pos_emb', pos.Copy())
    self.CreateChild('src_dropout', drop.Copy())
    if p.enc_task_emb is not None:
      self.CreateChild('src_task_emb', p.enc_task_emb.Copy())
    if p.add_tgt_embedding_layer:
      ttok = tok.Copy()
      if p.target_vocab_size:
        ttok.vocab_size = p.target_vocab_size
      self.CreateChild('tgt_token_emb', ttok)
      self.CreateChild('tgt_pos_emb', pos.Copy())
      self.CreateChild('tgt_dropout', drop.Copy())
      if p.dec_task_emb is not None:
        self.CreateChild('tgt_task_emb', p.dec_task_emb.Copy())

  def GetEmbeddings(self, theta, side, ids, segment_pos=None, task_ids=None, t=None):
    """Embeds `ids` with the `side` ('src' | 'tgt') tables; `t`: single decode position."""
    p = self.params
    tok = self.children[side + '_token_emb']
    x = tok.EmbLookup(theta[side + '_token_emb'], ids.long())
    if p.scale_sqrt_depth:
      x = x * (self._dim ** 0.5)
    pos_layer, pos_theta = self.children[side + '_pos_emb'], theta[side + '_pos_emb']
    if t is not None:
      pos = pos_layer.FProp(pos_theta, p.max_seq_len)[int(t):int(t) + 1]
      pos = pos.unsqueeze(p.batch_dim)
    elif p.packed_input and segment_pos is not None:
      pos = pos_layer.FPropWithPosition(pos_theta, segment_pos)
    else:
      time_dim = 0 if p.batch_dim else 1
      pos = pos_layer.FProp(pos_theta, ids.shape[time_dim]).unsqueeze(p.batch_dim)
    x = x + pos.to(device=x.device, dtype=x.dtype)
    task_name = side + '_task_emb'
    if task_ids is not None and task_name in self.children:
      x = x + self.children[task_name].EmbLookup(theta[task_name], task_ids.long())
    return self.children[side + '_dropout'].FProp(theta[side + '_dropout'], x)

  def GetEncoderEmbeddingsDefaultTheta(self, input_ids, task_ids=None):
    return self.GetEmbeddings(self.theta, 'src', input_ids, task_ids=task_ids)

  def GetDecoderEmbeddingsDefaultTheta(self, input_ids, task_ids=None, t=None):
    side = 'tgt' if self.params.add_tgt_embedding_layer else 'src'
    return self.GetEmbeddings(self.theta, side, input_ids, task_ids=task_ids, t=t)

  def FProp(self, theta, source_id, source_paddings, target_id=None, target_paddings=None,
            source_segment_id=None, target_segment_id=None, source_segment_pos=None,
            target_segment_pos=None, source_task_id=None, target_task_id=None):
    p = self.params
    src = self.GetEmbeddings(theta, 'src', source_id, source_segment_pos, source_task_id)
    tgt = None
    if target_id is not None:
      side = 'tgt' if p.add_tgt_embedding_layer else 'src'
      tgt = self.GetEmbeddings(theta, side, target_id, target_segment_pos, target_task_id)
    rets = (src, source_paddings, tgt, target_paddings, source_segment_id, target_segment_id,
            None, None)
    return rets + ((source_task_id, target_task_id) if p.ret_task_ids else (None, None))

  @classmethod
  def FPropMeta(cls, p, inputs, *args):
    py_utils.CheckShapes((inputs,))
    d0, d1 = inputs[0], inputs[1]
    dim = p.model_dim or p.token_emb.embedding_dim
    shape = type(inputs)([d0, d1, dim])
    args = list(args)
    if p.add_tgt_embedding_layer and len(args) > 1 and args[1] is not None:
      args[1] = type(inputs)([args[1][0], args[1][1], dim])
    args = args[:5] + [None, None] + (args[7:] if p.ret_task_ids else [])
    return NestedMap(flops=d0 * d1 * dim * 3, out_shapes=(shape,) + tuple(args))


class GPipeTransformerStack(PipeliningLayer):
  """Encoder (+decoder) Transformer stack pipelined over cells (ref :576): optional
  pipelined embeddings and softmax, transparent encoder merging, final-encoder layer norm,
  deterministic dropout whenever a forward may be re-run (several cells or micro-batches)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('splits', 1, 'Number of cells, or the list of last layer indices per cell '
             '(ascending; the last entry is the number of layers).')
    p.Define('num_splits', 1, 'Alias: number of cells when `splits` is an int.')
    p.Define('model_dim', 1024, 'Model dim.')
    p.Define('num_encoder_layers', 0, 'Encoder layers.')
    p.Define('num_decoder_layers', 0, 'Decoder layers.')
    p.Define('use_pipelined_embeddings', False, 'Embeddings and softmax inside the pipeline.')
    p.Define('emb_tpl', GPipeTransformerEmbeddingLayer.Params(), 'Embedding tpl.')
    p.Define('softmax_tpl', GPipeTransformerSoftmaxLayer.Params(), 'Softmax tpl.')
    p.Define('label_smoothing', None, 'Label smoother params.')
    p.Define('encoder_tpl', GPipeTransformerLayer.Params(), 'Encoder layer tpl.')
    p.Define('decoder_tpl', GPipeTransformerLayer.Params(), 'Decoder layer tpl.')
    p.Define('transparent_merger_dropout_prob', 0.1, 'Dropout on the transparent weights.')
    p.Define('is_transparent', False, 'Encoder output = learned mix of all encoder layers.')
    p.Define('transparent_merger_tpl', DeterministicWeightsLayer.Params(), 'Mixing weights.')
    p.Define('packed_input', False, 'Packed inputs.')
    p.Define('normalize_encoder', False, 'Layer-normalise the final encoder output.')
    p.encoder_tpl.has_aux_atten = False
    p.decoder_tpl.has_aux_atten = True
    p.decoder_tpl.mask_self_atten = True
    p.batch_dim = 1
    return p

  def __init__(self, params):
    p = params
    num_layers = p.num_encoder_layers + p.num_decoder_layers
    splits = p.splits
    if isinstance(splits, (list, tuple)):
      splits = list(splits)
      assert splits[-1] == num_layers and all(a <= b for a, b in zip(splits, splits[1:]))
    else:
      k = max(int(splits), int(p.num_splits), 1)
      per = (num_layers - 1) // k + 1
      splits = [min((i + 1) * per, num_layers) for i in range(k)]
      splits[-1] = num_layers
    self._splits = splits
    rerun = len(splits) > 1 or p.num_micro_batches > 1
    merger = None
    if p.is_transparent:
      merger = p.transparent_merger_tpl.Copy().Set(num_sources=p.num_encoder_layers + 1)
      merger.dropout_tpl.keep_prob = 1.0 - p.transparent_merger_dropout_prob
    body = []
    for i in range(p.num_encoder_layers):
      lp = p.encoder_tpl.Copy().Set(name='encoder_%d' % i, source_dim=p.model_dim,
                                    packed_input=p.packed_input)
      last = i == p.num_encoder_layers - 1
      if p.is_transparent:
        lp.is_transparent = True
        lp.final_enc_layer = last
        if i == 0:
          lp.transparent_merger_tpl = merger
      if p.normalize_encoder and last:
        lp.normalize_output = True
        lp.final_enc_layer = True
      if rerun:
        lp = lp.cls.SetupDeterministicDropout(lp)
      assert not lp.has_aux_atten
      body.append(lp)
    for i in range(p.num_decoder_layers):
      lp = p.decoder_tpl.Copy().Set(name='decoder_%d' % i, source_dim=p.model_dim,
                                    packed_input=p.packed_input)
      if 'mask_self_atten' in lp:
        lp.mask_self_atten = True
      if rerun:
        lp = lp.cls.SetupDeterministicDropout(lp)
      assert lp.has_aux_atten
      body.append(lp)
    emb = softmax = None
    if p.use_pipelined_embeddings:
      emb = p.emb_tpl.Copy().Set(name='emb', packed_input=p.packed_input,
                                 is_transparent=p.is_transparent, batch_dim=p.batch_dim,
                                 add_tgt_embedding_layer=p.num_decoder_layers > 0)
      if not (emb.model_dim or emb.token_emb.embedding_dim):
        emb.model_dim = p.model_dim
      if rerun:
        emb.dropout_tpl = layers.DeterministicDropoutLayer.Params()
      if p.softmax_tpl is not None and p.softmax_tpl.num_classes:
        softmax = p.softmax_tpl.Copy().Set(name='softmax', input_dim=p.model_dim,
                                           inputs_from_decoder=p.num_decoder_layers > 0)
    cells, start = [], 0
    for si, end in enumerate(splits):
      sub = body[start:end]
      if si == 0 and emb is not None:
        sub = [emb] + sub
      if si == len(splits) - 1 and softmax is not None:
        sub = sub + [softmax]
      cells.append(gpipe.FeatureExtractionLayer.Params().Set(name='cell_%d' % si, sub=sub))
      start = end
    p.cell_tpl = cells
    super().__init__(p)
    if p.label_smoothing is not None:
      self.CreateChild('smoother', p.label_smoothing)

  def _LayersByPrefix(self, prefix):
    out = []
    for si in range(len(self._splits)):
      cell = self.children['cell_%d' % si]
      out += [(int(n[len(prefix):]), cell.children[n]) for n in cell.children
              if n.startswith(prefix)]
    return [l for _, l in sorted(out, key=lambda kv: kv[0])]

  def GetEncoders(self):
    return self._LayersByPrefix('encoder_')

  def GetDecoders(self):
    decoders = self._LayersByPrefix('decoder_')
    assert len(decoders) == self.params.num_decoder_layers
    return decoders

  def Logits(self, theta, inputs):
    """Softmax logits of the pipelined softmax layer for `[..., model_dim]` inputs."""
    last = 'cell_%d' % (len(self._splits) - 1)
    return self.children[last].softmax.Logits(theta[last].softmax, inputs)

  def _Emb(self):
    return self.children['cell_0'].children['emb']

  def EncoderEmbedFPropDefaultTheta(self, source_id, source_task_id=None):
    return self._Emb().GetEncoderEmbeddingsDefaultTheta(source_id, source_task_id)

  def DecoderEmbedFPropDefaultTheta(self, tgt_id, tgt_task_id=None, t=None):
    return self._Emb().GetDecoderEmbeddingsDefaultTheta(tgt_id, tgt_task_id, t)

  def EncoderFPropDefaultTheta(self, source_vecs, source_paddings, source_segment_id=None):
    """Runs the encoder layers only, outside the pipeline (decoding)."""
    state = (source_vecs, source_paddings, None, None, source_segment_id, None, None, None)
    for layer in self.GetEncoders():
      state = layer.FProp(layer.theta, *state)[:8]
    return state[0]

  def FProp(self, theta, source_input, source_paddings, target_input=None, target_paddings=None,
            source_segment_id=None, target_segment_id=None, labels=None, label_weights=None,
            source_segment_pos=None, target_segment_pos=None, source_task_id=None,
            target_task_id=None):
    """Ids (`use_pipelined_embeddings`) or vectors in → logits (pipelined softmax) or the
    decoder (else encoder) output vectors. With `labels`, returns (xent [T,B], logits)."""
    p = self.params
    if p.use_pipelined_embeddings:
      args = (source_input, source_paddings, target_input, target_paddings, source_segment_id,
              target_segment_id, source_segment_pos, target_segment_pos, source_task_id,
              target_task_id)
    else:
      args = (source_input, source_paddings, target_input, target_paddings, source_segment_id,
              target_segment_id, None, None, source_task_id, target_task_id)
    out = super().FProp(theta, *args)
    if isinstance(out, tuple):
      out = out[2] if p.num_decoder_layers > 0 and out[2] is not None else out[0]
    if labels is None:
      return out
    logits = out.float()
    targets = torch.nn.functional.one_hot(labels.long(), logits.shape[-1]).float()
    if p.label_smoothing is not None:
      targets = self.smoother.FProp(theta.smoother, target_paddings.t(), labels.t().long(),
                                    target_ids=None).transpose(0, 1)
    del label_weights          # per-example xent is unweighted, as in XentLossFromLogits
    xent = -(targets * torch.log_softmax(logits, -1)).sum(-1)
    return xent, logits


class GPipeEvolvedTransformerStack(GPipeTransformerStack):
  """The same pipeline built from Evolved Transformer layers (ref :891)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.encoder_tpl = GPipeEvolvedTransformerEncoderLayer.Params()
    p.decoder_tpl = GPipeEvolvedTransformerDecoderLayer.Params()
    return p


class GPipeBatchMajorTransformerSoftmaxLayer(layers.SimpleFullSoftmax):
  """Softmax closing a batch-major pipeline (segment *masks* in the tuple) (ref :976)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('inputs_from_decoder', False, 'Read target_vecs instead of source_vecs.')
    return p

  def FProp(self, theta, source_vecs, source_paddings, target_vecs=None, target_paddings=None,
            encoder_self_atten_segment_mask=None, decoder_self_atten_segment_mask=None,
            decoder_cross_atten_segment_mask=None):
    x = target_vecs if self.params.inputs_from_decoder else source_vecs
    shp = x.shape
    return self.Logits(theta, x.reshape(-1, shp[-1])).reshape(*shp[:-1], -1)

  @classmethod
  def FPropMeta(cls, p, inputs, *args):
    d0, d1 = (args[1][:2] if p.inputs_from_decoder else inputs[:2])
    return NestedMap(flops=2 * d0 * d1 * p.input_dim * p.num_classes,
                     out_shapes=(type(inputs)([d0, d1, p.num_classes]),))


class GPipeBatchMajorTransformerEmbeddingLayer(GPipeTransformerEmbeddingLayer):
  """Embeddings opening a batch-major pipeline (ref :1011): ids `[B, T]`; packed inputs
  turn segment ids into the three attention segment masks the batch-major layers take."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.batch_dim = 0
    return p

  def FProp(self, theta, source_id, source_paddings, target_id=None, target_paddings=None,
            source_segment_id=None, target_segment_id=None, source_segment_pos=None,
            target_segment_pos=None, source_task_id=None, target_task_id=None):
    p = self.params
    src = self.GetEmbeddings(theta, 'src', source_id, source_segment_pos, source_task_id)
    tgt = None
    if target_id is not None:
      side = 'tgt' if p.add_tgt_embedding_layer else 'src'
      tgt = self.GetEmbeddings(theta, side, target_id, target_segment_pos, target_task_id)
    enc_mask = dec_self_mask = dec_cross_mask = None
    if p.packed_input and source_segment_id is not None:
      enc_mask = bma.SegmentMask(source_segment_id, source_segment_id, dtype=src.dtype)
      if tgt is not None:
        dec_self_mask = bma.SegmentMask(target_segment_id, target_segment_id, dtype=src.dtype)
        dec_cross_mask = bma.SegmentMask(target_segment_id, source_segment_id,
                                         dtype=src.dtype)
    return (src, source_paddings, tgt, target_paddings, enc_mask, dec_self_mask,
            dec_cross_mask)

  @classmethod
  def FPropMeta(cls, p, inputs, *args):
    d0, d1 = inputs[0], inputs[1]
    dim = p.model_dim or p.token_emb.embedding_dim
    args = list(args)
    if p.add_tgt_embedding_layer and len(args) > 1 and args[1] is not None:
      args[1] = type(inputs)([args[1][0], args[1][1], dim])
    return NestedMap(flops=d0 * d1 * dim * 3,
                     out_shapes=(type(inputs)([d0, d1, dim]),) + tuple(args[:6]))


class GPipeBatchMajorTransformerStack(PipeliningLayer):
  """Batch-major variant built from `bma.GPipeBatchMajorTransformerLayer` (ref :1147), with
  optional pipelined embeddings / softmax."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('model_dim', 1024, 'Model dim.')
    p.Define('num_encoder_layers', 0, 'Encoder layers.')
    p.Define('num_decoder_layers', 0, 'Decoder layers.')
    p.Define('emb_tpl', None, 'GPipeBatchMajorTransformerEmbeddingLayer params (optional).')
    p.Define('softmax_tpl', None, 'GPipeBatchMajorTransformerSoftmaxLayer params (optional).')
    p.Define('encoder_tpl', bma.GPipeBatchMajorTransformerLayer.Params(), 'Encoder tpl.')
    p.Define('decoder_tpl', bma.GPipeBatchMajorTransformerLayer.Params(), 'Decoder tpl.')
    p.Define('packed_input', False, 'Packed inputs.')
    p.Define('num_splits', 1, 'Pipeline stages.')
    p.Define('splits', None, 'Optional list of last layer indices per stage.')
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
    n = len(layer_ps)
    if p.splits:
      splits = list(p.splits)
      assert splits[-1] == n
    else:
      k = max(p.num_splits, 1)
      per = -(-n // k)
      splits = [min((i + 1) * per, n) for i in range(k)]
    cells, start = [], 0
    for si, end in enumerate(splits):
      sub = layer_ps[start:end]
      if si == 0 and p.emb_tpl is not None:
        emb = p.emb_tpl.Copy().Set(name='emb', packed_input=p.packed_input,
                                   add_tgt_embedding_layer=p.num_decoder_layers > 0)
        if not (emb.model_dim or emb.token_emb.embedding_dim):
          emb.model_dim = p.model_dim
        sub = [emb] + sub
      if si == len(splits) - 1 and p.softmax_tpl is not None:
        sub = sub + [p.softmax_tpl.Copy().Set(
            name='softmax', input_dim=p.model_dim,
            inputs_from_decoder=p.num_decoder_layers > 0)]
      cells.append(gpipe.FeatureExtractionLayer.Params().Set(name='cell_%d' % si, sub=sub))
      start = end
    p.cell_tpl = cells
    super().__init__(p)

  def FProp(self, theta, source_vecs, source_paddings, target_vecs=None, target_paddings=None,
            encoder_self_atten_segment_mask=None, decoder_self_atten_segment_mask=None,
            decoder_cross_atten_segment_mask=None, source_segment_pos=None,
            target_segment_pos=None):
    """With `emb_tpl`: (source_ids, paddings, target_ids, paddings, source_segment_id,
    target_segment_id[, *_segment_pos]); without: vectors and the three segment masks."""
    p = self.params
    args = (source_vecs, source_paddings, target_vecs, target_paddings,
            encoder_self_atten_segment_mask, decoder_self_atten_segment_mask)
    if p.emb_tpl is not None:
      args += (source_segment_pos, target_segment_pos)
    else:
      args += (decoder_cross_atten_segment_mask,)
    out = super().FProp(theta, *args)
    if isinstance(out, tuple):
      return out[2] if p.num_decoder_layers > 0 else out[0]
    return out
