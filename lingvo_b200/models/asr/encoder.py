"""ASR encoder (ref `lingvo/tasks/asr/encoder.py:32`): SpecAugment → strided
conv+BN stack (time/frequency subsampling) → [conv-LSTM] → bidirectional LSTM
stack with optional per-layer projections → `[T, B, D]`.
"""

from __future__ import annotations

import torch

from lingvo_b200.core import base_layer
from lingvo_b200.core import bn_layers
from lingvo_b200.core import conv_layers_with_time_padding as conv_lib
from lingvo_b200.core import layers
from lingvo_b200.core import rnn_cell
from lingvo_b200.core import rnn_layers
from lingvo_b200.core import spectrum_augmenter
from lingvo_b200.core.nested_map import NestedMap


class AsrEncoder(base_layer.BaseLayer):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('specaugment_network', spectrum_augmenter.SpectrumAugmenter.Params(),
             'SpecAugment params.')
    p.Define('use_specaugment', False, 'Apply SpecAugment in training.')
    p.Define('lstm_tpl', rnn_cell.LSTMCellSimple.Params(), 'LSTM cell template.')
    p.Define('cnn_tpl', conv_lib.Conv2DLayerWithPadding.Params(), 'Conv template.')
    p.Define('proj_tpl', layers.ProjectionLayer.Params(), 'Projection template.')
    p.Define('highway_skip', False, 'Kept for parity.')
    p.Define('conv_lstm_tpl', rnn_cell.ConvLSTMCell.Params(), 'Conv-LSTM template.')
    p.Define('conv_filter_shapes', [(3, 3, 1, 32), (3, 3, 32, 32)], 'Conv filters.')
    p.Define('conv_filter_strides', [(2, 2), (2, 2)], 'Conv strides.')
    p.Define('input_shape', [None, None, 80, 1], '[B, T, F, C].')
    p.Define('lstm_cell_size', 256, 'LSTM cell size per direction.')
    p.Define('num_cnn_layers', 2, 'Conv layers.')
    p.Define('num_conv_lstm_layers', 0, 'Conv-LSTM layers.')
    p.Define('num_lstm_layers', 3, 'Bi-LSTM layers.')
    p.Define('project_lstm_output', True, 'Projection between LSTM layers.')
    p.Define('pad_steps', 6, 'Extra padded frames appended to the input.')
    p.Define('residual_start', 0, 'First LSTM layer with a residual (0: none).')
    p.Define('residual_stride', 1, 'Residual every n layers.')
    p.Define('bidi_rnn_type', 'func', 'Kept for parity.')
    p.Define('extra_per_layer_outputs', False, 'Kept for parity.')
    p.Define('final_proj', None, 'Optional final projection.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    assert p.num_cnn_layers == len(p.conv_filter_shapes) == len(p.conv_filter_strides)
    if p.use_specaugment:
      self.CreateChild('specaugment', p.specaugment_network)
    convs, bns = [], []
    f, c = p.input_shape[2], p.input_shape[3]
    for i in range(p.num_cnn_layers):
      shape, stride = p.conv_filter_shapes[i], p.conv_filter_strides[i]
      convs.append(p.cnn_tpl.Copy().Set(name='conv_L%d' % i, filter_shape=tuple(shape),
                                        filter_stride=tuple(stride)))
      bns.append(bn_layers.BatchNormLayer.Params().Set(name='bn_L%d' % i, dim=shape[3]))
      f = -(-f // stride[1])
      c = shape[3]
    self.CreateChildren('conv', convs)
    self.CreateChildren('conv_bn', bns)
    self._conv_out_dim = f * c
    rnns, projs = [], []
    idim = self._conv_out_dim
    for i in range(p.num_lstm_layers):
      cell = p.lstm_tpl.Copy().Set(num_input_nodes=idim, num_output_nodes=p.lstm_cell_size)
      rnns.append(rnn_layers.BidirectionalFRNN.Params().Set(
          name='brnn_L%d' % i, fwd=cell.Copy(), bak=cell.Copy()))
      idim = 2 * p.lstm_cell_size
      if p.project_lstm_output and i < p.num_lstm_layers - 1:
        projs.append(p.proj_tpl.Copy().Set(name='proj_L%d' % i, input_dim=idim,
                                           output_dim=idim, batch_norm=True,
                                           activation='RELU'))
    self.CreateChildren('rnn', rnns)
    self.CreateChildren('proj', projs)
    self._out_dim = idim
    if p.final_proj is not None:
      self.CreateChild('final_proj', p.final_proj.Copy().Set(input_dim=idim))

  @property
  def output_dim(self):
    return self._out_dim

  def FProp(self, theta, batch, state0=None):
    """batch.src_inputs [B,T,F,C], batch.paddings [B,T] → encoded [T',B,D]."""
    p = self.params
    x, pad = batch.src_inputs.float(), batch.paddings.float()
    if p.use_specaugment and not self.do_eval:
      x, pad = self.specaugment.FProp(theta.specaugment, x, pad)
    if p.pad_steps > 0:
      x = torch.nn.functional.pad(x, (0, 0, 0, 0, 0, p.pad_steps))
      pad = torch.nn.functional.pad(pad, (0, p.pad_steps), value=1.0)
    for i, conv in enumerate(self.conv):
      x, pad = conv.FProp(theta.conv[i], x, pad)
      x = self.conv_bn[i].FProp(theta.conv_bn[i], x, pad.view(pad.shape[0], -1, 1, 1))
      x = torch.relu(x)
    b, t = x.shape[:2]
    xs = x.reshape(b, t, -1).transpose(0, 1)                 # [T,B,F·C]
    pad_t = pad.t().unsqueeze(-1)
    for i, rnn in enumerate(self.rnn):
      ys = rnn.FProp(theta.rnn[i], xs, pad_t)
      if p.project_lstm_output and i < len(self.proj):
        ys = self.proj[i].FProp(theta.proj[i], ys, pad_t)
      if p.residual_start > 0 and i + 1 >= p.residual_start and \
          (i + 1 - p.residual_start) % p.residual_stride == 0 and xs.shape == ys.shape:
        ys = ys + xs
      xs = ys
    if p.final_proj is not None:
      xs = self.final_proj.FProp(theta.final_proj, xs, pad_t)
    xs = xs * (1.0 - pad_t)
    return NestedMap(encoded=xs, padding=pad.t(), state=None)


class ConformerEncoder(base_layer.BaseLayer):
  """Conformer ASR encoder (Gulati et al. 2020), assembled from the reference building
  block `core/conformer_layer.py:471 ConformerLayer` (no registered reference model uses
  it — SURVEY §0.4): SpecAugment → 2× strided conv subsampling (time/4) → linear →
  dropout → N × ConformerLayer (½FFN, MHSA with relative positions, LConv, ½FFN, LN) →
  `[T', B, D]`, the interface of `AsrEncoder`.

  On a GPU the LConv module runs the fused GLU+mask+depthwise-conv kernel
  (`ops/csrc/conv_kernels.cu`), the norms the fused LN kernels, attention the fused
  attention path; `remat` recomputes each block in the backward pass.
  """

  @classmethod
  def Params(cls):
    from lingvo_b200.core import conformer_layer  # pylint: disable=g-import-not-at-top
    p = super().Params()
    p.Define('specaugment_network', spectrum_augmenter.SpectrumAugmenter.Params(),
             'SpecAugment params.')
    p.Define('use_specaugment', False, 'Apply SpecAugment in training.')
    p.Define('input_shape', [None, None, 80, 1], '[B, T, F, C].')
    p.Define('conv_filter_shapes', [(3, 3, 1, 32), (3, 3, 32, 32)], 'Subsampling convs.')
    p.Define('conv_filter_strides', [(2, 2), (2, 2)], 'Subsampling strides.')
    p.Define('cnn_tpl', conv_lib.Conv2DLayerWithPadding.Params(), 'Conv template.')
    p.Define('model_dim', 512, 'Encoder dim D.')
    p.Define('num_layers', 17, 'Conformer blocks.')
    p.Define('num_heads', 8, 'Attention heads.')
    p.Define('kernel_size', 32, 'Depthwise conv kernel.')
    p.Define('ff_hidden_dim', None, 'FFN hidden dim (default 4·D).')
    p.Define('atten_left_context', None, 'Attention left context (None: full).')
    p.Define('atten_right_context', None, 'Attention right context (None: full).')
    p.Define('use_relative_atten', True, 'Transformer-XL relative attention.')
    p.Define('dropout_prob', 0.1, 'Dropout.')
    p.Define('is_causal', False, 'Streaming (causal conv, limited right context).')
    p.Define('remat', False, 'Rematerialise every block in backward.')
    p.Define('conformer_tpl', conformer_layer.ConformerLayer.Params(), 'Block template.')
    p.Define('pad_steps', 0, 'Extra padded frames appended to the input.')
    return p

  def __init__(self, params):
    from lingvo_b200.core import conformer_layer  # pylint: disable=g-import-not-at-top
    super().__init__(params)
    p = self.params
    if p.use_specaugment:
      self.CreateChild('specaugment', p.specaugment_network)
    convs = []
    f, c = p.input_shape[2], p.input_shape[3]
    for i, (shape, stride) in enumerate(zip(p.conv_filter_shapes, p.conv_filter_strides)):
      convs.append(p.cnn_tpl.Copy().Set(name='conv_L%d' % i, filter_shape=tuple(shape),
                                        filter_stride=tuple(stride)))
      f = -(-f // stride[1])
      c = shape[3]
    self.CreateChildren('conv', convs)
    self.CreateChild('input_proj', layers.FCLayer.Params().Set(
        input_dim=f * c, output_dim=p.model_dim, activation='NONE'))
    self.CreateChild('input_dropout', layers.DropoutLayer.Params().Set(
        keep_prob=1.0 - p.dropout_prob))
    blocks = []
    for i in range(p.num_layers):
      bp = conformer_layer.ConformerLayer.CommonParams(
          input_dim=p.model_dim, atten_num_heads=p.num_heads,
          atten_left_context=p.atten_left_context, atten_right_context=p.atten_right_context,
          use_relative_atten=p.use_relative_atten, kernel_size=p.kernel_size,
          fflayer_hidden_dim=p.ff_hidden_dim or 4 * p.model_dim,
          dropout_prob=p.dropout_prob, is_causal=p.is_causal)
      bp.name = 'conformer_%d' % i
      bp.remat = p.remat
      blocks.append(bp)
    self.CreateChildren('blocks', blocks)

  @property
  def output_dim(self):
    return self.params.model_dim

  def FProp(self, theta, batch, state0=None):
    """batch.src_inputs [B,T,F,C], batch.paddings [B,T] → encoded [T',B,D]."""
    p = self.params
    x, pad = batch.src_inputs.float(), batch.paddings.float()
    if p.use_specaugment and not self.do_eval:
      x, pad = self.specaugment.FProp(theta.specaugment, x, pad)
    if p.pad_steps > 0:
      x = torch.nn.functional.pad(x, (0, 0, 0, 0, 0, p.pad_steps))
      pad = torch.nn.functional.pad(pad, (0, p.pad_steps), value=1.0)
    for i, conv in enumerate(self.conv):
      x, pad = conv.FProp(theta.conv[i], x, pad)
      x = torch.relu(x)
    b, t = x.shape[:2]
    h = self.input_proj.FProp(theta.input_proj, x.reshape(b, t, -1))
    h = self.input_dropout.FProp(theta.input_dropout, h)
    h = self._CastToFPropDtype(h)
    nm = NestedMap(features=h, paddings=pad)
    for i, blk in enumerate(self.blocks):
      nm = blk.FProp(theta.blocks[i], nm)
    out = nm.features * (1.0 - pad).unsqueeze(-1).to(nm.features.dtype)
    return NestedMap(encoded=out.transpose(0, 1), padding=pad.t(), state=None)
