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
from lingvo_b200.core import model_helper
from lingvo_b200.core import py_utils
from lingvo_b200.core import rnn_cell
from lingvo_b200.core import rnn_layers
from lingvo_b200.core import spectrum_augmenter
from lingvo_b200.core import summary_utils
from lingvo_b200.core.nested_map import NestedMap


class AsrEncoder(base_layer.BaseLayer):
  """The LAS listener (ref :32): SpecAugment → strided conv stack → bidirectional conv-LSTM
  blocks → bidirectional LSTM stack with projections, (highway) residuals and optional frame
  stacking. Returns NestedMap(encoded `[T', B, D]`, padding `[T', B]`, state) plus, with
  `extra_per_layer_outputs`, one `conv_i` / `conv_lstm_i` / `rnn_i` entry per layer."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('specaugment_network', spectrum_augmenter.SpectrumAugmenter.Params(),
             'SpecAugment params.')
    p.Define('use_specaugment', False, 'Apply SpecAugment in training.')
    p.Define('lstm_tpl', rnn_cell.LSTMCellSimple.Params(), 'LSTM cell template.')
    p.Define('cnn_tpl', conv_lib.Conv2DLayerWithPadding.Params(), 'Conv template.')
    p.Define('proj_tpl', layers.ProjectionLayer.Params(), 'Projection template.')
    p.Define('highway_skip', False,
             'Residual connections are gated (HighwaySkipLayer); needs residual_start.')
    p.Define('highway_skip_tpl', layers.HighwaySkipLayer.Params(), 'Highway skip template.')
    p.Define('conv_lstm_tpl', rnn_cell.ConvLSTMCell.Params(), 'Conv-LSTM cell template.')
    p.Define('after_conv_lstm_cnn_tpl', conv_lib.Conv2DLayerWithPadding.Params(),
             'Conv merging the two conv-LSTM directions back to the channel count.')
    p.Define('conv_filter_shapes', [(3, 3, 1, 32), (3, 3, 32, 32)], 'Conv filters.')
    p.Define('conv_filter_strides', [(2, 2), (2, 2)], 'Conv strides.')
    p.Define('input_shape', [None, None, 80, 1], '[B, T, F, C].')
    p.Define('lstm_cell_size', 256, 'LSTM cell size per direction.')
    p.Define('num_cnn_layers', 2, 'Conv layers.')
    p.Define('num_conv_lstm_layers', 0, 'Bidirectional conv-LSTM blocks.')
    p.Define('num_lstm_layers', 3, 'Bi-LSTM layers.')
    p.Define('project_after_last_lstm', False, 'Also project the last LSTM layer.')
    p.Define('project_lstm_output', True, 'Projection after every LSTM layer (but the last).')
    p.Define('pad_steps', 6, 'Extra padded frames appended to the input.')
    p.Define('residual_start', 0, 'First LSTM layer with a residual (0: none).')
    p.Define('residual_stride', 1, 'LSTM layers spanned by one residual connection.')
    p.Define('bidi_rnn_type', 'func', 'func: BidirectionalFRNN.')
    p.Define('extra_per_layer_outputs', False, 'Also return every layer\'s output.')
    p.Define('stacking_layer_tpl', layers.StackingOverTime.Params(), 'Frame stacking.')
    p.Define('layer_index_before_stacking', -1,
             'LSTM layer after which frames are stacked (< 0: no stacking).')
    p.Define('final_proj', None, 'Optional final projection.')
    p.Define('pad_first_lstm_input_to_multiple', 16,
             'The flattened conv output is zero-padded to a multiple of this (GEMM-friendly K).')
    p.lstm_tpl.params_init = py_utils.WeightInit.Uniform(0.1)
    p.conv_lstm_tpl.filter_shape = [1, 3]          # (time, frequency)
    p.after_conv_lstm_cnn_tpl.filter_stride = (1, 1)
    p.proj_tpl.batch_norm = True
    p.proj_tpl.activation = 'RELU'
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
    conv_output_shape = [None, None, f, c]

    clstm, clstm_cnn, clstm_bn = [], [], []
    for i in range(p.num_conv_lstm_layers):
      fwd = p.conv_lstm_tpl.Copy().Set(name='f_conv_lstm_%d' % i,
                                        inputs_shape=[None, 1, f, c],
                                        cell_shape=[None, 1, f, c])
      rnn_p = self.CreateConvLstmLayerParams().Set(name='conv_lstm_rnn_%d' % i, fwd=fwd,
                                                   bak=fwd.Copy().Set(
                                                       name='b_conv_lstm_%d' % i))
      clstm.append(rnn_p)
      kh, kw = 3, 3
      clstm_cnn.append(p.after_conv_lstm_cnn_tpl.Copy().Set(
          name='conv_lstm_cnn_%d' % i, filter_shape=(kh, kw, 2 * c, c),
          filter_stride=(1, 1)))
      clstm_bn.append(bn_layers.BatchNormLayer.Params().Set(name='conv_lstm_bn_%d' % i, dim=c))
    self.CreateChildren('conv_lstm_rnn', clstm)
    self.CreateChildren('conv_lstm_cnn', clstm_cnn)
    self.CreateChildren('conv_lstm_bn', clstm_bn)

    self._first_lstm_input_dim, self._first_lstm_input_dim_pad = \
        self.FirstLstmLayerInputDimAndPadding(conv_output_shape,
                                              p.pad_first_lstm_input_to_multiple)
    self._conv_out_dim = f * c
    rnns, projs, skips = [], [], []
    odim = self._first_lstm_input_dim
    num_proj = p.num_lstm_layers if p.project_after_last_lstm else p.num_lstm_layers - 1
    for i in range(p.num_lstm_layers):
      fwd = p.lstm_tpl.Copy().Set(name='fwd_rnn_L%d' % i, num_input_nodes=odim,
                                  num_output_nodes=p.lstm_cell_size)
      bak = fwd.Copy().Set(name='bak_rnn_L%d' % i)
      rnns.append(self.CreateBidirectionalRNNParams(fwd, bak).Set(name='brnn_L%d' % i))
      odim = 2 * p.lstm_cell_size
      if p.project_lstm_output and i < num_proj:
        projs.append(p.proj_tpl.Copy().Set(name='proj_L%d' % i, input_dim=odim,
                                           output_dim=odim))
      residual_index = i - p.residual_start + 1
      if p.residual_start > 0 and residual_index >= 0 and p.highway_skip:
        skips.append(p.highway_skip_tpl.Copy().Set(name='enc_hwskip_%d' % len(skips),
                                                   input_dim=odim))
      if p.layer_index_before_stacking == i:
        self.CreateChild('stacking', p.stacking_layer_tpl.Copy().Set(name='stacking_%d' % i))
        odim *= p.stacking_layer_tpl.left_context + 1 + p.stacking_layer_tpl.right_context
    self.CreateChildren('rnn', rnns)
    self.CreateChildren('proj', projs)
    self.CreateChildren('highway_skip', skips)
    self._out_dim = odim
    if p.final_proj is not None:
      fp = p.final_proj.Copy()
      if not fp.input_dim:
        fp.input_dim = odim
      self.CreateChild('final_proj', fp)
      self._out_dim = fp.output_dim or odim

  # -- construction hooks (subclasses override to change the recurrent flavour) ----------
  def CreateBidirectionalRNNParams(self, forward_p, backward_p):
    return model_helper.CreateBidirectionalRNNParams(self.params, forward_p, backward_p)

  def CreateConvLstmLayerParams(self):
    return rnn_layers.BidirectionalFRNN.Params()

  def FirstLstmLayerInputDimAndPadding(self, conv_output_shape, pad_to_multiple=16):
    """(padded, padding) width of the first LSTM's input = frequency × channels."""
    unpadded = conv_output_shape[2] * conv_output_shape[3]
    if pad_to_multiple and unpadded % pad_to_multiple:
      padded = -(-unpadded // pad_to_multiple) * pad_to_multiple
    else:
      padded = unpadded
    return padded, padded - unpadded

  @property
  def input_shape(self):
    return self.params.input_shape

  @property
  def output_dim(self):
    return self._out_dim

  @property
  def supports_streaming(self):
    return False

  def zero_state(self, theta, batch_size):
    return NestedMap()

  def FProp(self, theta, batch, state0=None):
    """batch.src_inputs [B,T,F,C], batch.paddings [B,T] → encoded [T',B,D]."""
    p = self.params
    x, pad = batch.src_inputs.float(), batch.paddings.float()
    outputs = NestedMap()
    if p.use_specaugment and not self.do_eval:
      x, pad = self.specaugment.FProp(theta.specaugment, x, pad)
    if p.pad_steps > 0:
      x = torch.nn.functional.pad(x, (0, 0, 0, 0, 0, p.pad_steps))
      pad = torch.nn.functional.pad(pad, (0, p.pad_steps), value=1.0)
    plot = summary_utils._ShouldAddSummary()   # pylint: disable=protected-access
    plots = []
    if plot:
      plots.append(summary_utils.PrepareSequenceForPlot(x.transpose(2, 3), pad, 'inputs'))

    def Masked(t, tp):
      return t * (1.0 - tp).reshape(tp.shape[0], tp.shape[1], 1, 1)

    for i, conv in enumerate(self.conv):
      x, pad = conv.FProp(theta.conv[i], x, pad)
      x = self.conv_bn[i].FProp(theta.conv_bn[i], x, pad.view(pad.shape[0], -1, 1, 1))
      x = torch.relu(x)
      if p.extra_per_layer_outputs:
        outputs['conv_%d' % i] = NestedMap(encoded=Masked(x, pad).transpose(0, 1),
                                           padding=pad.t())
      if plot:
        plots.append(summary_utils.PrepareSequenceForPlot(x.transpose(2, 3), pad,
                                                          'conv_%d_out' % i))

    for i, (rnn, cnn) in enumerate(zip(self.conv_lstm_rnn, self.conv_lstm_cnn)):
      # time-major `[T, B, 1, F, C]` through the bidirectional conv-LSTM, then a conv
      # brings the concatenated directions back to C channels
      rin = x.transpose(0, 1).unsqueeze(2)
      rout = rnn.FProp(theta.conv_lstm_rnn[i], rin, pad.t().unsqueeze(-1))
      cin = rout.squeeze(2).transpose(0, 1)
      x, pad = cnn.FProp(theta.conv_lstm_cnn[i], cin, pad)
      x = self.conv_lstm_bn[i].FProp(theta.conv_lstm_bn[i], x,
                                     pad.view(pad.shape[0], -1, 1, 1))
      x = torch.relu(x)
      if p.extra_per_layer_outputs:
        outputs['conv_lstm_%d' % i] = NestedMap(encoded=Masked(x, pad).transpose(0, 1),
                                                padding=pad.t())
      if plot:
        plots.append(summary_utils.PrepareSequenceForPlot(x, pad, 'conv_lstm_%d_out' % i))

    b, t = x.shape[:2]
    flat = x.reshape(b, t, -1)
    if self._first_lstm_input_dim_pad:
      flat = torch.nn.functional.pad(flat, (0, self._first_lstm_input_dim_pad))
    xs = flat.transpose(0, 1)                                # [T, B, F·C (padded)]
    pad_t = pad.t().unsqueeze(-1)
    num_proj = p.num_lstm_layers if p.project_after_last_lstm else p.num_lstm_layers - 1
    num_skips = 0
    residual_in = None
    for i, rnn in enumerate(self.rnn):
      ys = rnn.FProp(theta.rnn[i], xs, pad_t)
      residual_index = i - p.residual_start + 1
      if p.residual_start > 0 and residual_index >= 0:
        if residual_index % p.residual_stride == 0:
          residual_in = xs
        if residual_index % p.residual_stride == p.residual_stride - 1:
          if p.highway_skip:
            ys = self.highway_skip[num_skips].FProp(theta.highway_skip[num_skips],
                                                    residual_in, ys)
            num_skips += 1
          else:
            assert residual_in.shape == ys.shape, (
                'residual from a layer of another width: %s vs %s' %
                (tuple(residual_in.shape), tuple(ys.shape)))
            ys = ys + residual_in
      if p.project_lstm_output and i < num_proj:
        ys = self.proj[i].FProp(theta.proj[i], ys, pad_t)
      if i == p.num_lstm_layers - 1:
        ys = ys * (1.0 - pad_t)
      if p.extra_per_layer_outputs:
        ys = ys * (1.0 - pad_t)
        outputs['rnn_%d' % i] = NestedMap(encoded=ys, padding=pad_t.squeeze(2))
      if p.layer_index_before_stacking == i:
        st, st_pad = self.stacking.FProp(ys.transpose(0, 1), pad_t.transpose(0, 1))
        ys, pad_t = st.transpose(0, 1), st_pad.transpose(0, 1)
      if plot:
        plots.append(summary_utils.PrepareSequenceForPlot(
            ys.transpose(0, 1), pad_t.squeeze(2).t(), 'rnn_%d_out' % i))
      xs = ys
    if p.final_proj is not None:
      xs = self.final_proj.FProp(theta.final_proj, xs, pad_t)
      xs = xs * (1.0 - pad_t)
    if plot:
      summary_utils.PlotSequenceFeatures(list(reversed(plots)), 'encoder_example',
                                         xlabel='Time')
    outputs.encoded = xs
    outputs.padding = pad_t.squeeze(2)
    outputs.state = NestedMap()
    return outputs


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
