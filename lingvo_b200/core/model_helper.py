"""RNN-layer param factories (ref `lingvo/core/model_helper.py`)."""
from lingvo_b200.core import rnn_layers


def CreateUnidirectionalRNNParams(layer_params, cell_params):
  t = layer_params.unidi_rnn_type
  if t not in ('func', 'quasi_ifo', 'sru'):
    raise ValueError('Invalid unidi_rnn_type: %s' % t)
  return rnn_layers.FRNN.Params().Set(cell=cell_params)


def CreateBidirectionalRNNParams(layer_params, forward_cell_params, backward_cell_params):
  t = layer_params.bidi_rnn_type
  if t == 'func':
    cls = rnn_layers.BidirectionalFRNN
  elif t == 'native_cudnn':
    cls = rnn_layers.BidirectionalFRNN
  elif t == 'quasi_ifo':
    cls = rnn_layers.BidirectionalFRNNQuasi
  else:
    raise ValueError('Invalid bidi_rnn_type: %s' % t)
  return cls.Params().Set(fwd=forward_cell_params, bak=backward_cell_params)
