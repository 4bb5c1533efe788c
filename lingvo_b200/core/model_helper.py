"""RNN-layer param factories (ref `lingvo/core/model_helper.py`).

Task encoders/decoders carry a string knob (`unidi_rnn_type` / `bidi_rnn_type`) instead of a
layer class so that experiment configs can switch the recurrent implementation; these
factories map the knob to layer params around the given cell params.
"""

from lingvo_b200.core import rnn_layers

_UNIDI = {
    'func': rnn_layers.FRNN,
    # quasi-RNN / SRU cells run through the same functional RNN: their input half is hoisted
    # out of the time loop by `cell.ProjectInput`, the loop only does the pooling recurrence
    'quasi_ifo': rnn_layers.FRNN,
    'sru': rnn_layers.FRNN,
}

_BIDI = {
    'func': rnn_layers.BidirectionalFRNN,
    'quasi_ifo': rnn_layers.BidirectionalFRNNQuasi,
    'sru': rnn_layers.BidirectionalFRNNQuasi,
}


def CreateUnidirectionalRNNParams(layer_params, cell_params):
  """Params of a uni-directional RNN layer chosen by `layer_params.unidi_rnn_type`."""
  assert hasattr(layer_params, 'unidi_rnn_type'), 'layer params must contain unidi_rnn_type'
  t = layer_params.unidi_rnn_type
  if t not in _UNIDI:
    raise ValueError('Invalid unidi_rnn_type: %s' % t)
  return _UNIDI[t].Params().Set(cell=cell_params)


def CreateBidirectionalRNNParams(layer_params, forward_cell_params, backward_cell_params):
  """Params of a bi-directional RNN layer chosen by `layer_params.bidi_rnn_type`."""
  assert hasattr(layer_params, 'bidi_rnn_type'), 'layer params must contain bidi_rnn_type'
  t = layer_params.bidi_rnn_type
  if t not in _BIDI:
    raise ValueError('Invalid bidi_rnn_type: %s' % t)
  return _BIDI[t].Params().Set(fwd=forward_cell_params, bak=backward_cell_params)
