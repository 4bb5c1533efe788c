"""Base class for `StreamStep` ≡ `FProp` equivalence tests (ref
`lingvo/core/stream_step_test_base.py`).

A streaming layer processes the sequence in chunks of `stride` frames with an explicit
state; running `StreamStep` over all chunks must reproduce `FProp` on the whole sequence
(possibly delayed by the layer's right context). Subclasses provide `_GetParams`,
`_FProp`, `_StreamStep`, `_GetFPropOutput`; `_TestStreamStepHelper` does the rest.
"""

from __future__ import annotations

import numpy as np
import torch

from lingvo_b200.core import py_utils
from lingvo_b200.core import test_utils


class StreamStepTestBase(test_utils.TestCase):

  @property
  def input_rank(self):
    """3 for [B, T, D] inputs, 4 for [B, T, F, C]."""
    return 3

  def _GetInputs(self, batch_size, max_seqlen, input_dim, full_seq=False):
    g = torch.Generator().manual_seed(123)
    if self.input_rank == 3:
      x = torch.randn(batch_size, max_seqlen, input_dim, generator=g)
    else:
      x = torch.randn(batch_size, max_seqlen, input_dim, 1, generator=g)
    if full_seq:
      lens = torch.full((batch_size,), max_seqlen)
    else:
      lens = torch.randint(max(1, max_seqlen // 2), max_seqlen + 1, (batch_size,), generator=g)
    pad = (torch.arange(max_seqlen).unsqueeze(0) >= lens.unsqueeze(1)).float()
    return x, pad

  def _PadInput(self, inputs, paddings, num_frames):
    """Appends `num_frames` padded frames (to flush a right-context delay)."""
    if num_frames <= 0:
      return inputs, paddings
    z = torch.zeros((inputs.shape[0], num_frames) + tuple(inputs.shape[2:]), dtype=inputs.dtype)
    return torch.cat([inputs, z], 1), torch.cat(
        [paddings, torch.ones(paddings.shape[0], num_frames)], 1)

  def _NormalizeStreamStepOutput(self, outputs, paddings, right_context, max_seqlen,
                                 num_layers=1):
    """Drops the first `right_context · num_layers` delayed frames and trims to max_seqlen."""
    d = right_context * num_layers
    return outputs[:, d:d + max_seqlen], paddings[:, d:d + max_seqlen]

  # -- hooks -------------------------------------------------------------------------
  def _GetParams(self, **kwargs):
    raise NotImplementedError()

  def _FProp(self, layer, inputs, paddings):
    return layer.FProp(layer.theta, inputs, paddings)

  def _StreamStep(self, layer, step_inputs, step_paddings, state):
    return layer.StreamStep(layer.theta, step_inputs, step_paddings, state)

  def _GetFPropOutput(self, fprop_out):
    """→ (outputs, paddings) from whatever FProp returns."""
    return fprop_out[0], fprop_out[1]

  # -- the test ----------------------------------------------------------------------
  def _TestStreamStepHelper(self, batch_size=2, max_seqlen=16, input_dim=8, stride=1,
                            right_context=0, tol=1e-5, **kwargs):
    p = self._GetParams(input_dim=input_dim, stride=stride, right_context=right_context,
                        **kwargs)
    with self.SetEval(True):
      layer = p.Instantiate()
    inputs, paddings = self._GetInputs(batch_size, max_seqlen, input_dim)
    with torch.no_grad():
      base_out, base_pad = self._GetFPropOutput(self._FProp(layer, inputs, paddings))
      s_in, s_pad = self._PadInput(inputs, paddings, right_context)
      state = layer.zero_state(batch_size)
      outs, pads = [], []
      for t in range(0, s_in.shape[1], stride):
        o, pd, state = self._StreamStep(layer, s_in[:, t:t + stride], s_pad[:, t:t + stride],
                                        state)
        outs.append(o)
        pads.append(pd)
      out, pad = self._NormalizeStreamStepOutput(torch.cat(outs, 1), torch.cat(pads, 1),
                                                 right_context, base_out.shape[1])
    mask = (1.0 - base_pad).reshape(base_pad.shape + (1,) * (base_out.dim() - 2))
    np.testing.assert_allclose((out * mask).numpy(), (base_out * mask).numpy(), rtol=tol,
                               atol=tol)
    np.testing.assert_array_equal(pad.numpy(), base_pad.numpy())
    return out

  def _TestRightContextStackingLayersHelper(self, num_layers=2, **kwargs):
    """Stacks `num_layers` identical layers; the streaming delay adds up."""
    batch_size, max_seqlen = kwargs.pop('batch_size', 2), kwargs.pop('max_seqlen', 16)
    input_dim, stride = kwargs.pop('input_dim', 8), kwargs.pop('stride', 1)
    right_context = kwargs.pop('right_context', 1)
    tol = kwargs.pop('tol', 1e-5)
    ps = [self._GetParams(input_dim=input_dim, stride=stride, right_context=right_context,
                          **kwargs).Set(name='l%d' % i) for i in range(num_layers)]
    with self.SetEval(True):
      layers = [p.Instantiate() for p in ps]
    inputs, paddings = self._GetInputs(batch_size, max_seqlen, input_dim, full_seq=True)
    with torch.no_grad():
      x, pd = inputs, paddings
      for l in layers:
        x, pd = self._GetFPropOutput(self._FProp(l, x, pd))
      base = x
      s_in, s_pad = self._PadInput(inputs, paddings, right_context * num_layers)
      states = [l.zero_state(batch_size) for l in layers]
      outs = []
      for t in range(0, s_in.shape[1], stride):
        o, opd = s_in[:, t:t + stride], s_pad[:, t:t + stride]
        for i, l in enumerate(layers):
          o, opd, states[i] = self._StreamStep(l, o, opd, states[i])
        outs.append(o)
      out, _ = self._NormalizeStreamStepOutput(torch.cat(outs, 1), s_pad, right_context,
                                               max_seqlen, num_layers)
    np.testing.assert_allclose(out.numpy(), base.numpy(), rtol=tol, atol=tol)

