"""Shared fixtures for quantization tests (ref `lingvo/core/quant_test_lib.py`).

`SampleQuantizedProjectionLayer` exercises every `QuantizableLayer` hook on a small
`x·W → tanh/sigmoid/relu` pipeline; `QuantUtilsBaseTest._testLayerHelper` runs it
through train / eval and checks the outputs against expectations.
"""

from __future__ import annotations

import numpy as np
import torch

from lingvo_b200.core import py_utils
from lingvo_b200.core import quant_utils
from lingvo_b200.core import test_utils


class SampleQuantizedProjectionLayer(quant_utils.QuantizableLayer):
  """ref :28."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('input_dim', 2, 'Depth of the input.')
    p.Define('output_dim', 3, 'Depth of the output.')
    p.Define('activation', 'TANH', 'TANH | SIGMOID | RELU | NONE.')
    return p

  def __init__(self, params):
    super().__init__(params)
    self.TrackQActs('inputs', 'transformed')

  def _CreateLayerVariables(self):
    super()._CreateLayerVariables()
    p = self.params
    self.CreateVariable('w', py_utils.WeightParams(
        [p.input_dim, p.output_dim], py_utils.WeightInit.Gaussian(1.0, seed=1), p.dtype))

  def FProp(self, theta, inputs, paddings):
    p = self.params
    w = self.QWeight(theta.w)
    inputs = self.QAct('inputs', inputs)
    out = self.QMatmul(inputs.reshape(-1, p.input_dim), w)
    out = self.QAct('transformed', out)
    if p.activation == 'TANH':
      out = self.QRTanh(out)
    elif p.activation == 'SIGMOID':
      out = self.QRSigmoid(out)
    elif p.activation == 'RELU':
      out = self.QRRelu(out)
    out = out.reshape(list(inputs.shape[:-1]) + [p.output_dim])
    return out * (1.0 - paddings)


class QuantUtilsBaseTest(test_utils.TestCase):
  """ref :114."""

  # [batch=1... ] fixed inputs shared by every quantization-domain test
  INPUTS = np.array([[[-1.0, 0.5], [0.25, 2.0]], [[0.0, -3.0], [1.5, 0.125]]], np.float32)
  PADDINGS = np.array([[[0.0], [1.0]], [[0.0], [0.0]]], np.float32)

  def _testLayerHelper(self, test_case, p, expected=None, not_expected=None,  # pylint: disable=invalid-name
                       global_step=None, tol=1e-5):
    """Builds the layer from `p`, runs FProp on the fixed inputs and compares."""
    del test_case
    if global_step is not None:
      py_utils.SetGlobalStep(int(global_step))
    layer = p.Instantiate()
    with torch.no_grad():
      out = layer.FProp(layer.theta, torch.from_numpy(self.INPUTS),
                        torch.from_numpy(self.PADDINGS))
    out = out.numpy()
    self.assertEqual(out.shape, (2, 2, p.output_dim))
    self.assertTrue(np.all(out[0, 1] == 0.0))                 # padded frame
    if expected is not None:
      np.testing.assert_allclose(out, np.asarray(expected), rtol=tol, atol=tol)
    if not_expected is not None:
      self.assertFalse(np.allclose(out, np.asarray(not_expected), rtol=tol, atol=tol))
    return out
