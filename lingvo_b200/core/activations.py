"""Activation registry (reference `core/activations.py:1-163`).

Names: RELU, RELU6, LEAKY_RELU, SIGMOID, TANH, GELU, GELU_APPROXIMATE,
GELU_RAW, SWISH, SILU, SOFTPLUS, SQUARED_RELU, EXP, NONE, plus the
`ActivationLayer` wrapper.
"""

import math

import torch
import torch.nn.functional as F

from lingvo_b200.core import base_layer

_ACTIVATIONS = {
    'RELU': F.relu,
    'RELU6': F.relu6,
    'LEAKY_RELU': lambda x: F.leaky_relu(x, 0.2),
    'SIGMOID': torch.sigmoid,
    'TANH': torch.tanh,
    'GELU': F.gelu,
    'GELU_APPROXIMATE': lambda x: F.gelu(x, approximate='tanh'),
    'GELU_RAW': lambda x: 0.5 * x * (1 + torch.tanh(
        math.sqrt(2 / math.pi) * (x + 0.044715 * torch.pow(x, 3)))),
    'SWISH': F.silu,
    'SILU': F.silu,
    'SOFTPLUS': F.softplus,
    'SQUARED_RELU': lambda x: torch.square(F.relu(x)),
    'EXP': torch.exp,
    'NONE': lambda x: x,
}

_FLOPS_PER_ELEMENT = {
    'NONE': 0, 'RELU': 1, 'RELU6': 1, 'LEAKY_RELU': 2, 'SIGMOID': 4,
    'TANH': 6, 'GELU': 15, 'GELU_APPROXIMATE': 15, 'GELU_RAW': 15,
    'SWISH': 4, 'SILU': 4, 'SOFTPLUS': 11, 'SQUARED_RELU': 2, 'EXP': 3,
}


def GLUVariants(x, activation_name):
  """`x1 * act(x2)` for the two halves of the last dim (https://arxiv.org/abs/2002.05202)."""
  x1, x2 = x.chunk(2, dim=-1)
  return x1 * _ACTIVATIONS[activation_name](x2)


_ACTIVATIONS.update({
    'GLU': lambda x: GLUVariants(x, 'SIGMOID'),
    'BILINEAR_GLU': lambda x: GLUVariants(x, 'NONE'),
    'RELU_GLU': lambda x: GLUVariants(x, 'RELU'),
    'GELU_GLU': lambda x: GLUVariants(x, 'GELU'),
    'SWISH_GLU': lambda x: GLUVariants(x, 'SWISH'),
})
_FLOPS_PER_ELEMENT.update({'GLU': 5, 'BILINEAR_GLU': 1, 'RELU_GLU': 2, 'GELU_GLU': 16,
                           'SWISH_GLU': 5})


def GetFn(activation_name):
  return _ACTIVATIONS[activation_name]


def GetFlops(activation_name):
  return _FLOPS_PER_ELEMENT[activation_name]


def IsSupported(activation_name):
  return activation_name in _ACTIVATIONS


def DimMultiplier(activation_name):
  """GLU variants consume 2 × the output dim."""
  return 2 if (activation_name.startswith('GATED_') or activation_name.endswith('GLU')) else 1


class ActivationLayer(base_layer.BaseLayer):
  """Layer wrapper around the activation registry."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('activation', 'RELU', 'Activation function name.')
    return p

  def FProp(self, theta, inputs, paddings=None):
    out = GetFn(self.params.activation)(inputs)
    return out if paddings is None else (out, paddings)

  @classmethod
  def FPropMeta(cls, p, inputs, *args):
    from lingvo_b200.core.nested_map import NestedMap
    return NestedMap(flops=inputs.num_elements() * GetFlops(p.activation),
                     out_shapes=(inputs,) + tuple(args))
