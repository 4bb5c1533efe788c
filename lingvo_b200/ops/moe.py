"""Front-end of the fused sm_100a MoE kernels (csrc/moe_kernels.cu)."""

import torch

from lingvo_b200 import ops


def available() -> bool:
  mod = ops.native(required=False)
  return mod is not None and hasattr(mod, '_has_moe')
