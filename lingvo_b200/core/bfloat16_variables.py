"""bf16 checkpoint values (ref `lingvo/core/bfloat16_variables.py`): store selected
variables as bfloat16 in checkpoints and widen them back on restore."""
import re

import torch


def ToBfloat16State(state_dict, var_name_regex='.*'):
  pat = re.compile(var_name_regex)
  return {k: (v.to(torch.bfloat16) if isinstance(v, torch.Tensor) and v.is_floating_point()
              and pat.match(k) else v) for k, v in state_dict.items()}


def FromBfloat16State(state_dict, dtype=torch.float32):
  return {k: (v.to(dtype) if isinstance(v, torch.Tensor) and v.dtype == torch.bfloat16 else v)
          for k, v in state_dict.items()}


def get_saver_spec_for_variables_with_bf16_overrides(variables_to_restore):  # pylint: disable=invalid-name
  """Name → loader that widens bf16 checkpoint tensors to the variable's dtype."""
  return {name: (lambda t, v=var: t.to(v.dtype)) for name, var in variables_to_restore.items()}
