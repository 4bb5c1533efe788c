"""Temporary variable wrappers (ref `lingvo/core/var_tmp_wrappers.py`).

`VarWrapperTrackAssign` records assignments to a variable made inside a scope (used by
the reference to collect update ops inside while-loops); `StackedVarWrapperWithManualSharding`
presents one slice `[i]` of a stacked `[N, …]` variable as if it were the variable.
"""
import torch


class VarWrapperTrackAssign:

  def __init__(self, var):
    self._var = var
    self._previous = None
    self.tracked = []

  @property
  def raw_var(self):
    return self._var

  def assign(self, value):  # pylint: disable=invalid-name
    self.tracked.append(value)
    with torch.no_grad():
      self._var.data.copy_(value)
    return self._var

  def assign_add(self, delta):  # pylint: disable=invalid-name
    return self.assign(self._var.data + delta)

  def assign_sub(self, delta):  # pylint: disable=invalid-name
    return self.assign(self._var.data - delta)

  def __getattr__(self, name):
    return getattr(self._var, name)


class StackedVarWrapperWithManualSharding:

  def __init__(self, stacked_var, index=0):
    self._var = stacked_var
    self._index = index

  @property
  def value(self):
    return self._var[self._index]

  def assign(self, value):  # pylint: disable=invalid-name
    with torch.no_grad():
      self._var.data[self._index].copy_(value)
    return self.value

  @property
  def shape(self):
    return self._var.shape[1:]

  @property
  def dtype(self):
    return self._var.dtype
