"""Multi-task models with shared sub-networks (ref `lingvo/core/multitask_model.py`)."""

import re

from lingvo_b200.core import base_model


def _Share(tasks, get_child, attr):
  """Makes every task's `attr` child the first task's object (weights shared)."""
  first = None
  for t in tasks:
    child = get_child(t)
    if child is None:
      continue
    if first is None:
      first = child
    else:
      t.children[attr] = first
      # theta / vars views follow the child object
      if hasattr(t, '_private_children'):
        t._private_children[attr] = first  # pylint: disable=protected-access


class SharedEncoderModel(base_model.MultiTaskModel):
  """All tasks share one encoder (ref :21)."""

  def __init__(self, params):
    super().__init__(params)
    _Share(self.tasks, lambda t: t.children.get('enc', t.children.get('encoder')),
           'enc' if 'enc' in self.tasks[0].children else 'encoder')


class SharedEncoderDecoderModel(base_model.MultiTaskModel):
  """Tasks share encoder and decoder (ref :45)."""

  def __init__(self, params):
    super().__init__(params)
    for attr in ('enc', 'encoder', 'dec', 'decoder'):
      if attr in self.tasks[0].children:
        _Share(self.tasks, lambda t, a=attr: t.children.get(a), attr)


class RegExSharedVariableModel(base_model.MultiTaskModel):
  """Variables whose names match a rule are shared across tasks (ref :80).

  `variable_renaming_rules`: list of (regex, replacement); two variables that map to
  the same renamed name share storage (the first one created wins)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('variable_renaming_rules', None, 'List of (regex, format string).')
    return p

  def _CreateChildrenVariables(self):
    # Variables are created lazily (after __init__), so sharing is applied right after the
    # children's variables exist.
    super()._CreateChildrenVariables()
    rules = [(re.compile(r), fmt) for r, fmt in (self.params.variable_renaming_rules or [])]
    canon = {}
    for task in self.tasks:
      for _, layer in task.Walk():
        for key, var in list(layer._private_vars.items()):  # pylint: disable=protected-access
          name = var.var_name
          for rx, fmt in rules:
            m = rx.match(name)
            if m:
              name = fmt % m.groups()
              break
          if name in canon and canon[name] is not var and canon[name].shape == var.shape:
            layer._private_vars[key] = canon[name]  # pylint: disable=protected-access
          else:
            canon.setdefault(name, var)
