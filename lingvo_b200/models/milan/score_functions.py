"""Score functions (ref `lingvo/tasks/milan/score_functions.py`)."""

import torch

from lingvo_b200.core import base_layer


class DotProductScoreFunction(base_layer.BaseLayer):
  """score[..i, ..j] = <x_i, y_j>: `[x_batch…, D]` × `[y_batch…, D]` → `[x_batch…, y_batch…]`."""

  def FProp(self, theta, x_batch, y_batch):
    xs, ys = x_batch.shape[:-1], y_batch.shape[:-1]
    s = torch.matmul(x_batch.reshape(-1, x_batch.shape[-1]),
                     y_batch.reshape(-1, y_batch.shape[-1]).t())
    return s.reshape(*xs, *ys)
