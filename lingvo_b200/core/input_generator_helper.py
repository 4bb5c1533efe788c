"""Batch splitting helpers (ref `lingvo/core/input_generator_helper.py`)."""
import torch

from lingvo_b200.core.nested_map import NestedMap


def ComputeSplits(batch_size, num_splits):
  """Sizes of `num_splits` nearly equal parts of `batch_size`."""
  base, rem = divmod(int(batch_size), int(num_splits))
  return [base + (1 if i < rem else 0) for i in range(num_splits)]


def SplitTensors(xs, num_splits):
  """Splits every tensor in `xs` along dim 0 into `num_splits` parts."""
  sizes = ComputeSplits(xs[0].shape[0], num_splits)
  parts = [torch.split(x, sizes, 0) for x in xs]
  return [[p[i] for p in parts] for i in range(num_splits)]


def SplitDictOfTensors(t_dict, num_splits):
  keys = sorted(t_dict)
  splits = SplitTensors([t_dict[k] for k in keys], num_splits)
  return [NestedMap(dict(zip(keys, s))) for s in splits]
