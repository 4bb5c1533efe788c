"""Batch-size scaling between split / infeed / global (reference
`core/batch_utils.py:20-100`)."""

from lingvo_b200.core import cluster_factory


def scale_infeed_to_global(infeed_batch_size, use_per_host_infeed):
  """Global batch = infeed batch × number of input-producing processes."""
  cluster = cluster_factory.Current()
  if use_per_host_infeed and cluster.world_size > 1:
    return infeed_batch_size * cluster.world_size
  return infeed_batch_size


def scale_split_to_infeed(split_batch_size, use_per_host_infeed):
  """Infeed batch = per-split batch × splits handled by this process."""
  cluster = cluster_factory.Current()
  global_batch_size = split_batch_size * cluster.num_splits_per_client
  if cluster.job in ('trainer', 'trainer_client', 'executor_tpu') and (
      cluster.mode == 'sync'):
    return global_batch_size
  return split_batch_size * cluster.num_splits_per_client
