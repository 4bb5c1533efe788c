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


def scale_global_to_infeed(global_batch_size, use_per_host_infeed):
  """Infeed batch of one input-producing process (ref :42)."""
  cluster = cluster_factory.Current()
  if use_per_host_infeed and cluster.world_size > 1:
    return global_batch_size // cluster.world_size
  return global_batch_size


def scale_global_to_worker(global_batch_size):
  """Per-device batch; the global batch must divide evenly (ref :84)."""
  cluster = cluster_factory.Current()
  n = max(int(cluster.total_worker_devices), 1)
  q, r = divmod(global_batch_size, n)
  if r:
    raise ValueError('global_batch_size %d did not divide evenly by %d workers.' %
                     (global_batch_size, n))
  return q
