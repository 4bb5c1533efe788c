"""Helpers to obtain / build clusters (reference `core/cluster_factory.py`)."""

from lingvo_b200.core import cluster as cluster_lib

Cluster = cluster_lib._Cluster  # pylint: disable=protected-access
_DEFAULT = [None]


def SetCluster(cls):
  """Swaps the cluster implementation (`cls.Params().Instantiate()` builds clusters from
  now on) (ref :23)."""
  global Cluster  # pylint: disable=invalid-name,global-statement
  Cluster = cls
  _DEFAULT[0] = None


def Current():
  """The innermost active cluster, or a default single-process one."""
  top = cluster_lib._Cluster._TopOrNone()  # pylint: disable=protected-access
  if top is not None:
    return top
  if _DEFAULT[0] is None:
    _DEFAULT[0] = Cluster(Cluster.Params())
  return _DEFAULT[0]


def ForTestingWorker(mode=None, job=None, gpus=None, split_size=None, tpus=None,
                     add_summary=None, cpus=None, do_eval=None,
                     num_tpu_hosts=None):
  """Fake topology for tests (reference :60-112)."""
  p = Current().params.Copy()
  if mode is not None:
    p.mode = mode
  if job is not None:
    p.job = job
  if do_eval is not None:
    p.do_eval = do_eval
  if gpus is not None:
    p.worker.gpus_per_replica = gpus
    p.worker.cpus_per_replica = 0
    p.worker.tpus_per_replica = 0
    if split_size is not None:
      p.worker.devices_per_split = split_size
  elif cpus is not None:
    p.worker.cpus_per_replica = cpus
    p.worker.gpus_per_replica = 0
    if split_size is not None:
      p.worker.devices_per_split = split_size
  elif split_size is not None:
    p.worker.devices_per_split = split_size
  if add_summary is not None:
    p.add_summary = add_summary
  return p.Instantiate()


def SetEval(mode=True):
  p = Current().params.Copy()
  p.do_eval = mode
  return p.Instantiate()


def SetImmediatelyInstantiateVariables(mode=True):
  p = Current().params.Copy()
  p.immediately_instantiate_variables = mode
  return p.Instantiate()


def SetModelSplit(split_id):
  p = Current().params.Copy()
  p.split_id = split_id
  return p.Instantiate()


def SetRequireSequentialInputOrder(mode=True):
  p = Current().params.Copy()
  p.require_sequential_input_order = mode
  return p.Instantiate()
