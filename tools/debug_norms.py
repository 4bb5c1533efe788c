"""Prints which tensors still go through the generic SumSquared reduction in a train step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lingvo_b200 import model_registry
from lingvo_b200.core import cluster_factory, py_utils
import lingvo_b200.models.lm.params.synthetic_packed_input  # noqa

cfg = model_registry.GetParams('lm.synthetic_packed_input.MoELm8E', 'Train')
cfg.cluster.worker.gpus_per_replica = 1
dev = torch.device('cuda', 0)
with cluster_factory.Cluster(cfg.cluster):
  model = cfg.Instantiate(); model.to(dev); task = model.tasks[0]
  batch = task._MoveBatch(task.input.GetPreprocessedInputBatch(), dev)
  task.TrainStep([batch])
  names = {v.data_ptr(): v.var_name for v in task.vars.Flatten()}
  orig = py_utils.SumSquared
  def spy(ts):
    ts = list(ts)
    print('SumSquared over', len(ts), 'tensors:',
          [(names.get(t.data_ptr(), '?'), tuple(t.shape), str(t.dtype)) for t in ts][:60])
    return orig(ts)
  py_utils.SumSquared = spy
  task.TrainStep([batch])
  opt = task.learners[0].optimizer
  for v in task.vars.Flatten():
    dims = opt._FactoredDims(list(v.shape))
    if dims is not None and not opt._FusedEligible(v, dims):
      print('factored but not fused:', v.var_name, tuple(v.shape), dims)
