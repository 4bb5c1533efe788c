"""Checks the fused global-norm path (Σg² from the Adafactor stats pass + multi-tensor small
vars, carried Σw²) against plain torch reductions over the same gradients / variables."""
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
  calls = []
  orig = py_utils.SumSquared
  def spy(ts):
    ts = list(ts); calls.append(len(ts)); return orig(ts)
  py_utils.SumSquared = spy
  for step in range(3):
    calls.clear()
    w_before = (sum(float(v.detach().double().square().sum()) for v in task.vars.Flatten())) ** 0.5
    m, _ = task.TrainStep([batch])
    vg = task._last_var_grads
    leaves = [x for x in vg.Flatten() if isinstance(x, py_utils.VarGrad)]
    ref_g = (sum(float(x.grad.double().square().sum()) for x in leaves)) ** 0.5
    em = task._eval_metrics
    def get(k):
      for kk, v in em.items():
        if kk.startswith(k):
          return float(v[0])
      return float('nan')
    print('step %d: SumSquared calls %s | grad_norm fused %.6f ref %.6f | var_norm fused %.6f ref(before step) %.6f | loss %.5f' % (
        step, calls, get('grad_norm/all'), ref_g, get('var_norm/all'), w_before, float(m['loss'][0])))
