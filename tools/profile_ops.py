"""Which Python call sites the residual elementwise/copy kernels of a train step come from."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from lingvo_b200 import model_registry
from lingvo_b200.core import cluster_factory
import lingvo_b200.models.lm.params.synthetic_packed_input  # noqa


def main():
  cfg = model_registry.GetParams('lm.synthetic_packed_input.MoELm8E', 'Train')
  cfg.cluster.worker.gpus_per_replica = 1
  dev = torch.device('cuda', 0)
  with cluster_factory.Cluster(cfg.cluster):
    model = cfg.Instantiate(); model.to(dev); task = model.tasks[0]
    batches = [task._MoveBatch(task.input.GetPreprocessedInputBatch(), dev) for _ in range(2)]
    for i in range(3):
      task.TrainStep([batches[i % 2]])
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True,
                 record_shapes=True) as prof:
      task.TrainStep([batches[0]])
      torch.cuda.synchronize()
  ka = prof.key_averages(group_by_stack_n=6, group_by_input_shape=True)
  rows = [e for e in ka if e.key.startswith('aten::') and e.self_device_time_total > 150]
  rows.sort(key=lambda e: -e.self_device_time_total)
  out = []
  for e in rows[:45]:
    stack = [s for s in e.stack if 'lingvo_b200' in s or 'bench' in s][:3]
    out.append('%8.0f us  x%-3d %-28s %s\n      %s' % (
        e.self_device_time_total, e.count, e.key, str(e.input_shapes)[:90],
        '\n      '.join(s[-110:] for s in stack)))
  txt = '\n'.join(out)
  os.makedirs('gpurun_out', exist_ok=True)
  open('gpurun_out/profile_ops.txt', 'w').write(txt)
  print(txt[-7000:])


if __name__ == '__main__':
  main()
