"""Per-kernel breakdown of one training step (torch.profiler, not for timing)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from lingvo_b200 import model_registry
from lingvo_b200.core import cluster_factory
import lingvo_b200.models.lm.params.synthetic_packed_input  # noqa

def main():
  name = sys.argv[1] if len(sys.argv) > 1 else 'lm.synthetic_packed_input.MoELm8E'
  cfg = model_registry.GetParams(name, 'Train')
  cfg.cluster.worker.gpus_per_replica = 1
  dev = torch.device('cuda', 0)
  with cluster_factory.Cluster(cfg.cluster):
    model = cfg.Instantiate(); model.to(dev); task = model.tasks[0]
    batches = [task._MoveBatch(task.input.GetPreprocessedInputBatch(), dev) for _ in range(2)]
    for i in range(3):
      task.TrainStep([batches[i % 2]])
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
      for i in range(2):
        task.TrainStep([batches[i % 2]])
      torch.cuda.synchronize()
  os.makedirs('gpurun_out', exist_ok=True)
  tbl = prof.key_averages().table(sort_by='cuda_time_total', row_limit=70, max_name_column_width=70)
  open('gpurun_out/profile_step.txt', 'w').write(tbl)
  print(tbl[-9000:])

if __name__ == '__main__':
  main()
