"""Per-kernel breakdown of the *product* train step at N ≥ 1 GPUs (torch.profiler on rank 0
around CUDA-graph replays; a diagnostic, never a bench number).

  python tools/profile_step_mgpu.py                     # N = 1
  torchrun --nproc-per-node N tools/profile_step_mgpu.py

Writes gpurun_out/profile_step_n<N>.txt (kernel table, sorted by total CUDA time) and
gpurun_out/profile_step_n<N>.json (per-kernel {calls, total_us, avg_us} + step ms).
"""
import json
import os
import sys
import tempfile
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

import bench


def main():
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  lr = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(lr)
  dev = torch.device('cuda', lr)
  if world > 1:
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=dev)
  from lingvo_b200.parallel import mesh as mesh_lib
  mesh_lib.Reset()
  args = types.SimpleNamespace(model=os.environ.get('LB_MODEL', bench.MODEL),
                               cuda_graph=os.environ.get('LB_GRAPH', 'auto'))
  runner = bench._CreateTrainer(args, tempfile.mkdtemp())   # pylint: disable=protected-access
  task = runner.task
  steps = 4
  with runner._cluster:   # pylint: disable=protected-access
    engine = runner.engine
    batches = [task._MoveBatch(task.input.GetPreprocessedInputBatch(), dev) for _ in range(4)]  # pylint: disable=protected-access
    for i in range(6):
      engine.Step(batches[i % 4])
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
      engine.Step(batches[i % 4])
    e1.record()
    torch.cuda.synchronize()
    ms_plain = e0.elapsed_time(e1) / steps
    if world > 1:
      dist.barrier()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
      for i in range(steps):
        engine.Step(batches[i % 4])
      torch.cuda.synchronize()
  if rank == 0:
    os.makedirs('gpurun_out', exist_ok=True)
    ka = prof.key_averages()
    tbl = ka.table(sort_by='cuda_time_total', row_limit=60, max_name_column_width=90)
    tag = 'n%d%s' % (world, os.environ.get('LB_TAG', ''))
    open('gpurun_out/profile_step_%s.txt' % tag, 'w').write(tbl)
    rows = {}
    for e in ka:
      t = getattr(e, 'device_time_total', None)
      if t is None:
        t = getattr(e, 'cuda_time_total', 0.0)
      if t <= 0:
        continue
      rows[e.key] = {'calls_per_step': e.count / steps, 'us_per_step': t / steps,
                     'avg_us': t / max(e.count, 1)}
    total = sum(r['us_per_step'] for r in rows.values())
    json.dump({'n_gpus': world, 'ms_per_step_unprofiled': ms_plain,
               'kernel_ms_per_step_sum': total / 1e3, 'cuda_graph': engine.cuda_graph,
               'kernels': dict(sorted(rows.items(), key=lambda kv: -kv[1]['us_per_step']))},
              open('gpurun_out/profile_step_%s.json' % tag, 'w'), indent=1)
    print('step %.2f ms (unprofiled), kernel-time sum %.2f ms' % (ms_plain, total / 1e3))
    print(tbl[:6000])
  if world > 1:
    engine._graphed = None   # pylint: disable=protected-access
    import gc
    gc.collect()
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
