"""GPipe relative throughput of the Transformer-big MT pipeline (BASELINE config #5).

  python tools/gpipe_bench.py                                   # 1 GPU, no pipelining
  torchrun --nproc-per-node N tools/gpipe_bench.py              # N pipeline stages

The only GPU scaling curve the reference publishes is GPipe relative throughput
1.0 / 0.93 / 0.85 / 0.775 at 1 / 2 / 4 / 8 V100 (`lm/params/one_billion_wds.py:169-179`,
same total batch, more stages ⇒ more bubbles). This tool measures the same quantity for
`mt.wmt14_en_de.WmtEnDeTransformerBigGPipe`: tokens/s at fixed global batch with N stages,
device-timed (CUDA events, max over ranks), synthetic token ids. Rank 0 appends one JSON
line to gpurun_out/gpipe_bench.jsonl.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from lingvo_b200.core.nested_map import NestedMap


def main():
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  lr = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(lr)
  dev = torch.device('cuda', lr)
  if world > 1:
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=dev)
  from lingvo_b200 import model_registry
  import lingvo_b200.models.mt.params.wmt14_en_de  # noqa: F401
  cfg = model_registry.GetParams('mt.wmt14_en_de.WmtEnDeTransformerBigGPipe', 'Train')
  tp = cfg.task
  tp.input = None
  tp.random_seed = 3
  micro = int(os.environ.get('LB_MICRO', '8'))
  tp.stack.Set(num_splits=world, splits=world, num_micro_batches=micro)
  tp.remat = os.environ.get('LB_REMAT', '1') == '1'
  if os.environ.get('LB_DTYPE', 'bf16') == 'bf16':
    tp.fprop_dtype = torch.bfloat16
    tp.stack.fprop_dtype = torch.bfloat16
  task = tp.Instantiate()
  task.to(dev)
  b, t, v = int(os.environ.get('LB_BATCH', '128')), 96, 32000
  g = torch.Generator().manual_seed(0)
  ids = torch.randint(1, v, (b, t), generator=g).to(dev)
  pad = torch.zeros(b, t, device=dev)
  batch = NestedMap(src=NestedMap(ids=ids, paddings=pad),
                    tgt=NestedMap(ids=torch.roll(ids, 1, 1), labels=ids, paddings=pad,
                                  weights=1 - pad))

  def step():
    task.FPropDefaultTheta(batch)
    task.BProp()

  for _ in range(3):
    step()
  torch.cuda.synchronize()
  if world > 1:
    dist.barrier()
  k = int(os.environ.get('LB_STEPS', '8'))
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(k):
    step()
  e1.record()
  torch.cuda.synchronize()
  ms = torch.tensor([e0.elapsed_time(e1) / k], device=dev)
  if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
  if rank == 0:
    out = {'model': 'mt.wmt14_en_de.WmtEnDeTransformerBigGPipe', 'stages': world,
           'global_batch': b, 'seq_len': t, 'micro_batches': micro, 'remat': bool(tp.remat),
           'dtype': os.environ.get('LB_DTYPE', 'bf16'), 'ms_per_step': float(ms),
           'tokens_per_s': 2 * b * t / (float(ms) / 1e3),
           'loss': float(task._eval_metrics['loss'][0])}  # pylint: disable=protected-access
    os.makedirs('gpurun_out', exist_ok=True)
    with open('gpurun_out/gpipe_bench.jsonl', 'a') as f:
      f.write(json.dumps(out) + '\n')
    print(json.dumps(out))
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
