"""Per-phase timing of the fused TP FFN forward (run under torchrun, 2+ GPUs)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
  rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
  dev = torch.device('cuda', int(os.environ['LOCAL_RANK']))
  torch.cuda.set_device(dev)
  dist.init_process_group('nccl', device_id=dev)
  from lingvo_b200 import ops
  from lingvo_b200.ops import gemm as G
  from lingvo_b200.parallel import tp
  T, M, H = 8192, 2048, 8192
  ms, hs = M // world, H // world
  bf = torch.bfloat16
  eng = tp.TpEngine(T, M, H, dev)
  xfull = torch.randn(T, M, device=dev, dtype=bf)
  wi = torch.randn(M, hs, device=dev, dtype=bf) * 0.02
  wo = torch.randn(hs, M, device=dev, dtype=bf) * 0.02
  h = torch.randn(T, hs, device=dev, dtype=bf)
  st = eng.sets[0]
  nat = ops.native()

  def timeit(fn, n=30):
    for _ in range(3):
      fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
      fn()
    e1.record(); dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round(float(t) * 1e3, 1)

  r = {}
  r['gemm1_local_us'] = timeit(lambda: G.gemm(xfull, wi, True, False, act=1))
  r['gemm1_agprologue_us'] = timeit(lambda: G.gemm(st['x'], wi, True, False, act=1, a_peer_ptrs=st['x_peers']))
  r['gemm1_cublas_us'] = timeit(lambda: torch.relu(xfull @ wi))
  r['gemm2_local_us'] = timeit(lambda: G.gemm(h, wo, True, False))
  r['gemm2_rs_epilogue_us'] = timeit(lambda: G.gemm(h, wo, True, False, nblk_ptrs=st['nblk'], nblk_ld=ms))
  r['gemm2_cublas_us'] = timeit(lambda: h @ wo)
  r['sync_us'] = timeit(lambda: eng.chan.Sync(0))
  r['reduce_us'] = timeit(lambda: nat.tp_reduce_slabs(st['slab'], world))
  xs = xfull[:, :ms].contiguous()
  r['stage_copy_us'] = timeit(lambda: st['x'].copy_(xs))
  parts = [torch.empty_like(xs) for _ in range(world)]
  r['nccl_allgather_us'] = timeit(lambda: dist.all_gather(parts, xs))
  y = torch.randn(T, M, device=dev, dtype=bf)
  out = torch.empty(T, ms, device=dev, dtype=bf)
  r['nccl_reduce_scatter_us'] = timeit(lambda: dist.reduce_scatter_tensor(out, y.view(world, -1, ms) if False else y.t().contiguous().view(world * ms, T)[:world * ms].reshape(world, ms * T).reshape(-1)[:T * M].view(-1)) if False else dist.reduce_scatter(out, [c.contiguous() for c in y.split(ms, dim=1)]))
  if rank == 0:
    print(json.dumps(r))
  dist.barrier(); dist.destroy_process_group()


if __name__ == '__main__':
  main()
