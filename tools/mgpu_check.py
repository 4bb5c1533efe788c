"""Multi-GPU numerics checks (run under torchrun): fused EP/DP vs NCCL baseline."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist

def run(mode, steps=3, dense_adam=False):
  from lingvo_b200 import model_registry
  from lingvo_b200.core import cluster_factory
  from lingvo_b200.parallel import mesh as mesh_lib, dp as dp_lib
  import lingvo_b200.models.lm.params.synthetic_packed_input  # noqa
  mesh_lib.Reset(mode=mode)
  name = 'DenseLmTiny' if dense_adam else 'MoELm8ETiny'
  cfg = model_registry.GetParams('lm.synthetic_packed_input.' + name, 'Train')
  if dense_adam:
    from lingvo_b200.core import optimizer
    cfg.task.train.optimizer = optimizer.Adam.Params().Set(beta1=0.9, beta2=0.98, epsilon=1e-8)
    cfg.task.train.learning_rate = 1e-3
    cfg.task.train.clip_gradient_norm_to_value = 1.0
  cfg.task.random_seed = 1
  cfg.input.random_seed = 5
  cfg.cluster.worker.gpus_per_replica = 1
  losses = []
  with cluster_factory.Cluster(cfg.cluster):
    m = cfg.Instantiate(); m.to(torch.device('cuda', torch.cuda.current_device()))
    task = m.tasks[0]
    dp_lib.Attach(task)
    for _ in range(steps):
      metrics, _ = task.TrainStep()
      losses.append(float(metrics['loss'][0].detach()))
    # The global gradient norm must agree on every rank (expert grads are rank-local and
    # are summed over the EP group inside the learner).
    gn = metrics.get('grad_norm/all')
    if gn is not None:
      g = gn[0].detach().float().reshape(1).cuda()
      gs = [torch.zeros_like(g) for _ in range(dist.get_world_size())]
      dist.all_gather(gs, g)
      vals = [float(x) for x in gs]
      assert max(vals) - min(vals) <= 1e-4 * max(vals), ('grad norm differs across ranks', vals)
  return losses

def main():
  lr = int(os.environ.get('LOCAL_RANK', 0))
  torch.cuda.set_device(lr)
  dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
  a = run('nccl'); b = run('fused')
  c = run('nccl', 4, True); d = run('fused', 4, True)        # dense LM + ZeRO-Adam
  if dist.get_rank() == 0:
    print(json.dumps({'nccl': a, 'fused': b, 'adam_nccl': c, 'zero_adam': d}))
    assert all(abs(x - y) < 5e-2 for x, y in zip(a, b)), (a, b)
    assert all(abs(x - y) < 5e-2 for x, y in zip(c, d)), (c, d)
    assert d[-1] < d[0], d
    print('MGPU_OK')
  dist.barrier(); dist.destroy_process_group()

if __name__ == '__main__':
  main()
