"""Multi-GPU checks of the *product* training path (run under torchrun).

1. `runners.Trainer` (what `python -m lingvo_b200.trainer --job=trainer_client` runs) for 10
   steps of MoELm8ETiny: replicated weights must be bit-identical on every rank, expert
   shards must differ; the checkpoint is one bundle with a data shard per rank.
2. Kill-and-resume: a fresh Trainer restores the sharded checkpoint and continues.
3. Gradient-level agreement between the fused peer-memory path and the NCCL baseline
   for the same weights and batch (per-variable relative error), not just the loss.
Prints TRAINER_MGPU_OK on success.
"""
import glob
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

MODEL = 'lm.synthetic_packed_input.MoELm8ETiny'


def MakeTrainer(logdir, max_steps, mode='fused', graph='auto'):
  from lingvo_b200 import model_registry
  from lingvo_b200 import runners
  from lingvo_b200.parallel import mesh
  import lingvo_b200.models.lm.params.synthetic_packed_input  # noqa: F401
  mesh.Reset(mode=mode)
  cfg = model_registry.GetParams(MODEL, 'Train')
  cfg.cluster.mode = 'sync'
  cfg.cluster.job = 'trainer_client'
  cfg.cluster.worker.replicas = dist.get_world_size()
  cfg.cluster.worker.gpus_per_replica = 1
  for tp in (cfg.train, cfg.task.train):
    tp.max_steps = max_steps
    tp.save_interval_steps = 5
    tp.async_checkpointing = True
    tp.summary_interval_steps = 5
  cfg.task.train.use_cuda_graph = graph
  cfg.task.train.lr_schedule.warmup_steps = 16
  return runners.Trainer(cfg, '', logdir, '', None)


def Checksums(task):
  rep, exp = [], []
  for v in sorted(task.vars.Flatten(), key=lambda v: v.var_name):
    s = v.data.double().sum().reshape(1)
    (exp if getattr(v, 'expert_parallel', False) else rep).append(s)
  return torch.cat(rep), torch.cat(exp)


def Gather(t):
  out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
  dist.all_gather(out, t)
  return torch.stack(out)


def Grads(mode, seed_state, steps=1):
  """Per-variable synchronised gradients (no optimizer) under `mode`, same weights and the
  same batch every time. `steps` > 1 re-runs fprop/bprop: from the second run on the fused
  engine is in its overlapped (bucketed, comm-stream) schedule."""
  from lingvo_b200 import model_registry
  from lingvo_b200.core import cluster_factory
  from lingvo_b200.core import py_utils
  from lingvo_b200.parallel import dp as dp_lib
  from lingvo_b200.parallel import mesh
  mesh.Reset(mode=mode)
  cfg = model_registry.GetParams(MODEL, 'Train')
  cfg.cluster.worker.gpus_per_replica = 1
  cfg.cluster.worker.replicas = dist.get_world_size()
  with cluster_factory.Cluster(cfg.cluster):
    model = cfg.Instantiate()
    model.to(torch.device('cuda', torch.cuda.current_device()))
    task = model.tasks[0]
    with torch.no_grad():
      for v in task.vars.Flatten():
        v.data.copy_(seed_state[v.var_name])
    dp_lib.Attach(task)
    batch = task.GetInputBatch()
    lrn = task.learners[0]
    for _ in range(steps):
      task.FPropDefaultTheta(batch)
      _, var_grads, _ = lrn._ComputeLossesAndGradients(task._metrics, task.vars)  # pylint: disable=protected-access
      if lrn.grad_sync is not None:
        var_grads = lrn.grad_sync(var_grads)
    out = {}
    for vg in var_grads.Flatten():
      if isinstance(vg, py_utils.VarGrad) and vg.grad is not None:
        out[vg.var.var_name] = vg.grad.detach().float().clone()
    torch.cuda.synchronize()
  return out


def main():
  lr = int(os.environ.get('LOCAL_RANK', 0))
  torch.cuda.set_device(lr)
  dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
  rank, world = dist.get_rank(), dist.get_world_size()
  holder = [tempfile.mkdtemp() if rank == 0 else None]
  dist.broadcast_object_list(holder, src=0)
  logdir = holder[0]
  report = {}

  # 1. product trainer, 10 steps
  torch.manual_seed(1000 + rank)
  t = MakeTrainer(logdir, 10)
  t.Start()
  assert t.task.global_step == 10
  report['cuda_graph'] = t.engine.cuda_graph
  rep, exp = Checksums(t.task)
  reps, exps = Gather(rep), Gather(exp)
  assert torch.equal(reps, reps[0:1].expand_as(reps)), 'replicated weights differ across ranks'
  assert not torch.equal(exps[0], exps[1]), 'expert shards should differ'
  dist.barrier()
  if rank == 0:
    train_dir = os.path.join(logdir, 'train')
    shards = sorted(os.path.basename(p) for p in glob.glob(
        os.path.join(train_dir, 'ckpt-00000010.data-*')))
    assert len(shards) == world, shards
    assert os.path.exists(os.path.join(train_dir, 'ckpt-00000010.index'))
    report['shards'] = shards
  state = {v.var_name: v.data.clone() for v in t.task.vars.Flatten()}
  rep10 = rep.clone()
  del t

  # 2. resume
  t2 = MakeTrainer(logdir, 14)
  t2._checkpointer.Restore()   # pylint: disable=protected-access
  rep_r, exp_r = Checksums(t2.task)
  assert torch.equal(rep_r, rep10), 'restored replicated weights differ from the saved ones'
  assert torch.equal(exp_r, exp), 'restored expert shard differs'
  t2.Start()
  assert t2.task.global_step == 14
  rep2, _ = Checksums(t2.task)
  reps2 = Gather(rep2)
  assert torch.equal(reps2, reps2[0:1].expand_as(reps2))
  assert not torch.equal(rep2, rep10)
  del t2

  # 3. gradients: fused vs NCCL baseline, same weights, same batch
  g_nccl = Grads('nccl', state)
  g_fused = Grads('fused', state)
  # overlapped schedule (3rd run) ≡ monolithic schedule (1st run): same kernel, same values
  g_ovl = Grads('fused', state, steps=3)
  ovl_err = max(float((g_fused[k] - g_ovl[k]).abs().max() / (g_fused[k].abs().max() + 1e-20))
                for k in g_fused)
  report['overlap_vs_monolithic_max_rel'] = ovl_err
  assert ovl_err < 1e-6, ovl_err
  worst = 0.0
  worst_name = None
  for k, a in g_nccl.items():
    b = g_fused[k]
    err = float((a - b).norm() / (a.norm() + 1e-20))
    if err > worst:
      worst, worst_name = err, k
  report['grad_rel_err_max'] = worst
  report['grad_rel_err_var'] = worst_name
  assert worst < 8e-2, (worst, worst_name)     # bf16 GEMMs + different MoE kernels upstream
  w = torch.tensor([worst], device='cuda')
  dist.all_reduce(w, op=dist.ReduceOp.MAX)
  if rank == 0:
    report['grad_rel_err_max_over_ranks'] = float(w)
    print(json.dumps(report))
    print('TRAINER_MGPU_OK')
  dist.barrier()
  dist.destroy_process_group()


if __name__ == '__main__':
  main()
