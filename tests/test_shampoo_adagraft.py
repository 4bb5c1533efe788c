"""Distributed Shampoo + AdaGraft (ref `distributed_shampoo_test.py`, `adagraft_test.py`)."""

import numpy as np
import pytest
import torch

from lingvo_b200.core import adagraft
from lingvo_b200.core import distributed_shampoo as ds
from lingvo_b200.core import optimizer
from lingvo_b200.core import py_utils


def _Var(shape, name='layer/w/var', seed=0):
  g = torch.Generator().manual_seed(seed)
  w = torch.nn.Parameter(torch.randn(*shape, generator=g))
  w.var_name = name
  return w


def _InvRoot(m, p):
  """m^{-1/p} through a float64 eigendecomposition (oracle)."""
  s, u = np.linalg.eigh(m.astype(np.float64))
  return (u * np.maximum(s, 1e-30) ** (-1.0 / p)) @ u.T


# ----------------------------------------------------------------------- partitioner --
def test_tensor_partitioner_round_trip_and_metadata():
  cfg = ds.PartitionConfig(max_dim_size=5, partition_size=3)
  t = torch.arange(8 * 4 * 7, dtype=torch.float32).reshape(8, 4, 7)
  meta = ds.TensorPartitioner.partition_metadata(t, cfg)
  assert meta.split_sizes_per_dim == [[3, 3, 2], [4], [3, 3, 1]]
  assert meta.num_splits_per_dim == [3, 1, 3]
  parts = ds.TensorPartitioner.partition_tensor(t, cfg)
  assert len(parts) == 9 and parts[0].shape == (3, 4, 3) and parts[-1].shape == (2, 4, 1)
  assert parts[0].data_ptr() == t.data_ptr()                    # views, no copies
  back = ds.TensorPartitioner.reform_tensor(parts, meta.num_splits_per_dim)
  assert torch.equal(back, t)
  with pytest.raises(ValueError):
    ds.PartitionConfig(4, 5)


# -------------------------------------------------------------------------- shampoo --
def test_shampoo_first_step_matches_the_closed_form():
  """momentum 0, preconditioning from step 0: the update is L^{-1/4} G R^{-1/4} rescaled to
  the norm of the AdaGrad step (= sign(G))."""
  w = _Var((5, 3))
  w0 = w.detach().clone()
  g = torch.randn(5, 3, generator=torch.Generator().manual_seed(1))
  opt = optimizer.DistributedShampoo.Params().Set(
      name='sh', momentum=0.0, start_preconditioning_steps=0, matrix_epsilon=1e-12).Instantiate()
  assert isinstance(opt, ds.DistributedShampoo)
  opt.Apply(0.1, [py_utils.VarGrad(w, g)])
  gn = g.numpy().astype(np.float64)
  l, r = gn @ gn.T, gn.T @ gn
  # rank-deficient L (5×5 from a 5×3 G): compare through the damped oracle
  damp = lambda m: m + 1e-12 * np.linalg.eigvalsh(m).max() * np.eye(len(m))
  pg = _InvRoot(damp(l), 4) @ gn @ _InvRoot(damp(r), 4)
  diag = gn / np.sqrt(gn * gn + 1e-30)
  want = w0.numpy() - 0.1 * pg * (np.linalg.norm(diag) / np.linalg.norm(pg))
  np.testing.assert_allclose(w.detach().numpy(), want, rtol=2e-3, atol=2e-3)
  slots = opt.GetOptimizerSlots()
  assert {'layer/w/accumulator', 'layer/w/mat_statistics_0', 'layer/w/mat_statistics_1',
          'layer/w/mat_preconditioner_0', 'layer/w/mat_preconditioner_1'} <= set(slots)
  np.testing.assert_allclose(slots['layer/w/mat_statistics_0'].numpy(), l, rtol=1e-5, atol=1e-5)


def test_shampoo_diagonal_until_start_then_warms_up():
  kw = dict(momentum=0.0, start_preconditioning_steps=2)
  w_s, w_a = _Var((4, 4)), _Var((4, 4))
  sh = optimizer.DistributedShampoo.Params().Set(name='sh', **kw).Instantiate()
  ada = optimizer.Adagrad.Params().Set(name='ada', initial_accumulator_value=0.0).Instantiate()
  gen = torch.Generator().manual_seed(3)
  for step in range(4):
    g = torch.randn(4, 4, generator=gen)
    before = w_s.detach().clone()
    sh.Apply(0.05, [py_utils.VarGrad(w_s, g)])
    acc = sh._slots['layer/w/var']['accumulator']   # pylint: disable=protected-access
    diag_step = 0.05 * g * torch.rsqrt(acc + 1e-30)
    if step <= 2:      # steps 0,1: diagonal; step 2: warm-up weight (2-2)/2 = 0 → still diagonal
      torch.testing.assert_close(before - w_s.detach(), diag_step, atol=1e-6, rtol=1e-5)
    else:              # step 3: half preconditioned
      assert not torch.allclose(before - w_s.detach(), diag_step, atol=1e-5)
  del w_a, ada


def test_shampoo_rank3_and_partial_preconditioning():
  """A rank-3 tensor with one axis above fallback_to_diagonal_dim: exponents are −1/(2·2),
  the large axis is left alone (ref `_compute_preconditioned_raw_grad` tensordot rotation)."""
  w = _Var((3, 9, 4))
  opt = optimizer.DistributedShampoo.Params().Set(
      name='sh', momentum=0.0, start_preconditioning_steps=0, fallback_to_diagonal_dim=8,
      matrix_epsilon=1e-10).Instantiate()
  g = torch.randn(3, 9, 4, generator=torch.Generator().manual_seed(5))
  opt.Apply(1.0, [py_utils.VarGrad(w, g)])
  st = opt._slots['layer/w/var']   # pylint: disable=protected-access
  assert 'mat_statistics_0' in st and 'mat_statistics_2' in st and 'mat_statistics_1' not in st
  gn = g.numpy().astype(np.float64)
  s0 = np.tensordot(gn, gn, axes=([1, 2], [1, 2]))
  s2 = np.tensordot(gn, gn, axes=([0, 1], [0, 1]))
  np.testing.assert_allclose(st['mat_statistics_0'].numpy(), s0, rtol=1e-4, atol=1e-4)
  damp = lambda m: m + 1e-10 * np.linalg.eigvalsh(m).max() * np.eye(len(m))
  pg = np.einsum('ia,ibc->abc', _InvRoot(damp(s0), 4), gn)
  pg = np.einsum('abc,cd->abd', pg, _InvRoot(damp(s2), 4))
  raw = opt._PreconditionedRawGrad(w, opt._Blocks(g)).numpy()   # pylint: disable=protected-access
  np.testing.assert_allclose(raw, pg, rtol=5e-3, atol=5e-3)


def test_shampoo_blocks_large_axes_and_falls_back():
  opt = optimizer.DistributedShampoo.Params().Set(
      name='sh', momentum=0.9, start_preconditioning_steps=0, block_size=4,
      block_partition_threshold_size=6, max_any_dim=64).Instantiate()
  big, vec, huge = _Var((10, 5), 'a/w/var'), _Var((7,), 'b/b/var'), _Var((70, 2), 'c/w/var')
  vg = [py_utils.VarGrad(v, torch.randn_like(v)) for v in (big, vec, huge)]
  opt.Apply(0.01, vg)
  a = opt._slots['a/w/var']   # pylint: disable=protected-access
  # 10 → blocks of 4,4,2 along axis 0; axis 1 (5 ≤ 6) is not cut
  assert a['0_mat_statistics_0'].shape == (4, 4) and a['2_mat_statistics_0'].shape == (2, 2)
  assert a['1_mat_statistics_1'].shape == (5, 5)
  assert 'precond_grad_momentum' in a and 'momentum' in a
  for key in ('b/b/var', 'c/w/var'):                    # rank 1 / dim > max_any_dim: diagonal
    assert not any('mat_' in k for k in opt._slots[key])   # pylint: disable=protected-access


def test_shampoo_moving_average_and_stat_frequency():
  w = _Var((3, 3))
  opt = optimizer.DistributedShampoo.Params().Set(
      name='sh', momentum=0.0, second_moment_averaging=0.5,
      statistics_computation_frequency=2).Instantiate()
  g = torch.eye(3)
  for _ in range(3):                                    # stats at steps 0 and 2 only
    opt.Apply(0.0, [py_utils.VarGrad(w, g)])
  stat = opt._slots['layer/w/var']['mat_statistics_0']   # pylint: disable=protected-access
  torch.testing.assert_close(stat, torch.eye(3) * 0.75)  # 0.5·(0.5·0 + 0.5) + 0.5


def test_shampoo_converges_faster_than_adagrad_on_an_ill_conditioned_quadratic():
  torch.manual_seed(0)
  a = torch.diag(torch.tensor([1.0, 0.01, 1.0, 0.01, 1.0, 0.01]))
  q = torch.linalg.qr(torch.randn(6, 6))[0]
  h = q @ a @ q.t()
  target = torch.randn(6, 6)

  def Run(p, steps=150):
    w = _Var((6, 6), seed=7)
    opt = p.Instantiate()
    for _ in range(steps):
      loss = 0.5 * ((w - target).t() @ h @ (w - target)).trace()
      g, = torch.autograd.grad(loss, w)
      opt.Apply(0.3, [py_utils.VarGrad(w, g)])
    return float(0.5 * ((w - target).t() @ h @ (w - target)).trace())

  sh = Run(optimizer.DistributedShampoo.Params().Set(name='sh', momentum=0.0,
                                                     start_preconditioning_steps=1))
  ad = Run(optimizer.Adagrad.Params().Set(name='ada', initial_accumulator_value=0.0))
  assert sh < 0.8 * ad, (sh, ad)


def test_shampoo_checkpoint_round_trip_including_matrix_slots():
  def Make():
    return optimizer.DistributedShampoo.Params().Set(
        name='sh', momentum=0.9, start_preconditioning_steps=1, block_size=3,
        block_partition_threshold_size=4).Instantiate()
  gen = torch.Generator().manual_seed(11)
  grads = [torch.randn(6, 4, generator=gen) for _ in range(5)]
  w1, o1 = _Var((6, 4)), Make()
  for g in grads[:3]:
    o1.Apply(0.05, [py_utils.VarGrad(w1, g)])
  saved = {k: v.clone() for k, v in o1.GetOptimizerSlots().items()}
  assert any('1_mat_preconditioner_0' in k for k in saved)
  w2, o2 = _Var((6, 4)), Make()
  with torch.no_grad():
    w2.copy_(w1)
  used = o2.LoadOptimizerSlots(saved)
  assert set(used) == set(saved)
  for g in grads[3:]:
    o1.Apply(0.05, [py_utils.VarGrad(w1, g)])
    o2.Apply(0.05, [py_utils.VarGrad(w2, g)])
  torch.testing.assert_close(w1, w2)


# ------------------------------------------------------------------------- adagraft --
def _Graft(**kw):
  return optimizer.AdaGraft.Params().Set(
      name='graft', magnitude_optimizer=optimizer.SGD.Params(),
      direction_optimizer=optimizer.Adam.Params().Set(beta1=0.9, beta2=0.999, epsilon=1e-8),
      **kw).Instantiate()


def test_adagraft_takes_norm_from_sgd_and_direction_from_adam():
  w, wa = _Var((4, 3)), _Var((4, 3))
  g = torch.randn(4, 3, generator=torch.Generator().manual_seed(2))
  opt = _Graft()
  assert isinstance(opt, adagraft.AdaGraft)
  adam = optimizer.Adam.Params().Set(name='adam', beta1=0.9, beta2=0.999,
                                     epsilon=1e-8).Instantiate()
  w0 = w.detach().clone()
  opt.Apply(0.1, [py_utils.VarGrad(w, g)])
  adam.Apply(0.1, [py_utils.VarGrad(wa, g)])
  step = w.detach() - w0
  adam_step = wa.detach() - w0
  assert float(step.norm()) == pytest.approx(float((0.1 * g).norm()), rel=1e-5)   # SGD's size
  cos = float((step * adam_step).sum() / (step.norm() * adam_step.norm()))
  assert cos == pytest.approx(1.0, abs=1e-5)                                      # Adam's way
  # child slots are checkpointed, the scratch copy is not
  slots = opt.GetOptimizerSlots()
  assert any(k.endswith('/Adam') for k in slots) and not any('scratch' in k for k in slots)


def test_adagraft_global_norm_and_direction_lr_and_zero_direction():
  a, b = _Var((3, 3), 'a/w/var', 1), _Var((2, 5), 'b/w/var', 2)
  a0, b0 = a.detach().clone(), b.detach().clone()
  ga, gb = torch.randn(3, 3) * 5, torch.randn(2, 5) * 0.1
  opt = _Graft(use_global_norm=True, direction_optimizer_lr=1.0, diagnostic=True)
  opt.Apply(0.2, [py_utils.VarGrad(a, ga), py_utils.VarGrad(b, gb)])
  sa, sb = a.detach() - a0, b.detach() - b0
  total = torch.sqrt(sa.norm() ** 2 + sb.norm() ** 2)
  want = torch.sqrt((0.2 * ga).norm() ** 2 + (0.2 * gb).norm() ** 2)
  assert float(total) == pytest.approx(float(want), rel=1e-4)
  # per-layer norms are NOT those of SGD (only the global one is grafted) …
  assert abs(float(sb.norm()) - float((0.2 * gb).norm())) > 1e-3
  # … and the first Adam step has equal per-element size, so the split follows √numel
  assert float(sa.norm() / sb.norm()) == pytest.approx((9 / 10) ** 0.5, rel=1e-3)
  assert set(opt.m_step_norm) == {'a/w/var', 'b/w/var'}
  # zero gradient → zero direction step → no movement and no NaNs
  c = _Var((2, 2), 'c/w/var')
  c0 = c.detach().clone()
  o2 = _Graft()
  o2.Apply(0.1, [py_utils.VarGrad(c, torch.zeros(2, 2))])
  assert torch.equal(c.detach(), c0)


def test_adagraft_state_round_trip():
  gen = torch.Generator().manual_seed(4)
  grads = [torch.randn(3, 4, generator=gen) for _ in range(4)]
  w1, o1 = _Var((3, 4)), _Graft()
  for g in grads[:2]:
    o1.Apply(0.1, [py_utils.VarGrad(w1, g)])
  saved = {k: v.clone() for k, v in o1.GetOptimizerSlots().items()}
  w2, o2 = _Var((3, 4)), _Graft()
  with torch.no_grad():
    w2.copy_(w1)
  o2.LoadOptimizerSlots(saved)
  for g in grads[2:]:
    o1.Apply(0.1, [py_utils.VarGrad(w1, g)])
    o2.Apply(0.1, [py_utils.VarGrad(w2, g)])
  torch.testing.assert_close(w1, w2)
