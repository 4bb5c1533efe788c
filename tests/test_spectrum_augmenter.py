"""SpecAugment (ref `lingvo/core/spectrum_augmenter_test.py` and `..._on_device_test.py`):
mask geometry, warp semantics, noise, block masks, domains, stateless seeds, and
on-device ≡ base for the same random stream."""

import numpy as np
import pytest
import torch

from lingvo_b200.core import py_utils
from lingvo_b200.core import spectrum_augmenter as sa
from lingvo_b200.core import spectrum_augmenter_on_device as sa_dev


def _Aug(cls=sa.SpectrumAugmenter, **kw):
  base = dict(name='aug', freq_mask_max_bins=0, time_mask_max_frames=0, random_seed=1234)
  base.update(kw)
  return cls.Params().Set(**base).Instantiate()


def _Spans(row):
  """Lengths of the zero runs of a 1-D 0/1 array."""
  runs, cur = [], 0
  for v in row:
    if v == 0:
      cur += 1
    elif cur:
      runs.append(cur)
      cur = 0
  if cur:
    runs.append(cur)
  return runs


def test_time_mask_respects_lengths_widths_and_count():
  aug = _Aug(time_mask_max_frames=5, time_mask_count=2)
  b, t = 16, 40
  x = torch.ones(b, t, 6, 1)
  pad = torch.zeros(b, t)
  pad[:, 30:] = 1.0
  y, p2 = aug.FPropDefaultTheta(x, pad)
  assert p2 is pad
  m = y[:, :, 0, 0].numpy()
  assert (y == y[:, :, :1, :1].expand_as(y)).all()          # whole frames are masked
  total = 0
  for r in m:
    runs = _Spans(r)
    assert len(runs) <= 2 and all(w <= 2 * 4 for w in runs)  # each width < 5 (two may touch)
    assert r[30:].all()                                     # never beyond the valid length
    total += sum(runs)
  assert total > 0


def test_dynamic_time_mask_scales_with_length_and_ratio():
  aug = _Aug(use_dynamic_time_mask_max_frames=True, time_mask_max_ratio=0.25, time_mask_count=1)
  b, t = 64, 80
  x = torch.ones(b, t, 2, 1)
  pad = torch.zeros(b, t)
  pad[: b // 2, 20:] = 1.0                                    # short utterances: ≤ 5 frames
  y, _ = aug.FPropDefaultTheta(x, pad)
  masked = (y[:, :, 0, 0] == 0).sum(1)
  assert int(masked[: b // 2].max()) <= 5
  assert int(masked[b // 2:].max()) <= 20
  assert int(masked[b // 2:].max()) > 5


def test_time_masks_per_frame_adapts_multiplicity():
  aug = _Aug(time_mask_max_frames=3, time_mask_count=10, time_masks_per_frame=0.05)
  b, t = 8, 100
  x = torch.ones(b, t, 2, 1)
  pad = torch.zeros(b, t)
  pad[:4, 20:] = 1.0                                          # 20 frames → 1 mask; 100 → 5
  y, _ = aug.FPropDefaultTheta(x, pad)
  m = y[:, :, 0, 0].numpy()
  assert all(len(_Spans(r)) <= 1 for r in m[:4])
  assert max(len(_Spans(r)) for r in m[4:]) > 1


def test_frequency_mask_geometry():
  aug = _Aug(freq_mask_max_bins=4, freq_mask_count=3)
  x = torch.ones(10, 7, 30, 2)
  y, _ = aug.FPropDefaultTheta(x, torch.zeros(10, 7))
  m = y[:, 0, :, 0].numpy()
  assert (y == y[:, :1, :, :1].expand_as(y)).all()          # same bins on every frame/channel
  for r in m:
    assert len(_Spans(r)) <= 3 and sum(_Spans(r)) <= 9
  assert (m == 0).any()


def test_masks_with_noise_fill_only_masked_frames():
  aug = _Aug(time_mask_max_frames=8, time_mask_count=1, use_noise=True, gaussian_noise=True)
  x = torch.full((6, 30, 5, 1), 3.0)
  y, _ = aug.FPropDefaultTheta(x, torch.zeros(6, 30))
  changed = (y != 3.0).any(-1).any(-1)
  assert changed.any()
  vals = y[changed]
  assert float(vals.std()) > 0.5 and abs(float(vals.mean())) < 0.5      # N(0, 1) noise
  small = _Aug(time_mask_max_frames=8, time_mask_count=1, use_noise=True)
  y2, _ = small.FPropDefaultTheta(x, torch.zeros(6, 30))
  vals2 = y2[(y2 != 3.0).any(-1).any(-1)]
  assert 0.05 < float(vals2.std()) < 0.25                    # stddev ∈ (0.1, 0.2)


def test_warp_matrix_rows_interpolate_and_fix_the_ends():
  aug = _Aug()
  n = 12
  origin = torch.tensor([4.0, 6.0])
  dest = torch.tensor([7.0, 6.0])
  cr = torch.tensor([10.0, 12.0])
  w = aug._ConstructWarpMatrix(2, n, origin, dest, cr)
  assert w.shape == (2, n, n)
  np.testing.assert_allclose(w[:, 1:10].sum(-1).numpy(), 1.0, atol=1e-5)
  # identity map when destination == origin
  np.testing.assert_allclose(w[1].numpy(), np.eye(n), atol=1e-6)
  # row `dest` reads exactly pixel `origin`; rows ≥ choose_range are identity
  assert float(w[0, 7, 4]) == pytest.approx(1.0)
  np.testing.assert_allclose(w[0, 10:, 10:].numpy(), np.eye(2), atol=1e-6)
  assert float(w[0, 0, 0]) == pytest.approx(1.0)
  # a linear ramp warps into the piecewise-linear source coordinate itself
  ramp = torch.arange(n, dtype=torch.float32)
  src = aug._SourceCoordinates(n, origin, dest, cr)
  np.testing.assert_allclose((w[0] @ ramp).numpy(), src[0].clamp(max=n - 1).numpy(), atol=1e-4)


def test_time_warp_moves_energy_but_keeps_padding_region():
  aug = _Aug(time_warp_max_frames=6, time_warp_max_ratio=1.0)
  b, t = 12, 40
  x = torch.zeros(b, t, 3, 1)
  x[:, :, :, 0] = torch.arange(t, dtype=torch.float32).view(1, t, 1)
  pad = torch.zeros(b, t)
  pad[:, 32:] = 1.0
  y, _ = aug.FPropDefaultTheta(x, pad)
  assert not torch.allclose(y, x)
  torch.testing.assert_close(y[:, 32:], x[:, 32:], atol=1e-3, rtol=1e-4)    # beyond the length
  torch.testing.assert_close(y[:, 0], x[:, 0])                            # left end fixed
  d = (y[:, :32, 0, 0] - x[:, :32, 0, 0]).abs().max()
  assert float(d) <= 6.0 + 1e-4                                           # bounded shift
  assert (y[:, 1:32, 0, 0] >= y[:, :31, 0, 0] - 1e-4).all()               # monotone


def test_dynamic_time_warp_bound():
  aug = _Aug(time_warp_bound='dynamic', time_warp_max_ratio=0.1)
  b, t = 32, 50
  x = torch.arange(t, dtype=torch.float32).view(1, t, 1, 1).expand(b, t, 2, 1).contiguous()
  y, _ = aug.FPropDefaultTheta(x, torch.zeros(b, t))
  # (the map fixes 0 and `length`, one past the last frame, so the last frame may blend with
  # the out-of-range pixel — same as the reference's warp matrix; with a stretched tail
  # (slope < 1) the last few frames read past `length - 1` too: compare the interior)
  d = float((y - x)[:, :-6].abs().max())
  assert 0.0 < d <= 5.0 + 1e-4


def test_frequency_warp_is_along_frequency_only():
  aug = _Aug(freq_warp_max_bins=4)
  b, t, f = 8, 5, 24
  x = torch.arange(f, dtype=torch.float32).view(1, 1, f, 1).expand(b, t, f, 1).contiguous()
  y, _ = aug.FPropDefaultTheta(x, torch.zeros(b, t))
  assert not torch.allclose(y, x)
  assert (y == y[:, :1].expand_as(y)).all()
  assert float((y - x)[:, :, :-1].abs().max()) <= 4.0 + 1e-4


def test_frequency_noise_scales_bins_and_warms_up():
  aug = _Aug(freq_noise_max_stddev=0.5, freq_noise_warmup_steps=100)
  x = torch.ones(16, 6, 20, 1)
  py_utils.SetGlobalStep(0)
  try:
    y0, _ = aug.FPropDefaultTheta(x, torch.zeros(16, 6))
    torch.testing.assert_close(y0, x)                                      # weight 0 at step 0
    assert aug.augment_weight == 0.0
    py_utils.SetGlobalStep(50)
    assert aug.augment_weight == 0.5
    py_utils.SetGlobalStep(1000)
    assert aug.augment_weight == 1.0
    y, _ = aug.FPropDefaultTheta(x, torch.zeros(16, 6))
  finally:
    py_utils.SetGlobalStep(0)
  assert (y == y[:, :1].expand_as(y)).all()                               # constant over time
  s = y[:, 0, :, 0].std(1)
  assert float(s.max()) < 0.5 * 2.5 and float(s.max()) > 0.02
  assert abs(float(y.mean()) - 1.0) < 0.1


def test_block_mask():
  aug = _Aug(block_mask_prob=0.6, block_mask_size=dict(t=4, f=3))
  b, t, f = 32, 18, 10                                                     # ragged last blocks
  x = torch.ones(b, t, f, 1)
  y, _ = aug.FPropDefaultTheta(x, torch.zeros(b, t))
  m = y[..., 0]
  frac = float((m == 0).float().mean())
  assert 0.1 < frac < 0.5                                                  # E[p] = 0.3
  blocks = m[:, :16, :9].reshape(b, 4, 4, 3, 3)
  assert (blocks == blocks[:, :, :1, :, :1].expand_as(blocks)).all()       # block-constant


def test_unstack_masks_in_the_unstacked_domain():
  aug = _Aug(time_mask_max_frames=2, time_mask_count=1, unstack=True, stack_height=3)
  b, t, f = 8, 10, 12                                                      # 12 = 3 frames × 4
  x = torch.ones(b, t, f, 1)
  pad = torch.zeros(b, t)
  pad[:, 8:] = 1
  y, p2 = aug.FPropDefaultTheta(x, pad)
  assert y.shape == x.shape and p2 is pad
  sub = y.reshape(b, t * 3, 4)                                             # unstacked frames
  masked = (sub == 0).all(-1).sum(1)
  assert int(masked.max()) in (1,) and (sub[:, 24:] == 1).all()            # width < 2, in range
  ux, up = aug.UnstackFeatures(x, pad)
  assert ux.shape == (b, 30, 4, 1) and int(up.sum(1)[0]) == 6


def test_domains_use_their_own_settings_and_others_pass_through():
  aug = _Aug(domain_ids=[2, 7], time_mask_max_frames=[6, 0], time_mask_count=[1, 1],
             freq_mask_max_bins=[0, 5], freq_mask_count=[1, 2])
  b, t, f = 30, 20, 16
  x = torch.ones(b, t, f, 1)
  dom = torch.tensor([2, 7, 1] * 10).view(b, 1)
  y, _ = aug.FPropDefaultTheta(x, torch.zeros(b, t), domain_ids=dom)
  m = y[..., 0]
  d2, d7, d1 = m[0::3], m[1::3], m[2::3]
  assert (d1 == 1).all()                                                   # not listed
  assert (d2 == 0).any() and (d2 == d2[:, :, :1].expand_as(d2)).all()      # time masks only
  assert (d7 == 0).any() and (d7 == d7[:, :1].expand_as(d7)).all()         # freq masks only
  with pytest.raises(AssertionError):
    _Aug(domain_ids=[1, 2], time_mask_count=[1, 2, 3])


def test_seeded_runs_are_reproducible_and_step_dependent():
  kw = dict(time_mask_max_frames=5, time_mask_count=2, freq_mask_max_bins=3)
  x = torch.ones(4, 30, 10, 1)
  a, _ = _Aug(**kw).FPropDefaultTheta(x, torch.zeros(4, 30))
  b_, _ = _Aug(**kw).FPropDefaultTheta(x, torch.zeros(4, 30))
  assert torch.equal(a, b_)
  py_utils.SetGlobalStep(1)
  try:
    c, _ = _Aug(**kw).FPropDefaultTheta(x, torch.zeros(4, 30))
  finally:
    py_utils.SetGlobalStep(0)
  assert not torch.equal(a, c)


def test_input_dependent_seed_is_a_pure_function_of_the_features():
  kw = dict(time_mask_max_frames=5, time_mask_count=2, freq_mask_max_bins=3, random_seed=None,
            use_input_dependent_random_seed=True, use_noise=True)
  g = torch.Generator().manual_seed(0)
  x1 = torch.rand(4, 30, 10, 1, generator=g) + 1.0
  x2 = x1 + 0.5
  aug = _Aug(**kw)
  a, _ = aug.FPropDefaultTheta(x1, torch.zeros(4, 30))
  torch.manual_seed(999)                                                    # global RNG unused
  b_, _ = aug.FPropDefaultTheta(x1, torch.zeros(4, 30))
  assert torch.equal(a, b_)
  c, _ = aug.FPropDefaultTheta(x2, torch.zeros(4, 30))
  assert not torch.equal((a == 0), (c == 0)) or not torch.allclose(a - x1, c - x2)


def test_stateless_streams_are_uniform_and_independent():
  seed = torch.tensor(123456789, dtype=torch.int64)
  u = sa.StatelessUniform((20000,), seed, 1, 'cpu')
  v = sa.StatelessUniform((20000,), seed, 2, 'cpu')
  assert 0.0 <= float(u.min()) and float(u.max()) < 1.0
  assert abs(float(u.mean()) - 0.5) < 0.01 and abs(float(u.var()) - 1 / 12) < 0.005
  assert abs(float(((u - 0.5) * (v - 0.5)).mean())) < 0.003
  assert abs(float(((u[1:] - 0.5) * (u[:-1] - 0.5)).mean())) < 0.003
  n = sa.StatelessNormal((20000,), seed, 3, 'cpu')
  assert abs(float(n.mean())) < 0.03 and abs(float(n.std()) - 1.0) < 0.03
  assert torch.equal(u, sa.StatelessUniform((20000,), seed, 1, 'cpu'))
  assert not torch.equal(u, sa.StatelessUniform((20000,), seed + 1, 1, 'cpu'))


@pytest.mark.parametrize('kw', [
    dict(time_warp_max_frames=5, time_warp_max_ratio=1.0),
    dict(freq_warp_max_bins=3),
    dict(time_mask_max_frames=6, time_mask_count=2, freq_mask_max_bins=4, freq_mask_count=2,
         freq_noise_max_stddev=0.3, use_noise=True, block_mask_prob=0.3,
         block_mask_size=dict(t=4, f=4)),
    dict(time_warp_max_frames=4, time_warp_max_ratio=1.0, freq_warp_max_bins=2,
         time_mask_max_frames=6, time_mask_count=1, unstack=True, stack_height=2),
])
def test_on_device_variant_matches_base_for_the_same_stream(kw):
  g = torch.Generator().manual_seed(1)
  x = torch.randn(6, 24, 16, 2, generator=g)
  pad = torch.zeros(6, 24)
  pad[::2, 18:] = 1.0
  a, _ = _Aug(**kw).FPropDefaultTheta(x, pad)
  b_, _ = _Aug(sa_dev.SpectrumAugmenterOnDevice, **kw).FPropDefaultTheta(x, pad)
  assert not torch.allclose(a, x)
  torch.testing.assert_close(a, b_, atol=1e-5, rtol=1e-5)


def test_rank3_inputs_and_eval_identity():
  from lingvo_b200.core import cluster_factory
  aug = _Aug(time_mask_max_frames=5, time_mask_count=1)
  x = torch.ones(3, 20, 8)
  y, _ = aug.FPropDefaultTheta(x, torch.zeros(3, 20))
  assert y.shape == x.shape and (y == 0).any()
  with cluster_factory.SetEval(True):
    z, _ = aug.FPropDefaultTheta(x, torch.zeros(3, 20))
  assert torch.equal(z, x)
