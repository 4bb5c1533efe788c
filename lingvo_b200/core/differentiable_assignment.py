"""Differentiable assignment (ref `lingvo/core/differentiable_assignment.py:28-200`).

`max_assignment` finds a soft assignment `p [B, N, M]` that maximises `Σ score·p` plus an
entropy term subject to

    Σ_m p[b, n, m] = row_sums[b, n],   Σ_n p[b, n, m] = col_sums[b, m],   0 ≤ p ≤ upper bound.

The box constraint is handled by lifting: a second "slack" slice `s = upper − p` turns the
problem into a three-marginal transport problem over `q[b, k, n, m]` (k = 0 assignment,
k = 1 slack) with marginals  rows `[row_sums, Σ_m upper − row_sums]`, columns
`[col_sums, Σ_n upper − col_sums]` and slices `upper`; iterative Bregman projections (one
log-domain Sinkhorn update per marginal) solve it, with the temperature annealed from 1 down
to `epsilon` when `use_epsilon_scaling` is set. Everything is differentiable w.r.t. `score`.
"""
import torch

_NEG = -1e30


def _SafeLog(x):
  return torch.log(x.clamp_min(1e-38)).clamp_min(-1e36)


def max_assignment(score, *, elementwise_upper_bound, row_sums, col_sums, epsilon=0.1,  # pylint: disable=invalid-name
                   num_iterations=50, use_epsilon_scaling=True):
  """score / elementwise_upper_bound `[B,N,M]`; row_sums `[B,N]`; col_sums `[B,M]`.

  Returns (assignment `[B,N,M]`, iterations used, final temperature, delta) — delta is the
  largest relative violation of the row / column marginals and relative change of the
  solution in the last iteration (the reference's stopping statistic)."""
  b, n, m = score.shape
  ub = torch.as_tensor(elementwise_upper_bound, dtype=score.dtype, device=score.device)
  ub = ub.expand(b, n, m)
  rows = row_sums.reshape(b, n, 1).to(score.dtype)
  cols = col_sums.reshape(b, 1, m).to(score.dtype)
  assert bool(((rows.sum(1) - cols.sum(2)).abs() < 1e-4 * (1 + rows.sum(1).abs())).all()), (
      'row_sums and col_sums must have the same total')
  scores = torch.stack([score, torch.zeros_like(score)], 1)                     # [B, 2, N, M]
  marg_r = torch.stack([rows, ub.sum(-1, keepdim=True) - rows], 1)             # [B, 2, N, 1]
  marg_c = torch.stack([cols, ub.sum(-2, keepdim=True) - cols], 1)             # [B, 2, 1, M]
  marg_k = ub.unsqueeze(1)                                                      # [B, 1, N, M]
  assert bool((marg_r >= -1e-6).all()) and bool((marg_c >= -1e-6).all()) and bool((marg_k >= 0).all())
  log_r, log_c, log_k = _SafeLog(marg_r), _SafeLog(marg_c), _SafeLog(marg_k)
  u, v, w = torch.zeros_like(marg_r), torch.zeros_like(marg_c), torch.zeros_like(marg_k)
  eps = 1.0 if use_epsilon_scaling else float(epsilon)
  prev = None
  for it in range(int(num_iterations)):
    prev = (eps, u, v, w)
    eps = max(float(epsilon), eps * min(0.6 * 1.04 ** it, 0.85))
    s = scores / eps
    u = (log_r - torch.logsumexp(s + (w + v) / eps, -1, keepdim=True).clamp_min(_NEG)) * eps
    v = (log_c - torch.logsumexp(s + (w + u) / eps, -2, keepdim=True).clamp_min(_NEG)) * eps
    w = (log_k - torch.logsumexp(s + (u + v) / eps, -3, keepdim=True).clamp_min(_NEG)) * eps
  q = torch.exp((scores + u + v + w) / eps)
  with torch.no_grad():
    delta = torch.maximum(((marg_r - q.sum(-1, keepdim=True)).abs() / (marg_r + 1e-6)).max(),
                          ((marg_c - q.sum(-2, keepdim=True)).abs() / (marg_c + 1e-6)).max())
    if prev is not None:
      pe, pu, pv, pw = prev
      now, before = torch.exp((u + v + w) / eps), torch.exp((pu + pv + pw) / pe)
      delta = torch.maximum(delta, ((before - now).abs() / (now + 1e-6)).max())
  return q[:, 0], int(num_iterations), eps, delta
