"""Differentiable assignment (ref `lingvo/core/differentiable_assignment.py:28`).

`max_assignment(score [B,N,M], elementwise upper bound, row/col capacities)` solves the
entropy-regularised transport problem with Sinkhorn iterations in log space; the result
is a soft assignment matrix whose row sums ≤ row capacity and column sums ≤ column
capacity, differentiable w.r.t. the scores."""
import torch


def max_assignment(score, *, elementwise_upper_bound, row_sums, col_sums, epsilon=0.1,  # pylint: disable=invalid-name
                   num_iterations=50, use_epsilon_scaling=True):
  """score `[B,N,M]`; row_sums `[B,N]`; col_sums `[B,M]` → (assignment `[B,N,M]`, diff)."""
  ub = torch.as_tensor(elementwise_upper_bound, dtype=score.dtype, device=score.device)
  log_r = torch.log(row_sums.clamp_min(1e-30)).unsqueeze(-1)
  log_c = torch.log(col_sums.clamp_min(1e-30)).unsqueeze(-2)
  log_ub = torch.log(ub.clamp_min(1e-30))
  u = torch.zeros_like(log_r)
  v = torch.zeros_like(log_c)
  phases = [8.0, 4.0, 2.0, 1.0] if use_epsilon_scaling else [1.0]
  per_phase = max(num_iterations // len(phases), 1)
  x = prev = None
  for mult in phases:
    k = score / (epsilon * mult)
    for _ in range(per_phase):
      prev = x
      x = torch.minimum(k + u + v, log_ub)
      u = u + torch.clamp(log_r - torch.logsumexp(x, -1, keepdim=True), max=0.0)
      x = torch.minimum(k + u + v, log_ub)
      v = v + torch.clamp(log_c - torch.logsumexp(x, -2, keepdim=True), max=0.0)
      x = torch.minimum(k + u + v, log_ub)
  out = torch.exp(x)
  # final projection: scaling down can only reduce sums, so both capacities hold exactly
  out = out * torch.clamp(row_sums.unsqueeze(-1) / out.sum(-1, keepdim=True).clamp_min(1e-30), max=1.0)
  out = out * torch.clamp(col_sums.unsqueeze(-2) / out.sum(-2, keepdim=True).clamp_min(1e-30), max=1.0)
  diff = (torch.exp(x) - torch.exp(prev)).abs().max() if prev is not None else torch.zeros(())
  return out, diff
