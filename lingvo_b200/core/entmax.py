"""α-entmax (α = 1.5) and sparsemax with their gradients (ref `lingvo/core/entmax.py`).

entmax15(z) = argmax_p <p, z> + H^T_1.5(p): sparse probabilities computed exactly by
sorting (Peters et al. 2019)."""
import torch


class _Entmax15(torch.autograd.Function):

  @staticmethod
  def forward(ctx, x, dim):
    x = x / 2
    x = x - x.max(dim, keepdim=True).values
    xs, _ = torch.sort(x, dim=dim, descending=True)
    k = torch.arange(1, x.shape[dim] + 1, device=x.device, dtype=x.dtype)
    shape = [1] * x.dim()
    shape[dim] = -1
    k = k.view(shape)
    mean = xs.cumsum(dim) / k
    mean_sq = (xs * xs).cumsum(dim) / k
    ss = k * (mean_sq - mean * mean)
    delta = ((1 - ss) / k).clamp_min(0)
    tau = mean - torch.sqrt(delta)
    support = (tau <= xs).sum(dim, keepdim=True)
    tau_star = tau.gather(dim, support - 1)
    y = torch.clamp(x - tau_star, min=0) ** 2
    ctx.save_for_backward(y)
    ctx.dim = dim
    return y

  @staticmethod
  def backward(ctx, dy):
    y, = ctx.saved_tensors
    g = y.sqrt()
    dx = dy * g
    q = dx.sum(ctx.dim, keepdim=True) / g.sum(ctx.dim, keepdim=True)
    return dx - q * g, None


def entmax15(x, dim=-1):  # pylint: disable=invalid-name
  return _Entmax15.apply(x, dim)


def sparsemax(x, dim=-1):  # pylint: disable=invalid-name
  """Euclidean projection onto the simplex (autograd-differentiable)."""
  xs, _ = torch.sort(x, dim=dim, descending=True)
  k = torch.arange(1, x.shape[dim] + 1, device=x.device, dtype=x.dtype)
  shape = [1] * x.dim()
  shape[dim] = -1
  k = k.view(shape)
  css = xs.cumsum(dim) - 1
  support = ((xs - css / k) > 0).sum(dim, keepdim=True)
  tau = css.gather(dim, support - 1) / support.to(x.dtype)
  return torch.clamp(x - tau, min=0)


def entmax_support(x, dim=-1):  # pylint: disable=invalid-name
  return entmax15(x, dim) > 0
