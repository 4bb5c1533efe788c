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


def _EntmaxProb(x, alpha_m1):
  return torch.clamp(x, min=0) ** (1.0 / alpha_m1)


class _EntmaxBisect(torch.autograd.Function):
  """α-entmax for any α > 1 by bisection on the threshold τ (ref :35-92)."""

  @staticmethod
  def forward(ctx, inputs, alpha, axis, n_iter, ensure_sum_one):
    d = inputs.shape[axis]
    am1 = alpha - 1.0
    x = inputs * am1
    max_val = x.max(axis, keepdim=True).values
    tau_lo = max_val - 1.0
    tau_hi = max_val - (1.0 / d) ** am1
    f_lo = _EntmaxProb(x - tau_lo, am1).sum(axis, keepdim=True) - 1.0
    dm = tau_hi - tau_lo
    p_m = None
    for _ in range(n_iter):
      dm = dm / 2
      tau_m = tau_lo + dm
      p_m = _EntmaxProb(x - tau_m, am1)
      f_m = p_m.sum(axis, keepdim=True) - 1.0
      tau_lo = torch.where(f_m * f_lo > 0, tau_m, tau_lo)
    if ensure_sum_one:
      p_m = p_m / p_m.sum(axis, keepdim=True)
    ctx.save_for_backward(p_m)
    ctx.alpha, ctx.axis = alpha, axis
    return p_m

  @staticmethod
  def backward(ctx, dy):
    p_m, = ctx.saved_tensors
    gppr = torch.where(p_m > 0, p_m ** (2.0 - ctx.alpha), torch.zeros_like(p_m))
    dx = dy * gppr
    q = dx.sum(ctx.axis, keepdim=True) / gppr.sum(ctx.axis, keepdim=True)
    return dx - q * gppr, None, None, None, None


def entmax_support(inputs, alpha=1.5, axis=-1, n_iter=50, ensure_sum_one=True, dim=None):  # pylint: disable=invalid-name
  """α-entmax probabilities (sparse: entries below the threshold are exactly 0) (ref :35)."""
  axis = axis if dim is None else dim
  return _EntmaxBisect.apply(inputs, float(alpha), axis, int(n_iter), bool(ensure_sum_one))


class _EntmaxLoss(torch.autograd.Function):

  @staticmethod
  def forward(ctx, labels, inputs, alpha, n_iter, ensure_sum_one):
    p_star = _EntmaxBisect.forward(_Ctx(), inputs, alpha, -1, n_iter, ensure_sum_one)
    loss = (1.0 - (p_star ** alpha).sum(-1)) / (alpha * (alpha - 1.0))
    diff = p_star - labels.to(inputs.dtype)
    ctx.save_for_backward(diff)
    return loss + (diff * inputs).sum(-1)

  @staticmethod
  def backward(ctx, dy):
    diff, = ctx.saved_tensors
    g = dy.unsqueeze(-1) * diff
    return None, g, None, None, None


class _Ctx:
  def save_for_backward(self, *a):
    pass


def entmax_loss(labels, inputs, alpha=1.5, n_iter=50, ensure_sum_one=True):  # pylint: disable=invalid-name
  """Fenchel-Young loss of α-entmax (ref :95): `labels` are one-hot / probability targets
  `[..., V]`, `inputs` logits; d loss / d inputs = entmax(inputs) − labels."""
  assert labels.shape[0] == inputs.shape[0]
  return _EntmaxLoss.apply(labels, inputs, float(alpha), int(n_iter), bool(ensure_sum_one))
