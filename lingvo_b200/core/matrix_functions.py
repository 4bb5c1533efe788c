"""Matrix functions for Shampoo (reference `core/matrix_functions.py:1-174`).

`inlined_matrix_inverse_pth_root`: coupled Newton iteration for A^{-1/p} with
ridge regularisation scaled by the max eigenvalue (power iteration).
"""

import torch


def matrix_square_root(mat_a, mat_a_size=None, iter_count=100, ridge_epsilon=1e-4):
  """Newton–Schulz iteration for the matrix square root."""
  n = mat_a.shape[0]
  ident = torch.eye(n, dtype=mat_a.dtype, device=mat_a.device)
  a = mat_a + ridge_epsilon * ident
  norm = a.norm()
  y = a / norm
  z = ident.clone()
  for _ in range(iter_count):
    t = 0.5 * (3.0 * ident - z @ y)
    y_new = y @ t
    z = t @ z
    if (y_new - y).abs().max() < 1e-7:
      y = y_new
      break
    y = y_new
  return y * norm.sqrt()


def _max_eigen(mat, iters=50):
  v = torch.ones(mat.shape[0], dtype=mat.dtype, device=mat.device)
  v = v / v.norm()
  for _ in range(iters):
    w = mat @ v
    n = w.norm()
    if n == 0:
      return torch.zeros((), dtype=mat.dtype, device=mat.device)
    v = w / n
  return v @ (mat @ v)


def inlined_matrix_inverse_pth_root(mat_g, p, mat_g_size=None, iter_count=100,
                                    epsilon=1e-6, ridge_epsilon=1e-6):
  """A^{-1/p} via the coupled iteration of the reference (:88-174)."""
  mat_g = mat_g.float()
  n = mat_g.shape[0]
  ident = torch.eye(n, dtype=mat_g.dtype, device=mat_g.device)
  max_ev = _max_eigen(mat_g)
  ridge = ridge_epsilon * torch.clamp(max_ev, min=1e-16)
  damped = mat_g + ridge * ident
  alpha = -1.0 / p
  z = (1 + p) / (2 * damped.norm())
  mat_m = damped * z
  mat_h = ident * (z**(1.0 / p))
  err = (mat_m - ident).abs().max()
  for _ in range(iter_count):
    if err <= epsilon:
      break
    m_i = (1 - alpha) * ident + alpha * mat_m
    new_m = torch.linalg.matrix_power(m_i, p) @ mat_m
    new_h = mat_h @ m_i
    new_err = (new_m - ident).abs().max()
    if new_err > err * 1.2:   # diverging: keep the previous iterate
      break
    mat_m, mat_h, err = new_m, new_h, new_err
  return mat_h


def inverse_pth_root_no_sync(mat_g, p, iter_count=40, epsilon=1e-6, ridge_epsilon=1e-6):
  """Same coupled iteration with a fixed trip count and device-side convergence masking, so
  no iteration reads a value back to the host — safe to enqueue on a side stream (used by
  `preconditioner_captain`). Once converged (or diverging) the iterate is frozen by `where`."""
  mat_g = mat_g.float()
  n = mat_g.shape[0]
  ident = torch.eye(n, dtype=mat_g.dtype, device=mat_g.device)
  v = torch.ones(n, dtype=mat_g.dtype, device=mat_g.device) / (n ** 0.5)
  for _ in range(30):                                   # power iteration, no early exit
    w = mat_g @ v
    v = w / torch.clamp(w.norm(), min=1e-30)
  max_ev = v @ (mat_g @ v)
  damped = mat_g + ridge_epsilon * torch.clamp(max_ev, min=1e-16) * ident
  alpha = -1.0 / p
  z = (1 + p) / (2 * damped.norm())
  mat_m = damped * z
  mat_h = ident * (z ** (1.0 / p))
  err = (mat_m - ident).abs().max()
  live = torch.ones((), dtype=torch.bool, device=mat_g.device)
  for _ in range(iter_count):
    m_i = (1 - alpha) * ident + alpha * mat_m
    new_m = torch.linalg.matrix_power(m_i, p) @ mat_m
    new_h = mat_h @ m_i
    new_err = (new_m - ident).abs().max()
    live = live & (err > epsilon) & (new_err <= err * 1.2)
    mat_m = torch.where(live, new_m, mat_m)
    mat_h = torch.where(live, new_h, mat_h)
    err = torch.where(live, new_err, err)
  return mat_h
