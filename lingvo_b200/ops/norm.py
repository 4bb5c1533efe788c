"""Fused RMS-norm / LayerNorm (csrc/norm_kernels.cu) with autograd."""

import torch

from lingvo_b200 import ops


def available() -> bool:
  mod = ops.native(required=False)
  return mod is not None and hasattr(mod, '_has_norm')


class _NormFn(torch.autograd.Function):

  @staticmethod
  def forward(ctx, x, scale, bias, eps, center):
    shape = x.shape
    x2 = x.contiguous().reshape(-1, shape[-1])
    y, stats, _ = ops.native().norm_fwd(x2, None, scale, bias, eps, center, False)
    ctx.save_for_backward(x2, scale if scale is not None else torch.empty(0),
                          stats)
    ctx.center = center
    ctx.has_scale = scale is not None
    ctx.has_bias = bias is not None
    ctx.scale_dtype = scale.dtype if scale is not None else None
    ctx.bias_dtype = bias.dtype if bias is not None else None
    return y.reshape(shape)

  @staticmethod
  def backward(ctx, dy):
    x2, scale, stats = ctx.saved_tensors
    shape = dy.shape
    dy2 = dy.contiguous().reshape(-1, shape[-1])
    dx, ds, db = ops.native().norm_bwd(
        x2, dy2, None, scale if ctx.has_scale else None, stats, ctx.center,
        ctx.has_scale and ctx.needs_input_grad[1],
        ctx.has_bias and ctx.needs_input_grad[2])
    if ds is not None and ctx.scale_dtype is not None:
      ds = ds.to(ctx.scale_dtype)
    if db is not None and ctx.bias_dtype is not None:
      db = db.to(ctx.bias_dtype)
    return dx.reshape(shape), ds, db, None, None


def rms_norm(x, scale, eps):
  return _NormFn.apply(x, scale, None, float(eps), False)


def layer_norm(x, scale, bias, eps, center=True):
  if x.dtype != torch.bfloat16:
    x = x.to(torch.bfloat16)
  return _NormFn.apply(x, scale, bias if center else None, float(eps), center)


def rms_norm_ref(x, scale, eps):
  xf = x.float()
  y = xf * torch.rsqrt(xf.square().mean(-1, keepdim=True) + eps)
  if scale is not None:
    y = y * scale.float()
  return y.to(x.dtype)
