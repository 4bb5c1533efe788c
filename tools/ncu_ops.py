"""Launches the hot kernels once each at flagship shapes (driver for `ncu --set full`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lingvo_b200.ops import gate, gemm as G, norm

dev = torch.device('cuda')
bf = torch.bfloat16
t, m, hdim, e = 8192, 2048, 8192, 8
x = torch.randn(t, m, device=dev).to(bf).requires_grad_(True)
w_out = torch.randn(hdim, m, device=dev).to(bf)           # ffn out: [T, H] x [H, M]
h = torch.randn(t, hdim, device=dev).to(bf)
wo = torch.randn(m, m, device=dev).to(bf)
scale = torch.ones(m, device=dev, requires_grad=True)
gw = (torch.randn(m, e, device=dev) * 0.05).to(bf).requires_grad_(True)
for _ in range(3):
  G.gemm(h, w_out, True, False)                            # K = 8192, 512 tiles
  G.gemm(x.detach(), wo, True, False)                      # K = 2048
  G.gemm(h, x.detach(), False, False)                      # wgrad: [H, M] = hᵀ · x
  y = norm.rms_norm(x, scale, 1e-6)
  torch.autograd.grad(y, [x, scale], torch.randn_like(y))
  yl = gate.gate_logits(x, gw)
  torch.autograd.grad(yl, [x, gw], torch.randn_like(yl))
torch.cuda.synchronize()
