"""Runs one forward+backward of the relative-bias attention at the flagship shape (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lingvo_b200.ops import attention as A

dev = torch.device('cuda')
b, l, h, d = 8, 1024, 16, 128
bf = torch.bfloat16
q = (torch.randn(b, l, h, d, device=dev) * 0.3).to(bf).requires_grad_()
k = (torch.randn(b, l, h, d, device=dev) * 0.3).to(bf).requires_grad_()
v = torch.randn(b, l, h, d, device=dev).to(bf).requires_grad_()
rel = (torch.randn(h, 2 * l - 1, device=dev) * 0.5).requires_grad_()
mask = torch.triu(torch.ones(l, l, device=dev), 1).unsqueeze(0).expand(b, l, l) * -1e9
for _ in range(2):
  o = A.rel_bias_attention(q, k, v, rel, mask.contiguous(), 1.0, causal=True)
  o.backward(torch.randn_like(o))
torch.cuda.synchronize()
