"""One forward + backward of the flash attention at the benchmark shape (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lingvo_b200.ops import attention as A
b, l, h, d = 8, 1024, 16, 128
torch.manual_seed(0)
qkv = (torch.randn(b, l, 3 * h * d, device='cuda') * 0.5).bfloat16()
q, k, v = [t.reshape(b, l, h, d).detach().requires_grad_(True) for t in qkv.split(h * d, dim=-1)]
rel = torch.randn(h, 2 * l - 1, device='cuda').requires_grad_(True)
seg = torch.ones(b, l, dtype=torch.int32, device='cuda')
pos = torch.arange(l, dtype=torch.int32, device='cuda').unsqueeze(0).expand(b, l).contiguous()
d_o = (torch.randn(b, l, h, d, device='cuda') * 0.5).bfloat16()
for _ in range(2):
  o = A.flash_attention(q, k, v, rel, seg, pos, 1.0, True)
  torch.autograd.grad(o, [q, k, v, rel], d_o)
torch.cuda.synchronize()
