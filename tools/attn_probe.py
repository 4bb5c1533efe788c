"""Times SDPA backends on the flagship attention shape with an additive bias."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from torch.nn.attention import sdpa_kernel, SDPBackend

dev = torch.device('cuda'); bf = torch.bfloat16
B, H, L, D = 8, 16, 1024, 128
q = torch.randn(B, H, L, D, device=dev, dtype=bf, requires_grad=True)
k = torch.randn(B, H, L, D, device=dev, dtype=bf, requires_grad=True)
v = torch.randn(B, H, L, D, device=dev, dtype=bf, requires_grad=True)
bias0 = (torch.randn(1, H, L, L, device=dev) * 0.5)
causal = torch.tril(torch.ones(L, L, device=dev, dtype=torch.bool))
bias0 = bias0.masked_fill(~causal, -1e9).to(bf)
do = torch.randn(B, H, L, D, device=dev, dtype=bf)


def timeit(fn, n=10):
  for _ in range(3): fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(n): fn()
  e1.record(); torch.cuda.synchronize()
  return round(e0.elapsed_time(e1) / n * 1e3, 1)

res = {}
for name, be in [('efficient', SDPBackend.EFFICIENT_ATTENTION), ('cudnn', SDPBackend.CUDNN_ATTENTION),
                 ('math', SDPBackend.MATH)]:
  for bias_grad in (False, True):
    bias = bias0.clone().requires_grad_(bias_grad)
    def fb():
      with sdpa_kernel(be):
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=bias.expand(B, H, L, L), scale=1.0 / D ** 0.5)
      o.backward(do)
      q.grad = k.grad = v.grad = None; bias.grad = None
      return o
    try:
      res['%s_biasgrad%d_us' % (name, bias_grad)] = timeit(fb)
    except Exception as e:  # pylint: disable=broad-except
      res['%s_biasgrad%d_us' % (name, bias_grad)] = 'ERR ' + str(e)[:120]
# causal-only flash for reference
def fl():
  with sdpa_kernel(SDPBackend.FLASH_ATTENTION):
    o = F.scaled_dot_product_attention(q, k, v, is_causal=True, scale=1.0 / D ** 0.5)
  o.backward(do); q.grad = k.grad = v.grad = None
try:
  res['flash_causal_nobias_us'] = timeit(fl)
except Exception as e:
  res['flash_causal_nobias_us'] = 'ERR ' + str(e)[:100]
def cd():
  with sdpa_kernel(SDPBackend.CUDNN_ATTENTION):
    o = F.scaled_dot_product_attention(q, k, v, is_causal=True, scale=1.0 / D ** 0.5)
  o.backward(do); q.grad = k.grad = v.grad = None
try:
  res['cudnn_causal_nobias_us'] = timeit(cd)
except Exception as e:
  res['cudnn_causal_nobias_us'] = 'ERR ' + str(e)[:100]
print(json.dumps(res))
