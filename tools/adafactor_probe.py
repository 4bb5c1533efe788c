"""Times the factored Adafactor phases on the flagship's variable shapes (1 GPU).

  python tools/adafactor_probe.py            # CUDA-event timings + effective GB/s
  python tools/adafactor_probe.py ncu        # one pass of each kernel for an ncu capture
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from lingvo_b200 import ops
from lingvo_b200.ops import optim


def main():
  ncu = len(sys.argv) > 1 and sys.argv[1] == 'ncu'
  dev = torch.device('cuda:0')
  C = ops.native()
  shapes = [(8, 2048, 8192), (8, 8192, 2048), (1, 2048, 2048), (1, 32000, 2048), (1, 2048, 8192)]
  if ncu:
    shapes = shapes[:1]
  out = []
  for (B, R, Cc) in shapes:
    w = torch.randn(B, R, Cc, device=dev)
    g = torch.randn(B, R, Cc, device=dev, dtype=torch.bfloat16) * 1e-2
    wb = w.to(torch.bfloat16)
    vr_rows = Cc >= R
    vr = torch.zeros(B, R if vr_rows else Cc, device=dev)
    vc = torch.zeros(B, Cc if vr_rows else R, device=dev)
    br4, bc4 = (B * R + 3) // 4 * 4, (B * Cc + 3) // 4 * 4
    scratch = torch.zeros(4 + 2 * br4 + 2 * bc4, device=dev)
    tot = torch.zeros(1, device=dev)
    gs = torch.ones((), device=dev)

    def stats():
      C.adafactor_stats(w, g, scratch, B, R, Cc, True, False, tot)

    def update():
      C.adafactor_update(w, g, vr, vc, scratch, wb, B, R, Cc, vr_rows, 1e-3, 0.99, 1e-30, 1e-3,
                         1.0, True, gs, False, None)

    C.adafactor_stats(w, g, scratch, B, R, Cc, True, True, tot)
    update()
    torch.cuda.synchronize()
    if ncu:
      stats(); update(); torch.cuda.synchronize()
      return
    rec = {'shape': [B, R, Cc]}
    n = B * R * Cc
    for name, fn, bytes_ in (('stats', stats, 2 * n), ('update(factors+rms+apply)', update, (2 + 2 + 4 + 4 + 2) * n)):
      for _ in range(3):
        fn()
      ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
      ev[0].record()
      for _ in range(10):
        fn()
      ev[1].record()
      torch.cuda.synchronize()
      ms = ev[0].elapsed_time(ev[1]) / 10
      rec[name] = {'us': round(ms * 1e3, 1), 'GBps': round(bytes_ / ms / 1e6, 0)}
    out.append(rec)
    print(json.dumps(rec), flush=True)
  with open('gpurun_out/adafactor_probe.json', 'w') as f:
    json.dump(out, f, indent=1)


if __name__ == '__main__':
  main()
