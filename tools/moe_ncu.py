"""One forward + backward of the fused MoE layer at the benchmark shape on ONE GPU (driver for
`ncu`): launches `moe_gate_dispatch_kernel`, the grouped expert GEMMs with the row-pointer
epilogue, `moe_combine_kernel`, `moe_scatter_rows_kernel`, `moe_gather_rows_kernel`,
`moe_combine_bwd_gate_kernel`. On one GPU the "peers" are this GPU's own buffers, so the
kernels' instruction mix and HBM behaviour are what ncu sees; NVLink time is measured
separately (bench `exposed_a2a_ms_per_step`, tools/allreduce_bench.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lingvo_b200.core import gshard_builder as gb
from lingvo_b200.core import cluster_factory

torch.manual_seed(0)
dev = torch.device('cuda')
b = gb.MoEBuilder.Params().Set(model_dim=2048, attention_num_heads=16, attention_key_value_dim=128,
                               moe_hidden_dim=8192, e_dim=8, c_dim=0, capacity_factor=2.0,
                               num_groups=8, dtype=torch.float32, fprop_dtype=torch.bfloat16)
with cluster_factory.ForTestingWorker(mode='sync', job='trainer_client', gpus=1):
  p = gb.MoELayer.Params().Set(name='moe', b=b, dtype=torch.float32, fprop_dtype=torch.bfloat16)
  layer = p.Instantiate()
  layer.InstantiateVariables()
  layer.to(dev)
theta = layer.theta.Transform(lambda t: t.to(dev))
x = (torch.randn(8, 1024, 2048, device=dev) * 0.5).bfloat16().requires_grad_(True)
seg = torch.ones(8, 1024, dtype=torch.long, device=dev)
th = theta.Transform(lambda t: t.bfloat16() if t.dim() >= 3 else t)
for _ in range(3):
  y, aux = layer.FProp(th, x, seg)
  (y.float().sum() + aux).backward()
torch.cuda.synchronize()
print('MOE_NCU_DRIVER_OK', tuple(y.shape))
