"""The `bench.py --impl reference` comparator computes the same model as the framework.

The stock arm (`baseline/stock_moe_lm.py`) imports nothing from lingvo_b200; here its loss,
and one unfused-Adafactor step, are checked against `UniTransformer` (MoELm8ETiny) with
the same weights — so the headline ratio compares two implementations of one function.
"""

import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'baseline'))
import stock_moe_lm  # noqa: E402

from lingvo_b200 import model_registry  # noqa: E402
from lingvo_b200.core import cluster_factory  # noqa: E402
import lingvo_b200.models.lm.params.synthetic_packed_input  # noqa: F401,E402


class TinyCfg(stock_moe_lm.Config):
  vocab = 256
  model_dim = 64
  heads = 4
  head_dim = 16
  ff_dim = 128
  moe_hidden = 128
  experts = 8
  layers = 2
  seq_len = 64
  batch = 4
  compute_dtype = torch.float32


def _Ours(min_dim_to_factor=128):
  cfg = model_registry.GetParams('lm.synthetic_packed_input.MoELm8ETiny', 'Train')
  cfg.task.train.optimizer.min_dim_size_to_factor = min_dim_to_factor
  cfg.task.fprop_dtype = torch.float32
  cfg.task.builder.fprop_dtype = torch.float32
  with cluster_factory.Cluster(cfg.cluster):
    model = cfg.Instantiate()
  return model.tasks[0]


def _CopyWeights(task, stock):
  P = stock.params
  name = {}
  for i, kind in enumerate(stock.kinds):
    pre = 'transformer/decoder/layer_%03d/' % i
    name[pre + 'ln/scale/var'] = 'l%d/ln' % i
    if kind == 'attn':
      for w in ('wq', 'wk', 'wv', 'wo', 'wrb'):
        name[pre + 'dec_self_attention/%s/var' % w] = 'l%d/%s' % (i, w)
    elif kind == 'ffw':
      name[pre + 'dense_relu_dense/wi/var'] = 'l%d/wi' % i
      name[pre + 'dense_relu_dense/wo/var'] = 'l%d/wo' % i
    else:
      name[pre + 'moe/gw/var'] = 'l%d/gw' % i
      name[pre + 'moe/wi/var'] = 'l%d/moe_wi' % i
      name[pre + 'moe/wo/var'] = 'l%d/moe_wo' % i
  name['transformer/decoder/final_layer_norm/scale/var'] = 'final_ln'
  name['transformer/dec_emb/embedding/var'] = 'emb'
  with torch.no_grad():
    for v in task.vars.Flatten():
      v.data.copy_(P[name[v.var_name]])
  return name


@pytest.mark.parametrize('min_dim_to_factor', [128, 32])
def test_stock_arm_matches_framework_loss_and_update(min_dim_to_factor):
  """128: tiny dims ⇒ non-factored second moments; 32: the factored path (what the real
  2048/8192-wide matrices use)."""
  torch.manual_seed(0)
  cfg = TinyCfg()
  cfg.min_dim_size_to_factor = min_dim_to_factor
  stock = stock_moe_lm.StockMoeLm(cfg, torch.device('cpu'))
  task = _Ours(min_dim_to_factor)
  name = _CopyWeights(task, stock)
  hb = stock_moe_lm.SyntheticBatch(cfg, 0, 0, pin=False)
  loss_stock = stock.Loss(hb['ids'], hb['labels'], hb['segment_ids'], hb['segment_pos'])
  from lingvo_b200.core.nested_map import NestedMap
  batch = NestedMap(tgt=NestedMap(ids=hb['ids'], labels=hb['labels'],
                                  segment_ids=hb['segment_ids'],
                                  segment_pos=hb['segment_pos']))
  metrics, _ = task.FPropDefaultTheta(batch)
  loss_ours = metrics['loss'][0]
  assert abs(float(loss_stock) - float(loss_ours)) < 2e-4 * abs(float(loss_ours)), (
      float(loss_stock), float(loss_ours))
  # one optimizer step each: the updated weights agree
  task.BProp()
  stock.TrainStep(hb)
  worst = 0.0
  for v in task.vars.Flatten():
    a, b = v.data, stock.params[name[v.var_name]].data
    err = float((a - b).abs().max() / (b.abs().max() + 1e-12))
    worst = max(worst, err)
  assert worst < 1e-5, worst


def test_stock_arm_trains():
  cfg = TinyCfg()
  stock = stock_moe_lm.StockMoeLm(cfg, torch.device('cpu'))
  hb = stock_moe_lm.SyntheticBatch(cfg, 0, 0, pin=False)
  cfg.warmup_steps = 100           # lr 0.1 so a few steps visibly reduce the loss
  losses = [float(stock.TrainStep(hb)) for _ in range(6)]
  assert all(l == l for l in losses)
  assert losses[-1] < losses[0]


def test_stock_arm_does_not_import_framework():
  src = open(os.path.join(os.path.dirname(__file__), '..', 'baseline',
                          'stock_moe_lm.py')).read()
  assert 'import lingvo_b200' not in src and 'from lingvo_b200' not in src
