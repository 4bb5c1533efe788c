"""Magnitude pruning (ref `lingvo/core/pruning_utils.py`; the reference delegates to
`tensorflow_model_optimization`).

`PruningOp.ApplyPruning(hparams, layer, name, weight)` registers a mask for `weight`;
`UpdateMasks(step)` recomputes masks on the polynomial sparsity schedule
  s(t) = s_f + (s_i − s_f)(1 − (t − t0)/(t1 − t0))³
and `MaskedWeight` returns w ⊙ mask. Block sparsity `[bh, bw]` prunes whole blocks."""
import torch


def UsePruningInterface(pruning_hparams_dict):
  return bool(pruning_hparams_dict)


def ApplyCompression(params):
  return bool(getattr(params, 'pruning_hparams_dict', None))


class PruningHParams:

  def __init__(self, **kw):
    self.begin_pruning_step = kw.get('begin_pruning_step', 0)
    self.end_pruning_step = kw.get('end_pruning_step', -1)
    self.initial_sparsity = kw.get('initial_sparsity', 0.0)
    self.target_sparsity = kw.get('target_sparsity', 0.5)
    self.pruning_frequency = kw.get('pruning_frequency', 10)
    self.sparsity_function_exponent = kw.get('sparsity_function_exponent', 3.0)
    self.block_height = kw.get('block_height', 1)
    self.block_width = kw.get('block_width', 1)


class PruningOp:
  _masks = {}
  _weights = {}
  _hparams = None

  @classmethod
  def Reset(cls):
    cls._masks, cls._weights, cls._hparams = {}, {}, None

  @classmethod
  def Setup(cls, pruning_hparams_dict, global_step=None):
    del global_step
    cls._hparams = PruningHParams(**(pruning_hparams_dict or {}))
    return cls._hparams

  @classmethod
  def ApplyPruning(cls, pruning_hparams_dict, lstmobj, weight_name, wm_pc, dtype, scope=None):
    """Registers `lstmobj.vars[weight_name]`; returns the mask tensor."""
    del wm_pc, dtype, scope
    if cls._hparams is None:
      cls.Setup(pruning_hparams_dict)
    w = lstmobj.vars[weight_name]
    key = w.var_name
    cls._weights[key] = w
    cls._masks[key] = torch.ones_like(w.data)
    return cls._masks[key]

  @classmethod
  def Sparsity(cls, step):
    h = cls._hparams
    if step < h.begin_pruning_step:
      return 0.0
    end = h.end_pruning_step if h.end_pruning_step > 0 else h.begin_pruning_step + 1
    frac = min(max((step - h.begin_pruning_step) / max(end - h.begin_pruning_step, 1), 0.0), 1.0)
    return h.target_sparsity + (h.initial_sparsity - h.target_sparsity) * (
        (1.0 - frac) ** h.sparsity_function_exponent)

  @classmethod
  def UpdateMasks(cls, step):
    """Recomputes every mask for `step`'s target sparsity; returns the sparsity used."""
    h = cls._hparams
    s = cls.Sparsity(step)
    for key, w in cls._weights.items():
      mag = w.data.abs().float()
      if (h.block_height > 1 or h.block_width > 1) and mag.dim() == 2:
        bh, bw = h.block_height, h.block_width
        r, c = mag.shape
        pr, pc = -(-r // bh) * bh - r, -(-c // bw) * bw - c
        m = torch.nn.functional.pad(mag, (0, pc, 0, pr))
        blocks = m.reshape(m.shape[0] // bh, bh, m.shape[1] // bw, bw).mean((1, 3))
        k = int(s * blocks.numel())
        thr = blocks.flatten().kthvalue(k).values if k > 0 else -1.0
        bm = (blocks > thr).float()
        mask = bm.repeat_interleave(bh, 0).repeat_interleave(bw, 1)[:r, :c]
      else:
        k = int(s * mag.numel())
        thr = mag.flatten().kthvalue(k).values if k > 0 else -1.0
        mask = (mag > thr).float()
      cls._masks[key].copy_(mask.to(cls._masks[key].dtype))
    return s

  @classmethod
  def MaskedWeight(cls, w):
    m = cls._masks.get(getattr(w, 'var_name', None))
    return w if m is None else w * m

  @classmethod
  def GetMixResult(cls, theta, concat, lstmobj):
    return torch.matmul(concat, cls.MaskedWeight(lstmobj.vars.wm).to(concat.dtype))

  @classmethod
  def ApplyTensorflowUpdate(cls, *a, **k):
    return None
