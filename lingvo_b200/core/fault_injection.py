"""Fault injection for the recovery paths (SURVEY §5.3 — the reference has none).

The recovery model is restart-from-checkpoint plus numeric step skipping; these hooks make
both testable on demand, in unit tests and on a live job:

  LINGVO_B200_FAULTS="transient@5,nan_grad@7,transient@12x2,exit@20:rank=1,drop_ckpt@30"

  kind        effect at global step N
  ----------  -------------------------------------------------------------------
  transient   raises ConnectionError in the trainer loop → `_RunLoop` retries, the
              loop restores the latest checkpoint and continues
  fatal       raises ValueError → the job fails fast (no retry)
  nan_grad    poisons one gradient with NaN → the learner's NaN/Inf guard must zero
              the step (`grad_scale = 0`)
  inf_loss    multiplies the loss by +inf before backward
  exit        `os._exit(17)`: a hard rank death, for restart-by-scheduler tests
  drop_ckpt   deletes the newest checkpoint's data shard → restore must fall back
  stall       sleeps `seconds` (default 5) — heartbeat / timeout tests

`xK` repeats a fault K times (once per retry of that step), `:rank=R` limits it to one rank.
Every firing is logged and counted in `Injector.fired`.
"""

from __future__ import annotations

import glob
import logging
import os
import re
import time
from typing import Dict, List, Optional

import torch

_SPEC_RE = re.compile(
    r'^(?P<kind>[a-z_]+)@(?P<step>\d+)(?:x(?P<times>\d+))?(?::(?P<opts>[a-z_0-9=.,:]+))?$')
KINDS = ('transient', 'fatal', 'nan_grad', 'inf_loss', 'exit', 'drop_ckpt', 'stall')


class Fault:

  def __init__(self, kind: str, step: int, times: int = 1, **opts):
    assert kind in KINDS, 'unknown fault kind %r (have %s)' % (kind, KINDS)
    self.kind, self.step, self.remaining, self.opts = kind, int(step), int(times), opts

  def __repr__(self):
    return 'Fault(%s@%d x%d %s)' % (self.kind, self.step, self.remaining, self.opts)


def ParseSpec(spec: str) -> List[Fault]:
  faults = []
  for item in filter(None, (s.strip() for s in (spec or '').split(','))):
    m = _SPEC_RE.match(item)
    if not m:
      raise ValueError('bad fault spec %r (want kind@step[xN][:k=v])' % item)
    opts = {}
    for kv in filter(None, (m.group('opts') or '').split(':')):
      k, _, v = kv.partition('=')
      opts[k] = float(v) if re.fullmatch(r'[0-9.]+', v) and '.' in v else (
          int(v) if v.isdigit() else v)
    faults.append(Fault(m.group('kind'), int(m.group('step')), int(m.group('times') or 1), **opts))
  return faults


class Injector:
  """Holds the armed faults of this process; the hooks below consult it."""

  def __init__(self, faults: Optional[List[Fault]] = None, rank: Optional[int] = None):
    self.faults = list(faults or [])
    self.rank = int(os.environ.get('RANK', '0')) if rank is None else rank
    self.fired: Dict[str, int] = {}

  def _Take(self, kind: str, step: int) -> Optional[Fault]:
    for f in self.faults:
      if f.kind == kind and f.step == step and f.remaining > 0 and (
          'rank' not in f.opts or int(f.opts['rank']) == self.rank):
        f.remaining -= 1
        self.fired[kind] = self.fired.get(kind, 0) + 1
        logging.warning('[fault-injection] firing %s at step %d (rank %d)', kind, step, self.rank)
        return f
    return None

  # ---- hook: trainer loop, before the step ----
  def BeforeStep(self, step: int, train_dir: Optional[str] = None):
    f = self._Take('stall', step)
    if f:
      time.sleep(float(f.opts.get('seconds', 5)))
    if self._Take('drop_ckpt', step) and train_dir:
      shards = sorted(glob.glob(os.path.join(train_dir, 'ckpt-*.data-*')))
      if shards:
        newest = shards[-1].split('.data-')[0]
        for path in glob.glob(newest + '.data-*'):
          os.remove(path)
        logging.warning('[fault-injection] removed data shards of %s', newest)
    if self._Take('exit', step):
      os._exit(17)  # pylint: disable=protected-access
    if self._Take('fatal', step):
      raise ValueError('[fault-injection] fatal fault at step %d' % step)
    if self._Take('transient', step):
      raise ConnectionError('[fault-injection] transient fault at step %d' % step)

  # ---- hook: learner, on the loss and on the gradients ----
  def OnLoss(self, step: int, loss: torch.Tensor) -> torch.Tensor:
    if self._Take('inf_loss', step):
      return loss * float('inf')
    return loss

  def OnGradients(self, step: int, var_grads):
    if not self._Take('nan_grad', step):
      return var_grads
    from lingvo_b200.core import py_utils  # pylint: disable=g-import-not-at-top
    leaves = [vg for vg in var_grads.Flatten() if isinstance(vg, py_utils.VarGrad)]
    if leaves:
      g = leaves[0].grad
      g.reshape(-1)[0] = float('nan')
    return var_grads


_INJECTOR: Optional[Injector] = None


def Get() -> Optional[Injector]:
  """The process-wide injector (armed from LINGVO_B200_FAULTS on first use), or None."""
  global _INJECTOR
  if _INJECTOR is None:
    spec = os.environ.get('LINGVO_B200_FAULTS', '')
    if not spec:
      return None
    _INJECTOR = Injector(ParseSpec(spec))
  return _INJECTOR


def Arm(spec_or_faults, rank: Optional[int] = None) -> Injector:
  global _INJECTOR
  faults = ParseSpec(spec_or_faults) if isinstance(spec_or_faults, str) else list(spec_or_faults)
  _INJECTOR = Injector(faults, rank)
  return _INJECTOR


def Disarm():
  global _INJECTOR
  _INJECTOR = None
