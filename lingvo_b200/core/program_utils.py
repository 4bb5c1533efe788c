"""Executor helpers (reference `core/program_utils.py:45-129`)."""

import csv
import os
from typing import Dict


class DecodeStatusCache:
  """Persists which checkpoints a decode program has already processed."""

  def __init__(self, program_dir: str):
    self.cache_file = os.path.join(program_dir, 'decoded_datasets.csv')
    self.cached_ckpts = set()
    if os.path.exists(self.cache_file):
      with open(self.cache_file) as f:
        for row in csv.reader(f):
          if row:
            self.cached_ckpts.add(row[0])

  def UpdateCkpt(self, ckpt_key: str):
    with open(self.cache_file, 'a') as f:
      f.write(ckpt_key + '\n')
    self.cached_ckpts.add(ckpt_key)

  def TryLoadCache(self, ckpt_key: str) -> bool:
    return ckpt_key in self.cached_ckpts


class TriggerScheduler:
  """Offset/interval triggering of eval/decode programs (reference :111)."""

  def __init__(self, offset: int, interval: int):
    self.offset = offset
    self.interval = interval
    self.count = 0

  def Trigger(self):
    self.count += 1

  def ShouldRun(self) -> bool:
    if self.interval <= 0:
      return False
    if self.count < self.offset:
      return False
    return (self.count - self.offset) % self.interval == 0

  def State(self) -> Dict[str, int]:
    return {'offset': self.offset, 'interval': self.interval, 'count': self.count}

  def SetState(self, state: Dict[str, int]):
    self.count = int(state.get('count', 0))


def SummaryToCsv(summaries: Dict[str, float]) -> str:
  return '\n'.join('%s,%s' % (k, v) for k, v in sorted(summaries.items()))


def CsvToSummary(text: str) -> Dict[str, float]:
  out = {}
  for row in text.splitlines():
    if ',' in row:
      k, v = row.rsplit(',', 1)
      out[k] = float(v)
  return out
