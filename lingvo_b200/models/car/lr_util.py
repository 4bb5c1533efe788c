"""Epoch-based learning-rate helpers (ref `lingvo/tasks/car/lr_util.py`).

Car experiments specify schedules in epochs; these helpers convert to steps from the
training input's `num_samples` / batch size and install the schedule on `train_p`.
"""

from __future__ import annotations

from lingvo_b200.core import schedule


def _GetTrainingStatistics(train_input_p):
  """→ (steps_per_epoch, batch_size) (ref :25)."""
  batch = getattr(train_input_p, 'batch_size', 0) or 0
  if not batch and 'bucket_batch_limit' in train_input_p:
    batch = train_input_p.bucket_batch_limit[0]
  n = train_input_p.num_samples
  assert batch > 0 and n > 0, 'train input needs num_samples and a batch size'
  return max(1, n // batch), batch


def _GetSteps(steps_per_epoch, warmup_epoch, start_epoch, total_epoch):
  return (int(warmup_epoch * steps_per_epoch), int(start_epoch * steps_per_epoch),
          int(total_epoch * steps_per_epoch))


def SetExponentialLR(train_p, train_input_p, exp_start_epoch, total_epoch, warmup_epoch=0,
                     limit_epoch=None, multiplier_min=0.01, warmup_init=0.0):
  """Linear warm-up → constant → exponential decay to `multiplier_min` at
  `limit_epoch` (default total) (ref :52)."""
  spe, _ = _GetTrainingStatistics(train_input_p)
  warm, start, total = _GetSteps(spe, warmup_epoch, exp_start_epoch, total_epoch)
  limit = int((limit_epoch or total_epoch) * spe)
  assert start <= limit <= total or limit_epoch, (start, limit, total)
  train_p.max_steps = total
  train_p.lr_schedule = schedule.LinearRampupExponentialDecay.Params().Set(
      warmup=warm, decay_start=max(start, warm + 1), decay_end=limit, min=multiplier_min,
      warmup_init=warmup_init)
  return train_p


def SetCosineLR(train_p, train_input_p, total_epoch, warmup_epoch=0, warmup_init=0.0):
  """Linear warm-up then cosine decay to 0 at `total_epoch` (ref :112)."""
  spe, _ = _GetTrainingStatistics(train_input_p)
  warm, _, total = _GetSteps(spe, warmup_epoch, 0, total_epoch)
  train_p.max_steps = total
  train_p.lr_schedule = schedule.LinearRampupCosineSchedule.Params().Set(
      warmup_steps=warm, warmup_init=warmup_init, initial_value=1.0, final_value=0.0,
      total_steps=total)
  return train_p
