"""Eager runners (ref `lingvo/eager_runners.py`).

The reference has a graph-mode runner family (`runners.py`) and a TF2-eager family that wraps
each job's step in `tf.function`. This framework is eager by construction, so both surfaces
map to the same job classes; the `tf.function` role (removing per-step Python and launch
overhead) is played by CUDA-graph capture of the whole train step inside `TrainEngine`
(`p.train.use_cuda_graph`, `--use_cuda_graph`).

  Trainer         → runners.Trainer          (ref :31)   train loop on the TrainEngine
  TrainSummaries  → runners.TrainSummaries   (ref :118)  follows checkpoints, writes the
                                                          training summaries without updating
  Evaler          → runners.Evaler           (ref :181)  eval metrics per checkpoint
  Decoder         → runners.Decoder          (ref :341)  decode + decoder metrics per checkpoint
"""
from lingvo_b200 import runners

Trainer = runners.Trainer
TrainSummaries = runners.TrainSummaries
Evaler = runners.Evaler
Decoder = runners.Decoder


def GetRunnerClass(job: str):
  """Job name (as used by `--job`) → runner class of the eager family."""
  table = {'trainer': Trainer, 'trainer_client': Trainer, 'train_summaries': TrainSummaries,
           'evaler': Evaler, 'decoder': Decoder}
  for prefix, cls in table.items():
    if job == prefix or job.startswith(prefix + '_'):
      return cls
  raise ValueError('No eager runner for job %r' % job)
