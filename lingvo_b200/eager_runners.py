"""Eager runners (ref `lingvo/eager_runners.py`).

The reference has a graph-mode runner family (`runners.py`) and a TF2-eager family
that wraps the train step in `tf.function`. This framework is eager by
construction, so the eager runners are the regular runners; the `tf.function`
role (removing per-step Python/launch overhead) is played by CUDA-graph capture
of the train step (`Trainer.Params().use_cuda_graph`).
"""
from lingvo_b200 import runners

Trainer = runners.Trainer
TrainSummaries = getattr(runners, 'TrainSummaries', runners.Controller)
Evaler = runners.Evaler
Decoder = runners.Decoder
