"""Offline batch prediction over the newest checkpoints
(ref `lingvo/core/predictor_runner_base.py`): watches `checkpoint_dir`, and for each
new checkpoint runs `InputGenerator` → `predictor.Run` → `OutputWriter`."""

from __future__ import annotations

import os
import time

from absl import logging

from lingvo_b200.core import saver as saver_lib


class PredictorRunnerBase:

  def __init__(self, checkpoint_dir, output_dir, inference_graph_filename, predictor,
               subgraph_name='default', batch_size=32, max_inputs=0, watch=False,
               poll_secs=30):
    self._ckpt_dir, self._out_dir = checkpoint_dir, output_dir
    self._graph = inference_graph_filename
    self._pred = predictor
    self._subgraph = subgraph_name
    self._batch_size, self._max_inputs = batch_size, max_inputs
    self._watch, self._poll = watch, poll_secs
    os.makedirs(output_dir, exist_ok=True)

  # -- subclass API ---------------------------------------------------------------
  def InputGenerator(self):
    """Yields feed dicts."""
    raise NotImplementedError

  def RunBatch(self, output_dir, batch):
    """Runs the predictor on `batch`; returns the outputs."""
    return self._pred.Run(None, subgraph_name=self._subgraph, **batch)

  def OutputWriter(self, output_dir, outputs):
    """Persists one batch of outputs."""
    raise NotImplementedError

  # -- driver -----------------------------------------------------------------------
  def _Done(self, step):
    return os.path.exists(os.path.join(self._out_dir, 'step_%08d' % step, 'DONE'))

  def _PredictOneCheckpoint(self, path, step):
    out_dir = os.path.join(self._out_dir, 'step_%08d' % step)
    os.makedirs(out_dir, exist_ok=True)
    self._pred.Load(path)
    n = 0
    for batch in self.InputGenerator():
      self.OutputWriter(out_dir, self.RunBatch(out_dir, batch))
      n += 1
      if self._max_inputs and n * self._batch_size >= self._max_inputs:
        break
    with open(os.path.join(out_dir, 'DONE'), 'w') as f:
      f.write('%d batches\n' % n)
    logging.info('Predicted %d batches for step %d', n, step)

  def Run(self):
    seen = set()
    while True:
      latest = saver_lib.LatestCheckpoint(self._ckpt_dir)
      if latest and latest not in seen:
        step = int(latest.rsplit('-', 1)[-1]) if '-' in os.path.basename(latest) else 0
        if not self._Done(step):
          self._PredictOneCheckpoint(latest, step)
        seen.add(latest)
      if not self._watch:
        return
      time.sleep(self._poll)
