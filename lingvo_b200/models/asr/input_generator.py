"""ASR input (ref `lingvo/tasks/asr/input_generator.py:24`): tf.Example records with
`uttid`, `transcript`, `frames` (flattened `[T·F]` log-mel floats); the transcript is
tokenised on the fly; bucket key = number of frames."""

from __future__ import annotations

import numpy as np

from lingvo_b200.core import base_input_generator
from lingvo_b200.core.nested_map import NestedMap
from lingvo_b200.utils import tf_example


class AsrInput(base_input_generator.BaseSequenceInputGenerator):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('frame_size', 40, 'Feature dims per frame.')
    p.Define('append_eos_frame', True, 'Append an all-zero EOS frame.')
    p.source_max_length = 3000
    return p

  def ProcessRecord(self, record, source_id=0):
    p = self.params
    f = tf_example.ParseExample(record)
    frames = f['frames'].astype(np.float32).reshape(-1, p.frame_size)
    if p.append_eos_frame:
      frames = np.concatenate([frames, np.zeros((1, p.frame_size), np.float32)], 0)
    n = frames.shape[0]
    if n > p.bucket_upper_bound[-1]:
      return None
    text = f['transcript'][0]
    text = text.decode('utf-8') if isinstance(text, bytes) else str(text)
    ids, labels, pads = self.StringsToIds([text])
    nt = int((1 - pads[0]).sum())
    out = NestedMap(
        src=NestedMap(src_inputs=frames[:, :, None], mask=np.ones(n, np.float32)),
        tgt=NestedMap(ids=ids[0, :nt].numpy().astype(np.int32),
                      labels=labels[0, :nt].numpy().astype(np.int32),
                      weights=np.ones(nt, np.float32), mask=np.ones(nt, np.float32)))
    return out, n

  def _PreprocessInputBatch(self, batch):
    batch.src.paddings = 1.0 - batch.src.mask
    batch.tgt.paddings = 1.0 - batch.tgt.mask
    return batch
