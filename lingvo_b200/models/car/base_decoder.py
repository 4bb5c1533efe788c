"""Output decoders: dataset-specific post-processing + metrics of a detector (ref
`lingvo/tasks/car/base_decoder.py`)."""

from __future__ import annotations

import torch

from lingvo_b200.core import base_layer


class BaseDecoder(base_layer.BaseLayer):
  """ref :23."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('laser_sampling_rate', 0.05, 'Fraction of laser points kept for visualisation.')
    p.Define('ap_metric', None, 'AP metric params.')
    p.Define('draw_visualizations', False, 'Render top-down / camera visualisations.')
    p.name = 'decoder'
    return p

  def _SampleLaserForVisualization(self, points_xyz, points_padding):
    """Keeps a random `laser_sampling_rate` fraction of the points `[B,P,3]`, padded to a
    static count (ref :43)."""
    p = self.params
    b, n, _ = points_xyz.shape
    k = max(1, int(n * p.laser_sampling_rate))
    noise = torch.rand(b, n, device=points_xyz.device).masked_fill(points_padding > 0.5, -1.0)
    idx = noise.topk(k, -1).indices
    pts = points_xyz.gather(1, idx.unsqueeze(-1).expand(-1, -1, 3))
    pad = points_padding.gather(1, idx)
    return pts, pad

  def SaveTensors(self, tensor_map):
    """Hook to stash intermediate tensors for debugging (ref :70)."""
    self._saved_tensors = dict(tensor_map)

  def CreateDecoderMetrics(self):
    raise NotImplementedError()

  def ProcessOutputs(self, input_batch, model_outputs):
    raise NotImplementedError()

  def PostProcessDecodeOut(self, dec_out_dict, dec_metrics_dict):
    raise NotImplementedError()

  def DecodeFinalize(self, decode_finalize_args):
    """Called once per decode run with (out_path, decode_out list)."""
