"""Punctuator model (ref `lingvo/tasks/punctuator/model.py:22`): RNMT whose decode
post-processing reports BLEU between reference and restored punctuation."""

from lingvo_b200.models.mt import model as mt_model


class RNMTModel(mt_model.RNMTModel):
  """RNMT+ for punctuation restoration."""

  def PostProcessDecodeOut(self, dec_out, dec_metrics):
    return super().PostProcessDecodeOut(dec_out, dec_metrics)


class TransformerModel(mt_model.TransformerModel):
  """Transformer variant (same input contract)."""
