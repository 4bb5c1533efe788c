"""MLPerf sub-word tokenizer wrapper (ref `lingvo/core/ml_perf_tokenizer.py`)."""
from lingvo_b200.core import tokenizers


class MlPerfTokenizer(tokenizers.WpmTokenizer):
  """The MLPerf Transformer sub-token vocabulary is a longest-match word-piece table."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('vocab_filepath_override', None, 'Kept for parity.')
    return p
