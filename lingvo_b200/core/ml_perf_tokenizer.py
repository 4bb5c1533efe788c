"""MLPerf sub-word tokenizer used when decoding the MLPerf Transformer
(ref `lingvo/core/ml_perf_tokenizer.py`).

The MLPerf translation benchmark ships a tensor2tensor-style sub-token vocabulary file
(quoted tokens, `_` marks the end of a word, `\\u`/`\\<ord>;` escapes). The reference only
needs *id → string* for this vocabulary — targets arrive pre-tokenised — and implements it
as the native `ml_perf_subword_id_to_string` op; here the native decoder is
`ops/csrc_host` `MlPerfSubword` (one C++ call per sequence, vocabulary cached per file),
reached through `ops/host_ops.MlPerfSubwordIdToString`.
"""

import numpy as np
import torch

from lingvo_b200.core import tokenizers
from lingvo_b200.ops import host_ops


class MlPerfTokenizer(tokenizers.BaseTokenizer):
  """Id → string only, for MLPerf decoding (ref :21)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('vocab_filepath', None, 'Path of the MLPerf sub-token vocabulary file.')
    return p

  def __init__(self, params):
    super().__init__(params)
    assert self.params.vocab_filepath, 'MlPerfTokenizer needs p.vocab_filepath'

  def IdsToStrings(self, ids, lens, languages=None):
    """ids `[B, T]`, lens `[B]` → list of B strings (only the first lens[b] ids count)."""
    del languages
    if isinstance(ids, torch.Tensor):
      ids = ids.detach().cpu().numpy()
    if isinstance(lens, torch.Tensor):
      lens = lens.detach().cpu().numpy()
    ids = np.asarray(ids)
    if ids.ndim == 1:
      ids = ids[None]
    return host_ops.MlPerfSubwordIdToString(ids, np.asarray(lens).reshape(-1),
                                            self.params.vocab_filepath)

  def _Decode(self, ids) -> str:
    return self.IdsToStrings(np.asarray(ids)[None], [len(ids)])[0]

  def _Encode(self, text):
    raise NotImplementedError('The MLPerf tokenizer only converts ids to strings '
                              '(inputs are pre-tokenised), as in the reference.')
