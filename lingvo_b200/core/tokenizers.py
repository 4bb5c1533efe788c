"""Tokenizers (ref `lingvo/core/tokenizers.py`, native `tokenizer_ops_kernels.cc`).

Every tokenizer layer implements
  StringsToIds(strs, max_length) → (ids [B,T], labels [B,T], paddings [B,T])
  IdsToStrings(ids [B,T], lens [B]) → list[str]
with the reference's conventions: `ids` = <s> + tokens, `labels` = tokens +
</s>, both truncated/padded to `max_length`; padded label positions hold the
eos id. The string ↔ id maps are native (`ops/csrc_host/text_ops.cpp`).
"""

from __future__ import annotations

import numpy as np
import torch

from lingvo_b200 import ops
from lingvo_b200.core import base_layer


class BaseTokenizer(base_layer.BaseLayer):
  """Common params + the ids/labels/paddings assembly (ref :37-140)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.name = 'tokenizer'
    p.Define('vocab_size', 64, 'Size of the vocabulary.')
    p.Define('append_eos', True, 'Append </s> to the labels.')
    p.Define('pad_to_max_length', True, 'Pad to max_length (else to the batch max).')
    p.Define('target_unk_id', 0, 'Unknown id.')
    p.Define('target_sos_id', 1, 'Start-of-sentence id.')
    p.Define('target_eos_id', 2, 'End-of-sentence id.')
    p.Define('target_wb_id', -1, 'Word-boundary id (optional).')
    return p

  @property
  def sos_id(self):
    return self.params.target_sos_id

  @property
  def eos_id(self):
    return self.params.target_eos_id

  @property
  def unk_id(self):
    return self.params.target_unk_id

  # -- subclass hooks -------------------------------------------------------------
  def _Encode(self, text: str):
    raise NotImplementedError

  def _Decode(self, ids) -> str:
    raise NotImplementedError

  def _Assemble(self, token_lists, max_length, append_eos=None, extras=()):
    """Pads token lists into (ids, labels, paddings) [+ one int tensor per `extras` entry,
    each a per-sample list aligned with the labels; the eos label repeats the last entry]."""
    p = self.params
    append_eos = p.append_eos if append_eos is None else append_eos
    b = len(token_lists)
    if not p.pad_to_max_length:
      max_length = min(max_length, max([len(t) + 1 for t in token_lists] + [1]))
    ids = np.full((b, max_length), p.target_eos_id, np.int32)
    labels = np.full((b, max_length), p.target_eos_id, np.int32)
    paddings = np.ones((b, max_length), np.float32)
    aligned = [np.zeros((b, max_length), np.int32) for _ in extras]
    for i, toks in enumerate(token_lists):
      toks = list(toks)
      inp = ([p.target_sos_id] + toks)[:max_length]
      lab = (toks + ([p.target_eos_id] if append_eos else []))[:max_length]
      ids[i, :len(inp)] = inp
      labels[i, :len(lab)] = lab
      paddings[i, :max(len(lab), 1) if append_eos else len(lab)] = 0.0
      for out, per_sample in zip(aligned, extras):
        vals = list(per_sample[i])
        vals = (vals + [vals[-1] if vals else 0] * (len(lab) - len(vals)))[:len(lab)]
        out[i, :len(vals)] = vals
    return (torch.from_numpy(ids), torch.from_numpy(labels), torch.from_numpy(paddings)) + \
        tuple(torch.from_numpy(a) for a in aligned)

  @staticmethod
  def _ToText(strs):
    return [s.decode('utf-8') if isinstance(s, bytes) else s for s in strs]

  def _AppendEos(self, external_append_eos):
    return self.params.append_eos if external_append_eos is None else external_append_eos

  def StringsToIds(self, strs, max_length, external_append_eos=None, languages=None):
    return self._StringsToIdsImpl(strs, max_length, self._AppendEos(external_append_eos),
                                  languages)

  def _StringsToIdsImpl(self, strs, max_length, append_eos, languages):
    del languages
    return self._Assemble([self._Encode(s) for s in self._ToText(strs)], max_length, append_eos)

  # -- byte offsets (ref :134) -------------------------------------------------------------------
  def _EncodeWithOffsets(self, text):
    """(ids, start byte offsets, end byte offsets). Generic alignment: every id is decoded on
    its own and its surface form located in the UTF-8 bytes at or after the end of the
    previous token (tokenizer case-folding is honoured); tokens with no surface form in the
    text (unk, control pieces) get an empty span at the cursor."""
    ids = list(self._Encode(text))
    raw = text.encode('utf-8')
    low = raw.lower()
    starts, ends, cur = [], [], 0
    for i in ids:
      surface = self._Decode([int(i)]).replace('\u2581', ' ')
      piece = (surface.strip() or surface).encode('utf-8')   # whitespace tokens stay as is
      pos = -1
      if piece:
        pos = raw.find(piece, cur)
        if pos < 0:
          pos = low.find(piece.lower(), cur)
      if pos < 0:
        starts.append(cur)
        ends.append(cur)
      else:
        starts.append(pos)
        cur = pos + len(piece)
        ends.append(cur)
    return ids, starts, ends

  def StringsToIdsWithOffsets(self, strs, max_length, external_append_eos=None,
                              languages=None):
    """(ids, labels, paddings, start_offsets, end_offsets), all [batch, maxlen]; the offsets
    are byte positions of label j in the original string."""
    return self._StringsToIdsWithOffsetsImpl(strs, max_length,
                                             self._AppendEos(external_append_eos), languages)

  def _StringsToIdsWithOffsetsImpl(self, strs, max_length, append_eos, languages):
    del languages
    enc = [self._EncodeWithOffsets(s) for s in self._ToText(strs)]
    # The eos label sits at the end of the string: an empty span at the last token's end.
    ends = [e[2] for e in enc]
    starts = [e[1] + e[2][-1:] for e in enc]
    return self._Assemble([e[0] for e in enc], max_length, append_eos, extras=(starts, ends))

  # -- segments (ref :95) ------------------------------------------------------------------------
  SEGMENT_DELIMITER = '<segment>'

  def StringsToIdsWithSegments(self, strs, max_length, external_append_eos=None,
                               languages=None):
    """Strings made of `<segment>`-separated parts → (ids, labels, paddings, segment_ids);
    segment_ids[i, j] is the 0-based part label j came from (eos: the last part)."""
    return self._StringsToIdsWithSegmentsImpl(strs, max_length,
                                              self._AppendEos(external_append_eos), languages)

  def _StringsToIdsWithSegmentsImpl(self, strs, max_length, append_eos, languages):
    del languages
    toks, segs = [], []
    for s in self._ToText(strs):
      t, g = [], []
      for k, part in enumerate(s.split(self.SEGMENT_DELIMITER)):
        piece = list(self._Encode(part.strip()))
        t += piece
        g += [k] * len(piece)
      toks.append(t)
      segs.append(g)
    return self._Assemble(toks, max_length, append_eos, extras=(segs,))

  def Initialize(self, sess=None):
    """Tokenizers hold no deferred state here (vocab files load in the constructor)."""
    del sess

  def IdsToStrings(self, ids, lens, languages=None):
    del languages
    ids = np.asarray(ids.cpu() if isinstance(ids, torch.Tensor) else ids)
    lens = np.asarray(lens.cpu() if isinstance(lens, torch.Tensor) else lens)
    return [self._Decode([int(t) for t in row[:int(n)]]) for row, n in zip(ids, lens)]

  def IdsToTokens(self, ids, languages=None):
    return [self._Decode([int(i)]) for i in np.asarray(ids).reshape(-1)]


class AsciiTokenizer(BaseTokenizer):
  """Lower-cased character tokenizer with the fixed 76-symbol table (ref :143)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.vocab_size = 76
    return p

  def _Encode(self, text):
    return ops.host().ascii_to_ids(text)

  def _Decode(self, ids):
    return ops.host().ascii_to_string(ids)


class VocabFileTokenizer(BaseTokenizer):
  """Whitespace tokens looked up in a vocab file (ref :170)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('token_vocab_filepath', None, 'One token (optionally `token<TAB>id`) per line.')
    p.Define('ngram_vocab_filepath', None, 'Kept for parity.')
    p.Define('ngram_separator', '', 'Kept for parity.')
    p.Define('tokens_delimiter', ' ', 'Token delimiter.')
    p.Define('load_token_ids_from_vocab', True, 'Ids come from the file\'s second column.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    self._tok = ops.host().VocabTokenizer(p.token_vocab_filepath,
                                          p.load_token_ids_from_vocab)

  @property
  def sos_id(self):
    return self._tok.sos_id if self._tok.sos_id >= 0 else self.params.target_sos_id

  @property
  def eos_id(self):
    return self._tok.eos_id if self._tok.eos_id >= 0 else self.params.target_eos_id

  def _Encode(self, text):
    return self._tok.to_ids(text)

  def _Decode(self, ids):
    return self._tok.to_string(ids)


class BpeTokenizer(BaseTokenizer):
  """Byte-pair encoding from merge rules + vocab (ref :242)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('codes_filepath', None, 'BPE merge rules.')
    p.Define('words_to_ids_filepath', None, 'BPE vocabulary.')
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    self._tok = ops.host().BpeTokenizer(p.codes_filepath, p.words_to_ids_filepath)

  def _Encode(self, text):
    return self._tok.to_ids(text)

  def _Decode(self, ids):
    return self._tok.to_string(ids)


class WpmTokenizer(BaseTokenizer):
  """Greedy longest-match word-piece model over a vocab file (ref :300,
  `wpm_encoder.py`). Pieces starting a word carry the `▁` prefix."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('vocab_filepath', None, 'Word-piece vocabulary, one piece per line.')
    p.Define('merge_prob', 1.0, 'Kept for parity.')
    return p

  def __init__(self, params):
    super().__init__(params)
    with open(self.params.vocab_filepath, encoding='utf-8') as f:
      pieces = [l.rstrip('\n').split('\t')[0] for l in f if l.strip()]
    self._p2i = {p: i for i, p in enumerate(pieces)}
    self._i2p = pieces
    self._max = max(len(p) for p in pieces)

  def _Encode(self, text):
    out = []
    for word in text.split():
      w = '▁' + word
      i = 0
      while i < len(w):
        for j in range(min(len(w), i + self._max), i, -1):
          pid = self._p2i.get(w[i:j])
          if pid is not None:
            out.append(pid)
            i = j
            break
        else:
          out.append(self.params.target_unk_id)
          i += 1
    return out

  def _Decode(self, ids):
    s = ''.join(self._i2p[i] if 0 <= i < len(self._i2p) else '<unk>' for i in ids)
    return s.replace('▁', ' ').strip()


class SentencePieceTokenizer(BaseTokenizer):
  """SentencePiece model wrapper (ref :418)."""

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('spm_model', None, 'Path to the .model file.')
    p.Define('alpha', 1.0, 'Sampling smoothing.')
    p.Define('nbest_size', 0, 'n-best sampling size.')
    return p

  def __init__(self, params):
    super().__init__(params)
    import sentencepiece as spm  # pylint: disable=g-import-not-at-top
    self._sp = spm.SentencePieceProcessor(model_file=self.params.spm_model)

  def _Encode(self, text):
    p = self.params
    if p.nbest_size:
      return self._sp.encode(text, enable_sampling=True, alpha=p.alpha,
                             nbest_size=p.nbest_size)
    return self._sp.encode(text)

  def _Decode(self, ids):
    return self._sp.decode([int(i) for i in ids])
