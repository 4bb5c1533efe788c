"""BERT word-piece tokenizer (ref `lingvo/tasks/lm/tokenizer.py:26` `BertTokenizer`).

Basic tokenisation (lower-casing, accent stripping, punctuation splitting, CJK isolation)
followed by greedy longest-match-first word pieces with the `##` continuation prefix —
implemented here directly (the reference wraps `tensorflow_text`), so there is no TF
dependency and the vocabulary file format (`vocab.txt`, one piece per line) is the
standard one.
"""

from __future__ import annotations

import unicodedata

from lingvo_b200.core import tokenizers


def _IsPunct(ch):
  cp = ord(ch)
  if 33 <= cp <= 47 or 58 <= cp <= 64 or 91 <= cp <= 96 or 123 <= cp <= 126:
    return True
  return unicodedata.category(ch).startswith('P')


def _IsCjk(cp):
  return (0x4E00 <= cp <= 0x9FFF or 0x3400 <= cp <= 0x4DBF or 0xF900 <= cp <= 0xFAFF or
          0x20000 <= cp <= 0x2A6DF or 0x2A700 <= cp <= 0x2CEAF or 0x2F800 <= cp <= 0x2FA1F)


def BasicTokenize(text, lower_case=True):
  if lower_case:
    text = unicodedata.normalize('NFD', text.lower())
    text = ''.join(c for c in text if unicodedata.category(c) != 'Mn')
  out, cur = [], []
  def _Flush():
    if cur:
      out.append(''.join(cur))
      cur.clear()
  for ch in text:
    cp = ord(ch)
    if ch.isspace():
      _Flush()
    elif cp == 0 or cp == 0xFFFD or unicodedata.category(ch) in ('Cc', 'Cf'):
      continue
    elif _IsPunct(ch) or _IsCjk(cp):
      _Flush()
      out.append(ch)
    else:
      cur.append(ch)
  _Flush()
  return out


class BertTokenizer(tokenizers.BaseTokenizer):

  @classmethod
  def Params(cls):
    p = super().Params()
    p.Define('vocab_filepath', None, 'vocab.txt (one word piece per line).')
    p.Define('lower_case', True, 'Lower-case and strip accents (uncased models).')
    p.Define('unk_token', '[UNK]', 'Unknown token.')
    p.Define('max_chars_per_word', 100, 'Longer words map to unk.')
    p.Define('suffix_indicator', '##', 'Continuation prefix.')
    p.target_sos_id = 101      # [CLS]
    p.target_eos_id = 102      # [SEP]
    p.target_unk_id = 100
    p.append_eos = True
    return p

  def __init__(self, params):
    super().__init__(params)
    p = self.params
    with open(p.vocab_filepath, encoding='utf-8') as f:
      self._pieces = [l.rstrip('\n') for l in f]
    self._ids = {w: i for i, w in enumerate(self._pieces)}
    self._unk = self._ids.get(p.unk_token, p.target_unk_id)
    if not p.vocab_size:
      p.vocab_size = len(self._pieces)

  def WordPieces(self, word):
    p = self.params
    if len(word) > p.max_chars_per_word:
      return [self._unk]
    out, start = [], 0
    while start < len(word):
      end, cur = len(word), None
      while start < end:
        piece = word[start:end]
        if start > 0:
          piece = p.suffix_indicator + piece
        if piece in self._ids:
          cur = self._ids[piece]
          break
        end -= 1
      if cur is None:
        return [self._unk]
      out.append(cur)
      start = end
    return out

  def Encode(self, text):
    ids = []
    for w in BasicTokenize(text, self.params.lower_case):
      ids.extend(self.WordPieces(w))
    return ids

  def _Encode(self, text):
    return self.Encode(text)

  def _Decode(self, ids):
    p = self.params
    words = []
    for t in ids:
      piece = self._pieces[t] if 0 <= t < len(self._pieces) else p.unk_token
      if piece.startswith(p.suffix_indicator) and words:
        words[-1] += piece[len(p.suffix_indicator):]
      else:
        words.append(piece)
    return ' '.join(words)
