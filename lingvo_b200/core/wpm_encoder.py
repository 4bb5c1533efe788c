"""Word-piece model encoder (ref `lingvo/core/wpm_encoder.py:41`): greedy
longest-prefix matching over a vocab file; `▁` marks word starts."""

import torch

NO_TOKEN = 0
BOW_STR = '▁'


class WpmEncoder:

  def __init__(self, wpm_filepath, merge_prob=1.):
    del merge_prob
    with open(wpm_filepath, encoding='utf-8') as f:
      self._pieces = [l.rstrip('\n').split('\t')[0] for l in f if l.strip('\n')]
    self._piece2id = {p: i for i, p in enumerate(self._pieces)}
    self._max = max(len(p) for p in self._pieces)
    self._unk = self._piece2id.get('<unk>', 0)

  @property
  def sentence_start_id(self):
    return self._piece2id.get('<s>', 1)

  @property
  def sentence_end_id(self):
    return self._piece2id.get('</s>', 2)

  @property
  def unk_id(self):
    return self._unk

  @property
  def sentence_start_string(self):
    return '<s>'

  @property
  def sentence_end_string(self):
    return '</s>'

  def EncodeWord(self, word):
    w = BOW_STR + word
    out, i = [], 0
    while i < len(w):
      for j in range(min(len(w), i + self._max), i, -1):
        if w[i:j] in self._piece2id:
          out.append(w[i:j])
          i = j
          break
      else:
        out.append('<unk>')
        i += 1
    return out

  def Encode(self, text):
    """→ (ids list, pieces list)."""
    pieces = [p for word in text.split() for p in self.EncodeWord(word)]
    return [self._piece2id.get(p, self._unk) for p in pieces], pieces

  def EncodeToStringAndIds(self, text):
    ids, pieces = self.Encode(text)
    return pieces, ids

  def Decode(self, ids):
    s = ''.join(self._pieces[i] if 0 <= i < len(self._pieces) else '<unk>' for i in ids)
    return s.replace(BOW_STR, ' ').strip()
