"""BLEU scoring (ref `lingvo/core/scorers.py`)."""
import collections
import math


def _Tokenize(s):
  s = s.decode('utf-8') if isinstance(s, bytes) else s
  return s.split()


def NGrams(lst, order):
  return (lst[i:i + order] for i in range(len(lst) - order + 1))


class Unsegmenter:
  """Undoes BPE (`@@ `) / WPM-SPM (`▁`) segmentation."""

  def __init__(self, separator_type=None):
    self._t = separator_type

  def __call__(self, line):
    line = line.decode('utf-8') if isinstance(line, bytes) else line
    if self._t == 'bpe':
      return line.replace('@@ ', '').strip()
    if self._t in ('wpm', 'spm'):
      return line.replace(' ', '').replace('▁', ' ').strip()
    return line


class BleuScorer:
  """Corpus BLEU: geometric mean of clipped n-gram precisions × brevity penalty."""

  def __init__(self, max_ngram=4, separator_type=None):
    self._max = max_ngram
    self._unseg = Unsegmenter(separator_type)
    self._hyp_ngram_matches = [0] * max_ngram
    self._hyp_ngram_counts = [0] * max_ngram
    self._num_ref_tokens = 0
    self._num_hyp_tokens = 0

  @property
  def unsegmenter(self):
    return self._unseg

  def AddSentence(self, ref_str, hyp_str):
    ref = tuple(_Tokenize(self._unseg(ref_str)))
    hyp = tuple(_Tokenize(self._unseg(hyp_str)))
    self._num_ref_tokens += len(ref)
    self._num_hyp_tokens += len(hyp)
    for o in range(1, self._max + 1):
      r = collections.Counter(tuple(g) for g in NGrams(ref, o))
      h = collections.Counter(tuple(g) for g in NGrams(hyp, o))
      self._hyp_ngram_matches[o - 1] += sum(min(c, r[g]) for g, c in h.items())
      self._hyp_ngram_counts[o - 1] += max(len(hyp) - o + 1, 0)

  def ComputeOverallScore(self):
    score = 0.0
    num = 0
    for m, c in zip(self._hyp_ngram_matches, self._hyp_ngram_counts):
      if c == 0:
        break
      if m == 0:
        return 0.0
      score += math.log(m / c)
      num += 1
    if not num:
      return 0.0
    score = math.exp(score / num)
    if self._num_hyp_tokens < self._num_ref_tokens:
      score *= math.exp(1 - self._num_ref_tokens / max(self._num_hyp_tokens, 1))
    return score
