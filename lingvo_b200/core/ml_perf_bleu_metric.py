"""MLPerf (tensor2tensor-style) BLEU (ref `lingvo/core/ml_perf_bleu_metric.py`)."""
import collections
import math
import re
import sys
import unicodedata

from lingvo_b200.core import metrics


def _get_ngrams(segment, max_order):  # pylint: disable=invalid-name
  c = collections.Counter()
  for o in range(1, max_order + 1):
    for i in range(len(segment) - o + 1):
      c[tuple(segment[i:i + o])] += 1
  return c


def compute_bleu(reference_corpus, translation_corpus, max_order=4, use_bp=True):  # pylint: disable=invalid-name
  matches = [0] * max_order
  possible = [0] * max_order
  ref_len = hyp_len = 0
  for ref, hyp in zip(reference_corpus, translation_corpus):
    ref_len += len(ref)
    hyp_len += len(hyp)
    r, h = _get_ngrams(ref, max_order), _get_ngrams(hyp, max_order)
    for g, c in (r & h).items():
      matches[len(g) - 1] += c
    for g, c in h.items():
      possible[len(g) - 1] += c
  precisions = []
  smooth = 1.0
  for m, p in zip(matches, possible):
    if p == 0:
      precisions.append(0.0)
    elif m > 0:
      precisions.append(m / p)
    else:
      smooth *= 2
      precisions.append(1.0 / (smooth * p))
  geo = math.exp(sum(math.log(x) for x in precisions if x > 0) / max_order) if max(
      precisions) > 0 else 0.0
  bp = 1.0
  if use_bp and hyp_len:
    ratio = hyp_len / max(ref_len, 1)
    bp = math.exp(1 - 1.0 / ratio) if ratio < 1.0 else 1.0
  return geo * bp


class UnicodeRegex:
  """Punctuation / symbol handling of mteval-v14."""

  def __init__(self):
    punct = ''.join(chr(x) for x in range(sys.maxunicode)
                    if unicodedata.category(chr(x)).startswith('P'))
    self.nondigit_punct_re = re.compile(r'([^\d])([' + re.escape(punct) + r'])')
    self.punct_nondigit_re = re.compile(r'([' + re.escape(punct) + r'])([^\d])')
    sym = ''.join(chr(x) for x in range(sys.maxunicode)
                  if unicodedata.category(chr(x)).startswith('S'))
    self.symbol_re = re.compile('([' + re.escape(sym) + '])')


_UREGEX = None


def bleu_tokenize(string):  # pylint: disable=invalid-name
  global _UREGEX
  if _UREGEX is None:
    _UREGEX = UnicodeRegex()
  string = _UREGEX.nondigit_punct_re.sub(r'\1 \2 ', string)
  string = _UREGEX.punct_nondigit_re.sub(r' \1 \2', string)
  string = _UREGEX.symbol_re.sub(r' \1 ', string)
  return string.split()


def bleu_wrapper(ref_lines, hyp_lines, case_sensitive=False):  # pylint: disable=invalid-name
  if not case_sensitive:
    ref_lines = [x.lower() for x in ref_lines]
    hyp_lines = [x.lower() for x in hyp_lines]
  return compute_bleu([bleu_tokenize(x) for x in ref_lines],
                      [bleu_tokenize(x) for x in hyp_lines])


class MlPerfBleuMetric(metrics.BaseMetric):

  def __init__(self, **kwargs):
    self._refs, self._hyps = [], []

  def Update(self, ref_str, hyp_str, eval_weight=1.0):
    del eval_weight
    self._refs.append(ref_str)
    self._hyps.append(hyp_str)

  @property
  def value(self):
    return bleu_wrapper(self._refs, self._hyps) if self._refs else 0.0


def is_unicode(s):  # pylint: disable=invalid-name
  return isinstance(s, str)


def to_unicode(s, ignore_errors=False):  # pylint: disable=invalid-name
  if is_unicode(s):
    return s
  return s.decode('utf-8', errors='ignore' if ignore_errors else 'strict')


def native_to_unicode(s):  # pylint: disable=invalid-name
  """Bytes or str → str, dropping undecodable bytes (ref :120)."""
  try:
    return to_unicode(s)
  except UnicodeDecodeError:
    return to_unicode(s, ignore_errors=True)


def bleu_score(predictions, labels, max_order=4):  # pylint: disable=invalid-name
  """Corpus BLEU of token-id sequences `[batch, time]` (trailing 0 padding ignored) against
  `labels`; returns (bleu, weight 1.0) like the reference's metric fn."""
  import numpy as np  # pylint: disable=g-import-not-at-top

  def _Rows(x):
    x = np.asarray(x.detach().cpu() if hasattr(x, 'detach') else x)
    x = x.reshape(x.shape[0], -1)
    return [[int(t) for t in np.trim_zeros(row, 'b')] for row in x]

  return compute_bleu(_Rows(labels), _Rows(predictions), max_order), 1.0
