"""WER scorer, object-oriented version (ref `lingvo/tasks/asr/tools/simple_wer_v2.py`).

`SimpleWER` accumulates errors over many (hyp, ref) pairs, tracks key-phrase
precision/recall/F1 and the most frequent substitution / insertion / deletion patterns;
pluggable `HtmlHandler`s render the alignment.

  python -m lingvo_b200.models.asr.tools.simple_wer_v2 hyp.txt ref.txt [keyphrases.txt]
"""

from __future__ import annotations

import collections
import re
import sys

from lingvo_b200.models.asr import levenshtein_distance as lev


def TxtPreprocess(txt):
  txt = re.sub(r'[^\w\s\']', ' ', txt.lower(), flags=re.UNICODE)
  return ' '.join(txt.split())


def RemoveCommentTxtPreprocess(txt):
  return TxtPreprocess(re.sub(r'\[\w+\]', '', txt))


class HtmlHandler:
  """Renderer interface (ref :60)."""

  def Setup(self, hypothesis, reference):
    pass

  def Render(self, hyp_word, ref_word, err_type):
    return ''


class HighlightAlignedHtmlHandler(HtmlHandler):
  """Colour-coded alignment (ref :83)."""

  _STYLE = {'sub': 'yellow', 'del': 'red', 'ins': 'green'}

  def __init__(self, highlight_color=None):
    self._colors = dict(self._STYLE)
    if highlight_color:
      self._colors.update(highlight_color)

  def Render(self, hyp_word, ref_word, err_type):
    if err_type == 'none':
      return '%s ' % hyp_word
    c = self._colors[err_type]
    if err_type == 'sub':
      return ('<span style="background-color: %s"><del>%s</del></span>'
              '<span style="background-color: %s">%s </span> ' % (c, hyp_word, c, ref_word))
    if err_type == 'del':
      return '<span style="background-color: %s">%s </span> ' % (c, ref_word)
    return '<span style="background-color: %s"><del>%s</del> </span> ' % (c, hyp_word)


class SimpleWER:
  """Accumulating scorer (ref :160)."""

  def __init__(self, key_phrases=None, html_handler=None, preprocess_handler=None):
    self._pre = preprocess_handler or RemoveCommentTxtPreprocess
    self._html = html_handler or HighlightAlignedHtmlHandler()
    self.key_phrases = [self._pre(k) for k in (key_phrases or []) if k.strip()]
    self.aligned_htmls = []
    self.wer_info = {'sub': 0, 'ins': 0, 'del': 0, 'nw': 0}
    self.kp_stats = collections.OrderedDict(
        (k, {'ref': 0, 'hyp': 0, 'hit': 0}) for k in self.key_phrases)
    self.err_patterns = {'sub': collections.Counter(), 'ins': collections.Counter(),
                         'del': collections.Counter()}

  def AddHypRef(self, hypothesis, reference):
    hyp, ref = self._pre(hypothesis), self._pre(reference)
    h, r = hyp.split(), ref.split()
    self._html.Setup(hyp, ref)
    self.wer_info['nw'] += len(r)
    pieces = []
    for op, ri, hi in lev.Alignment(r, h):
      hw, rw = (h[hi] if hi >= 0 else ''), (r[ri] if ri >= 0 else '')
      if op == 'ok':
        pieces.append(self._html.Render(hw, rw, 'none'))
        continue
      self.wer_info[op] += 1
      self.err_patterns[op][(rw, hw) if op == 'sub' else (rw or hw)] += 1
      pieces.append(self._html.Render(hw, rw, op))
    self.aligned_htmls.append(''.join(pieces))
    pad_r, pad_h = ' %s ' % ref, ' %s ' % hyp
    for k, st in self.kp_stats.items():
      nr, nh = pad_r.count(' %s ' % k), pad_h.count(' %s ' % k)
      st['ref'] += nr
      st['hyp'] += nh
      st['hit'] += min(nr, nh)

  def GetWER(self):
    """→ (total WER %, dict of sub/ins/del %)."""
    nw = max(self.wer_info['nw'], 1)
    parts = {k: 100.0 * self.wer_info[k] / nw for k in ('sub', 'ins', 'del')}
    return sum(parts.values()), parts

  def GetBreakdownWER(self):
    return self.GetWER()[1]

  def GetKeyPhraseStats(self):
    """→ (jaccard, F1, precision, recall) over all key phrases (ref :291)."""
    ref = sum(s['ref'] for s in self.kp_stats.values())
    hyp = sum(s['hyp'] for s in self.kp_stats.values())
    hit = sum(s['hit'] for s in self.kp_stats.values())
    if not ref and not hyp:
      return 1.0, 1.0, 1.0, 1.0
    prec = hit / hyp if hyp else 0.0
    rec = hit / ref if ref else 0.0
    f1 = 2 * prec * rec / (prec + rec) if prec + rec else 0.0
    union = ref + hyp - hit
    return (hit / union if union else 1.0), f1, prec, rec

  def GetSummaries(self):
    wer, parts = self.GetWER()
    s = 'WER = %.2f%% (sub %.2f%%, ins %.2f%%, del %.2f%%) over %d words' % (
        wer, parts['sub'], parts['ins'], parts['del'], self.wer_info['nw'])
    kp = ''
    if self.key_phrases:
      j, f1, p, r = self.GetKeyPhraseStats()
      kp = 'key phrases: jaccard %.3f F1 %.3f precision %.3f recall %.3f' % (j, f1, p, r)
    return s, kp

  def GetMostFrequentErrPatterns(self, n=10):
    return {k: c.most_common(n) for k, c in self.err_patterns.items()}


def _ReadLines(path):
  with open(path, encoding='utf-8') as f:
    return [l.rstrip('\n') for l in f if l.strip()]


def main(argv):
  if len(argv) < 3:
    print(__doc__)
    return 1
  hyps, refs = _ReadLines(argv[1]), _ReadLines(argv[2])
  phrases = _ReadLines(argv[3]) if len(argv) > 3 else None
  scorer = SimpleWER(key_phrases=phrases)
  for h, r in zip(hyps, refs):
    scorer.AddHypRef(h, r)
  for line in scorer.GetSummaries():
    if line:
      print(line)
  return 0


if __name__ == '__main__':
  sys.exit(main(sys.argv))
