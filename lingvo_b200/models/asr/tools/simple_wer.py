"""Stand-alone word-error-rate scorer (ref `lingvo/tasks/asr/tools/simple_wer.py`).

  python -m lingvo_b200.models.asr.tools.simple_wer hyp.txt ref.txt [keyphrases.txt]

Each input line is `<utt_id> <transcript...>` or a bare transcript (then lines pair up
by order). Prints WER with its ins/del/sub split and, given key phrases, their
precision-style hit statistics. `ComputeWER` / `AnalyzeKeyPhrases` are importable.
"""

from __future__ import annotations

import re
import sys

from lingvo_b200.models.asr import levenshtein_distance as lev


def TxtPreprocess(txt):
  """Lower-case, keep word characters and apostrophes, squeeze blanks (ref :53)."""
  txt = re.sub(r'[^\w\s\']', ' ', txt.lower(), flags=re.UNICODE)
  return ' '.join(txt.split())


def RemoveCommentTxtPreprocess(txt):
  """Also removes [bracketed comments] such as [noise] (ref :47)."""
  return TxtPreprocess(re.sub(r'\[\w+\]', '', txt))


def HighlightAlignedHtml(hyp, ref, err_type):
  """One aligned position rendered as HTML (ref :65)."""
  if err_type == 'none':
    return '%s ' % hyp
  if err_type == 'sub':
    return '<span style="background-color: yellow"><del>%s</del></span>' \
           '<span style="background-color: yellow">%s </span> ' % (hyp, ref)
  if err_type == 'del':
    return '<span style="background-color: red">%s </span> ' % ref
  if err_type == 'ins':
    return '<span style="background-color: green"><del>%s</del> </span> ' % hyp
  raise ValueError('unknown err_type ' + err_type)


def ComputeWER(hyp, ref, diagnosis=False):
  """→ (wer_info dict, html string) (ref :96)."""
  h, r = hyp.split(), ref.split()
  info = {'sub': 0, 'ins': 0, 'del': 0, 'nw': len(r)}
  html = []
  for op, ri, hi in lev.Alignment(r, h):
    if op != 'ok':
      info[op] += 1
    if diagnosis:
      html.append(HighlightAlignedHtml(h[hi] if hi >= 0 else '', r[ri] if ri >= 0 else '',
                                       'none' if op == 'ok' else op))
  return info, ''.join(html)


# -- the v1 entry points (ref `simple_wer.py:46-292`) ------------------------------------------
def ComputeEditDistanceMatrix(hs, rs):
  """Word-level Levenshtein DP table `[len(rs) + 1, len(hs) + 1]` (int32 numpy): entry
  (r, h) is the distance between the first r reference and first h hypothesis words."""
  import numpy as np  # pylint: disable=g-import-not-at-top
  dr, dh = len(rs) + 1, len(hs) + 1
  dists = np.zeros((dr, dh), np.int32)
  dists[:, 0] = np.arange(dr)
  dists[0, :] = np.arange(dh)
  for i in range(1, dr):
    for j in range(1, dh):
      if rs[i - 1] == hs[j - 1]:
        dists[i, j] = dists[i - 1, j - 1]
      else:
        dists[i, j] = 1 + min(dists[i - 1, j - 1], dists[i, j - 1], dists[i - 1, j])
  return dists


def PreprocessTxtBeforeWER(txt):
  """Lower-cases, drops [noise]-style comments and " - ", squeezes blanks (v1 rules)."""
  txt = re.sub(r'\[\w+\]', '', txt.lower())
  txt = txt.replace(' - ', ' ')
  return ' '.join(txt.replace('\n', ' ').split())


def GenerateSummaryFromErrs(nref, errs):
  """→ ('total error = …, total word = …, wer = …%', 'Error breakdown: del …, ins …, sub …')."""
  total = sum(errs[k] for k in ('sub', 'ins', 'del'))
  nref = max(nref, 1)
  return ('total error = %d, total word = %d, wer = %.2f%%' % (total, nref,
                                                                total * 100.0 / nref),
          'Error breakdown: del = %.2f%%, ins=%.2f%%, sub=%.2f%%' % (
              errs['del'] * 100.0 / nref, errs['ins'] * 100.0 / nref,
              errs['sub'] * 100.0 / nref))


def AverageWERs(hyps, refs, verbose=True, diagnosis=False):
  """Corpus-level errors over paired lists → (errs dict, number of reference words, list of
  aligned diagnosis html strings)."""
  total = {'sub': 0, 'ins': 0, 'del': 0}
  totalw = 0
  htmls = []
  for hyp, ref in zip(hyps, refs):
    info, html = ComputeWER(PreprocessTxtBeforeWER(hyp), PreprocessTxtBeforeWER(ref), diagnosis)
    if diagnosis:
      htmls.append(html)
    totalw += info['nw']
    for k in total:
      total[k] += info[k]
  if verbose:
    for line in GenerateSummaryFromErrs(totalw, total):
      print(line)
  return total, totalw, htmls


def AnalyzeKeyPhrases(hyp, ref, keyphrases):
  """Counts key phrases present in ref (`ref_nkp`) and recovered in hyp (`hyp_nkp`)
  (ref :183)."""
  ret = {'matched': [], 'ref_nkp': 0, 'hyp_nkp': 0}
  pad_ref, pad_hyp = ' %s ' % ref, ' %s ' % hyp
  for kp in keyphrases:
    n_ref = pad_ref.count(' %s ' % kp)
    if n_ref:
      n_hyp = min(pad_hyp.count(' %s ' % kp), n_ref)
      ret['ref_nkp'] += n_ref
      ret['hyp_nkp'] += n_hyp
      ret['matched'].append(kp)
  return ret


def _ReadTranscripts(path):
  out = {}
  with open(path, encoding='utf-8') as f:
    for n, line in enumerate(f):
      line = line.strip()
      if not line:
        continue
      m = re.match(r'^(\S+)\s+(.*)$', line)
      if m and re.search(r'[\d_\-]', m.group(1)):
        out[m.group(1)] = m.group(2)
      else:
        out['#%d' % n] = line
  return out


def main(argv):
  if len(argv) < 3:
    print(__doc__)
    return 1
  hyps, refs = _ReadTranscripts(argv[1]), _ReadTranscripts(argv[2])
  phrases = []
  if len(argv) > 3:
    with open(argv[3], encoding='utf-8') as f:
      phrases = [TxtPreprocess(l) for l in f if l.strip()]
  tot = {'sub': 0, 'ins': 0, 'del': 0, 'nw': 0}
  kp = {'ref_nkp': 0, 'hyp_nkp': 0}
  for key, ref in refs.items():
    hyp = TxtPreprocess(hyps.get(key, ''))
    ref = RemoveCommentTxtPreprocess(ref)
    info, _ = ComputeWER(hyp, ref)
    for k in tot:
      tot[k] += info[k]
    if phrases:
      a = AnalyzeKeyPhrases(hyp, ref, phrases)
      kp['ref_nkp'] += a['ref_nkp']
      kp['hyp_nkp'] += a['hyp_nkp']
  errs = tot['sub'] + tot['ins'] + tot['del']
  nw = max(tot['nw'], 1)
  print('WER: %.2f%% (%d errors / %d words: %d sub, %d ins, %d del)' % (
      100.0 * errs / nw, errs, tot['nw'], tot['sub'], tot['ins'], tot['del']))
  if phrases:
    print('Key phrases: %d / %d recovered (%.1f%%)' % (
        kp['hyp_nkp'], kp['ref_nkp'], 100.0 * kp['hyp_nkp'] / max(kp['ref_nkp'], 1)))
  return 0


if __name__ == '__main__':
  sys.exit(main(sys.argv))
