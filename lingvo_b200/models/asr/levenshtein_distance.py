"""Edit-distance bookkeeping for ASR scoring (ref `lingvo/tasks/asr/levenshtein_distance.py`).

`LevenshteinDistance(ref, hyp)` returns the insertion / deletion / substitution split of
one minimum-cost alignment. Implemented as a full DP table plus a back-trace (the table
is also what `tools/simple_wer_v2` uses to render alignments).
"""

from __future__ import annotations

from typing import List, Sequence, Tuple


class ErrorStats:
  """Counts of one alignment (ref :22)."""

  __slots__ = ('insertions', 'deletions', 'subs', 'total')

  def __init__(self, ins=0, dels=0, subs=0, tot=0):
    self.insertions, self.deletions, self.subs, self.total = ins, dels, subs, tot

  def __repr__(self):
    return 'ErrorStats(ins=%d, dels=%d, subs=%d, tot=%d)' % (
        self.insertions, self.deletions, self.subs, self.total)

  def __eq__(self, other):
    return (self.insertions, self.deletions, self.subs, self.total) == (
        other.insertions, other.deletions, other.subs, other.total)


def CostTable(ref: Sequence, hyp: Sequence) -> List[List[int]]:
  """d[i][j] = edit distance between ref[:i] and hyp[:j]."""
  n, m = len(ref), len(hyp)
  d = [[0] * (m + 1) for _ in range(n + 1)]
  for i in range(1, n + 1):
    d[i][0] = i
  for j in range(1, m + 1):
    d[0][j] = j
  for i in range(1, n + 1):
    ri, row, up = ref[i - 1], d[i], d[i - 1]
    for j in range(1, m + 1):
      best = up[j - 1] + (ri != hyp[j - 1])
      if up[j] + 1 < best:
        best = up[j] + 1
      if row[j - 1] + 1 < best:
        best = row[j - 1] + 1
      row[j] = best
  return d


def Alignment(ref: Sequence, hyp: Sequence) -> List[Tuple[str, int, int]]:
  """Back-trace of one optimal alignment: list of (op, ref_idx, hyp_idx) with op in
  {'ok', 'sub', 'del', 'ins'} (idx −1 where the side is absent). Ties prefer
  substitution, then deletion, then insertion (matches the reference's counts)."""
  d = CostTable(ref, hyp)
  i, j = len(ref), len(hyp)
  ops = []
  while i > 0 or j > 0:
    if i > 0 and j > 0:
      diag = d[i - 1][j - 1] + (ref[i - 1] != hyp[j - 1])
      ins_c, del_c = d[i][j - 1] + 1, d[i - 1][j] + 1
      if diag < ins_c and diag < del_c:
        ops.append(('ok' if ref[i - 1] == hyp[j - 1] else 'sub', i - 1, j - 1))
        i, j = i - 1, j - 1
        continue
      if del_c < ins_c:
        ops.append(('del', i - 1, -1))
        i -= 1
        continue
      ops.append(('ins', -1, j - 1))
      j -= 1
    elif i > 0:
      ops.append(('del', i - 1, -1))
      i -= 1
    else:
      ops.append(('ins', -1, j - 1))
      j -= 1
  ops.reverse()
  return ops


def LevenshteinDistance(lst_ref: List[str], lst_hyp: List[str]) -> ErrorStats:
  """Error split of an optimal alignment of `lst_hyp` against `lst_ref` (ref :35)."""
  st = ErrorStats()
  for op, _, _ in Alignment(lst_ref, lst_hyp):
    if op == 'sub':
      st.subs += 1
    elif op == 'del':
      st.deletions += 1
    elif op == 'ins':
      st.insertions += 1
  st.total = st.subs + st.deletions + st.insertions
  return st
