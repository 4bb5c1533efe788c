"""ASR decoder utilities (ref `lingvo/tasks/asr/decoder_utils.py`)."""

from __future__ import annotations

import re


def SetRnnCellNodes(decoder_params, rnn_cell_params):
  rnn_cell_params.num_output_nodes = decoder_params.rnn_cell_dim
  if decoder_params.rnn_cell_hidden_dim > 0:
    rnn_cell_params.num_hidden_nodes = decoder_params.rnn_cell_hidden_dim


def Tokenize(string):
  """Whitespace tokens, case preserved (ref `decoder_utils.py:37`; callers lower-case
  explicitly for the case-insensitive rates)."""
  if isinstance(string, bytes):
    string = string.decode('utf-8')
  return string.split()


def EditDistance(ref_str, hyp_str):
  """Word-level Levenshtein → (ins, subs, dels, total)."""
  return EditDistanceInIds(Tokenize(ref_str), Tokenize(hyp_str))


def EditDistanceInIds(ref, hyp):
  """Levenshtein alignment counts between two token sequences."""
  n, m = len(ref), len(hyp)
  # dp[i][j] = (cost, ins, subs, dels)
  prev = [(j, j, 0, 0) for j in range(m + 1)]
  for i in range(1, n + 1):
    cur = [(i, 0, 0, i)]
    for j in range(1, m + 1):
      if ref[i - 1] == hyp[j - 1]:
        best = prev[j - 1]
      else:
        c = prev[j - 1]
        best = (c[0] + 1, c[1], c[2] + 1, c[3])
      d = prev[j]
      cand = (d[0] + 1, d[1], d[2], d[3] + 1)
      if cand[0] < best[0]:
        best = cand
      ins = cur[j - 1]
      cand = (ins[0] + 1, ins[1] + 1, ins[2], ins[3])
      if cand[0] < best[0]:
        best = cand
      cur.append(best)
    prev = cur
  cost, ins, subs, dels = prev[m]
  return ins, subs, dels, cost


def ComputeWer(hyps, refs, normalize_punct_and_cap=False):
  """→ list of (errors, ref_words) per pair."""
  out = []
  for h, r in zip(hyps, refs):
    if normalize_punct_and_cap:
      h, r = re.sub(r'[^\w\s]', '', h), re.sub(r'[^\w\s]', '', r)
    _, _, _, total = EditDistance(r, h)
    out.append((total, len(Tokenize(r))))
  return out


def FilterEpsilon(string):
  return ' '.join(string.replace('<epsilon>', ' ').split())


def FilterNoise(string):
  return ' '.join(t for t in string.split() if t != '<noise>')
