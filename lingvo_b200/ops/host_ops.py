"""Python facade over the native host ops (`_H.so`): packing, MASS, maps."""

from __future__ import annotations

import numpy as np

from lingvo_b200 import ops


def PackSequences(src_actual_seq_len, tgt_actual_seq_len, packed_batch_size,
                  packed_src_seq_len, packed_tgt_seq_len, seed=0):
  """→ (src_segment_ids, src_segment_pos, src_indices_in_input,
        tgt_segment_ids, tgt_segment_pos, tgt_indices_in_input), each int32
  `[packed_batch_size, seq_len]` (ref `pack_ops.cc:202-313`)."""
  return ops.host().pack_sequences(
      np.asarray(src_actual_seq_len, np.int32), np.asarray(tgt_actual_seq_len, np.int32),
      int(packed_batch_size), int(packed_src_seq_len), int(packed_tgt_seq_len), int(seed))


def PackSingleSequence(input_lengths, max_packed_length, require_sequential_order=False):
  """Packed-group id per input (-1: too long) (ref `pack_ops.cc:403-460`)."""
  return np.asarray(ops.host().pack_single_sequence(
      np.asarray(input_lengths, np.int32), int(max_packed_length),
      bool(require_sequential_order)), np.int32)


def ApplyPacking(inputs, padding, segment_ids, indices_in_input):
  """Gathers rows of `inputs [N, T, …]` into the packed layout described by
  `segment_ids` / `indices_in_input` `[B, L]`; empty slots get `padding`
  (ref `pack_ops.cc:480-700`)."""
  inputs = np.asarray(inputs)
  seg = np.asarray(segment_ids)
  idx = np.asarray(indices_in_input)
  b, _ = seg.shape
  if inputs.ndim == 1:                      # per-sequence scalars: sum over a row's items
    out = np.zeros((b,), inputs.dtype)
    for r in range(b):
      used = np.unique(idx[r][seg[r] > 0])
      out[r] = inputs[used].sum() if used.size else padding
    return out
  return ops.host().apply_packing(np.ascontiguousarray(inputs),
                                  np.asarray(padding, dtype=inputs.dtype).reshape(1),
                                  seg.astype(np.int32), idx.astype(np.int32))


def Mass(ids, weights, actual_seq_len, mask_id=3, mask_ratio=0.5, mask_minlen=0,
         span_len=100000, random_start_prob=0.6, keep_prob=0.1, rand_prob=0.1,
         mask_prob=0.8, mask_target=True, vocab_size=0, first_unreserved_id=4, seed=0):
  """MASS masking (ref `mass_op.cc`): → (src_ids, tgt_ids, tgt_labels, tgt_weights)."""
  return ops.host().mass(
      np.asarray(ids, np.int32), np.asarray(weights, np.float32),
      np.asarray(actual_seq_len, np.int32), mask_id, mask_ratio, mask_minlen, span_len,
      random_start_prob, keep_prob, rand_prob, mask_prob, mask_target, vocab_size,
      first_unreserved_id, seed)


def BestStep(hist_file, tol=0.0, minimize=True):
  """(best_step, last_step) of a `step value` history file (ref `best_step_op_kernels.cc`)."""
  return ops.host().best_step(hist_file, tol, minimize)


class StaticMap:
  """Immutable key→value lookup with a default (ref `static_map_op.cc`:
  StaticMapStringInt / StaticMapIntString / StaticMapIntInt). Backed by the native hash
  maps of `_H.so`; the flavour follows the key / value types."""

  def __init__(self, keys, vals=None, unk=None):
    keys = [k.decode() if isinstance(k, bytes) else k for k in keys]
    vals = None if vals is None else [v.decode() if isinstance(v, bytes) else v for v in vals]
    h = ops.host()
    str_keys = bool(keys) and isinstance(keys[0], str)
    str_vals = bool(vals) and isinstance(vals[0], str)
    if str_keys and str_vals:
      raise ValueError('string → string maps are not supported (ref has none either)')
    if str_keys:
      self._m = h.StaticMapStringInt(keys, [int(v) for v in vals or []], -1 if unk is None else int(unk))
    elif str_vals:
      self._m = h.StaticMapIntString([int(k) for k in keys], vals, '' if unk is None else str(unk))
    else:
      self._m = h.StaticMapIntInt([int(k) for k in keys], [int(v) for v in vals or []],
                                  -1 if unk is None else int(unk))
    self._str_keys, self._str_vals = str_keys, str_vals

  def __len__(self):
    return len(self._m)

  def Lookup(self, xs):
    arr = np.asarray(xs)
    flat = arr.reshape(-1).tolist()
    if self._str_keys:
      flat = [x.decode() if isinstance(x, bytes) else str(x) for x in flat]
    else:
      flat = [int(x) for x in flat]
    out = self._m.lookup(flat)
    if self._str_vals:
      return np.asarray(out, dtype=object).reshape(arr.shape)
    return np.asarray(out, np.int64).reshape(arr.shape)


def MlPerfSubwordIdToString(token_ids, seq_lengths, vocab_filepath):
  """Decodes `[B, T]` sub-word ids (first `seq_lengths[b]` of each row) to strings with the
  MLPerf transformer vocabulary (ref `ml_perf_subword_op.cc`)."""
  vocab = _CachedVocab(('mlperf', vocab_filepath), lambda: ops.host().MlPerfSubword(vocab_filepath))
  ids = np.asarray(token_ids)
  lens = np.maximum(np.asarray(seq_lengths).reshape(-1), 0)
  assert ids.ndim == 2 and lens.shape[0] == ids.shape[0], (ids.shape, lens.shape)
  return [vocab.decode(ids[i, :lens[i]].tolist()) for i in range(ids.shape[0])]


def NgramIdToToken(token_ids, seq_lengths, ngram_vocab_filepath, ngram_separator=''):
  """Ids → concatenated n-gram tokens, one string per row (ref `tokenizer_ops_kernels.cc:150`)."""
  vocab = _CachedVocab(('ngram', ngram_vocab_filepath),
                       lambda: ops.host().VocabTokenizer(ngram_vocab_filepath, False))
  ids = np.asarray(token_ids)
  lens = np.maximum(np.asarray(seq_lengths).reshape(-1), 0)
  assert ids.ndim == 2 and lens.shape[0] == ids.shape[0], (ids.shape, lens.shape)
  return [vocab.join_ids(ids[i, :lens[i]].tolist(), ngram_separator) for i in range(ids.shape[0])]


def TokenInVocab(token, vocab_filepath, load_token_ids_from_vocab=False):
  """Membership test for one token or a list of tokens (ref `x_ops.cc:649`)."""
  vocab = _CachedVocab(('vocab', vocab_filepath, bool(load_token_ids_from_vocab)),
                       lambda: ops.host().VocabTokenizer(vocab_filepath, bool(load_token_ids_from_vocab)))
  dec = lambda t: t.decode() if isinstance(t, bytes) else t
  if isinstance(token, (str, bytes)):
    return dec(token) in vocab
  return np.asarray([dec(t) in vocab for t in token], bool)


def StrToVocabTokens(labels, vocab_filepath, append_eos=True, maxlen=300, pad_to_maxlen=True,
                     load_token_ids_from_vocab=True, delimiter=' '):
  """Whitespace (or `delimiter`) tokenisation against a vocab file → (token_ids with a
  leading <s>, target_ids, paddings), each `[B, maxlen]` (ref `x_ops.cc:696`)."""
  vocab = _CachedVocab(('vocab', vocab_filepath, bool(load_token_ids_from_vocab)),
                       lambda: ops.host().VocabTokenizer(vocab_filepath, bool(load_token_ids_from_vocab)))
  b = len(labels)
  rows = []
  for lab in labels:
    lab = lab.decode() if isinstance(lab, bytes) else lab
    toks = lab.split(delimiter) if delimiter != ' ' else lab.split()
    if delimiter == '':
      toks = list(lab)
    rows.append([vocab.token_to_id(t) for t in toks if t != ''])
  width = maxlen if pad_to_maxlen else max(
      [min(len(r) + (1 if append_eos else 0), maxlen) for r in rows] + [1])
  ids = np.full((b, width), vocab.eos_id, np.int32)
  tgt = np.full((b, width), vocab.eos_id, np.int32)
  pad = np.ones((b, width), np.float32)
  for i, r in enumerate(rows):
    r = r[:width - 1] if append_eos else r[:width]
    labels_i = r + [vocab.eos_id] if append_eos else r
    n = len(labels_i)
    tgt[i, :n] = labels_i
    ids[i, 0] = vocab.sos_id
    ids[i, 1:n] = labels_i[:n - 1]
    pad[i, :n] = 0.0
  return ids, tgt, pad


_VOCABS = {}


def _CachedVocab(key, factory):
  if key not in _VOCABS:
    _VOCABS[key] = factory()
  return _VOCABS[key]
