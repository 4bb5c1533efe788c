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
  b, l = seg.shape
  out = np.full((b, l) + inputs.shape[2:], padding, dtype=inputs.dtype)
  if inputs.ndim == 1:                      # per-sequence scalars: sum over a row's items
    out = np.zeros((b,), inputs.dtype)
    for r in range(b):
      used = np.unique(idx[r][seg[r] > 0])
      out[r] = inputs[used].sum() if used.size else padding
    return out
  # position inside the source row = running count within the segment
  for r in range(b):
    valid = seg[r] > 0
    if not valid.any():
      continue
    pos = np.zeros(l, np.int64)
    prev_key, run = None, 0
    for c in range(l):
      if not valid[c]:
        continue
      key = (seg[r, c], idx[r, c])
      run = run + 1 if key == prev_key else 0
      prev_key = key
      pos[c] = run
    cols = np.nonzero(valid)[0]
    out[r, cols] = inputs[idx[r, cols], pos[cols]]
  return out


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
  """Immutable key→value lookup with a default (ref `static_map_op.cc`)."""

  def __init__(self, keys, vals=None, unk=None):
    keys = list(keys)
    vals = list(range(len(keys))) if vals is None else list(vals)
    self._m = dict(zip(keys, vals))
    self._unk = unk

  def Lookup(self, xs):
    arr = np.asarray(xs)
    flat = [self._m.get(x.decode() if isinstance(x, bytes) else
                        (x.item() if hasattr(x, 'item') else x), self._unk)
            for x in arr.reshape(-1)]
    return np.asarray(flat).reshape(arr.shape)
