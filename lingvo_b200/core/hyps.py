"""`Hypothesis` records and the host-side hyp ops (ref `core/ops/hyps.proto`,
`beam_search_step_op_kernels.cc`: `HypsFromBeamSearchOuts` :1188, `UnpackHyp`).

The device beam search (`ops/beam_search.py`) keeps everything in tensors; these helpers
exist for the serialised interchange format the reference's decoders / tools consume:
wire-compatible `tensorflow.lingvo.Hypothesis` bytes written with the in-repo protobuf
codec, so files produced here parse with the reference's `hyps_pb2` and vice versa.
"""

from __future__ import annotations

import dataclasses
import struct
from typing import List, Sequence

import numpy as np
import torch

from lingvo_b200.utils import protowire as pw


@dataclasses.dataclass
class Hypothesis:
  """beam_id=1, ids=2 (packed), scores=3 (packed), atten_vecs=4 {prob=1 packed},
  normalized_score=5."""
  beam_id: int = 0
  ids: List[int] = dataclasses.field(default_factory=list)
  scores: List[float] = dataclasses.field(default_factory=list)
  atten_vecs: List[List[float]] = dataclasses.field(default_factory=list)
  normalized_score: float = 0.0

  def SerializeToString(self) -> bytes:
    out = [pw.f_varint(1, self.beam_id)]
    if self.ids:
      out.append(pw.f_packed_varint(2, [i & 0xFFFFFFFFFFFFFFFF for i in self.ids]))
    if self.scores:
      out.append(pw.f_packed_float(3, self.scores))
    for vec in self.atten_vecs:
      out.append(pw.f_bytes(4, pw.f_packed_float(1, vec) if len(vec) else b''))
    if self.normalized_score:
      out.append(pw.f_float(5, self.normalized_score))
    return b''.join(out)

  @classmethod
  def FromString(cls, buf: bytes) -> 'Hypothesis':
    h = cls()
    for field, wire, val in pw.parse(buf):
      if field == 1:
        h.beam_id = _Int32(val)
      elif field == 2:
        h.ids.extend(_Int32(v) for v in (pw.parse_packed_varints(val) if wire == 2 else [val]))
      elif field == 3:
        h.scores.extend(_Floats(val) if wire == 2 else [pw.as_float(val)])
      elif field == 4:
        vec = []
        for f2, w2, v2 in pw.parse(val):
          if f2 == 1:
            vec.extend(_Floats(v2) if w2 == 2 else [pw.as_float(v2)])
        h.atten_vecs.append(vec)
      elif field == 5:
        h.normalized_score = pw.as_float(val)
    return h


def _Int32(v: int) -> int:
  v &= 0xFFFFFFFF
  return v - (1 << 32) if v >= (1 << 31) else v


def _Floats(buf: bytes) -> List[float]:
  return list(struct.unpack('<%df' % (len(buf) // 4), buf))


def HypsFromBeamSearchOuts(hyps, prev_hyps, done_hyps, scores, atten_probs, eos_scores,
                           eos_atten_probs, eos_id: int, num_hyps_per_beam: int):
  """Serialised `Hypothesis` for every terminated (step, hyp) cell, b'' elsewhere.

  hyps / prev_hyps / done_hyps / scores / eos_scores: `[T, K*B]`; atten_probs /
  eos_atten_probs: `[T, K*B, S]`. A hyp that terminates at step i in slot j is the path
  obtained by following `prev_hyps` back from j; its last token is `eos_id` with
  `eos_scores[i, j]` and `eos_atten_probs[i, j]`. Returns an object ndarray `[T, K*B]`.
  """
  to_np = lambda x: x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
  hyps, prev_hyps, done = to_np(hyps), to_np(prev_hyps), to_np(done_hyps).astype(bool)
  scores, eos_scores = to_np(scores).astype(np.float32), to_np(eos_scores).astype(np.float32)
  att, eos_att = to_np(atten_probs).astype(np.float32), to_np(eos_atten_probs).astype(np.float32)
  assert hyps.ndim == 2 and att.ndim == 3, (hyps.shape, att.shape)
  assert hyps.shape == prev_hyps.shape == done.shape == scores.shape == eos_scores.shape
  assert att.shape == eos_att.shape and att.shape[:2] == hyps.shape
  t, n = hyps.shape
  num_beams = n // num_hyps_per_beam
  out = np.full((t, n), b'', dtype=object)
  for i, j in zip(*np.nonzero(done)):
    # walk back through the parent pointers
    cur = j
    path = []
    for s in range(i - 1, -1, -1):
      path.append(cur)                       # slot used when reading step s
      cur = prev_hyps[s, cur]
    path = path[::-1]                        # path[s] = slot whose token was emitted at step s
    h = Hypothesis(beam_id=int(j % num_beams))
    for s in range(i):
      h.ids.append(int(hyps[s, path[s]]))
      h.scores.append(float(scores[s, path[s]]))
      h.atten_vecs.append(att[s, path[s]].tolist())
    h.ids.append(int(eos_id))
    h.scores.append(float(eos_scores[i, j]))
    h.atten_vecs.append(eos_att[i, j].tolist())
    out[i, j] = h.SerializeToString()
  return out


def UnpackHyp(in_hyps: Sequence[bytes], max_seq_length: int = 0):
  """→ (ids `[N, L]` zero padded, seq_lens `[N]`, scores `[N]` = normalized_score).
  L = max_seq_length, or the longest hyp when 0; longer hyps are truncated (ref `UnpackHyp`)."""
  parsed = [Hypothesis.FromString(bytes(h)) if h else Hypothesis() for h in in_hyps]
  width = max_seq_length or max([len(h.ids) for h in parsed] + [0])
  ids = np.zeros((len(parsed), width), np.int32)
  lens = np.zeros(len(parsed), np.int32)
  scores = np.zeros(len(parsed), np.float32)
  for r, h in enumerate(parsed):
    n = min(len(h.ids), width)
    ids[r, :n] = h.ids[:n]
    lens[r] = n
    scores[r] = h.normalized_score
  return ids, lens, scores
