"""Functional recurrence (ref `lingvo/core/recurrent.py`).

`Recurrent(theta, state0, inputs, cell_fn, …)` computes

    state = state0
    for t in range(T):
      state, extras = cell_fn(theta, state, inputs[t])
      acc_state[t] = state
    return acc_state, state

over the leading (time) axis of every tensor in `inputs` (ref :985). `theta`, the states and
`inputs` are NestedMaps. Feature parity with the reference:

  * **memory**: the reference's backward loop re-runs `cell_fn` per step, so the forward keeps
    no activations. Here `remat_steps=k` re-materialises the forward in chunks of k steps
    (`torch.utils.checkpoint`): k=1 is the reference's O(1)-per-step footprint, larger k
    trades memory for less recompute, 0 keeps everything (plain autograd).
  * **`cell_grad`**: a hand-written step gradient
    `dtheta, dstate0, dinputs, dcaptured = cell_grad(theta, state0, inputs_t, extras, dstate1)`
    is honoured through a per-step autograd Function (ref :736).
  * **`stop_fn(t, theta, state)`** ends the loop early; the remaining rows of the accumulated
    state repeat the last state (ref :1015).
  * **`extras` / `return_acc_extras`**: the per-step extras can be accumulated too.
  * **`accumulator_layer`**: the layer's accumulators ride along in the state under the key
    `accumulators`, are restored before every step, and land in the layer again after the
    loop (ref `_AugmentState` :808, `_WrapAccumulatorCellFn` :824); their values carry no
    gradient.
  * **step seeds**: every step sees its own `py_utils` step seed so stateless random ops in
    `cell_fn` differ per step and agree between forward and rematerialised forward
    (ref `_WrapCellFnWithStepSeed` :866).
  * a single-time-step input skips the loop machinery (ref :965).

The loop is a host loop that launches the same kernels every step — under
`GraphedTrainStep` the whole unrolled recurrence is captured into one CUDA graph, which is
what removes the per-step launch latency the reference hides inside `tf.While`.

`StackedRecurrent` (ref :1423) pipelines a stack of recurrences over devices: layer i runs on
`devices[i]`; step t of layer i+1 consumes step t of layer i. With several CUDA devices in one
process the loop is issued time-major, layer-minor: kernel launches are asynchronous, so
device i works on step t while device i+1 works on step t−1 — the skewed schedule of the
reference's Send/Recv pipeline (for one-process-per-GPU pipelining see `parallel/pp.py`).
"""

from __future__ import annotations

from typing import Callable, List, Optional

import torch
from torch.utils import checkpoint as _ckpt

from lingvo_b200.core import py_utils
from lingvo_b200.core.nested_map import NestedMap

_ACC_KEY = 'accumulators'


def _Index(nmap, t):
  return nmap.Transform(lambda x: x[t])


def _SeqLen(nmap):
  for x in nmap.Flatten():
    return x.shape[0]
  raise ValueError('Recurrent inputs must contain at least one tensor')


def FlattenPadding(padding):
  """[T, B, 1] / [T, B] → [T]: a step counts as padding when every batch row is padded."""
  if padding is None:
    return None
  return padding.reshape(padding.shape[0], -1).min(dim=1).values


def _SeqPaddingLength(inputs):
  """Number of trailing time steps that are padding for the whole batch (ref :178)."""
  pad = inputs.get('padding') if isinstance(inputs, dict) else None
  if pad is None:
    return 0
  all_pad = FlattenPadding(pad) > 0.5
  if not bool(all_pad.any()):
    return 0
  rev = torch.flip(all_pad, [0]).to(torch.int32)
  return int(torch.cumprod(rev, 0).sum().item())


def _IsSingleTimeStep(inputs):
  return all(x.shape[0] == 1 for x in inputs.Flatten())


# -- cell decoration --------------------------------------------------------------------
def _WrapAccumulators(accumulator_layer, cell_fn):
  """Accumulators travel in `state.accumulators` (detached: no gradient through them)."""
  if accumulator_layer is None:
    return cell_fn

  def Wrapped(theta, state0, inputs):
    accumulator_layer.SetAccumulatorValues(state0[_ACC_KEY])
    inner = NestedMap({k: v for k, v in state0.items() if k != _ACC_KEY})
    state1, extras = cell_fn(theta, inner, inputs)
    state1 = NestedMap(state1)
    state1[_ACC_KEY] = accumulator_layer.GetAccumulatorValues().Transform(
        lambda v: v.detach() if isinstance(v, torch.Tensor) else v)
    return state1, extras

  return Wrapped


def _WrapStepSeed(cell_fn, base_seed):
  """Step t runs with step seed `base_seed + t` (also when re-run for rematerialisation)."""

  def Wrapped(theta, state0, inputs, t):
    py_utils.ResetStepSeed(base_seed + int(t))
    return cell_fn(theta, state0, inputs)

  return Wrapped


class _CellWithGrad(torch.autograd.Function):
  """One step whose backward is the user's `cell_grad`."""

  @staticmethod
  def forward(ctx, cell_fn, cell_grad, packs, n_theta, n_state, *flat):
    theta_t, state_t, in_t = packs
    theta = theta_t.Pack(list(flat[:n_theta]))
    state0 = state_t.Pack(list(flat[n_theta:n_theta + n_state]))
    inputs = in_t.Pack(list(flat[n_theta + n_state:]))
    with torch.no_grad():
      state1, extras = cell_fn(theta, state0, inputs)
    ctx.cell_grad = cell_grad
    ctx.packs = packs
    ctx.ns = (n_theta, n_state)
    ctx.extras = extras
    ctx.state1_tpl = state1
    ctx.save_for_backward(*flat)
    ctx.mark_non_differentiable(*[x for x in (extras.Flatten() if extras else [])
                                  if isinstance(x, torch.Tensor)])
    return tuple(state1.Flatten()) + tuple(extras.Flatten() if extras else ())

  @staticmethod
  def backward(ctx, *grads):
    flat = ctx.saved_tensors
    theta_t, state_t, in_t = ctx.packs
    n_theta, n_state = ctx.ns
    theta = theta_t.Pack(list(flat[:n_theta]))
    state0 = state_t.Pack(list(flat[n_theta:n_theta + n_state]))
    inputs = in_t.Pack(list(flat[n_theta + n_state:]))
    n1 = len(ctx.state1_tpl.Flatten())
    dstate1 = ctx.state1_tpl.Pack([
        g if g is not None else torch.zeros_like(s)
        for g, s in zip(grads[:n1], ctx.state1_tpl.Flatten())])
    out = ctx.cell_grad(theta, state0, inputs, ctx.extras, dstate1)
    dtheta, dstate0, dinputs = out[0], out[1], out[2]

    def Flat(nm, like):
      if nm is None:
        return [None] * len(like.Flatten())
      return [g if isinstance(g, torch.Tensor) else None for g in nm.Flatten()]

    return (None, None, None, None, None, *Flat(dtheta, theta_t), *Flat(dstate0, state_t),
            *Flat(dinputs, in_t))


def _WithCellGrad(cell_fn, cell_grad):
  if cell_grad is None:
    return cell_fn

  def Wrapped(theta, state0, inputs):
    flat = theta.Flatten() + state0.Flatten() + inputs.Flatten()
    probe_state1 = None
    res = _CellWithGrad.apply(cell_fn, cell_grad, (theta, state0, inputs), len(theta.Flatten()),
                              len(state0.Flatten()), *flat)
    # recover the structures: run metadata from a no-grad probe is avoided by packing with
    # the templates recorded in the Function (state1 has state0's structure by contract)
    n1 = len(state0.Flatten())
    state1 = state0.Pack(list(res[:n1]))
    extras = NestedMap()
    rest = list(res[n1:])
    if rest:
      extras = NestedMap({'extra_%d' % i: v for i, v in enumerate(rest)})
    del probe_state1
    return state1, extras

  return Wrapped


# -- the loop ----------------------------------------------------------------------------
def Recurrent(theta, state0, inputs, cell_fn: Callable, cell_grad=None, cell_type=None,
              stop_fn=None, extras=None, max_input_length=None, check_stateful_ops=False,
              accumulator_layer=None, allow_implicit_capture=False,
              allowed_tensor_captures=None, backward_cleanup=None, return_acc_extras=False,
              remat_steps: int = 0, skip_trailing_padding: bool = False):
  """Returns `(acc_state, final_state)` — plus `acc_extras` with `return_acc_extras`.
  See the module docstring."""
  del cell_type, check_stateful_ops, allow_implicit_capture, allowed_tensor_captures
  slen = _SeqLen(inputs)
  if max_input_length is not None:
    slen = min(slen, int(max_input_length))
  run_len = slen
  if skip_trailing_padding:
    run_len = max(slen - _SeqPaddingLength(inputs), 1)

  state = NestedMap(state0)
  if accumulator_layer is not None:
    if _ACC_KEY in state:
      raise ValueError('state0 already has the key %r' % _ACC_KEY)
    state[_ACC_KEY] = accumulator_layer.GetAccumulatorValues()
  state_tpl = state
  step_fn = _WithCellGrad(_WrapAccumulators(accumulator_layer, cell_fn), cell_grad)
  base_seed = py_utils.GetStepSeed()
  step = _WrapStepSeed(step_fn, base_seed)

  accs: List[NestedMap] = []
  acc_extras: List[NestedMap] = []
  stopped_at = run_len

  def _Run(st, t0, t1):
    outs, exs = [], []
    for t in range(t0, t1):
      st, ex = step(theta, st, _Index(inputs, t), t)
      outs.append(st)
      exs.append(ex if ex is not None else NestedMap())
    return st, outs, exs

  single = _IsSingleTimeStep(inputs)
  use_remat = (remat_steps and remat_steps > 0 and torch.is_grad_enabled() and
               stop_fn is None and not single and cell_grad is None)
  if use_remat:
    n = len(state_tpl.Flatten())
    t = 0
    while t < run_len:
      t1 = min(t + remat_steps, run_len)
      shapes = {}

      def _Chunk(*flat_state, _t0=t, _t1=t1, _shapes=shapes):
        st = state_tpl.Pack(list(flat_state))
        st, outs, exs = _Run(st, _t0, _t1)
        flat = []
        for o in outs:
          flat.extend(o.Flatten())
        _shapes['ex_tpl'] = exs[0]
        for e in exs:
          flat.extend(e.Flatten())
        return tuple(st.Flatten()) + tuple(flat)

      res = _ckpt.checkpoint(_Chunk, *state.Flatten(), use_reentrant=False)
      state = state_tpl.Pack(list(res[:n]))
      k = t1 - t
      rest = res[n:]
      for i in range(k):
        accs.append(state_tpl.Pack(list(rest[i * n:(i + 1) * n])))
      ex_tpl = shapes.get('ex_tpl', NestedMap())
      ne = len(ex_tpl.Flatten())
      ex_flat = rest[k * n:]
      for i in range(k):
        acc_extras.append(ex_tpl.Pack(list(ex_flat[i * ne:(i + 1) * ne])) if ne else NestedMap())
      t = t1
  else:
    t = 0
    while t < run_len:
      if stop_fn is not None and bool(stop_fn(t, theta, state)):
        stopped_at = t
        break
      state, outs, exs = _Run(state, t, t + 1)
      accs.extend(outs)
      acc_extras.extend(exs)
      t += 1
    if not accs:                       # stopped before the first step
      accs.append(state)
      acc_extras.append(extras if extras is not None else NestedMap())
  py_utils.ResetStepSeed(base_seed + slen)
  if backward_cleanup is not None:
    backward_cleanup()

  # rows not computed (early stop / all-padding tail) repeat the last state
  while len(accs) < slen:
    accs.append(state)
    acc_extras.append(acc_extras[-1] if acc_extras else NestedMap())
  del stopped_at

  def _Stack(maps):
    tpl = maps[0]
    cols = list(zip(*[m.Flatten() for m in maps])) if tpl.Flatten() else []
    return tpl.Pack([torch.stack(list(c), 0) if isinstance(c[0], torch.Tensor) else c[-1]
                     for c in cols])

  final = state
  if accumulator_layer is not None:
    accumulator_layer.SetAccumulatorValues(final[_ACC_KEY])
    final = NestedMap({k: v for k, v in final.items() if k != _ACC_KEY})
    accs = [NestedMap({k: v for k, v in a.items() if k != _ACC_KEY}) for a in accs]
  acc_state = _Stack(accs)
  if return_acc_extras:
    return acc_state, final, _Stack(acc_extras)
  return acc_state, final


# -- stacked / pipelined -------------------------------------------------------------------
def _ToDevice(nmap, device):
  if device is None:
    return nmap
  return nmap.Transform(lambda x: x.to(device, non_blocking=True)
                        if isinstance(x, torch.Tensor) else x)


def StackedRecurrent(devices, cell_fns, cell_grads, cell_outs, cell_out_grads, thetas,
                     init_states, inputs, accumulator_layers=None, unused_acc_state=False):
  """A stack of recurrences, layer i on `devices[i]` (ref :1423):

      for t:  x = inputs[t]
              for i:  state_i, _ = cell_fns[i](thetas[i], state_i, x);  x = cell_outs[i](state_i)

  Returns `(acc_out, final_states)`: the last layer's outputs stacked over time (`None` with
  `unused_acc_state`) and every layer's final state. `cell_outs[i]` maps a layer's state to
  the next layer's input NestedMap. Issued time-major so that distinct devices pipeline
  (see the module docstring); `devices` entries may be None / equal (no transfers)."""
  del cell_grads, cell_out_grads
  n_layers = len(cell_fns)
  devices = list(devices) if devices else [None] * n_layers
  assert len(devices) == n_layers == len(cell_outs) == len(thetas) == len(init_states)
  accumulator_layers = accumulator_layers or [None] * n_layers
  thetas = [_ToDevice(th, d) for th, d in zip(thetas, devices)]
  states = [NestedMap(_ToDevice(s, d)) for s, d in zip(init_states, devices)]
  for i, layer in enumerate(accumulator_layers):
    if layer is not None:
      states[i][_ACC_KEY] = layer.GetAccumulatorValues()
  fns = [_WrapAccumulators(l, f) for l, f in zip(accumulator_layers, cell_fns)]
  slen = _SeqLen(inputs)
  base_seed = py_utils.GetStepSeed()
  outs = []
  for t in range(slen):
    x = _ToDevice(_Index(inputs, t), devices[0])
    for i in range(n_layers):
      py_utils.ResetStepSeed(base_seed + t * n_layers + i)
      states[i], _ = fns[i](thetas[i], states[i], x)
      visible = NestedMap({k: v for k, v in states[i].items() if k != _ACC_KEY})
      out = cell_outs[i](visible)
      x = out if isinstance(out, NestedMap) else NestedMap(x=out)
      if i + 1 < n_layers:
        x = _ToDevice(x, devices[i + 1])
        if 'padding' in inputs and 'padding' not in x:
          x.padding = _ToDevice(NestedMap(p=inputs.padding[t]), devices[i + 1]).p
    if not unused_acc_state:
      outs.append(x)
  py_utils.ResetStepSeed(base_seed + slen * n_layers)
  finals = []
  for i, layer in enumerate(accumulator_layers):
    st = states[i]
    if layer is not None:
      layer.SetAccumulatorValues(st[_ACC_KEY])
      st = NestedMap({k: v for k, v in st.items() if k != _ACC_KEY})
    finals.append(st)
  if unused_acc_state:
    return None, finals
  tpl = outs[0]
  acc = tpl.Pack([torch.stack([o.Flatten()[k] for o in outs], 0)
                  for k in range(len(tpl.Flatten()))])
  return acc, finals
