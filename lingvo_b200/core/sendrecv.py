"""Point-to-point tensor channel (ref `lingvo/core/sendrecv.py:37`).

`Channel(dtype, shape, send_rank, recv_rank).Send(x)` / `.Recv()` over
`torch.distributed` P2P (NCCL on GPUs → NVLink; gloo on CPU). Both ends construct
the same Channel; shapes are static so the receiver allocates without a handshake.
"""
import torch
import torch.distributed as dist


class Channel:

  def __init__(self, dtype, shape, send_device, recv_device, name=None, group=None):
    self._dtype, self._shape = dtype, tuple(shape)
    self._send, self._recv = int(send_device), int(recv_device)
    self._name, self._group = name, group

  def Send(self, tensor):
    assert tuple(tensor.shape) == self._shape, (tensor.shape, self._shape)
    if self._send == self._recv:
      self._loop = tensor
      return None
    return dist.isend(tensor.to(self._dtype).contiguous(), self._recv, group=self._group)

  def Recv(self, device=None):
    if self._send == self._recv:
      return self._loop
    out = torch.empty(self._shape, dtype=self._dtype, device=device)
    dist.recv(out, self._send, group=self._group)
    return out
