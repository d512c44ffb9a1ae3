"""Device-resident batches: torch owns the HBM buffers and the CUDA stream, the C ABI gets raw pointers.

PyTorch is plumbing here (device memory, streams, ``torch.distributed``); every kernel that runs is ours
(``smrt_b200/csrc``).  Nothing in this module computes on the CPU.
"""

from __future__ import annotations

import numpy as np

from . import capi
from .error import SMRTError
from .pack import MODE_PASSIVE, ProblemBatch


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise SMRTError("no CUDA device is available: the DORT path of smrt_b200 runs on the GPU only")
    return torch


class DeviceBatch:
    """A ProblemBatch uploaded to one GPU plus its output tensors, and the ``smrtb200_batch`` struct pointing at them."""

    def __init__(self, batch: ProblemBatch, n_max_stream: int, device: int = 0, pinned_source: bool = False):
        torch = _torch()
        self.torch = torch
        self.device = torch.device("cuda", device)
        self.batch = batch
        self.n_max_stream = int(n_max_stream)
        self.host = {}
        self.dev = {}
        self.h2d_bytes = 0
        for fld, attr, dt in capi.INPUT_FIELDS:
            if getattr(batch, attr) is None:  # optional input the batch does not carry: NULL in the struct
                continue
            arr = np.ascontiguousarray(getattr(batch, attr), dtype=dt)
            if dt == np.complex128:
                arr = arr.view(np.float64).reshape(arr.shape + (2,))
            t = torch.from_numpy(arr)
            if pinned_source:
                t = t.pin_memory()
            self.host[fld] = t
            self.h2d_bytes += t.numel() * t.element_size()
        self.host["theta"] = torch.from_numpy(np.ascontiguousarray(batch.theta, dtype=np.float64))
        self.host["theta_inc"] = torch.from_numpy(np.ascontiguousarray(
            batch.theta_inc if len(batch.theta_inc) else np.zeros(1), dtype=np.float64))
        B, L = batch.B, batch.L
        kw = dict(device=self.device)
        if batch.mode == MODE_PASSIVE:
            self.values = torch.zeros((B, 2, len(batch.theta)), dtype=torch.float64, **kw)
        else:
            self.values = torch.zeros((B, 3, 3, len(batch.theta_inc)), dtype=torch.float64, **kw)
        self.ks = torch.zeros((B, L), dtype=torch.float64, **kw)
        self.ka = torch.zeros((B, L), dtype=torch.float64, **kw)
        self.eps_eff = torch.zeros((B, L, 2), dtype=torch.float64, **kw)
        self.n_streams = torch.zeros(B, dtype=torch.int32, **kw)
        self.stream_angles = torch.zeros((B, self.n_max_stream), dtype=torch.float64, **kw)
        self.optical_depth = torch.zeros(B, dtype=torch.float64, **kw)
        self.status = torch.zeros(B, dtype=torch.int32, **kw)
        self.upload()

    def upload(self, non_blocking: bool = True):
        """host -> device copy of every input array (on torch's current stream)"""
        for k, t in self.host.items():
            if k in self.dev and self.dev[k].shape == t.shape:
                self.dev[k].copy_(t, non_blocking=non_blocking)
            else:
                self.dev[k] = t.to(self.device, non_blocking=non_blocking)

    def struct(self) -> capi.Batch:
        bt = capi.Batch()
        bt.B = self.batch.B
        for fld, _, _ in capi.INPUT_FIELDS:
            if fld in self.dev:
                setattr(bt, fld, self.dev[fld].data_ptr())
        bt.theta = self.dev["theta"].data_ptr()
        bt.theta_inc = self.dev["theta_inc"].data_ptr()
        bt.phi = float(self.batch.phi)
        bt.values, bt.ks, bt.ka = self.values.data_ptr(), self.ks.data_ptr(), self.ka.data_ptr()
        bt.eps_eff, bt.n_streams_out = self.eps_eff.data_ptr(), self.n_streams.data_ptr()
        bt.stream_angles, bt.optical_depth = self.stream_angles.data_ptr(), self.optical_depth.data_ptr()
        bt.status = self.status.data_ptr()
        return bt

    def outputs_to_host(self) -> capi.HostOutputs:
        out = capi.HostOutputs(self.batch, self.n_max_stream)
        out.values[...] = self.values.cpu().numpy()
        out.ks[...] = self.ks.cpu().numpy()
        out.ka[...] = self.ka.cpu().numpy()
        out.eps_eff[...] = self.eps_eff.cpu().numpy().view(np.complex128)[..., 0]
        out.n_streams[...] = self.n_streams.cpu().numpy()
        out.stream_angles[...] = self.stream_angles.cpu().numpy()
        out.optical_depth[...] = self.optical_depth.cpu().numpy()
        out.status[...] = self.status.cpu().numpy()
        return out
