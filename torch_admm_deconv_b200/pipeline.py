"""Host-side streaming helper: solve a sequence of HOST batches with copies overlapped with compute.

`HostPipeline` keeps three CUDA streams: pinned-host -> device copies, the solves (the ordinary public call
`fft_admm_tv`, back to back on ONE stream so kernels of different batches never interleave), and device -> host
copies.  Device input/output buffers are double-buffered and guarded by events, so the copy of batch i+1 and the
read-back of batch i-1 overlap with the solve of batch i.
"""
from __future__ import annotations

import torch

from .eops.deconv import fft_admm_tv

__all__ = ["HostPipeline"]


class HostPipeline:
    def __init__(self, device, lmbd: torch.Tensor, rho: torch.Tensor, kern: torch.Tensor, iso: bool = False,
                 maxit: int = 100, depth: int = 2):
        self.device = torch.device(device)
        self.lmbd, self.rho, self.kern = lmbd, rho, kern
        self.iso, self.maxit = iso, maxit
        self.depth = depth
        self.s_in = torch.cuda.Stream(self.device)
        self.s_solve = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self.streams = [self.s_in, self.s_solve, self.s_out]
        self.dev_in = [None] * depth
        self.dev_out = [None] * depth
        self.ev_in = [None] * depth        # H2D of the slot finished
        self.ev_solved = [None] * depth    # solve of the slot finished (input buffer reusable)
        self.ev_out = [None] * depth       # D2H of the slot finished (output buffer reusable)
        self.i = 0

    def submit(self, x_host: torch.Tensor, out_host: torch.Tensor) -> None:
        """Enqueue H2D(x_host) -> solve -> D2H(out_host); both host tensors must be pinned.  Returns immediately;
        `out_host` is valid after `synchronize()`."""
        k = self.i % self.depth
        self.i += 1
        with torch.cuda.stream(self.s_in):
            if self.dev_in[k] is None or self.dev_in[k].shape != x_host.shape:
                self.dev_in[k] = torch.empty(x_host.shape, dtype=x_host.dtype, device=self.device)
            if self.ev_solved[k] is not None:
                self.s_in.wait_event(self.ev_solved[k])            # previous user of this input slot is done
            self.dev_in[k].copy_(x_host, non_blocking=True)
            self.ev_in[k] = torch.cuda.Event(); self.ev_in[k].record(self.s_in)
        with torch.cuda.stream(self.s_solve):
            self.s_solve.wait_event(self.ev_in[k])
            if self.ev_out[k] is not None:
                self.s_solve.wait_event(self.ev_out[k])            # previous result of this slot has been read back
            self.dev_out[k] = fft_admm_tv(self.dev_in[k], self.lmbd, self.rho, self.kern, self.iso, self.maxit)
            self.ev_solved[k] = torch.cuda.Event(); self.ev_solved[k].record(self.s_solve)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_solved[k])
            out_host.copy_(self.dev_out[k], non_blocking=True)
            self.dev_out[k].record_stream(self.s_out)
            self.ev_out[k] = torch.cuda.Event(); self.ev_out[k].record(self.s_out)

    def synchronize(self) -> None:
        for s in self.streams:
            s.synchronize()
