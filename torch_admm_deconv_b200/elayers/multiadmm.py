"""Multi-solver fan-out: several `ADMMDeconv` layers applied to the SAME input, concatenated along channels.

Mirrors the reference containers `MultiADMM` (/root/reference/src/admmtor/modelbuild/blocks.py:252-261, attribute
`admms`) and `Deconvs` (modelbuild/deconver.py:8-23, attribute `blocks`); `ADMMFusion` (elayers/admmfusion.py:30-35)
builds the same list before its attention pooling.  The reference runs the solvers one after the other; with the
small batches it trains on (3 x 3 x 256 x 256) each solve fills only a fraction of a B200, so here the solvers are
enqueued on separate CUDA streams and overlap on the device.  Results are identical to the sequential loop.
"""
from __future__ import annotations

from typing import Dict, List

import torch

from .admmdeconv import ADMMDeconv

__all__ = ["MultiADMM", "Deconvs"]


def _fanout(mods, x: torch.Tensor, concurrent: bool) -> torch.Tensor:
    if not (concurrent and x.is_cuda and len(mods) > 1):
        return torch.cat([m(x) for m in mods], dim=1)
    cur = torch.cuda.current_stream(x.device)
    outs = []
    streams = [torch.cuda.Stream(x.device) for _ in mods]
    for m, s in zip(mods, streams):
        s.wait_stream(cur)                       # x is ready
        with torch.cuda.stream(s):
            o = m(x)
        x.record_stream(s)
        outs.append(o)
    for o, s in zip(outs, streams):
        cur.wait_stream(s)
        o.record_stream(cur)
    return torch.cat(outs, dim=1)


class MultiADMM(torch.nn.Module):
    def __init__(self, admm_dicts: List[Dict], concurrent: bool = True):
        super().__init__()
        self.admms = torch.nn.ModuleList(ADMMDeconv(**d) for d in admm_dicts)
        self.concurrent = concurrent

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return _fanout(list(self.admms), x, self.concurrent)


class Deconvs(torch.nn.Module):
    def __init__(self, admms_args: List[Dict], concurrent: bool = True):
        super().__init__()
        self.blocks = torch.nn.ModuleList(ADMMDeconv(**d) for d in admms_args)
        self.concurrent = concurrent

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return _fanout(list(self.blocks), x, self.concurrent)
