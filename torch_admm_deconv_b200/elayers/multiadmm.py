"""Multi-solver fan-out: several `ADMMDeconv` layers applied to the SAME input, concatenated along channels.

Mirrors the reference containers `MultiADMM` (/root/reference/src/admmtor/modelbuild/blocks.py:252-261, attribute
`admms`), `Deconvs` (modelbuild/deconver.py:8-23, attribute `blocks`) and `ADMMFusion` (elayers/admmfusion.py:9-40,
attributes `admms`, `acp`).  The reference runs the solvers one after the other and `torch.cat`s the results.  Here

  * F(x) -- the row R2C and column FFT of the shared input -- is computed ONCE (`shared_spectrum`, C ABI
    `admm_spectrum_forward`) and every solver starts from it with its own tables (`admm_ext.yhat_in`);
  * every solver writes its result straight into its channel slice of the concatenated output
    (`admm_ext.out_batch_stride`): no `torch.cat` copy (inference; under autograd the slice is filled with a tracked copy);
  * with the small batches the reference trains on (3 x 3 x 256 x 256) one solve fills a fraction of a B200, so the
    solvers are enqueued on side streams (created once per device and cached) and overlap on the device.

Results are identical to the sequential loop of the reference.
"""
from __future__ import annotations

from typing import Dict, List

import torch

from ..eops.deconv import shared_spectrum
from .admmdeconv import ADMMDeconv

__all__ = ["MultiADMM", "Deconvs", "ADMMFusion", "fanout"]

_STREAMS: Dict[tuple, List[torch.cuda.Stream]] = {}


def _side_streams(device: torch.device, n: int) -> List[torch.cuda.Stream]:
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    pool = _STREAMS.setdefault(key, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device))
    return pool[:n]


def fanout(mods, x: torch.Tensor, concurrent: bool = True) -> torch.Tensor:
    """torch.cat([m(x) for m in mods], dim=1) for `ADMMDeconv` modules, computed as described in the module docstring."""
    mods = list(mods)
    if not x.is_cuda or x.dim() != 4:
        return torch.cat([m(x) for m in mods], dim=1)
    B, C, H, W = x.shape
    big = torch.empty((B, len(mods) * C, H, W), dtype=torch.float32, device=x.device)
    yhat = shared_spectrum(x) if len(mods) > 1 else None
    # every slice view is taken right before its solver runs: under autograd the first tracked copy into `big` changes
    # the autograd state of the base, and views made before that would be stale
    if not (concurrent and len(mods) > 1):
        for i, m in enumerate(mods):
            m(x, out=big[:, i * C:(i + 1) * C], yhat=yhat)
        return big
    cur = torch.cuda.current_stream(x.device)
    streams = _side_streams(x.device, len(mods))
    for i, (m, s) in enumerate(zip(mods, streams)):
        s.wait_stream(cur)                       # x, the shared spectrum and the output buffer are ready
        with torch.cuda.stream(s):
            m(x, out=big[:, i * C:(i + 1) * C], yhat=yhat)
        for t in (x, big, yhat):
            if t is not None:
                t.record_stream(s)
    for s in streams:
        cur.wait_stream(s)
    return big


class MultiADMM(torch.nn.Module):
    """modelbuild/blocks.py:252-261."""

    def __init__(self, admm_dicts: List[Dict], concurrent: bool = True):
        super().__init__()
        self.admms = torch.nn.ModuleList(ADMMDeconv(**d) for d in admm_dicts)
        self.concurrent = concurrent

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return fanout(self.admms, x, self.concurrent)


class Deconvs(torch.nn.Module):
    """modelbuild/deconver.py:8-23."""

    def __init__(self, admms_args: List[Dict], concurrent: bool = True):
        super().__init__()
        self.blocks = torch.nn.ModuleList(ADMMDeconv(**d) for d in admms_args)
        self.concurrent = concurrent

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return fanout(self.blocks, x, self.concurrent)


class ADMMFusion(torch.nn.Module):
    """elayers/admmfusion.py:9-40: several solvers on the same input, concatenated, then attention channel pooling.

    The pooling (`AttentionChannelPooling` over `ChannelWiseAttention`, elayers/attentionpool.py, cwa.py) is a plain CNN
    block outside the solver path; it is not re-implemented here.  Pass the module as `acp`, or leave it None to build the
    reference's own class from an installed `admmtor` package with the reference's constructor arguments (same state-dict
    keys `admms.N.{w,lmbda,rho,b}`, `acp.*`, so reference checkpoints load with strict=True)."""

    def __init__(self, admms_cfgs: List[Dict], in_channels: int, compressions=None, probas_channels_factor: int = 2,
                 reduce_probas_space: bool = False, with_admms: bool = False, acp: torch.nn.Module = None,
                 concurrent: bool = True):
        super().__init__()
        self.in_channels = in_channels
        self.admms_cfgs = admms_cfgs
        self.with_admms = with_admms
        self.probas_channels_factor = probas_channels_factor
        self.reduce_probas_space = reduce_probas_space
        self.fusioned_channels_size = in_channels * len(admms_cfgs)
        self.admms = torch.nn.ModuleList(ADMMDeconv(**c) for c in admms_cfgs)       # admmfusion.py:28-30
        if acp is None:
            try:
                from admmtor.elayers.attentionpool import AttentionChannelPooling
                from admmtor.elayers.cwa import ChannelCompression
            except ImportError as e:  # pragma: no cover
                raise ImportError("ADMMFusion needs the reference's AttentionChannelPooling (package `admmtor`) or an "
                                  "`acp=` module; only the ADMM solvers are provided by torch_admm_deconv_b200") from e
            if compressions is None:                                               # admmfusion.py:13-14
                compressions = (ChannelCompression.STD, ChannelCompression.MEDIAN, ChannelCompression.MAX,
                                ChannelCompression.MEAN)
            acp = AttentionChannelPooling(self.fusioned_channels_size, in_channels, compressions, probas_channels_factor,
                                          reduce_probas_space)                      # admmfusion.py:31-32
        self.compressions = compressions
        self.acp = acp
        self.concurrent = concurrent

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = fanout(self.admms, x, self.concurrent)                                  # admmfusion.py:35
        if self.with_admms:
            return torch.cat([self.acp(x), x], dim=1)
        return self.acp(x)
