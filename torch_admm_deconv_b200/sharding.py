"""Multi-GPU plumbing for the ADMM-TV path: one process per GPU, contiguous batch split.

For iso=False the planes of a batch are independent (bit-identical to per-item calls), so the solve needs
NO collective: every rank runs `fft_admm_tv` / `ADMMDeconv` on its own slice.  The only exchange is the
optional gradient all-reduce of the layer's (tiny) parameters when the unrolled layer is trained:
`w` (k*k floats), `lmbda`, `rho`, `b` -- one NCCL call per step over NVLink.  iso=True couples the batch
through `pixelnorm` (reference deconv.py:23-24); here it is computed per shard (documented semantics).
"""
from __future__ import annotations

from typing import Iterable, Optional, Tuple

import torch
import torch.distributed as dist

__all__ = ["shard_range", "shard_batch", "allreduce_param_grads", "GradAllReducer", "gather_batch"]


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced split of n items: the first n % world ranks get one extra item."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank %r/%r" % (world, rank))
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(x: torch.Tensor, world: Optional[int] = None, rank: Optional[int] = None) -> torch.Tensor:
    """This rank's slice of the batch dimension (a view, no copy)."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_range(x.shape[0], world, rank)
    return x[lo:hi]


def allreduce_param_grads(params: Iterable[torch.nn.Parameter], group=None, average: bool = False) -> int:
    """Sum (or average) the gradients of the given parameters over all ranks with ONE all-reduce.
    Returns the number of floats exchanged.  Parameters without a gradient contribute zeros."""
    params = [p for p in params if p.requires_grad]
    if not params or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    dev = params[0].device
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).to(torch.float32) for p in params])
    flat = flat.to(dev)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off:off + n].reshape(p.shape).to(p.dtype)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
    return int(flat.numel())


class GradAllReducer:
    """All-reduce of the layer's (tiny) parameter gradients with the minimum number of launches: one multi-tensor copy
    into a persistent flat fp32 buffer, ONE NCCL all-reduce (ReduceOp.AVG / SUM, so no separate scaling kernel), one
    multi-tensor copy back -- enqueued on a side stream so it overlaps whatever the caller does next; `wait()` makes the
    current stream wait for it (call it before the optimizer step).  Replaces the per-step `torch.cat` / divide /
    per-parameter `copy_` sequence of `allreduce_param_grads` in training loops (etrain/trainer.py:53-70 has none: the
    reference is single-process)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None, average: bool = True):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.average = average
        self.enabled = bool(self.params) and dist.is_initialized() and dist.get_world_size(group) > 1
        if not self.enabled:
            return
        dev = self.params[0].device
        sizes = [p.numel() for p in self.params]
        self.flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
        self.views = [v.view(p.shape) for v, p in zip(self.flat.split(sizes), self.params)]
        self.stream = torch.cuda.Stream(dev) if dev.type == "cuda" else None
        # gloo (CPU tests) has no AVG
        self.op = dist.ReduceOp.AVG if (average and dist.get_backend(group) == "nccl") else dist.ReduceOp.SUM
        self.scale = (1.0 / dist.get_world_size(group)) if (average and self.op == dist.ReduceOp.SUM) else None

    def reduce(self) -> int:
        """Enqueue the exchange of the current `.grad`s.  Returns the number of floats exchanged."""
        if not self.enabled:
            return 0
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        if self.stream is None:
            torch._foreach_copy_(self.views, grads)
            dist.all_reduce(self.flat, op=self.op, group=self.group)
            if self.scale is not None:
                self.flat.mul_(self.scale)
            for p, v in zip(self.params, self.views):
                if p.grad is None:
                    p.grad = v.clone()
            torch._foreach_copy_([p.grad for p in self.params], self.views)
            return int(self.flat.numel())
        cur = torch.cuda.current_stream(self.flat.device)
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            torch._foreach_copy_(self.views, grads)
            dist.all_reduce(self.flat, op=self.op, group=self.group)
            if self.scale is not None:
                self.flat.mul_(self.scale)
            for p in self.params:
                if p.grad is None:
                    p.grad = torch.empty_like(p)
            torch._foreach_copy_([p.grad for p in self.params], self.views)
        for g in grads:
            g.record_stream(self.stream)
        return int(self.flat.numel())

    def wait(self) -> None:
        if self.enabled and self.stream is not None:
            torch.cuda.current_stream(self.flat.device).wait_stream(self.stream)


def gather_batch(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """All-gather the per-rank slices back into the full batch (utility for evaluation; not on the solve path)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(total, world, r) for r in range(world)]
    maxn = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((maxn,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(outs, sizes)], dim=0)
