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

__all__ = ["shard_range", "shard_batch", "allreduce_param_grads", "gather_batch"]


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
