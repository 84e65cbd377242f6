"""CPU tests of the multi-GPU host logic with the gloo backend, world_size 2 (no GPU needed): batch
sharding covers the batch exactly, sharded solves equal the unsharded solve (iso=False: no collective on
the solve path), and the parameter-gradient all-reduce sums over ranks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from torch_admm_deconv_b200.sharding import shard_range, shard_batch, allreduce_param_grads, gather_batch


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 4096):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from oracle import admm_oracle as O
        psf = O.make_psf("gauss", 5, 1.2)
        x = O.make_blurred((5, 2, 16, 16), psf, seed=3)              # 5 images over 2 ranks: 3 + 2
        full = O.admm_tv_spectral_form(x, 0.02, 0.04, psf[None, None], False, 6)
        mine = shard_batch(torch.from_numpy(x))
        lo, hi = shard_range(5, world, rank)
        assert mine.shape[0] == hi - lo
        part = O.admm_tv_spectral_form(mine.numpy(), 0.02, 0.04, psf[None, None], False, 6)   # no collective needed
        got = gather_batch(torch.from_numpy(part), 5).numpy()
        ok_solve = bool(np.array_equal(got, full))
        # gradient all-reduce of the layer parameters (w, lmbda, rho, b)
        from torch_admm_deconv_b200 import ADMMDeconv
        torch.manual_seed(0)
        m = ADMMDeconv((3, 3), max_iters=2, lmbda=None, rho=None, iso=False, bias=True)
        for i, p in enumerate(m.parameters()):
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        n = allreduce_param_grads(m.parameters())
        ok_grad = n == 9 + 3 and all(torch.allclose(p.grad, torch.full_like(p, 3.0 * (i + 1))) for i, p in enumerate(m.parameters()))
        # the launch-lean reducer: persistent flat buffer, one all-reduce, average over ranks
        from torch_admm_deconv_b200.sharding import GradAllReducer
        for i, p in enumerate(m.parameters()):
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        red = GradAllReducer(m.parameters(), average=True)
        n2 = red.reduce(); red.wait()
        ok_grad = ok_grad and n2 == 12 and all(torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 1))) for i, p in enumerate(m.parameters()))
        q.put((rank, ok_solve, ok_grad))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_solve_and_grad_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res), "sharded solve != unsharded solve"
    assert all(r[2] for r in res), "parameter gradient all-reduce wrong"
