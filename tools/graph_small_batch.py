"""Latency-bound small batches (the reference's eval / training shapes, iso=True, 100 iterations): eager launches vs a CUDA
graph of the whole forward, and per-iteration kernel times."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_admm_deconv_b200 import ADMMDeconv, _lib
dev = torch.device("cuda:0")


def med(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


for shape, iso in (((8, 3, 256, 256), True), ((3, 3, 256, 256), True), ((8, 3, 256, 256), False), ((1, 3, 512, 512), True)):
    m = ADMMDeconv((), max_iters=100, lmbda=0.02, rho=0.04, iso=iso).to(dev)
    x = torch.rand(shape, device=dev)
    with torch.inference_mode():
        eager = med(lambda: m(x))
        for pdl in (0, 1):
            _lib.set_option("use_pdl", pdl)
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                m(x)
            torch.cuda.current_stream().wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                y = m(x)
            t = med(g.replay)
            print("%s iso=%s: eager %.3f ms; CUDA graph (use_pdl=%d) %.3f ms" % (shape, iso, eager, pdl, t), flush=True)
        _lib.set_option("use_pdl", 1)
