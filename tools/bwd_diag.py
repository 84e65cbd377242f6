"""Backward diagnostics: gradient error vs the fp64 oracle adjoint as a function of the plane count and band height."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import admm_oracle as O
from torch_admm_deconv_b200 import fft_admm_tv, _lib

dev = torch.device("cuda:0")


def grads(x, gout, maxit, iso=False, k=None):
    xt = torch.tensor(x, device=dev, requires_grad=True)
    lt = torch.tensor([0.02], device=dev, requires_grad=True)
    rt = torch.tensor([0.04], device=dev, requires_grad=True)
    kt = torch.empty(0, device=dev) if k is None else torch.tensor(k, device=dev, requires_grad=True)
    out = fft_admm_tv(xt, lt, rt, kt, iso, maxit)
    (out * torch.tensor(gout.astype(np.float32), device=dev)).sum().backward()
    torch.cuda.synchronize()
    return xt.grad.cpu().numpy().astype(np.float64), float(lt.grad), float(rt.grad)


for B, H, W in [(1, 256, 256), (4, 256, 256), (16, 256, 256), (32, 256, 256), (8, 128, 128), (8, 512, 512)]:
    maxit = 10
    shape = (B, 3, H, W)
    rng = np.random.default_rng(40)
    x = O.make_blurred(shape, None, seed=4, noise=0.02)
    gout = rng.standard_normal(shape)
    g64 = O.admm_tv_backward(x.astype(np.float64), 0.02, 0.04, np.zeros((0,)), gout, False, maxit)
    for tag, opts in [("default", {}), ("generic", {"force_generic": 1}), ("nopdl", {"use_pdl": 0}), ("thin bands", {"rows_per_band": 4})]:
        for kk, vv in opts.items():
            _lib.set_option(kk, vv)
        try:
            gx, gl, gr = grads(x, gout, maxit)
        finally:
            for kk in opts:
                _lib.set_option(kk, {"use_pdl": 1}.get(kk, 0))
        d = np.abs(gx - g64[0]); m = np.abs(g64[0]).max()
        bad = np.argwhere(d > 1e-3 * m)
        rows = np.unique(bad[:, 2]) if len(bad) else []
        print("%s %-10s gx err %.2e  bad %d  planes %s rows %s  glam %.6g/%.6g grho %.6g/%.6g" % (
            shape, tag, d.max() / m, len(bad), np.unique(bad[:, 0] * 3 + bad[:, 1])[:8] if len(bad) else [],
            list(rows[:12]), gl, g64[1], gr, g64[2]), flush=True)
