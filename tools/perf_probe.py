"""Developer probe (not the bench): time the solve for a few shapes / tuning knobs with CUDA events.
usage: python tools/perf_probe.py [B C H W k maxit] ..."""
import itertools
import sys
import os
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_admm_deconv_b200 import fft_admm_tv, _lib  # noqa: E402

PEAK = 6551.0


def run(shape, k, maxit, reps=3):
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(1234)
    x = torch.rand(shape, generator=g).to(dev)
    kern = torch.rand(1, 1, k, k, generator=g).to(dev) if k else torch.empty(0, device=dev)
    if k:
        kern /= kern.sum()
    lam = torch.tensor([0.02], device=dev); rho = torch.tensor([0.04], device=dev)
    fft_admm_tv(x, lam, rho, kern, False, 3)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fft_admm_tv(x, lam, rho, kern, False, maxit)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    B, C, H, W = shape
    el = B * C * H * W
    gbs = 36.0 * el * maxit / (best * 1e-3) / 1e9
    mpix = B * H * W * maxit / (best * 1e-3) / 1e6
    return best, gbs, mpix, bool(torch.isfinite(out).all())


def main():
    cases = [((64, 3, 512, 512), 31, 20), ((256, 3, 256, 256), 15, 20), ((1, 1, 256, 256), 15, 50),
             ((1, 3, 2160, 3840), 63, 4)]
    sweeps = {"rows_per_band": [8, 16, 32], "cols_per_tile": [8, 16, 32], "threads": [256, 512]}
    for shape, k, maxit in cases:
        print("== shape", shape, "k", k, "maxit", maxit, flush=True)
        base, gbs, mpix, ok = run(shape, k, maxit)
        print("  default: %.3f ms  %.3f ms/it  %.1f GB/s algorithmic (%.1f%% of %.0f)  %.1f Mpix-it/s finite=%s"
              % (base, base / maxit, gbs, 100 * gbs / PEAK, PEAK, mpix, ok), flush=True)
        if shape[-1] > 1024:
            continue
        for key, vals in sweeps.items():
            for v in vals:
                _lib.set_option(key, v)
                t, gbs, mpix, ok = run(shape, k, maxit)
                print("  %s=%d: %.3f ms/it  %.1f GB/s (%.1f%%)" % (key, v, t / maxit, gbs, 100 * gbs / PEAK), flush=True)
            _lib.set_option(key, 0)


if __name__ == "__main__":
    main()
