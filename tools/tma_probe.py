import os, sys, json, subprocess, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_admm_deconv_b200 import fft_admm_tv, _lib
dev = torch.device("cuda:0")
def run(shape, k, maxit):
    g = torch.Generator().manual_seed(1)
    x = torch.rand(shape, generator=g).to(dev)
    kern = torch.rand(1, 1, k, k, generator=g).to(dev); kern /= kern.sum()
    lam = torch.tensor([0.02], device=dev); rho = torch.tensor([0.04], device=dev)
    outs = {}
    for tma in (0, 1):
        _lib.set_option("use_tma", tma)
        fft_admm_tv(x, lam, rho, kern, False, 3); torch.cuda.synchronize()
        _lib.set_option("profile", 1); _lib.profile_reset()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); out = fft_admm_tv(x, lam, rho, kern, False, maxit); e1.record(); torch.cuda.synchronize()
        rows = _lib.profile_read(0); cols = _lib.profile_read(1)
        _lib.set_option("profile", 0)
        outs[tma] = out
        print("shape %s tma=%d: %.3f ms total, rows %.1f us/launch, cols %.1f us/launch" %
              (shape, tma, e0.elapsed_time(e1), 1e3 * rows[0] / max(rows[1], 1), 1e3 * cols[0] / max(cols[1], 1)), flush=True)
    print("   max |tma - plain| = %.3e" % float((outs[0] - outs[1]).abs().max()), flush=True)
run((2, 3, 128, 128), 5, 10)
run((64, 3, 512, 512), 31, 30)
run((256, 3, 256, 256), 15, 30)
run((1, 1, 256, 256), 15, 30)
