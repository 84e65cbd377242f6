"""cfg1 and small batches: cluster-resident solver against the two-kernel path (ms per solve, CUDA events, L2 flushed)."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_inputs_torch, LAMBDA, RHO
from torch_admm_deconv_b200 import fft_admm_tv, _lib
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
lam = torch.tensor([LAMBDA], device=dev); rho = torch.tensor([RHO], device=dev)
for (B, C, H, W, maxit) in [(1, 1, 256, 256, 50), (1, 3, 256, 256, 50), (3, 3, 256, 256, 100), (8, 3, 256, 256, 50), (16, 3, 256, 256, 50),
                            (32, 3, 256, 256, 50), (1, 1, 128, 128, 50), (8, 3, 128, 128, 50)]:
    x, psf = make_inputs_torch((B, C, H, W), "gauss", 15, 2.5)
    x = x.to(dev); kern = psf.to(dev)
    res = {}
    for mode in (0, 2):
        _lib.set_option("use_cluster", mode)
        for _ in range(3):
            fft_admm_tv(x, lam, rho, kern, False, maxit)
        ts = []
        for _ in range(10):
            flush.fill_(1)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); fft_admm_tv(x, lam, rho, kern, False, maxit); e1.record()
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        res[mode] = sorted(ts)[len(ts) // 2]
    _lib.set_option("use_cluster", 1)
    print("%dx%dx%dx%d N=%d: two-kernel %.3f ms, cluster %.3f ms  (x%.2f); cluster: %.2f us per plane-iteration" %
          (B, C, H, W, maxit, res[0], res[2], res[0] / res[2], res[2] * 1e3 / maxit / (B * C)), flush=True)
