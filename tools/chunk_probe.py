"""L2-resident plane chunks: time a solve for several working-set budgets (option chunk_mb) and check bit-equality."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_inputs_torch, WORKLOADS, LAMBDA, RHO
from torch_admm_deconv_b200 import fft_admm_tv, _lib

dev = torch.device("cuda:0")
budgets = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 16, 28, 42, 56, 70, 84, 98, 112]
for wl in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["cfg2", "cfg5"]):
    B, C, H, W, kind, k, sigma, maxit = WORKLOADS[wl]
    x, psf = make_inputs_torch((B, C, H, W), kind, k, sigma)
    x = x.to(dev); kern = psf.to(dev)
    lam = torch.tensor([LAMBDA], device=dev); rho = torch.tensor([RHO], device=dev)
    ref = None
    for mb in budgets:
        _lib.set_option("chunk_mb", mb)
        out = fft_admm_tv(x, lam, rho, kern, False, maxit)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            out = fft_admm_tv(x, lam, rho, kern, False, maxit)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        if ref is None:
            ref = out.clone()
        same = bool(torch.equal(out, ref))
        per = H * W * 4 * 7 / 2**20
        print("%s chunk_mb=%3d (%5.1f planes of %d): %8.3f ms per solve  %7.1f Gpixel-it/s  bit-identical=%s"
              % (wl, mb, mb / per if mb else B * C, B * C, ms, B * H * W * maxit / ms / 1e6, same), flush=True)
    _lib.set_option("chunk_mb", -1)
