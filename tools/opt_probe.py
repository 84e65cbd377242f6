"""A/B of a library option on a benchmark workload: ms per solve and per-kernel ms (CUDA events, one stream).
usage: python tools/opt_probe.py <workload> <option> <v0,v1,...>"""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_inputs_torch, WORKLOADS, LAMBDA, RHO
from torch_admm_deconv_b200 import fft_admm_tv, _lib
from torch_admm_deconv_b200.eops import deconv as D
D.SPLIT_STREAMS = 1
dev = torch.device("cuda:0")
wl, opt, vals = sys.argv[1], sys.argv[2], [int(v) for v in sys.argv[3].split(",")]
# a workload name of bench.py, or an ad-hoc shape "B,C,H,W,k,maxit" (gauss PSF)
B, C, H, W, kind, k, sigma, maxit = WORKLOADS[wl] if wl in WORKLOADS else tuple(int(v) for v in wl.split(",")[:4]) + ("gauss", int(wl.split(",")[4]), 2.0, int(wl.split(",")[5]))
x, psf = make_inputs_torch((B, C, H, W), kind, k, sigma)
x = x.to(dev); kern = psf.to(dev)
lam = torch.tensor([LAMBDA], device=dev); rho = torch.tensor([RHO], device=dev)
ref = None
for rep in range(2):
    for v in vals:
        _lib.set_option(opt, v)
        for _ in range(2):
            out = fft_admm_tv(x, lam, rho, kern, False, maxit)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            out = fft_admm_tv(x, lam, rho, kern, False, maxit)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        _lib.set_option("profile", 1); _lib.profile_reset()
        for _ in range(2):
            fft_admm_tv(x, lam, rho, kern, False, maxit)
        torch.cuda.synchronize()
        r, c = _lib.profile_read(0), _lib.profile_read(1)
        _lib.set_option("profile", 0)
        if ref is None:
            ref = out.clone()
        print("%s %s=%d: %.3f ms per solve (%.1f Gpixel-it/s); rows %.1f us, cols %.1f us per launch; same=%s" %
              (wl, opt, v, ms, B * H * W * maxit / ms / 1e6, r[0] / max(r[1], 1) * 1e3, c[0] / max(c[1], 1) * 1e3,
               bool(torch.equal(out, ref))), flush=True)
