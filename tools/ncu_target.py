"""Small ncu target: one short solve of a bench workload with library options set.
usage: python tools/ncu_target.py <workload> <maxit> [opt=value ...]   (single stream, no split)"""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_inputs_torch, WORKLOADS, LAMBDA, RHO
from torch_admm_deconv_b200 import fft_admm_tv, _lib
from torch_admm_deconv_b200.eops import deconv as D
D.SPLIT_STREAMS = 1
dev = torch.device("cuda:0")
wl, maxit = sys.argv[1], int(sys.argv[2])
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    _lib.set_option(k, int(v))
B, C, H, W, kind, k, sigma, _ = WORKLOADS[wl]
x, psf = make_inputs_torch((B, C, H, W), kind, k, sigma)
x = x.to(dev); kern = psf.to(dev)
lam = torch.tensor([LAMBDA], device=dev); rho = torch.tensor([RHO], device=dev)
out = fft_admm_tv(x, lam, rho, kern, False, maxit)
torch.cuda.synchronize()
print("ok", float(out.mean()))
