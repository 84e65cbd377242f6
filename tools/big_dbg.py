import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_admm_deconv_b200 import fft_admm_tv, _lib
dev = torch.device("cuda:0")
torch.manual_seed(0)
Hh = int(os.environ.get('HH', '2160')); Ww = int(os.environ.get('WW', '3840'))
x = torch.rand(1, int(os.environ.get('PP', '2')), Hh, Ww, device=dev)
kern = torch.rand(1, 1, 63, 63, device=dev); kern /= kern.sum()
lam = torch.tensor([0.02], device=dev); rho = torch.tensor([0.04], device=dev)
def rel(a, b): return ((a - b).abs().max() / b.abs().max()).item()
for nit in (1, 2, 3):
    _lib.set_option("use_big", 0)
    ref = fft_admm_tv(x, lam, rho, kern, False, nit).clone()
    for ub in (1, 2, 3):
        _lib.set_option("use_big", ub)
        out = fft_admm_tv(x, lam, rho, kern, False, nit)
        d = (out - ref).abs()[0, 0]
        print("maxit %d use_big %d: rel err %.3e; worst row %d col %d; rows with err>1e-3: %d cols: %d" % (
            nit, ub, rel(out, ref), d.max(1).values.argmax().item(), d.max(0).values.argmax().item(),
            (d.max(1).values > 1e-3).sum().item(), (d.max(0).values > 1e-3).sum().item()), flush=True)
