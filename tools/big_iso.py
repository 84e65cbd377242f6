"""Large-frame sizes with iso=True and the plain R2C / C2R kernels: parity against the generic engine + timing."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_admm_deconv_b200 import fft_admm_tv, _lib
dev = torch.device("cuda:0")
torch.manual_seed(0)
Hh = int(os.environ.get('HH', '1080')); Ww = int(os.environ.get('WW', '1920'))
x = torch.rand(1, 3, Hh, Ww, device=dev)
kern = torch.rand(1, 1, 9, 9, device=dev); kern /= kern.sum()
lam = torch.tensor([0.02], device=dev); rho = torch.tensor([0.04], device=dev)
def rel(a, b): return ((a - b).abs().max() / b.abs().max()).item()
for iso in (False, True):
    _lib.set_option("use_big", 0)
    ref = fft_admm_tv(x, lam, rho, kern, iso, 5).clone()
    _lib.set_option("use_big", 3)
    out = fft_admm_tv(x, lam, rho, kern, iso, 5)
    print("iso=%s: big vs generic rel err %.3e" % (iso, rel(out, ref)), flush=True)
    for ub in (0, 3):
        _lib.set_option("use_big", ub)
        fft_admm_tv(x, lam, rho, kern, iso, 3); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fft_admm_tv(x, lam, rho, kern, iso, 20); e1.record(); torch.cuda.synchronize()
        print("   use_big %d: %.3f ms per iteration" % (ub, e0.elapsed_time(e1) / 20), flush=True)
_lib.set_option("use_big", 3)
# gradients with the large-frame plain kernels in the backward
xg = x[:, :1].clone().requires_grad_(True)
for iso in (False, True):
    outs = []
    for ub in (0, 3):
        _lib.set_option("use_big", ub)
        xg.grad = None
        fft_admm_tv(xg, lam, rho, kern, iso, 4).square().sum().backward()
        outs.append(xg.grad.clone())
    _lib.set_option("use_big", 3)
    d = (outs[1] - outs[0])
    print("iso=%s: grad big vs generic max-rel %.3e  l2-rel %.3e  pixels with |d|>1e-4*max: %d" % (iso, rel(outs[1], outs[0]), (d.norm() / outs[0].norm()).item(), (d.abs() > 1e-4 * outs[0].abs().max()).sum().item()), flush=True)
