import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_admm_deconv_b200 import fft_admm_tv, _lib
dev = torch.device("cuda:0")
B = int(os.environ.get("BB", "32")); Hh = int(os.environ.get("HH", "256")); Ww = int(os.environ.get("WW", "256"))
x = torch.rand(B, 3, Hh, Ww, device=dev)
lam = torch.tensor([0.02], device=dev); rho = torch.tensor([0.04], device=dev)
kern = torch.empty(0, device=dev)
def run(n=50):
    fft_admm_tv(x, lam, rho, kern, False, 5); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); fft_admm_tv(x, lam, rho, kern, False, n); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for R in [0] + [int(a) for a in sys.argv[1:]]:
    _lib.set_option("rows_per_band", R)
    us = run()
    print("B=%d %dx%d rows_per_band %2d: %.1f us per iteration (%.0f GB/s)" % (B, Hh, Ww, R, us, 36.0 * x.numel() / us / 1e3), flush=True)
