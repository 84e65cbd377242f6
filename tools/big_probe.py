"""cfg3 (2160x3840) large-FFT kernels: parity against the generic engine + per-kernel timing."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_admm_deconv_b200 import fft_admm_tv, _lib
dev = torch.device("cuda:0")
torch.manual_seed(0)
P = int(os.environ.get("PLANES", "3"))
x = torch.rand(1, P, 2160, 3840, device=dev)
kern = torch.rand(1, 1, 63, 63, device=dev); kern /= kern.sum()
lam = torch.tensor([0.02], device=dev); rho = torch.tensor([0.04], device=dev)

def rel(a, b):
    return ((a - b).abs().max() / b.abs().max()).item()

for nit in (2, 3, 6):
    _lib.set_option("force_generic", 1)
    ref = fft_admm_tv(x, lam, rho, kern, False, nit).clone()
    _lib.set_option("force_generic", 0)
    out = fft_admm_tv(x, lam, rho, kern, False, nit)
    torch.cuda.synchronize()
    print("maxit %d: big vs generic rel err %.3e" % (nit, rel(out, ref)), flush=True)

def run(n=20):
    fft_admm_tv(x, lam, rho, kern, False, 3); torch.cuda.synchronize()
    _lib.set_option("profile", 1); _lib.profile_reset()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); fft_admm_tv(x, lam, rho, kern, False, n); e1.record(); torch.cuda.synchronize()
    r = _lib.profile_read(0); c = _lib.profile_read(1)
    _lib.set_option("profile", 0)
    ms = e0.elapsed_time(e1) / n
    return ms, r[0] / max(r[1], 1), c[0] / max(c[1], 1)

for fg in (0, 1):
    _lib.set_option("force_generic", fg)
    ms, r, c = run()
    px = x.numel()
    print("force_generic %d: %.3f ms/it (rows %.3f ms = %.0f GB/s, cols %.3f ms = %.0f GB/s) whole %.1f%% of 6551"
          % (fg, ms, r, 24.0 * px / r / 1e6, c, 12.0 * px / c / 1e6, 36.0 * px / (r + c) / 1e6 / 65.51), flush=True)
_lib.set_option("force_generic", 0)
for R in (8, 12, 16, 24, 32):
    _lib.set_option("rows_per_band", R)
    ms, r, c = run()
    print("rows_per_band %d: rows %.3f ms" % (R, r), flush=True)
