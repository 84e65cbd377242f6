import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_admm_deconv_b200 import fft_admm_tv, _lib
dev = torch.device("cuda:0")
torch.manual_seed(0)
Hh = int(os.environ.get('HH', '2160')); Ww = int(os.environ.get('WW', '3840'))
x = torch.rand(int(os.environ.get('BB', '1')), 3, Hh, Ww, device=dev)
kern = torch.rand(1, 1, 63, 63, device=dev); kern /= kern.sum()
lam = torch.tensor([0.02], device=dev); rho = torch.tensor([0.04], device=dev)
def run(n=20):
    fft_admm_tv(x, lam, rho, kern, False, 3); torch.cuda.synchronize()
    _lib.set_option("profile", 1); _lib.profile_reset()
    fft_admm_tv(x, lam, rho, kern, False, n); torch.cuda.synchronize()
    r = _lib.profile_read(0); c = _lib.profile_read(1)
    _lib.set_option("profile", 0)
    return r[0] / max(r[1], 1), c[0] / max(c[1], 1)
r, c = run()
px = x.numel()
print("%s %dx%d: rows %.3f ms (%.0f GB/s) cols %.3f ms (%.0f GB/s) whole %.1f%%" % (os.environ.get("ADMM_B200_LIB", "default"), Hh, Ww, r, 24.0 * px / r / 1e6, c, 12.0 * px / c / 1e6, 36.0 * px / (r + c) / 1e6 / 65.51), flush=True)
for R in sys.argv[1:]:
    _lib.set_option("rows_per_band", int(R))
    r, c = run()
    print("   rows_per_band %s: rows %.3f ms" % (R, r), flush=True)
