import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_admm_deconv_b200 import fft_admm_tv, _lib
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.rand(1, 3, 2160, 3840, device=dev)
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
print("%s: rows %.3f ms  cols %.3f ms" % (os.environ.get("ADMM_B200_LIB", "default"), r, c), flush=True)
for R in sys.argv[1:]:
    _lib.set_option("rows_per_band", int(R))
    r, c = run()
    print("   rows_per_band %s: rows %.3f ms" % (R, r), flush=True)
