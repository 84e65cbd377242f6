"""Mixed dispatch: one axis on the power-of-two kernels, the other on the large-frame kernels (or generic): vs generic."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_admm_deconv_b200 import fft_admm_tv, _lib
dev = torch.device("cuda:0")
torch.manual_seed(1)
lam = torch.tensor([0.02], device=dev); rho = torch.tensor([0.04], device=dev)
kern = torch.rand(1, 1, 7, 7, device=dev); kern /= kern.sum()
bad = 0
for (H, W) in [(1024, 512), (1080, 256), (256, 1920), (512, 3840), (2048, 128), (720, 512), (128, 1280), (1440, 256),
               (1024, 1920), (2160, 1024), (720, 4096), (1080, 2560), (100, 1920), (1080, 100), (1024, 300)]:
    for iso in (False, True):
        x = torch.rand(1, 2, H, W, device=dev)
        _lib.set_option("force_generic", 1)
        ref = fft_admm_tv(x, lam, rho, kern, iso, 4).clone()
        _lib.set_option("force_generic", 0)
        out = fft_admm_tv(x, lam, rho, kern, iso, 4)
        e = ((out - ref).abs().max() / ref.abs().max()).item()
        flag = "" if e < 1e-5 else "   <-- MISMATCH"
        bad += e >= 1e-5
        print("%4d x %4d iso=%d: %.2e%s" % (H, W, iso, e, flag), flush=True)
print("mismatches:", bad)
