import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_admm_deconv_b200 import fft_admm_tv, _lib
dev = torch.device("cuda:0")
def run(shape, k, maxit, reps=5):
    g = torch.Generator().manual_seed(1)
    x = torch.rand(shape, generator=g).to(dev)
    kern = torch.rand(1, 1, k, k, generator=g).to(dev); kern /= kern.sum()
    lam = torch.tensor([0.02], device=dev); rho = torch.tensor([0.04], device=dev)
    outs = {}
    for pdl in (0, 1, 0, 1):
        _lib.set_option("use_pdl", pdl)
        fft_admm_tv(x, lam, rho, kern, False, 3); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): out = fft_admm_tv(x, lam, rho, kern, False, maxit)
        e1.record(); torch.cuda.synchronize()
        outs[pdl] = out
        print("shape %s pdl=%d: %.3f ms per solve (%.1f us/iteration)" % (shape, pdl, e0.elapsed_time(e1) / reps, 1e3 * e0.elapsed_time(e1) / reps / maxit), flush=True)
    print("   identical:", bool(torch.equal(outs[0], outs[1])), flush=True)
run((64, 3, 512, 512), 31, 100, 3)
run((256, 3, 256, 256), 15, 50, 3)
run((1, 1, 256, 256), 15, 50, 10)
run((8, 3, 256, 256), 15, 50, 10)
