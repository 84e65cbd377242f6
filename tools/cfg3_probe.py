import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_admm_deconv_b200 import fft_admm_tv, _lib
dev = torch.device("cuda:0")
x = torch.rand(1, 3, 2160, 3840, device=dev)
kern = torch.rand(1, 1, 63, 63, device=dev); kern /= kern.sum()
lam = torch.tensor([0.02], device=dev); rho = torch.tensor([0.04], device=dev)
def run(n=10):
    fft_admm_tv(x, lam, rho, kern, False, 3); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); fft_admm_tv(x, lam, rho, kern, False, n); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    return ms, 36.0 * x.numel() / (ms * 1e-3) / 1e9
for thr in (256, 512, 1024):
    _lib.set_option("threads", thr)
    for rows in (0, 4, 6, 8):
        for cols in (0, 4, 8):
            _lib.set_option("rows_per_band", rows); _lib.set_option("cols_per_tile", cols)
            try:
                ms, gbs = run()
                print("threads %4d rows %d cols %d: %.3f ms/it %.0f GB/s (%.1f%%)" % (thr, rows, cols, ms, gbs, gbs / 65.51), flush=True)
            except Exception as e:
                print("threads %4d rows %d cols %d: %s" % (thr, rows, cols, str(e)[:80]), flush=True)
