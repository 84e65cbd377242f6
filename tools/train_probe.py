"""Developer probe: cfg4 of BASELINE.json -- unrolled ADMM layer training step (fwd + bwd), batch 32 RGB 256x256,
learnable rho/lambda (and optionally w), 10 unrolled iterations."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_admm_deconv_b200 import ADMMDeconv
from oracle.admm_oracle import make_psf

def run(kern_size, iso, reps=5):
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m = ADMMDeconv(kern_size, max_iters=10, lmbda=None, rho=None, iso=iso).to(dev)
    with torch.no_grad():
        m.lmbda.fill_(0.02); m.rho.fill_(0.04)
        if kern_size:
            m.w.copy_(torch.from_numpy(make_psf("gauss", kern_size[0], 2.5)[None, None]).to(dev))
    x = torch.rand(32, 3, 256, 256, device=dev)
    def step():
        m.zero_grad(set_to_none=True)
        loss = (m(x) ** 2).mean()
        loss.backward()
        return loss
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    el = 32 * 3 * 256 * 256
    print("cfg4 kern=%s iso=%s: %.3f ms per fwd+bwd step, %.1f Mpix-it/s, %.1f GB/s at 96 B/element-iteration (%.1f%% of 6551)"
          % (kern_size, iso, ms, 32 * 256 * 256 * 10 / ms / 1e3, 96.0 * el * 10 / (ms * 1e-3) / 1e9, 96.0 * el * 10 / (ms * 1e-3) / 1e9 / 65.51))
    # forward only
    with torch.no_grad():
        for _ in range(2): m(x)
        torch.cuda.synchronize(); e0.record()
        for _ in range(reps): m(x)
        e1.record(); torch.cuda.synchronize()
    print("      forward only (inference): %.3f ms" % (e0.elapsed_time(e1) / reps))

if __name__ == "__main__":
    run((), False); run((15, 15), False); run((), True); run((15, 15), True)
