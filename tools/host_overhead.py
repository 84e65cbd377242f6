"""Host-side cost of one fft_admm_tv / ADMMDeconv call (enqueue only, no device sync) on a tiny problem."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_admm_deconv_b200 import fft_admm_tv, ADMMDeconv
dev = torch.device("cuda:0")
x = torch.rand(1, 1, 64, 64, device=dev)
kern = torch.rand(1, 1, 5, 5, device=dev); kern /= kern.sum()
lam = torch.tensor([0.02], device=dev); rho = torch.tensor([0.04], device=dev)
for maxit in (2, 50):
    for _ in range(20): fft_admm_tv(x, lam, rho, kern, False, maxit)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 200
    for _ in range(n): fft_admm_tv(x, lam, rho, kern, False, maxit)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("maxit %d: host enqueue %.1f us per call, wall %.1f us per call" % (maxit, (t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6))
m = ADMMDeconv((5, 5), max_iters=2, lmbda=None, rho=None, iso=False).to(dev)
xg = x.clone().requires_grad_(True)
for _ in range(20): m(xg).sum().backward()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(200): m(xg).sum().backward()
torch.cuda.synchronize(); t1 = time.perf_counter()
print("module fwd+bwd maxit 2: wall %.1f us per step" % ((t1 - t0) / 200 * 1e6))
