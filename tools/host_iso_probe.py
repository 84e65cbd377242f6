"""Host enqueue cost vs wall time of one small-batch solve (the reference's own shapes): is the path launch-bound on the host?"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_admm_deconv_b200 import fft_admm_tv
dev = torch.device("cuda:0")
for shape, iso in (((8,3,256,256), True), ((8,3,256,256), False), ((3,3,256,256), True)):
    x = torch.rand(*shape, device=dev)
    kern = torch.empty(0, device=dev)
    lam = torch.tensor([0.02], device=dev); rho = torch.tensor([0.04], device=dev)
    for _ in range(5): fft_admm_tv(x, lam, rho, kern, iso, 100)
    torch.cuda.synchronize()
    enq = wall = 0.0; n = 20
    for _ in range(n):                      # one solve at a time: the launch queue is empty, so enqueue time is host cost
        torch.cuda.synchronize(); t0 = time.perf_counter()
        fft_admm_tv(x, lam, rho, kern, iso, 100)
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        enq += t1 - t0; wall += t2 - t0
    print(shape, "iso", iso, ": host enqueue %.0f us per solve, wall %.0f us per solve" % (enq / n * 1e6, wall / n * 1e6))
