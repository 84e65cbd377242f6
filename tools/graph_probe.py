import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_admm_deconv_b200 import fft_admm_tv
dev = torch.device("cuda:0")
x = torch.rand(1, 1, 256, 256, device=dev)
kern = torch.rand(1, 1, 15, 15, device=dev); kern /= kern.sum()
lam = torch.tensor([0.02], device=dev); rho = torch.tensor([0.04], device=dev)
def timeit(f, n=20):
    f(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
direct = timeit(lambda: fft_admm_tv(x, lam, rho, kern, False, 50))
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    out = fft_admm_tv(x, lam, rho, kern, False, 50)
graph = timeit(lambda: g.replay())
print("cfg1 (1x1x256x256, 50 it): direct %.3f ms (%.1f Mpix-it/s), CUDA graph %.3f ms (%.1f Mpix-it/s)"
      % (direct, 65536 * 50 / direct / 1e3, graph, 65536 * 50 / graph / 1e3))
