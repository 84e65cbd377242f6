import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_admm_deconv_b200 import fft_admm_tv
dev = torch.device("cuda:0")
x = torch.rand(32, 3, 256, 256, device=dev)
lam = torch.tensor([0.02], device=dev); rho = torch.tensor([0.04], device=dev)
for _ in range(3):
    fft_admm_tv(x, lam, rho, torch.empty(0, device=dev), True, 10)
torch.cuda.synchronize()
print("ok")
