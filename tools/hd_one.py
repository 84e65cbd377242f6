import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_admm_deconv_b200 import fft_admm_tv
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.rand(8, 3, 1080, 1920, device=dev)
kern = torch.rand(1, 1, 31, 31, device=dev); kern /= kern.sum()
lam = torch.tensor([0.02], device=dev); rho = torch.tensor([0.04], device=dev)
fft_admm_tv(x, lam, rho, kern, False, 4); torch.cuda.synchronize()
