import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_admm_deconv_b200 import ADMMDeconv
dev = torch.device("cuda:0")
m = ADMMDeconv((), max_iters=10, lmbda=None, rho=None, iso=True).to(dev)
with torch.no_grad():
    m.lmbda.fill_(0.02); m.rho.fill_(0.04)
x = torch.rand(32, 3, 256, 256, device=dev)
for _ in range(3):
    m.zero_grad(set_to_none=True)
    (m(x) ** 2).mean().backward()
torch.cuda.synchronize()
print("ok")
