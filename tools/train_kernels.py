"""In-situ kernel times of one training step (torch.profiler / CUPTI; warm caches, unlike the ncu launch list)."""
import os, sys, torch, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_admm_deconv_b200 import ADMMDeconv
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
iso = os.environ.get("ISO", "1") == "1"
m = ADMMDeconv((), max_iters=10, lmbda=None, rho=None, iso=iso).to(dev)
with torch.no_grad():
    m.lmbda.fill_(0.02); m.rho.fill_(0.04)
x = torch.rand(32, 3, 256, 256, device=dev)
def step():
    m.zero_grad(set_to_none=True)
    (m(x) ** 2).mean().backward()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        agg[e.name[:60]][0] += 1; agg[e.name[:60]][1] += e.device_time
tot = sum(v[1] for v in agg.values())
print("iso=%s: %.3f ms of kernels per step" % (iso, tot / 3e3))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print("%-62s n=%3d  %8.1f us/step  avg %6.1f us  %4.1f%%" % (k, v[0] // 3, v[1] / 3, v[1] / v[0], 100 * v[1] / tot))
