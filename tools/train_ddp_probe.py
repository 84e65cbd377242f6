"""2-GPU training step: where does the time go (all-reduce of the layer gradients)?  torchrun --nproc-per-node 2"""
import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_admm_deconv_b200 import ADMMDeconv
from torch_admm_deconv_b200.sharding import allreduce_param_grads
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
m = ADMMDeconv((), max_iters=10, lmbda=None, rho=None, iso=False).to(dev)
with torch.no_grad(): m.lmbda.fill_(0.02); m.rho.fill_(0.04)
x = torch.rand(32, 3, 256, 256, device=dev)
def step(ar):
    m.zero_grad(set_to_none=True)
    (m(x) ** 2).mean().backward()
    if ar: allreduce_param_grads(m.parameters(), average=True)
for ar in (False, True, False, True):
    for _ in range(5): step(ar)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(30): step(ar)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    if rank == 0:
        print("allreduce=%s: host enqueue %.3f ms/step, wall %.3f ms/step" % (ar, (t1 - t0) / 30 * 1e3, (t2 - t0) / 30 * 1e3), flush=True)
if world > 1: dist.destroy_process_group()
