"""The reference's OWN workloads (scripts/train.py:19-24, configs/train_cfg.json:6,11,13): ADMMDeconv(kern_size=(), max_iters=100,
iso=True, learnable lmbda / rho) on (3,3,256,256) training batches and (8,3,256,256) evaluation batches -- this package
against the reference's eager CUDA path (baseline/_ref) on the same GPU.  ms per call, CUDA events, median of 10."""
import sys, os, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
from admmtor.elayers.admmdeconv import ADMMDeconv as RefLayer
from torch_admm_deconv_b200 import ADMMDeconv
dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


out = {}
for name, shape, train in (("train_3x3x256x256", (3, 3, 256, 256), True), ("eval_8x3x256x256", (8, 3, 256, 256), False)):
    x = torch.rand(shape, device=dev)
    res = {}
    for tag, cls in (("ours", ADMMDeconv), ("reference_cuda_eager", RefLayer)):
        torch.manual_seed(0)
        m = cls((), max_iters=100, iso=True).to(dev)
        with torch.no_grad():
            m.lmbda.fill_(0.02); m.rho.fill_(0.04)
        if train:
            def step():
                m.zero_grad(set_to_none=True)
                (m(x) ** 2).mean().backward()
            res[tag] = timeit(step)
        else:
            def fwd():
                with torch.inference_mode():
                    m(x)
            res[tag] = timeit(fwd)
    res["speedup"] = res["reference_cuda_eager"] / res["ours"]
    out[name] = res
    print(name, res, flush=True)
print(json.dumps(out))
