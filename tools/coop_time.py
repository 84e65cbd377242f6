"""Small batches: cooperative kernel (coop_small.cu) vs separate launches vs CUDA graph of the separate launches; ms per solve."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_admm_deconv_b200 import ADMMDeconv, _lib
dev = torch.device("cuda:0")


def med(fn, n=15):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


_lib.set_option("use_cluster", 0)
for shape, iso in (((8, 3, 256, 256), True), ((3, 3, 256, 256), True), ((8, 3, 256, 256), False), ((16, 3, 256, 256), True), ((16, 3, 256, 256), False),
                   ((32, 3, 256, 256), True), ((32, 3, 256, 256), False), ((1, 3, 512, 512), True), ((4, 3, 512, 512), False), ((8, 3, 128, 128), True)):
    m = ADMMDeconv((), max_iters=100, lmbda=0.02, rho=0.04, iso=iso).to(dev)
    x = torch.rand(shape, device=dev)
    res = {}
    with torch.inference_mode():
        for mode in (0, 2):
            _lib.set_option("use_coop", mode)
            res[mode] = med(lambda: m(x))
    _lib.set_option("use_coop", 0)
    print("%s iso=%s: separate launches %.3f ms, cooperative %.3f ms (x%.2f)  [%.1f us per iteration]" %
          (shape, iso, res[0], res[2], res[0] / res[2], res[2] * 10), flush=True)
_lib.set_option("use_cluster", 1)
