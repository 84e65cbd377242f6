"""L2-resident chunks on several streams: the batch is split over S streams, each solving its share chunk by chunk."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_inputs_torch, WORKLOADS, LAMBDA, RHO
from torch_admm_deconv_b200 import fft_admm_tv, _lib

dev = torch.device("cuda:0")
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
B, C, H, W, kind, k, sigma, maxit = WORKLOADS[wl]
x, psf = make_inputs_torch((B, C, H, W), kind, k, sigma)
x = x.to(dev); kern = psf.to(dev)
lam = torch.tensor([LAMBDA], device=dev); rho = torch.tensor([RHO], device=dev)
per = H * W * 4 * 7 / 2**20


def run(S, mb):
    _lib.set_option("chunk_mb", mb)
    streams = [torch.cuda.Stream(dev) for _ in range(S)]
    parts = torch.chunk(x, S, dim=0)
    cur = torch.cuda.current_stream()

    def go():
        outs = []
        for s, p in zip(streams, parts):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                outs.append(fft_admm_tv(p, lam, rho, kern, False, maxit))
        for s in streams:
            cur.wait_stream(s)
        return outs
    go(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        outs = go()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("%s streams=%d chunk_mb=%3d (%.1f planes/chunk): %8.3f ms  %7.1f Gpixel-it/s" %
          (wl, S, mb, mb / per if mb else B * C / S, ms, B * H * W * maxit / ms / 1e6), flush=True)


run(1, 0)
for S, mb in [(2, 56), (2, 42), (4, 28), (4, 21), (4, 14), (8, 14), (8, 7), (3, 35), (2, 0), (4, 0)]:
    run(S, mb)
_lib.set_option("chunk_mb", -1)
