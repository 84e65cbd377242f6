"""Accuracy probe on harder settings (many iterations, small rho, noise input) against the fp64 oracle."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_admm_deconv_b200 import fft_admm_tv
from oracle import admm_oracle as O
dev = torch.device("cuda:0")
def run(name, x, lam, rho, kern, iso, maxit):
    ref = O.admm_tv_spectral_form(x.astype(np.float64), lam, rho, kern, iso, maxit)
    ref32 = O.admm_tv_spectral_form(x.astype(np.float32), lam, rho, kern, iso, maxit)
    kt = torch.from_numpy(np.asarray(kern, np.float32)).to(dev) if np.size(kern) else torch.empty(0, device=dev)
    out = fft_admm_tv(torch.from_numpy(x).to(dev), torch.tensor([lam], device=dev), torch.tensor([rho], device=dev), kt, iso, maxit).cpu().numpy()
    print("%-46s gpu-vs-ref64 %.2e   numpy-fp32-vs-ref64 %.2e" % (name, O.rel_err(out, ref), O.rel_err(ref32, ref)), flush=True)
psf7 = O.make_psf("gauss", 7, 1.5)
x = O.make_blurred((2, 3, 256, 256), psf7, seed=1, noise=0.08)
run("notebook: iso, k7, lam=rho=0.02, 300 it", x, 0.02, 0.02, psf7[None, None], True, 300)
run("aniso, k7, 300 it", x, 0.02, 0.04, psf7[None, None], False, 300)
run("small rho 0.005, 150 it", x, 0.02, 0.005, psf7[None, None], False, 150)
run("large rho 1.0, 100 it", x, 0.02, 1.0, psf7[None, None], False, 100)
rng = np.random.default_rng(0)
w = rng.standard_normal((1, 1, 256, 256)).astype(np.float32)
run("white noise input, k15, 100 it", w, 0.02, 0.04, O.make_psf("gauss", 15, 2.5)[None, None], False, 100)
run("white noise, denoise, lam=0.5 rho=0.1, 200 it", w, 0.5, 0.1, np.zeros((0,), np.float32), False, 200)
psf31 = O.make_psf("motion", 31)
run("512^2 motion31 200 it", O.make_blurred((1, 3, 512, 512), psf31, seed=2), 0.02, 0.04, psf31[None, None], False, 200)
