import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import admm_oracle as O
import test_gpu_parity as T
for B in (2, 8, 32):
    for iso in (True, False):
        shape, maxit = (B, 3, 256, 256), 10
        rng = np.random.default_rng(40 + int(iso))
        x = O.make_blurred(shape, None, seed=4, noise=0.02)
        kern = np.zeros((0,), np.float32)
        gout = rng.standard_normal(shape)
        out, gx, gl, gr, gk, state = T._grads(x, 0.02, 0.04, kern, gout, iso, maxit, return_state=True)
        fx, fl, fr, fk = O.admm_tv_backward(x.astype(np.float64), 0.02, 0.04, kern, gout, iso, maxit)
        taubar = fl * 0.04
        print("B=%d iso=%s: glam %.6f / %.6f   grho %.6f / %.6f   (tau part %.4f, spectral part %.4f)" %
              (B, iso, gl[0], fl, gr[0], fr, taubar * 0.02 / 0.04 ** 2, fr + taubar * 0.02 / 0.04 ** 2), flush=True)
