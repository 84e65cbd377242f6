"""Fixtures with NEGATIVE tau = lmbd / rho from the unmodified reference (build container only).

lmbda and rho are unconstrained learnable U(0,1) parameters of ADMMDeconv (admmdeconv.py:26-41); without the clipper
of scripts/train.py:27-38 they can cross zero.  soft_thresh(q, tau) = sign(q) max(|q| - tau, 0) (deconv.py:15-16) then
returns sign(q)(|q| + |tau|) and the iteration still runs; a drop-in has to follow it.

    python tests/golden/make_golden_negtau.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import save_case, grad_case, make_blurred  # noqa: E402  (imports the reference from /root/reference/src)

rng = np.random.default_rng(77)
k5 = rng.random((5, 5)); k5 /= k5.sum()
save_case("negtau_k5_32x48_n10", make_blurred((2, 3, 32, 48), k5.astype(np.float32), seed=51), k5, -0.01, 0.04, False, 10)
save_case("negtau_rho_64x64_n8", make_blurred((1, 2, 64, 64), None, seed=52, noise=0.05), None, -0.02, 0.05, False, 8)
save_case("iso_negtau_32x32_n8", make_blurred((3, 2, 32, 32), k5.astype(np.float32), seed=53), k5, -0.01, 0.04, True, 8)
grad_case("grad_aniso_negtau_k3_16x16_n4", (2, 2, 16, 16), 3, False, 4, 54, lam=-0.01, rho=0.04)
