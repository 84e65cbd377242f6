"""Golden fixtures for the CALLERS of the hot path, produced by the UNMODIFIED reference on the CPU (build container):

    python tests/golden/make_golden_models.py

  fanout_multiadmm.npz   MultiADMM (modelbuild/blocks.py:252-261): three solvers (iso, kernel, bias) on one input
  fanout_fusion.npz      ADMMFusion (elayers/admmfusion.py:9-40) with its AttentionChannelPooling
  restorer_e2e.npz       DivergentRestorer (modelbuild/denoiser.py:7-63) built exactly as scripts/train.py:19-24,70-73
                         (two ADMMDeconv branches, kern_size=(), 100 iterations, iso=True, learnable lmbda / rho),
                         random weights (seeded), one forward on a 2 x 3 x 64 x 64 batch
  act_u8.npz             ADMMDeconv with relu / sigmoid / tanh activations and a uint8 image scaled by /255 (etransforms.py:29-31)

Each file holds the input, the state dict of the reference module (keys prefixed `sd.`; for the 7 M-parameter restorer a
per-tensor fingerprint instead) and its output, so the GPU tests can load the weights with strict=True into the same
architecture with this package's solver dropped in.
"""
import os
import sys

import numpy as np
import torch

REF_SRC = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF_SRC)
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from admmtor.elayers.admmdeconv import ADMMDeconv  # noqa: E402
from admmtor.elayers.admmfusion import ADMMFusion  # noqa: E402
from admmtor.modelbuild.blocks import MultiADMM  # noqa: E402
from admmtor.modelbuild.denoiser import DivergentRestorer  # noqa: E402
from oracle.admm_oracle import make_blurred, make_psf  # noqa: E402

torch.set_num_threads(8)

FANOUT_CFGS = [dict(kern_size=(), max_iters=12, lmbda=0.02, rho=0.04, iso=True),
               dict(kern_size=(5, 5), max_iters=8, lmbda=None, rho=None, iso=False),
               dict(kern_size=(), max_iters=10, lmbda=None, rho=None, iso=False, bias=True)]


def sd_arrays(m):
    return {"sd." + k: v.detach().cpu().numpy().copy() for k, v in m.state_dict().items()}


def tame(m):
    """Well-conditioned solver parameters (xavier `w` is zero-mean: den(0,0) ~ 0; U(0,1) lmbda/rho are fine)."""
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, ADMMDeconv) and mod.w.numel():
                k = mod.w.shape[-1]
                mod.w.copy_(torch.from_numpy(make_psf("gauss", k, 1.0)[None, None]))


def main():
    torch.manual_seed(11)
    x = make_blurred((2, 3, 32, 32), None, seed=31, noise=0.05)
    m = MultiADMM(FANOUT_CFGS); tame(m)
    with torch.no_grad():
        y = m(torch.from_numpy(x)).numpy()
    np.savez_compressed(os.path.join(HERE, "fanout_multiadmm.npz"), x=x, out32=y, **sd_arrays(m))
    print("fanout_multiadmm", y.shape)

    torch.manual_seed(12)
    f = ADMMFusion(FANOUT_CFGS, in_channels=3, with_admms=True); tame(f)
    f.eval()
    with torch.no_grad():
        yf = f(torch.from_numpy(x)).numpy()
        cat = torch.cat([a(torch.from_numpy(x)) for a in f.admms], dim=1)
        probs = f.acp.cwa(cat).numpy()
    np.savez_compressed(os.path.join(HERE, "fanout_fusion.npz"), x=x, out32=yf, probs=probs, **sd_arrays(f))
    print("fanout_fusion", yf.shape, "channel probabilities", np.round(probs, 4))

    torch.manual_seed(13)
    d1 = {'kern_size': (), 'max_iters': 100, 'iso': True}
    d2 = {'kern_size': (), 'max_iters': 100, 'iso': True}
    model = DivergentRestorer([2, 8, 32], 3, 3, 86, 86, 8, output_activation=torch.nn.Sigmoid(), admms=[d1, d2])
    model.eval()
    xr = make_blurred((2, 3, 64, 64), None, seed=32, noise=0.05)
    with torch.no_grad():
        yr = model(torch.from_numpy(xr)).numpy()
        admm_out = [a(torch.from_numpy(xr)).numpy() for a in model.blocks[0].admms]
    # 7.1 M parameters: the weights are NOT stored; the test rebuilds the reference model with the same seed (CPU generator,
    # same torch build on the GPU box) and checks this fingerprint before loading its checkpoint into the patched model
    sd = model.state_dict()
    np.savez_compressed(os.path.join(HERE, "restorer_e2e.npz"), x=xr, out32=yr, admm0=admm_out[0], admm1=admm_out[1], seed=np.int32(13),
                        sd_keys=np.array(list(sd.keys())), sd_sums=np.array([float(v.double().sum()) for v in sd.values()]),
                        sd_abs=np.array([float(v.double().abs().sum()) for v in sd.values()]))
    print("restorer_e2e", yr.shape, "params", sum(p.numel() for p in model.parameters()))

    # fused activation epilogues and the uint8 / 255 prologue
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, size=(2, 3, 48, 64), dtype=np.uint8)
    xs = torch.from_numpy(img).to(torch.float32) / 255.0                 # etransforms.py:29-31 after dataload.py:31
    out = {}
    for name, act in (("relu", torch.relu), ("sigmoid", torch.nn.Sigmoid()), ("tanh", torch.tanh)):
        torch.manual_seed(14)
        a = ADMMDeconv((3, 3), max_iters=9, lmbda=0.02, rho=0.04, iso=False, bias=True, activation=act); tame(a)
        with torch.no_grad():
            a.b.fill_(-0.4)                                               # so that relu actually clips
            out["out_" + name] = a(xs).numpy()
    np.savez_compressed(os.path.join(HERE, "act_u8.npz"), img=img, w=a.w.detach().numpy(), **out)
    print("act_u8 written")


if __name__ == "__main__":
    main()
