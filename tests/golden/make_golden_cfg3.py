"""Golden fixture for BASELINE.json configs[2] at its REAL shape and length: 1 x 3 x 2160 x 3840, 63 x 63 Gaussian
PSF, 200 ADMM iterations, produced by the UNMODIFIED reference (`/root/reference/src/admmtor/eops/deconv.py`,
`fft_admm_tv`) on the CPU in float32 (in float64 its 63 x 63 depthwise `F.conv2d` falls back to an im2col path that wants
263 GB of memory, so the reference itself cannot run this size in double).  Build container only (about an hour of CPU):

    python tests/golden/make_golden_cfg3.py [--skip-reference | --reference-only]

The full output is 200 MB, so the fixture keeps (i) 64 x 64 crops of all three channels at six positions
(the four wrap-around corners are among them), (ii) the per-row and per-column means of every channel and
(iii) global sums; the input is regenerated on the GPU box by `oracle.admm_oracle.make_blurred` (numpy
generator, seed 1234).  The fp64 oracle (spectral form) is run on the same input first and stored beside
the reference so the two can be compared without the reference (tests/test_oracle.py).
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.admm_oracle import make_psf, make_blurred, admm_tv_spectral_form  # noqa: E402

SHAPE = (1, 3, 2160, 3840)
K, SIGMA, MAXIT, LAM, RHO, SEED = 63, 8.0, 200, 0.02, 0.04, 1234
CROPS = [(0, 0), (0, 3776), (2096, 0), (2096, 3776), (1048, 1888), (517, 2931)]   # (row, col) of 64 x 64 windows
OUT = os.path.join(HERE, "cfg3_2160x3840_gauss63_n200.npz")


def summarise(out, tag):
    d = {}
    d[tag + "_crops"] = np.stack([out[0, :, r:r + 64, c:c + 64] for r, c in CROPS]).astype(np.float64)
    d[tag + "_rowmean"] = out[0].mean(axis=2).astype(np.float64)
    d[tag + "_colmean"] = out[0].mean(axis=1).astype(np.float64)
    d[tag + "_sum"] = np.float64(out.sum(dtype=np.float64))
    d[tag + "_sumsq"] = np.float64((out.astype(np.float64) ** 2).sum())
    d[tag + "_absmax"] = np.float64(np.abs(out).max())
    return d


def main():
    psf = make_psf("gauss", K, SIGMA)
    x = make_blurred(SHAPE, psf, seed=SEED)
    d = dict(shape=np.array(SHAPE), k=np.int32(K), sigma=np.float64(SIGMA), maxit=np.int32(MAXIT), lam=np.float64(LAM),
             rho=np.float64(RHO), seed=np.int32(SEED), crops=np.array(CROPS),
             x_sum=np.float64(x.sum(dtype=np.float64)), x_crop0=x[0, :, :8, :8].copy())
    o = None
    if "--reference-only" in sys.argv:
        d = dict(np.load(OUT))
    else:
        t0 = time.time()
        o = admm_tv_spectral_form(x.astype(np.float64), LAM, RHO, psf[None, None], False, MAXIT, workers=4)
        print("oracle fp64 spectral form: %.0f s" % (time.time() - t0), flush=True)
        d.update(summarise(o, "oracle64"))
        np.savez_compressed(OUT, **d)
    if "--skip-reference" in sys.argv:
        return
    import torch
    sys.path.insert(0, "/root/reference/src")
    from admmtor.eops.deconv import fft_admm_tv  # the reference
    torch.set_num_threads(6)
    t0 = time.time()
    with torch.no_grad():
        r = fft_admm_tv(torch.from_numpy(x), torch.tensor([LAM]), torch.tensor([RHO]), torch.from_numpy(psf[None, None]),
                        False, MAXIT).numpy()
    print("reference fp32: %.0f s" % (time.time() - t0), flush=True)
    d.update(summarise(r, "ref32"))
    if o is not None:
        d["oracle_vs_ref32"] = np.float64(np.abs(o - r).max() / np.abs(r).max())
    else:                                   # full oracle output not kept: compare on the stored crops
        d["oracle_vs_ref32"] = np.float64(np.abs(d["oracle64_crops"] - d["ref32_crops"]).max() / d["ref32_absmax"])
    print("oracle fp64 vs reference fp32 (max|a-b|/max|b|): %.3e" % d["oracle_vs_ref32"], flush=True)
    np.savez_compressed(OUT, **d)


if __name__ == "__main__":
    main()
