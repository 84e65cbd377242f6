"""B200-native drop-in for the ADMM-TV deconvolution path of georgegrosu1/torch-admm-deconv.

    reference                                    this package
    admmtor.eops.deconv.fft_admm_tv         ->   torch_admm_deconv_b200.eops.deconv.fft_admm_tv
    admmtor.elayers.admmdeconv.ADMMDeconv   ->   torch_admm_deconv_b200.elayers.admmdeconv.ADMMDeconv

The arithmetic runs in hand-written sm_100a CUDA kernels reached through the C ABI of
include/admm_b200.h; there is no CPU or PyTorch fallback.
"""
from .eops.deconv import fft_admm_tv, identity, soft_thresh, block_thresh, pixelnorm, hard_thresh, torch_abs2
from .elayers.admmdeconv import ADMMDeconv
from .elayers.multiadmm import MultiADMM, Deconvs

__all__ = ["fft_admm_tv", "ADMMDeconv", "MultiADMM", "Deconvs", "identity", "soft_thresh", "block_thresh", "pixelnorm", "hard_thresh",
           "torch_abs2"]
__version__ = "0.1.0"
