from .admmdeconv import ADMMDeconv
from .multiadmm import MultiADMM, Deconvs, ADMMFusion

__all__ = ["ADMMDeconv", "MultiADMM", "Deconvs", "ADMMFusion"]
