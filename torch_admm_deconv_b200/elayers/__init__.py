from .admmdeconv import ADMMDeconv
from .multiadmm import MultiADMM, Deconvs

__all__ = ["ADMMDeconv", "MultiADMM", "Deconvs"]
