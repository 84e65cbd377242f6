from .admmdeconv import ADMMDeconv

__all__ = ["ADMMDeconv"]
