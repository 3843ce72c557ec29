"""B200-native GLOM column-update engine behind the glom-pytorch `Glom` API."""
from ._native import GlomB200Error, LIB_PATH
from .glom import Glom

__all__ = ["Glom", "GlomB200Error", "LIB_PATH"]
