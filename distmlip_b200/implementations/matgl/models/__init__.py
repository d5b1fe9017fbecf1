"""Mirror of DistMLIP.implementations.matgl.models (same public names)."""
from .chgnet import CHGNet_Dist
from .tensornet import TensorNet_Dist

__all__ = ["CHGNet_Dist", "TensorNet_Dist"]
