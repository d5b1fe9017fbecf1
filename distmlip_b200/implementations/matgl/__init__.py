"""Mirror of DistMLIP.implementations.matgl (same public names)."""
from .models.chgnet import CHGNet_Dist
from .models.tensornet import TensorNet_Dist
from .pes import Potential_Dist
from .ase import PESCalculator_Dist, Relaxer, MolecularDynamics, TrajectoryObserver

__all__ = ["CHGNet_Dist", "TensorNet_Dist", "Potential_Dist", "PESCalculator_Dist", "Relaxer", "MolecularDynamics", "TrajectoryObserver"]
