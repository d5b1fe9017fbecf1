from ase.md.md import MolecularDynamics


class NPTBerendsen(MolecularDynamics):
    pass


class Inhomogeneous_NPTBerendsen(NPTBerendsen):
    pass
