from ase.md.md import MolecularDynamics


class NVTBerendsen(MolecularDynamics):
    pass
