from ase.md.md import MolecularDynamics


class Langevin(MolecularDynamics):
    pass
