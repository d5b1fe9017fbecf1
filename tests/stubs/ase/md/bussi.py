from ase.md.md import MolecularDynamics


class Bussi(MolecularDynamics):
    pass
