from ase.md.md import MolecularDynamics


class Andersen(MolecularDynamics):
    pass
