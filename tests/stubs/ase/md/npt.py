from ase.md.md import MolecularDynamics


class NPT(MolecularDynamics):
    pass
