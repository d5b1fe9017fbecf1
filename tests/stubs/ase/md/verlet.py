from ase.md.md import MolecularDynamics


class VelocityVerlet(MolecularDynamics):
    pass
