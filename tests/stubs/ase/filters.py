"""Cell filters: positions + a 3x3 strain block as generalised coordinates; the generalised forces on the cell are
-V * stress (stress in eV/A^3 as ASE calculators return it) -- enough to check that the Relaxer hands ASE the right units."""
import numpy as np


class FrechetCellFilter:
    def __init__(self, atoms, mask=None, scalar_pressure=0.0, **kwargs):
        self.atoms, self.kwargs = atoms, kwargs
        self.orig_cell = atoms.get_cell()
        self.last_stress = None

    def __len__(self):
        return len(self.atoms) + 3

    def get_positions(self):
        strain = np.linalg.solve(self.orig_cell, self.atoms.get_cell()) - np.eye(3)
        return np.vstack([self.atoms.get_positions(), strain])

    def set_positions(self, p):
        n = len(self.atoms)
        self.atoms.set_cell(self.orig_cell @ (np.eye(3) + p[n:]), scale_atoms=True)
        self.atoms.set_positions(p[:n] @ np.linalg.solve(self.orig_cell, self.atoms.get_cell()))

    def get_forces(self):
        f = self.atoms.get_forces()
        s = self.atoms.get_stress(voigt=False)
        self.last_stress = s
        return np.vstack([f, -self.atoms.get_volume() * s / len(self.atoms)])

    def get_potential_energy(self):
        return self.atoms.get_potential_energy()


ExpCellFilter = FrechetCellFilter
