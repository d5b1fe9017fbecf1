import numpy as np


class Atoms:
    def __init__(self, symbols, positions, cell, pbc=True, masses=None):
        self.symbols = list(symbols)
        self.positions = np.array(positions, dtype=float)
        self.cell = np.array(cell, dtype=float)
        self.pbc = np.array([bool(pbc)] * 3 if np.isscalar(pbc) else pbc, dtype=bool)
        self.numbers = np.full(len(self.symbols), 14)
        self.masses = np.full(len(self.symbols), 28.085) if masses is None else np.asarray(masses, float)
        self.momenta = np.zeros_like(self.positions)
        self.calc = None

    def __len__(self):
        return len(self.symbols)

    # --- geometry
    def get_positions(self, wrap=False):
        return self.positions.copy()

    def set_positions(self, p):
        self.positions = np.array(p, dtype=float)

    def get_cell(self):
        return self.cell.copy()

    def set_cell(self, cell, scale_atoms=False):
        cell = np.array(cell, dtype=float)
        if scale_atoms:
            self.positions = np.linalg.solve(self.cell.T, self.positions.T).T @ cell
        self.cell = cell

    def get_pbc(self):
        return self.pbc.copy()

    def get_scaled_positions(self, wrap=True):
        f = np.linalg.solve(self.cell.T, self.positions.T).T
        return f % 1.0 if wrap else f

    def get_chemical_symbols(self):
        return list(self.symbols)

    def get_atomic_numbers(self):
        return self.numbers.copy()

    def get_volume(self):
        return abs(np.linalg.det(self.cell))

    def get_masses(self):
        return self.masses.copy()

    # --- dynamics state
    def get_momenta(self):
        return self.momenta.copy()

    def set_momenta(self, m):
        self.momenta = np.array(m, dtype=float)

    def get_velocities(self):
        return self.momenta / self.masses[:, None]

    def get_kinetic_energy(self):
        return float(0.5 * np.sum(self.momenta**2 / self.masses[:, None]))

    def get_temperature(self):
        from ase import units

        return 2.0 * self.get_kinetic_energy() / (3.0 * len(self) * units.kB)

    # --- calculator plumbing
    def set_calculator(self, calc):
        self.calc = calc

    def _calc(self, props):
        self.calc.calculate(self, props, None)
        return self.calc.results

    def get_potential_energy(self):
        return float(self._calc(["energy"])["energy"])

    def get_forces(self):
        return np.array(self._calc(["forces"])["forces"], dtype=float)

    def get_stress(self, voigt=True):
        s = np.array(self._calc(["stress"])["stress"], dtype=float)
        if voigt and s.shape == (3, 3):
            s = np.array([s[0, 0], s[1, 1], s[2, 2], s[1, 2], s[0, 2], s[0, 1]])
        if not voigt and s.shape == (6,):
            s = np.array([[s[0], s[5], s[4]], [s[5], s[1], s[3]], [s[4], s[3], s[2]]])
        return s
