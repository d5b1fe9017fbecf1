import numpy as np


class MolecularDynamics:
    """velocity Verlet core shared by every stub ensemble; thermostats / barostats only record their parameters"""

    def __init__(self, atoms, timestep, trajectory=None, logfile=None, loginterval=1, append_trajectory=False, **kwargs):
        self.atoms, self.dt, self.params, self.nsteps = atoms, timestep, dict(kwargs), 0
        self.trajectory, self.logfile, self.loginterval = trajectory, logfile, loginterval
        self.energies = []

    def run(self, steps):
        a = self.atoms
        f = a.get_forces()
        for _ in range(steps):
            p = a.get_momenta() + 0.5 * self.dt * f
            a.set_positions(a.get_positions() + self.dt * p / a.get_masses()[:, None])
            f = a.get_forces()
            a.set_momenta(p + 0.5 * self.dt * f)
            self.energies.append(a.get_potential_energy() + a.get_kinetic_energy())
            self.nsteps += 1
