"""Optimizers: one damped steepest-descent loop under every name the reference's OPTIMIZERS enum lists."""
import numpy as np


class Optimizer:
    def __init__(self, atoms, maxstep=0.05, **kwargs):
        self.atoms, self.maxstep, self.observers, self.nsteps = atoms, maxstep, [], 0
        self.kwargs = kwargs

    def attach(self, fn, interval=1):
        self.observers.append((fn, interval))

    def run(self, fmax=0.1, steps=500):
        for it in range(steps):
            f = self.atoms.get_forces()
            for fn, interval in self.observers:
                if it % interval == 0:
                    fn()
            if np.sqrt((f**2).sum(axis=1)).max() < fmax:
                return True
            step = 0.1 * f
            norm = np.sqrt((step**2).sum(axis=1)).max()
            if norm > self.maxstep:
                step *= self.maxstep / norm
            self.atoms.set_positions(self.atoms.get_positions() + step)
            self.nsteps += 1
        return False


FIRE = BFGS = LBFGS = LBFGSLineSearch = MDMin = BFGSLineSearch = Optimizer
