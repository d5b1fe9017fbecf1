import numpy as np


def MaxwellBoltzmannDistribution(atoms, temperature_K=300.0, rng=None):
    from ase import units

    rng = rng or np.random.default_rng(0)
    m = atoms.get_masses()[:, None]
    atoms.set_momenta(rng.normal(size=(len(atoms), 3)) * np.sqrt(m * units.kB * temperature_K))
