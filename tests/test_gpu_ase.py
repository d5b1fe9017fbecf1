"""Relaxer and MolecularDynamics on the real engine (reference usage: examples/chgnet_example.ipynb cells 5-6), over the
`ase` package that is available (tests/stubs/ase when ASE itself is not installed)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
try:
    import ase  # noqa: F401
except ImportError:
    sys.path.insert(0, os.path.join(HERE, "stubs"))
    import ase  # noqa: F401

from ase import Atoms  # noqa: E402

from distmlip_b200.structures import si_diamond  # noqa: E402
from tests._util import make_model  # noqa: E402

pytestmark = pytest.mark.gpu


def test_relax_and_md_on_the_engine():
    from distmlip_b200.implementations.matgl import CHGNet_Dist, MolecularDynamics, Potential_Dist, Relaxer

    dm = CHGNet_Dist.from_existing(make_model(scale=1.6))
    dm.enable_distributed_mode([0])
    pot = Potential_Dist(model=dm)
    s = si_diamond(2, seed=3)
    atoms = Atoms(s.get_chemical_symbols(), s.get_positions(), s.get_cell())
    out = Relaxer(potential=pot, relax_cell=False).relax(atoms, fmax=1e-4, steps=25)
    e = out["trajectory"].energies
    assert len(e) >= 3 and e[-1] <= e[0] + 1e-9  # steepest descent on the model's own forces lowers its energy
    f0 = np.abs(out["trajectory"].forces[0]).max()
    f1 = np.abs(out["trajectory"].forces[-1]).max()
    assert f1 < f0
    md = MolecularDynamics(atoms, pot, ensemble="nve", timestep=1.0)
    md.run(10)
    en = md.dyn.energies
    assert max(en) - min(en) < 5e-4 * len(atoms)  # forces are the gradient of the energy: NVE conserves E + K
    dm._engine.close()
