"""Relaxer / MolecularDynamics / PESCalculator_Dist execute (reference: DistMLIP/implementations/matgl/ase.py:53-491).

ASE is not installed in the build image, so these tests put tests/stubs (a minimal `ase` stand-in, see its docstring) on
sys.path when `import ase` fails.  A harmonic Potential_Dist double stands in for the GPU engine here (CPU test); the same
classes run on the real engine in tests/test_gpu_ase.py.
"""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
try:
    import ase  # noqa: F401
except ImportError:
    sys.path.insert(0, os.path.join(HERE, "stubs"))
    import ase  # noqa: F401

from ase import Atoms  # noqa: E402

import distmlip_b200.implementations.matgl.ase as mase  # noqa: E402
from distmlip_b200.implementations.matgl.pes import Potential_Dist  # noqa: E402

GPA_PER_EVA3 = 160.21766208


class HarmonicPotential(Potential_Dist):
    """Potential_Dist double: E = k/2 |x - x0|^2 + hydrostatic stress p0 (GPa), the return convention of pes.py:50-146"""

    def __init__(self, x0, k=2.0, p0_gpa=3.0):
        self.x0, self.k, self.p0 = np.array(x0, float), k, p0_gpa
        self.calc_forces = self.calc_stresses = True
        self.calc_hessian = self.calc_site_wise = False
        self.calls = 0

    def forward(self, atoms, state_attr=None, tol=1e-8):
        self.calls += 1
        d = atoms.get_positions() - self.x0
        e = 0.5 * self.k * float((d * d).sum())
        return (torch.tensor([e], dtype=torch.float64), torch.tensor(-self.k * d, dtype=torch.float32),
                torch.eye(3, dtype=torch.float32) * self.p0, None)


def cell_atoms(n=8, seed=0):
    rng = np.random.default_rng(seed)
    x0 = rng.random((n, 3)) * 5.0
    return Atoms(["Si"] * n, x0 + rng.normal(0, 0.2, (n, 3)), np.eye(3) * 6.0), x0


def test_relaxer_hands_ase_the_stress_in_ev_per_a3():
    atoms, x0 = cell_atoms()
    pot = HarmonicPotential(x0)
    rel = mase.Relaxer(potential=pot, optimizer="FIRE", relax_cell=True)
    # the reference passes ONLY stress_weight = 1/160.2 (ase.py:153-157): GPa -> eV/A^3 exactly once
    assert abs(rel.calculator.stress_weight - 1.0 / GPA_PER_EVA3) < 1e-15
    out = rel.relax(atoms, fmax=1e-3, steps=5)  # (the double's stress is constant: only the plumbing is checked here)
    final, obs = out["final_structure"], out["trajectory"]
    assert len(final) == len(atoms) and len(obs.energies) >= 2 and len(obs.cells) == len(obs.energies)
    s = np.array(obs.stresses[-1])
    assert np.allclose(s[:3], 3.0 / GPA_PER_EVA3, rtol=1e-6)  # 3 GPa reaches the optimizer as 0.0187 eV/A^3
    rel2 = mase.Relaxer(potential=pot, optimizer="LBFGS", relax_cell=False)
    atoms2, _ = cell_atoms(seed=1)
    out2 = rel2.relax(atoms2, fmax=1e-3, steps=400, ase_cellfilter="Exp")
    assert np.sqrt((out2["final_structure"].get_forces() ** 2).sum(1)).max() < 1e-3
    with pytest.raises(KeyError):
        mase.Relaxer(potential=pot, optimizer="no_such_optimizer")


@pytest.mark.parametrize("ensemble", ["nve", "nvt", "nvt_langevin", "nvt_andersen", "nvt_bussi", "npt", "npt_berendsen",
                                      "npt_nose_hoover"])
def test_all_eight_ensembles_construct_and_run(ensemble):
    atoms, x0 = cell_atoms()
    pot = HarmonicPotential(x0)
    md = mase.MolecularDynamics(atoms, pot, ensemble=ensemble, temperature=300, timestep=0.5, loginterval=1)
    assert isinstance(atoms.calc, mase.PESCalculator_Dist)
    # MD asks the calculator for eV/A^3 (ase.py:291-296: stress_unit="eV/A3", stress_weight=1)
    assert abs(atoms.calc.stress_weight - 1.0 / GPA_PER_EVA3) < 1e-12
    md.run(5)
    assert pot.calls >= 5 and md.dyn.nsteps == 5
    if ensemble == "nve":
        e = md.dyn.energies
        assert abs(e[-1] - e[0]) < 1e-3 * max(1.0, abs(e[0]))  # velocity Verlet conserves the harmonic energy
    if ensemble == "nvt_bussi":
        assert atoms.get_kinetic_energy() > 0  # Maxwell-Boltzmann start (ase.py:360-362)
    other, _ = cell_atoms(seed=3)
    md.set_atoms(other)
    assert other.calc is not None and md.dyn.atoms is other


def test_unsupported_inputs_raise_like_the_reference():
    atoms, x0 = cell_atoms()
    with pytest.raises(ValueError):
        mase.MolecularDynamics(atoms, HarmonicPotential(x0), ensemble="nph")
    with pytest.raises(Exception):
        mase.MolecularDynamics(atoms, potential=object())
    with pytest.raises(ValueError):
        mase.PESCalculator_Dist(potential=HarmonicPotential(x0), stress_unit="kbar")


def test_upper_triangular_cell():
    atoms, x0 = cell_atoms()
    lower = np.array([[6.0, 0, 0], [1.0, 5.5, 0], [0.5, 0.7, 6.2]])
    atoms.set_cell(lower, scale_atoms=True)
    frac = atoms.get_scaled_positions(wrap=False)
    md = mase.MolecularDynamics(atoms, HarmonicPotential(x0), ensemble="npt_nose_hoover")
    cell = md.atoms.get_cell()
    assert np.allclose(cell, np.triu(cell))
    assert np.allclose(np.linalg.norm(cell, axis=1), np.linalg.norm(lower, axis=1))  # same lengths ...
    cosang = lambda c: [c[1] @ c[2] / np.linalg.norm(c[1]) / np.linalg.norm(c[2]), c[0] @ c[2] / np.linalg.norm(c[0]) / np.linalg.norm(c[2]),
                        c[0] @ c[1] / np.linalg.norm(c[0]) / np.linalg.norm(c[1])]
    assert np.allclose(cosang(cell), cosang(lower))  # ... and angles
    assert np.allclose(md.atoms.get_scaled_positions(wrap=False), frac)
    assert abs(abs(np.linalg.det(cell)) - abs(np.linalg.det(lower))) < 1e-9
