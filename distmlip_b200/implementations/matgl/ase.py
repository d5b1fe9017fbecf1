"""ASE adapters -- thin mirrors of DistMLIP/implementations/matgl/ase.py (PESCalculator_Dist :53-127,
Relaxer :130-223, MolecularDynamics :228-491, TrajectoryObserver from matgl.ext.ase).

ASE / pymatgen are imported lazily: PESCalculator_Dist works on any Atoms-like object (tests use
distmlip_b200.structures.SimpleAtoms); Relaxer and MolecularDynamics need a real ASE install and keep
the reference's keyword arguments.
"""
from __future__ import annotations

import numpy as np

from distmlip_b200.implementations.matgl.pes import Potential_Dist

try:  # pragma: no cover - ASE is absent in the build image
    from ase.calculators.calculator import Calculator as _Calculator, all_changes as _all_changes
    _HAVE_ASE = True
except Exception:  # noqa: BLE001
    _HAVE_ASE = False
    _all_changes = ["positions", "numbers", "cell", "pbc", "initial_charges", "initial_magmoms"]

    class _Calculator:  # minimal stand-in with the attributes PESCalculator_Dist touches
        def __init__(self, **kwargs):
            self.results = {}
            self.atoms = None

        def calculate(self, atoms=None, properties=None, system_changes=None):
            self.atoms = atoms


def _voigt6(s):
    s = np.asarray(s)
    return np.array([s[0, 0], s[1, 1], s[2, 2], (s[1, 2] + s[2, 1]) / 2, (s[0, 2] + s[2, 0]) / 2,
                     (s[0, 1] + s[1, 0]) / 2])


class PESCalculator_Dist(_Calculator):
    """Machine Learning Interatomic Potential calculator for ASE."""

    implemented_properties = ("energy", "free_energy", "forces", "stress", "hessian", "magmoms")

    def __init__(self, potential=None, state_attr=None, stress_unit="GPa", stress_weight=1.0, use_voigt=False,
                 **kwargs):
        super().__init__(**kwargs)
        assert isinstance(potential, Potential_Dist), "PESCalculatorDist requires using a Potential_Dist."
        self.potential = potential
        self.compute_stress = potential.calc_stresses
        self.compute_hessian = potential.calc_hessian
        self.compute_magmom = potential.calc_site_wise
        if stress_unit == "eV/A3":
            conversion_factor = 0.006241509125883258  # GPa -> eV/A^3 (ase.units)
        elif stress_unit == "GPa":
            conversion_factor = 1.0
        else:
            raise ValueError(f"Unsupported stress_unit: {stress_unit}. Must be 'GPa' or 'eV/A3'.")
        self.stress_weight = stress_weight * conversion_factor
        self.state_attr = state_attr
        self.use_voigt = use_voigt
        self.last_count = None

    def calculate(self, atoms, properties=None, system_changes=None):
        """ase.py:80-127."""
        properties = properties or ["energy"]
        system_changes = system_changes or _all_changes
        _Calculator.calculate(self, atoms=atoms, properties=properties, system_changes=system_changes)
        calc_result = self.potential(atoms, self.state_attr)
        self.results.update(
            energy=calc_result[0].detach().cpu().numpy().item(),
            free_energy=calc_result[0].detach().cpu().numpy(),
            forces=calc_result[1].detach().cpu().numpy(),
        )
        if self.compute_stress:
            st = calc_result[2].detach().cpu().numpy()
            self.results.update(stress=(_voigt6(st) if self.use_voigt else st) * self.stress_weight)
        if self.compute_magmom:
            self.results.update(magmoms=calc_result[4].detach().cpu().numpy())


class TrajectoryObserver:
    """matgl.ext.ase.TrajectoryObserver look-alike (collections of energies/forces/stresses per step)."""

    def __init__(self, atoms):
        self.atoms = atoms
        self.energies, self.forces, self.stresses = [], [], []
        self.atom_positions, self.cells = [], []

    def __call__(self):
        self.energies.append(float(self.atoms.get_potential_energy()))
        self.forces.append(self.atoms.get_forces())
        self.stresses.append(self.atoms.get_stress())
        self.atom_positions.append(self.atoms.get_positions())
        self.cells.append(self.atoms.get_cell()[:])


class Relaxer:
    """ase.py:130-223: Relaxer(potential, state_attr, optimizer="FIRE", relax_cell=True, stress_weight=1/160.21766208)."""

    def __init__(self, potential=None, state_attr=None, optimizer="FIRE", relax_cell=True,
                 stress_weight=1 / 160.21766208):
        if not _HAVE_ASE:
            raise ImportError("Relaxer needs ASE (not installed in this image)")
        import ase.optimize as opt

        self.optimizer = getattr(opt, optimizer) if isinstance(optimizer, str) else optimizer
        self.calculator = PESCalculator_Dist(potential=potential, state_attr=state_attr, stress_unit="eV/A3",
                                             stress_weight=stress_weight)
        self.relax_cell = relax_cell

    def relax(self, atoms, fmax=0.1, steps=500, traj_file=None, interval=1, verbose=False,
              ase_cellfilter="Frechet", params_asecellfilter=None, **kwargs):
        import contextlib
        import io
        import sys

        from ase.constraints import ExpCellFilter
        from ase.filters import FrechetCellFilter

        atoms.set_calculator(self.calculator)
        stream = sys.stdout if verbose else io.StringIO()
        params_asecellfilter = params_asecellfilter or {}
        with contextlib.redirect_stdout(stream):
            obs = TrajectoryObserver(atoms)
            if self.relax_cell:
                atoms = (FrechetCellFilter(atoms, **params_asecellfilter) if ase_cellfilter == "Frechet"
                         else ExpCellFilter(atoms, **params_asecellfilter))
            optimizer = self.optimizer(atoms, **kwargs)
            optimizer.attach(obs, interval=interval)
            optimizer.run(fmax=fmax, steps=steps)
            obs()
        if self.relax_cell:
            atoms = atoms.atoms
        return {"final_structure": atoms, "trajectory": obs}


class MolecularDynamics:
    """ase.py:228-491: same 17 keyword arguments; ensembles delegate to ase.md."""

    def __init__(self, atoms, potential, state_attr=None, stress_weight=1.0, ensemble="nvt", temperature=300,
                 timestep=1.0, pressure=1.01325 * 1e-4, taut=None, taup=None, friction=1.0e-3, andersen_prob=1.0e-2,
                 ttime=25.0, pfactor=75.0**2.0, external_stress=None, compressibility_au=None, trajectory=None,
                 logfile=None, loginterval=1, append_trajectory=False, mask=None):
        if not _HAVE_ASE:
            raise ImportError("MolecularDynamics needs ASE (not installed in this image)")
        from ase import units
        from ase.md import Langevin
        from ase.md.andersen import Andersen
        from ase.md.nvtberendsen import NVTBerendsen
        from ase.md.verlet import VelocityVerlet

        self.atoms = atoms
        self.atoms.set_calculator(PESCalculator_Dist(potential=potential, state_attr=state_attr, stress_unit="eV/A3",
                                                     stress_weight=stress_weight))
        taut = taut if taut is not None else 100 * timestep * units.fs
        common = dict(trajectory=trajectory, logfile=logfile, loginterval=loginterval,
                      append_trajectory=append_trajectory)
        ens = ensemble.lower()
        if ens == "nve":
            self.dyn = VelocityVerlet(self.atoms, timestep * units.fs, **common)
        elif ens in ("nvt", "nvt_berendsen"):
            self.dyn = NVTBerendsen(self.atoms, timestep * units.fs, temperature_K=temperature, taut=taut, **common)
        elif ens == "nvt_langevin":
            self.dyn = Langevin(self.atoms, timestep * units.fs, temperature_K=temperature, friction=friction, **common)
        elif ens == "nvt_andersen":
            self.dyn = Andersen(self.atoms, timestep * units.fs, temperature_K=temperature,
                                andersen_prob=andersen_prob, **common)
        else:
            raise ValueError("Ensemble not supported by this thin mirror (nve/nvt/nvt_langevin/nvt_andersen)")
        self.trajectory, self.logfile, self.loginterval, self.timestep = trajectory, logfile, loginterval, timestep

    def run(self, steps):
        self.dyn.run(steps)

    def set_atoms(self, atoms):
        calculator = self.atoms.calc
        self.atoms = atoms
        self.dyn.atoms = atoms
        self.dyn.atoms.set_calculator(calculator)
