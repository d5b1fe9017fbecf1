"""Minimal stand-in for the parts of ASE the reference's matgl/ase.py mirror touches (tests only).

ASE is not installed in the build image and cannot be fetched.  This package lets `Relaxer` and `MolecularDynamics`
EXECUTE in tests: it implements just enough behaviour (an Atoms container wired to a calculator, a steepest-descent
"FIRE", velocity Verlet, pass-through thermostats, a cell filter that exposes the stress as extra degrees of freedom in
eV/A^3) to check call signatures, unit handling and result plumbing.  It is never imported by product code when a real
ASE is present (tests put this directory on sys.path only if `import ase` fails).
"""
from ase.atoms import Atoms  # noqa: F401

__version__ = "0.0-stub"
IS_STUB = True
