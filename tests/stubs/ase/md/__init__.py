from ase.md.langevin import Langevin  # noqa: F401
