from ase.filters import ExpCellFilter  # noqa: F401  (older ASE location, used by the reference's import)
