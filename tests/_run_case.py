"""Helper (not a test): run one CHGNet evaluation in THIS process and save E, F, S to an .npz.
Used by test_gpu_experimental.py to evaluate process-wide kernel switches (read once per process) in a child process.
Usage: python tests/_run_case.py out.npz [ncells] [seed]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests._util import engine_from_model, make_model, si_diamond  # noqa: E402

out = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 5
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 31
model = make_model()
eng = engine_from_model(model)
atoms = si_diamond(n, seed=seed)
species = np.array([model.element_types.index(x) for x in atoms.get_chemical_symbols()], dtype=np.int32)
eng.set_structure(atoms.get_positions(), atoms.get_cell(), species, atoms.get_pbc().astype(np.int32))
E, F, S = eng.compute(True, True)
np.savez(out, E=E, F=F, S=S)
eng.close()
