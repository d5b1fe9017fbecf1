"""Diagnostic (not a test): time ONE slab of an N-way split of the 1M-atom cell on one GPU with the halo exchanges skipped
(B2M_DEBUG_NO_HALO=1: wrong numbers, right amount of per-partition work).  Separates per-partition compute (halo rows,
smaller grids) from inter-rank waiting in the N-GPU step time.   usage: B2M_DEBUG_NO_HALO=1 python tests/slab_timing.py [world] [rank]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distmlip_b200 import _lib  # noqa: E402
from distmlip_b200.random_init import RandomCHGNet  # noqa: E402
from distmlip_b200.structures import si_diamond  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 3
m = RandomCHGNet(seed=0)
sd = m.state_dict()
eng = _lib.Engine(n_elem=sd["atom_embedding.weight"].shape[0], dim=64, max_n=9, max_f=4, n_blocks=4, cutoff=5.0,
                  three_body_cutoff=3.0, cutoff_exponent=5)
eng.load_state_dict({k: v.float() for k, v in sd.items()})
eng.finalize()
atoms = si_diamond(50)
eng.set_partition(rank, world)
eng.set_structure(atoms.get_positions(), atoms.get_cell(), np.zeros(len(atoms), dtype=np.int32), atoms.get_pbc().astype(np.int32))
for _ in range(3):
    eng.compute_resident(1)
ts = []
for _ in range(8):
    _e, ms = eng.compute_resident(1)
    ts.append((ms, eng.timings()["fwd_ms"], eng.timings()["bwd_ms"]))
c = eng.counts()
t = np.mean(ts, axis=0)
print(f"slab {rank}/{world}: own {c['n_own']} halo {c['n_halo']} edges {c['n_edges']} bonds {c['n_bond_own']}+{c['n_bond_halo']} "
      f"angles {c['n_angles']}: {t[0]:.2f} ms/step (fwd {t[1]:.2f}, bwd {t[2]:.2f}), graph {eng.timings()['graph_ms']:.2f} ms")
