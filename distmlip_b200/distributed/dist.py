"""Distributed -- host-side view of the GPU-resident partitioned graph.

Mirror of DistMLIP/distributed/dist.py (class Distributed).  In the reference this object holds the
19 host arrays produced by the C extension for *all* partitions and moves halo rows between GPUs
with cross-device slice assignment (dist.py:323-388).  Here the graph of *this rank's* partition is
built and kept on the GPU by libb200mlip (b2m_set_structure); halo exchange is NCCL point-to-point
inside b2m_compute.  The accessors below expose the same counters (dist.py:462-551) and, for tests,
the partition content in canonical form.
"""
from __future__ import annotations

import numpy as np


class Distributed:
    "Distributed Graph for parallelized MLIP inference (one rank's partition)"

    def __init__(self, engine, species, total_num_nodes, use_bond_graph, num_partitions):
        self.engine = engine
        self.species = species
        self.total_num_nodes = total_num_nodes
        self.use_bond_graph = use_bond_graph
        self.num_partitions = num_partitions
        c = engine.counts()
        self.counts = c
        self.rank = c["rank"]
        self.total_num_edges = None  # global count needs a reduction over ranks; see num_atom_edges
        self.forces = None
        self.stress = None

    @staticmethod
    def cartesian_to_wrapped_fractional(positions_cartesian, lattice, pbc):
        """dist.py:128-156."""
        if not pbc[0] and not pbc[1] and not pbc[2]:
            return positions_cartesian
        frac = np.linalg.solve(lattice.T, np.transpose(positions_cartesian)).T
        for i, periodic in enumerate(pbc):
            if periodic:
                frac[:, i] %= 1.0
                frac[:, i] %= 1.0
        return frac

    @classmethod
    def create_distributed(cls, cart_coords, frac_coords, lattice_matrix, num_partitions, pbc, cutoff,
                           three_body_cutoff=0, tol=1e-8, use_bond_graph=False, num_threads=1, *, engine=None,
                           species=None):
        """dist.py:158-275.  `engine` (a distmlip_b200._lib.Engine) and `species` are the two extra
        keyword-only arguments: the graph is built on that engine's GPU.  frac_coords / num_threads are
        accepted for signature compatibility (wrapping is recomputed on the device; no host threads)."""
        if engine is None:
            raise RuntimeError("create_distributed needs engine=: the graph lives on the GPU, there is no CPU builder")
        if num_partitions != engine.world:
            raise ValueError(f"num_partitions={num_partitions} but the engine's communicator has {engine.world} ranks")
        cart_coords = np.ascontiguousarray(cart_coords, dtype=float)
        lattice_matrix = np.ascontiguousarray(lattice_matrix, dtype=float)
        if species is None:
            species = np.zeros(len(cart_coords), dtype=np.int32)
        engine.set_structure(cart_coords, lattice_matrix, species, np.asarray(pbc).astype(np.int32), tol)
        obj = cls(engine, np.asarray(species), len(cart_coords), use_bond_graph, num_partitions)
        obj.cart = cart_coords
        return obj

    # ---- counters (dist.py:462-551), for this rank's partition ----
    def num_atoms(self, partition=None):
        return self.counts["n_own"] + self.counts["n_halo"]

    def num_atom_edges(self, partition=None):
        return self.counts["n_edges"]

    def num_bonds(self, partition=None):
        assert self.use_bond_graph, "num_bonds only works when bond graph is enabled"
        return self.counts["n_bond_own"] + self.counts["n_bond_halo"]

    def num_bond_edges(self, partition=None):
        assert self.use_bond_graph, "num_bond_edges only works when bond graph is enabled"
        return self.counts["n_angles"]

    def num_atom_border_nodes(self, partition=None):
        return self.counts["n_halo"]

    def num_bond_border_nodes(self, partition=None):
        assert self.use_bond_graph, "num_bond_border_nodes only works when bond graph is enabled"
        return self.counts["n_bond_halo"]

    # ---- canonical partition content (tests) ----
    def owned_gids(self):
        return self.engine.partition_info(0)

    def halo_gids(self):
        return self.engine.partition_info(1), self.engine.partition_info(2)

    def edges(self):
        return self.engine.partition_info(3)

    def bonds(self):
        return self.engine.partition_info(4)

    def angles(self):
        return self.engine.partition_info(5)

    def to_lists(self):
        return self.engine.partition_info(6)

    def walls(self):
        return self.engine.partition_info(7)

    def __repr__(self):
        c = self.counts
        val = f"""Distributed:
    Total num atoms: {self.total_num_nodes}
    Bond graph exists: {self.use_bond_graph}\n"""
        val += f"Partition {self.rank} (of {self.num_partitions}):\n"
        val += f"\t# of atom graph nodes: {self.num_atoms()} ({c['n_halo']} border nodes)\n"
        val += f"\t# of atom graph edges: {c['n_edges']}"
        if self.use_bond_graph:
            val += f"\t# of bond graph nodes: {self.num_bonds()}. ({c['n_bond_halo']} border nodes)\n"
            val += f"\t# of bond graph edges: {c['n_angles']}"
        return val + "\n"
