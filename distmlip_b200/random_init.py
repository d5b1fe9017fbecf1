"""Random-init CHGNet weights (matgl attribute tree / state_dict names) for benchmarks.

There is no network for pretrained checkpoints, so bench.py times the architecture with seeded random
weights.  This container is product-side (nothing from oracle/ is imported): it only has to look like a
matgl `CHGNet` to `CHGNet_Dist.from_existing` -- attributes + `state_dict()` + `to()`.
Initialisation follows the PyTorch defaults of the corresponding layers (Linear: U(-1/sqrt(in), 1/sqrt(in)),
Embedding: N(0,1)); the radial / Fourier frequencies start at k*pi and k as in matgl.
"""
from __future__ import annotations

import math

import torch

DEFAULT_ELEMENTS = (
    "H", "He", "Li", "Be", "B", "C", "N", "O", "F", "Ne", "Na", "Mg", "Al", "Si", "P", "S", "Cl", "Ar",
    "K", "Ca", "Sc", "Ti", "V", "Cr", "Mn", "Fe", "Co", "Ni", "Cu", "Zn", "Ga", "Ge", "As", "Se", "Br",
    "Kr", "Rb", "Sr", "Y", "Zr", "Nb", "Mo", "Tc", "Ru", "Rh", "Pd", "Ag", "Cd", "In", "Sn", "Sb", "Te",
    "I", "Xe", "Cs", "Ba", "La", "Ce", "Pr", "Nd", "Pm", "Sm", "Eu", "Gd", "Tb", "Dy", "Ho", "Er", "Tm",
    "Yb", "Lu", "Hf", "Ta", "W", "Re", "Os", "Ir", "Pt", "Au", "Hg", "Tl", "Pb", "Bi", "Ac", "Th", "Pa",
    "U", "Np", "Pu",
)


class RandomCHGNet:
    """Duck-typed stand-in for a matgl CHGNet carrying seeded random weights."""

    def __init__(self, seed=0, dim=64, max_n=9, max_f=4, num_blocks=4, cutoff=5.0, threebody_cutoff=3.0,
                 cutoff_exponent=5, element_types=DEFAULT_ELEMENTS):
        self.element_types = tuple(element_types)
        self.cutoff, self.three_body_cutoff, self.cutoff_exponent = cutoff, threebody_cutoff, cutoff_exponent
        self.n_blocks, self.use_bond_graph = num_blocks, True
        self.readout_field, self.readout_operation, self.state_embedding = "atom_feat", "sum", None
        g = torch.Generator().manual_seed(seed)

        def lin(out_f, in_f, bias=True, prefix=""):
            b = 1.0 / math.sqrt(in_f)
            d = {prefix + "weight": (torch.rand(out_f, in_f, generator=g) * 2 - 1) * b}
            if bias:
                d[prefix + "bias"] = (torch.rand(out_f, generator=g) * 2 - 1) * b
            return d

        sd = {
            "bond_expansion.frequencies": math.pi * torch.arange(1, max_n + 1, dtype=torch.float32),
            "threebody_bond_expansion.frequencies": math.pi * torch.arange(1, max_n + 1, dtype=torch.float32),
            "angle_expansion.frequencies": torch.arange(0, max_f + 1, dtype=torch.float32),
            "atom_embedding.weight": torch.randn(len(self.element_types), dim, generator=g),
        }
        nf = 2 * max_f + 1
        sd.update(lin(dim, max_n, False, "bond_embedding.layers.0."))
        sd.update(lin(dim, nf, False, "angle_embedding.layers.0."))
        for name in ("atom_bond_weights", "bond_bond_weights", "threebody_bond_weights"):
            sd.update(lin(dim, max_n, False, name + "."))
        for l in range(num_blocks):
            p = f"atom_graph_layers.{l}.conv_layer."
            for br in ("layers", "gates"):
                sd.update(lin(dim, 3 * dim, True, p + f"node_update_func.{br}.layers.0."))
                sd.update(lin(dim, dim, True, p + f"node_update_func.{br}.layers.1."))
            sd.update(lin(dim, dim, False, p + "node_out_func."))
        for l in range(num_blocks - 1):
            p = f"bond_graph_layers.{l}.conv_layer."
            for br in ("layers", "gates"):
                sd.update(lin(dim, 4 * dim, True, p + f"node_update_func.{br}.layers.0."))
                sd.update(lin(dim, dim, True, p + f"node_update_func.{br}.layers.1."))
                sd.update(lin(dim, 4 * dim, True, p + f"edge_update_func.{br}.layers.0."))
            sd.update(lin(dim, dim, False, p + "node_out_func."))
        sd.update(lin(1, dim, True, "sitewise_readout."))
        for i, (o, n) in enumerate(((dim, dim), (dim, dim), (1, dim))):
            sd.update(lin(o, n, True, f"final_layer.layers.{i}."))
        self._sd = sd

    def to(self, *_args, **_kwargs):
        return self

    def state_dict(self):
        return dict(self._sd)
