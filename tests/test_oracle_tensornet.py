"""CPU: the stage-by-stage TensorNet mirror (oracle/tensornet_manual.py, the layout and reverse pass the CUDA kernels
follow) against autograd through the module restatement (oracle/tensornet_ref.py)."""
import numpy as np
import pytest
import torch

from distmlip_b200.structures import SimpleAtoms, rough_cell, si_diamond
from oracle import graph_ref as G
from oracle import tensornet_manual as TM
from oracle.manual_ref import forces_from_gvec
from oracle.tensornet_ref import TensorNetRef, potential_ref


def make_tn(seed=0, scale=1.0, **kw):
    torch.manual_seed(seed)
    m = TensorNetRef(**kw)
    if scale != 1.0:
        with torch.no_grad():
            for n, p in m.named_parameters():
                if "weight" in n and p.ndim == 2 and ".emb." not in n:
                    p.mul_(scale)
    return m


def tn_graph(atoms, rc=5.0):
    cart, lat = atoms.get_positions(), np.array(atoms.get_cell())
    i1, i2, off, _d2, _b = G.neighbor_list(cart, lat, atoms.get_pbc().astype(np.int64), rc, 0.0)
    return dict(i1=i1, i2=i2, off=off, vec=cart[i2] + off @ lat - cart[i1])


@pytest.mark.parametrize("group", ["O(3)", "SO(3)"])
def test_manual_mirror_equals_autograd(group):
    a0 = si_diamond(2, sigma=0.15, seed=1)
    atoms = SimpleAtoms(["Si" if i % 3 else "O" for i in range(len(a0))], a0.get_positions(), a0.get_cell())
    m = make_tn(seed=3, scale=1.5, equivariance_invariance_group=group)
    og = tn_graph(atoms)
    E, F, S = potential_ref(m, atoms, graph=(og["i1"], og["i2"], og["off"]), dtype=torch.float64, data_std=1.3)
    types = np.array([m.element_types.index(s) for s in atoms.get_chemical_symbols()])
    out = TM.run(m, types, og["vec"], og["i1"], og["i2"], data_std=1.3)
    assert abs(1.3 * float(out["energy"]) - float(E)) < 1e-10 * max(1.0, abs(float(E)))
    Fm, Sm = forces_from_gvec(out["gvec"], og["vec"], og["i1"], og["i2"], len(atoms), atoms.get_volume())
    assert float((Fm - F).abs().max()) < 1e-10 * max(1.0, float(F.abs().max()))
    assert float((Sm - S).abs().max()) < 1e-10 * max(1.0, float(S.abs().max()))


def test_oracle_symmetries():
    """rotation invariance of the energy / equivariance of the forces, and translation invariance."""
    atoms = rough_cell(40, seed=2)
    m = make_tn(seed=1, scale=1.5)
    E, F, _ = potential_ref(m, atoms, dtype=torch.float64)
    th = 0.7
    R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    cell = np.array(atoms.get_cell()) @ R.T
    b = SimpleAtoms(atoms.get_chemical_symbols(), atoms.get_positions() @ R.T + np.array([0.37, 0.11, -0.2]) @ cell, cell)
    E2, F2, _ = potential_ref(m, b, dtype=torch.float64)
    assert abs(float(E2 - E)) < 1e-9
    assert np.abs(F2.numpy() - F.numpy() @ R.T).max() < 1e-9


def test_random_init_model_has_the_oracle_state_dict():
    """bench.py's product-side random TensorNet carries exactly the tensors of the restated module tree"""
    from distmlip_b200.random_init import RandomTensorNet

    a, b = RandomTensorNet(seed=0).state_dict(), TensorNetRef().state_dict()
    assert set(a) == set(b)
    assert all(tuple(a[k].shape) == tuple(b[k].shape) for k in a)
    assert torch.allclose(a["bond_expansion.rbf.centers"], b["bond_expansion.rbf.centers"])


def _partitioned_energy(model, atoms, P, exchange_after_embedding):
    """The reference's partitioned schedule (tensornet.py:92-147) restated on the oracle modules: per partition the
    subgraph of the edges it owns (destination-owned) over [owned | halo] atoms, `tensor_embedding` and every layer applied
    per partition, `atom_transfer` (halo rows <- owner rows) after each layer -- and, optionally, after the embedding."""
    from oracle.tensornet_ref import decompose_tensor, tensor_norm

    cart, lat = atoms.get_positions(), np.array(atoms.get_cell())
    o = G.GraphOracle(cart, lat, atoms.get_pbc().astype(np.int64), P, float(model.cutoff), 0.0, False)
    assert o.accepts
    model = model.double()
    types_all = torch.tensor([model.element_types.index(s) for s in atoms.get_chemical_symbols()])
    parts = []
    for p in range(P):
        own = o.owned(p)
        halo = np.concatenate([o.from_list(p, q) for q in range(P) if q != p]) if P > 1 else np.zeros(0, dtype=np.int64)
        local = np.concatenate([own, halo]).astype(np.int64)
        g2l = -np.ones(len(cart), dtype=np.int64)
        g2l[local] = np.arange(len(local))
        s, d, off = o.edges_of(p)
        vec = torch.tensor(cart[d] + off @ lat - cart[s])
        parts.append(dict(own=own, local=local, n_own=len(own), src=torch.tensor(g2l[s]), dst=torch.tensor(g2l[d]), vec=vec,
                          dist=torch.linalg.norm(vec, dim=1), types=types_all[local]))

    def transfer(feats):  # Distributed.atom_transfer (dist.py:323-388)
        glob = torch.zeros(len(cart), *feats[0].shape[1:], dtype=feats[0].dtype)
        for pt, f in zip(parts, feats):
            glob[pt["own"]] = f[: pt["n_own"]]
        return [torch.cat([f[: pt["n_own"]], glob[pt["local"][pt["n_own"]:]]]) for pt, f in zip(parts, feats)]

    with torch.no_grad():
        X, attr = [], []
        for pt in parts:
            ea = model.bond_expansion(pt["dist"])
            attr.append(ea)
            X.append(model.tensor_embedding(pt["src"], pt["dst"], pt["types"], pt["dist"], pt["vec"], ea, len(pt["local"])))
        if exchange_after_embedding:
            X = transfer(X)
        for layer in model.layers:
            X = [layer(pt["src"], pt["dst"], pt["dist"], ea, x) for pt, ea, x in zip(parts, attr, X)]
            X = transfer(X)
        Xg = torch.zeros(len(cart), *X[0].shape[1:], dtype=X[0].dtype)  # Distributed.aggregate (dist.py:277-321)
        for pt, x in zip(parts, X):
            Xg[pt["own"]] = x[: pt["n_own"]]
        I, A, S = decompose_tensor(Xg)
        x = torch.cat((tensor_norm(I), tensor_norm(A), tensor_norm(S)), dim=-1)
        return float(model.final_layer.gated(model.linear(model.out_norm(x))).sum())


def test_reference_schedule_depends_on_the_partition_count_and_the_engines_does_not():
    """the reference's first atom_transfer follows layer 0 (tensornet.py:119-127) although the embedding
    already aggregates edges, so its layer 0 reads zero tensors for the halo sources.  With the exchange the engine adds
    after the embedding, every partition count reproduces the single-graph energy; without it the energy moves."""
    atoms = si_diamond(2, sigma=0.15, seed=3, nz=8)
    m = make_tn(seed=4, scale=1.5)
    e1 = _partitioned_energy(m, atoms, 1, False)
    og = tn_graph(atoms)
    E, _, _ = potential_ref(m, atoms, graph=(og["i1"], og["i2"], og["off"]), dtype=torch.float64, calc_forces=False,
                            calc_stresses=False)
    assert abs(e1 - float(E)) < 1e-10
    for P in (2, 3):
        assert abs(_partitioned_energy(m, atoms, P, True) - e1) < 1e-10
        assert abs(_partitioned_energy(m, atoms, P, False) - e1) > 1e-4
