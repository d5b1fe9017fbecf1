"""GPU: the TensorNet path (csrc/kernels_tn.cu through the C-ABI) against the oracle (oracle/tensornet_ref.py, autograd)
and, stage by stage, against the mirror of the kernels' schedule (oracle/tensornet_manual.py)."""
import numpy as np
import pytest
import torch

from distmlip_b200.structures import SimpleAtoms, rough_cell, si_diamond
from tests._util import key5
from tests.test_oracle_tensornet import make_tn, tn_graph

pytestmark = pytest.mark.gpu


def tn_engine(model, device=0, data_mean=0.0, data_std=1.0, element_refs=None):
    from distmlip_b200 import _lib

    sd = model.state_dict()
    eng = _lib.Engine(n_elem=sd["tensor_embedding.emb.weight"].shape[0], n_blocks=model.nblocks,
                      cutoff=float(model.cutoff), device=device,
                      tensornet=dict(units=model.units, num_rbf=model.num_rbf,
                                     so3=model.equivariance_invariance_group == "SO(3)",
                                     rbf_width=float(model.bond_expansion.rbf.width)))
    eng.load_state_dict({k: v.float() for k, v in sd.items()})
    eng.set_scaling(data_mean, data_std)
    if element_refs is not None:
        eng.set_element_refs(element_refs)
    eng.finalize()
    return eng


def mixed(atoms, other="O", every=3):
    sym = [other if i % every == 0 else s for i, s in enumerate(atoms.get_chemical_symbols())]
    return SimpleAtoms(sym, atoms.get_positions(), atoms.get_cell())


def set_structure(eng, model, atoms):
    sp = np.array([model.element_types.index(s) for s in atoms.get_chemical_symbols()], dtype=np.int32)
    eng.set_structure(atoms.get_positions(), np.array(atoms.get_cell()), sp, atoms.get_pbc().astype(np.int32))
    return sp


def oracle_efs(model, atoms, **kw):
    from oracle.tensornet_ref import potential_ref

    og = tn_graph(atoms, float(model.cutoff))
    return potential_ref(model, atoms, graph=(og["i1"], og["i2"], og["off"]), dtype=torch.float64, **kw), og


def tap_table(eng, model, atoms, og, data_std=1.0):
    """(name, max |engine - mirror|, max |mirror|) for every exported intermediate of the forward and reverse pass"""
    from oracle import tensornet_manual as TM

    types = np.array([model.element_types.index(s) for s in atoms.get_chemical_symbols()])
    out = TM.run(model, types, og["vec"], og["i1"], og["i2"], data_std=data_std)
    taps = {k: v.numpy() for k, v in out["taps"].items()}
    ekey = {k: i for i, k in enumerate(key5(np.column_stack([og["i1"], og["i2"], og["off"]])))}
    eperm = np.array([ekey[k] for k in key5(eng.partition_info(3))], dtype=np.int64)
    own, halo = eng.partition_info(0), eng.partition_info(1)
    loc = np.concatenate([own, halo]).astype(np.int64)
    nb = model.nblocks
    rows = []

    def cmp(name, got, ref):
        rows.append((name, float(np.abs(got - ref).max()), float(np.abs(ref).max())))

    def node(name, ref, owned_only):
        got = eng.debug_tensor(name)
        idx = own if owned_only or got.shape[0] == len(own) else loc
        ref = ref[idx]
        cmp(name, got.reshape(ref.shape) if got.size == ref.size else got.reshape(len(idx), 10, -1), ref)

    cmp("rbf", eng.debug_tensor("rbf")[:, : model.num_rbf], taps["rbf"][eperm])
    cmp("cut", eng.debug_tensor("cut")[:, 0], taps["cut"][eperm])
    cmp("P", eng.debug_tensor("P"), taps["P"][eperm])
    for name in ("T0", "ln0", "s2p", "T0m"):
        node(name, taps[name], True)
    for l in range(nb):
        node(f"X{l}", taps[f"X{l}"], False)
        cmp(f"f3p{l}", eng.debug_tensor(f"f3p{l}"), taps[f"f3p{l}"][eperm])
        node(f"Xh{l}", taps[f"Xh{l}"], False)
        node(f"Y{l}", taps[f"Y{l}"], False)
        for name in (f"msg{l}", f"Pn{l}", f"dX{l}"):
            node(name, taps[name], True)
    node(f"X{nb}", taps[f"X{nb}"], True)
    node("inv", taps["inv"], True)
    node("xr", taps["xr"], True)
    node("e_atom", taps["e_atom"], True)
    for name in ("gdX0", "gmsg0", "gT0"):
        node(name, taps[name], True)
    if len(halo) == 0:
        node("gY0", taps["gY0"], False)
        node("gX0", taps["gX0"], False)
    cmp("gP", eng.debug_tensor("gP"), taps["gP"][eperm])
    cmp("g_rbf", eng.debug_tensor("g_rbf")[:, : model.num_rbf], taps["g_rbf"][eperm])
    cmp("gC", eng.debug_tensor("gC")[:, 0], taps["gC"][eperm])
    cmp("gvh", eng.debug_tensor("gvh"), taps["gvh"][eperm])
    cmp("gd", eng.debug_tensor("gd")[:, 0], taps["gd"][eperm])
    return rows


def check_efs(eng, model, atoms, ref, tol_e=2e-6, tol_f=2e-5):
    (E, F, S), _og = ref
    e, f, s = eng.compute(forces=True, stress=True)
    n = len(atoms)
    fs = max(1.0, float(F.abs().max()))
    assert abs(e - float(E)) / n < tol_e, (e, float(E))
    assert np.abs(f - F.numpy()).max() < tol_f * fs, np.abs(f - F.numpy()).max()
    assert np.abs(s - S.numpy()).max() < 2e-4 * max(1.0, float(S.abs().max())), (s, S)
    return e, f, s


@pytest.mark.parametrize("group", ["O(3)", "SO(3)"])
def test_stage_taps_and_efs_small(group):
    atoms = mixed(si_diamond(2, sigma=0.15, seed=1))
    model = make_tn(seed=3, scale=1.5, equivariance_invariance_group=group)
    eng = tn_engine(model, data_std=1.3)
    set_structure(eng, model, atoms)
    ref = oracle_efs(model, atoms, data_std=1.3)
    e, f, s = eng.compute(forces=True, stress=True)
    rows = tap_table(eng, model, atoms, ref[1], data_std=1.3)
    bad = [r for r in rows if r[1] > 2e-4 * max(r[2], 1e-3)]
    print("\n".join(f"{n:8s} err {a:.3e}  ref {b:.3e}" for n, a, b in rows))
    assert not bad, bad
    check_efs(eng, model, atoms, ref)


def test_efs_larger_cells_and_options():
    model = make_tn(seed=5, scale=1.5, nblocks=3)
    eng = tn_engine(model, data_mean=0.7, data_std=0.9, element_refs=np.linspace(-0.5, 0.5, len(model.element_types)))
    for atoms in (mixed(si_diamond(4, sigma=0.15, seed=2)), mixed(rough_cell(300, seed=4), other="Ge", every=2)):
        set_structure(eng, model, atoms)
        ref = oracle_efs(model, atoms, data_mean=0.7, data_std=0.9,
                         element_refs=np.linspace(-0.5, 0.5, len(model.element_types)))
        check_efs(eng, model, atoms, ref)


def test_translation_and_repeat_are_stable():
    atoms = mixed(si_diamond(3, sigma=0.12, seed=7))
    model = make_tn(seed=2, scale=1.5)
    eng = tn_engine(model)
    set_structure(eng, model, atoms)
    e0, f0, _ = eng.compute()
    e1, f1, _ = eng.compute()
    assert abs(e0 - e1) < 2e-6 * len(atoms) and np.abs(f0 - f1).max() < 1e-5
    shifted = SimpleAtoms(atoms.get_chemical_symbols(), atoms.get_positions() + np.array([0.3, -1.1, 2.2]), atoms.get_cell())
    set_structure(eng, model, shifted)
    e2, f2, _ = eng.compute()
    assert abs(e0 - e2) < 5e-6 * len(atoms) and np.abs(f0 - f2).max() < 5e-5


@pytest.mark.parametrize("parts", [2, 3])
def test_single_process_group_matches_oracle(parts):
    atoms = mixed(si_diamond(2, sigma=0.15, seed=3, nz=4 * parts))
    model = make_tn(seed=4, scale=1.5)
    eng = tn_engine(model, device=[0] * parts)
    set_structure(eng, model, atoms)
    ref = oracle_efs(model, atoms)
    check_efs(eng, model, atoms, ref)
    single = tn_engine(model)
    set_structure(single, model, atoms)
    e1, f1, s1 = single.compute()
    e2, f2, s2 = eng.compute()
    assert abs(e1 - e2) < 2e-6 * len(atoms) and np.abs(f1 - f2).max() < 2e-5


def test_plugin_surface_tensornet_dist():
    from distmlip_b200.implementations.matgl import Potential_Dist, TensorNet_Dist

    atoms = mixed(si_diamond(2, sigma=0.15, seed=5, nz=8))
    model = make_tn(seed=6, scale=1.5)
    (E, F, S), _ = oracle_efs(model, atoms)
    dist = TensorNet_Dist.from_existing(model)
    dist.enable_distributed_mode([0, 0])
    pot = Potential_Dist(model=dist, calc_stresses=True)
    e, f, s, h = pot(atoms)
    assert h is None and pot.last_dist_info.num_partitions == 2
    assert abs(float(e) - float(E)) < 2e-6 * len(atoms)
    assert np.abs(f.numpy() - F.numpy()).max() < 2e-5 * max(1.0, float(F.abs().max()))
    assert np.abs(s.numpy() - S.numpy()).max() < 2e-4 * max(1.0, float(S.abs().max()))
    with pytest.raises(NotImplementedError):
        bad = make_tn(seed=6)
        bad.is_intensive = True
        TensorNet_Dist.from_existing(bad).enable_distributed_mode([0])
