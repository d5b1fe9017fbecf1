"""GPU parity tests (run with `-m gpu` on a B200): the CUDA path through the C-ABI vs the oracle.

Tolerances: BASELINE.json's north_star asks for 1e-4 eV/atom and 1e-3 eV/A; with random-init weights the
forces are ~1e-2 eV/A, so the tests hold the engine to much tighter bounds (fp32 round-off level):
  dE/atom < 2e-7, dF < 2e-6 eV/A, dStress < 2e-6 GPa.  Integer / index work is compared bit-exactly.
"""
import json
import os

import numpy as np
import pytest
import torch

from distmlip_b200.structures import SimpleAtoms, rough_cell, si_diamond
from oracle import graph_ref as G
from oracle import manual_ref as M
from oracle.chgnet_ref import potential_ref
from tests._util import (digest, engine_from_model, engine_partition_digests, golden_cases, make_model, manual_run,
                         maps_to_oracle, oracle_graph, oracle_partition_digests)

pytestmark = pytest.mark.gpu
TOL_E, TOL_F, TOL_S = 2e-7, 2e-6, 2e-6
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "graph_golden.json")))


def species_of(model, atoms):
    return np.array([model.element_types.index(s) for s in atoms.get_chemical_symbols()], dtype=np.int32)


def run_engine(eng, model, atoms, forces=True, stress=True):
    eng.set_structure(atoms.get_positions(), atoms.get_cell(), species_of(model, atoms),
                      atoms.get_pbc().astype(np.int32))
    return eng.compute(forces, stress)


@pytest.fixture(scope="module")
def model():
    return make_model()


@pytest.fixture(scope="module")
def eng(model):
    e = engine_from_model(model)
    yield e
    e.close()


def check_vs_oracle(eng, model, atoms):
    E, F, S = run_engine(eng, model, atoms)
    Eo, Fo, So, _ = potential_ref(model, atoms, dtype=torch.float32)
    E64, F64, S64, _ = potential_ref(make_model().double(), atoms, dtype=torch.float64)
    n = len(atoms)
    assert abs(E - Eo.item()) / n < TOL_E and abs(E - E64.item()) / n < TOL_E
    assert np.abs(F - Fo.numpy()).max() < TOL_F and np.abs(F - F64.numpy()).max() < TOL_F
    assert np.abs(S - So.numpy()).max() < TOL_S and np.abs(S - S64.numpy()).max() < TOL_S
    return E, F, S


# ------------------------------------------------------------------ end-to-end parity
@pytest.mark.parametrize("n", [2, 4])  # 64 atoms; 512 atoms = BASELINE config[0]
def test_energy_forces_stress_match_oracle(eng, model, n):
    check_vs_oracle(eng, model, si_diamond(n))


def test_many_tiles_per_cta_match_oracle(eng, model):
    """8 000 atoms = 1 750 edge tiles and 750 angle tiles: every persistent CTA (grid = 2 x 148) runs 5-6 tiles, so the
    mbarrier phase flips, the TMEM reuse across tiles and the destination runs that straddle tile boundaries are checked
    against the oracle directly (not only through self-consistency properties)."""
    atoms = si_diamond(10, seed=31)
    check_vs_oracle(eng, model, atoms)
    c = eng.counts()
    assert c["n_edges"] // 128 > 5 * 2 * 148 and c["n_angles"] // 128 > 2 * 148


def test_rough_cell_8000_matches_oracle(eng, model):
    """degree-imbalanced structure (random sequential addition): 0-12 bonds and 10-40 edges per atom, so
    destination runs of every length occur inside and across tiles; > 5 tiles per CTA."""
    atoms = rough_cell(8000, seed=11)
    check_vs_oracle(eng, model, atoms)


def test_stage_taps_match_manual_mirror(eng, model):
    atoms = si_diamond(3)
    og = oracle_graph(atoms)
    run_engine(eng, model, atoms)
    ep, bp, ap = maps_to_oracle(eng, og)
    assert sorted(ep) == list(range(len(og["i1"]))) and sorted(bp) == list(range(len(og["bond_edges"])))
    assert sorted(ap) == list(range(len(og["la"])))
    gid = eng.partition_info(0)
    taps = manual_run(model, atoms, og)["taps"]
    for l in range(model.n_blocks + 1):
        assert np.abs(eng.debug_tensor(f"x{l}") - taps[f"x{l}"][gid].numpy()).max() < 5e-6
    for l in range(model.n_blocks):
        assert np.abs(eng.debug_tensor(f"h{l}") - taps[f"h{l}"][bp].numpy()).max() < 5e-6
    for l in range(model.n_blocks - 1):
        assert np.abs(eng.debug_tensor(f"ang{l}") - taps[f"ang{l}"][ap].numpy()).max() < 5e-6
    assert np.abs(eng.debug_tensor("gh") - taps["gh0"][bp].numpy()).max() < 1e-8
    assert np.abs(eng.debug_tensor("gang") - taps["gang0"][ap].numpy()).max() < 1e-10
    ev = eng.debug_tensor("e_vec")
    assert np.abs(ev[:, :3] - og["vec"][ep]).max() < 1e-6


def test_larger_weights_relative_parity(eng):
    """weights scaled so forces are O(0.1-1 eV/A): relative error stays at fp32 round-off."""
    m = make_model(scale=1.6)
    e2 = engine_from_model(m)
    atoms = si_diamond(3, seed=4)
    E, F, S = run_engine(e2, m, atoms)
    E64, F64, S64, _ = potential_ref(make_model(scale=1.6).double(), atoms, dtype=torch.float64)
    fmax = F64.abs().max().item()
    assert fmax > 0.05
    assert np.abs(F - F64.numpy()).max() < 2e-5 * max(1.0, fmax)
    assert abs(E - E64.item()) / len(atoms) < 1e-6 * max(1.0, abs(E64.item()) / len(atoms))
    e2.close()


# ------------------------------------------------------------------ edge cases of the graph builder
def sheared(atoms, a=3.0, b=1.5):
    lat = atoms.get_cell()
    lat[2, 0], lat[1, 0] = a, b
    return SimpleAtoms(atoms.get_chemical_symbols(), atoms.get_scaled_positions() @ lat, lat)


def test_triclinic_cell(eng, model):
    check_vs_oracle(eng, model, sheared(si_diamond(3, seed=5)))


def test_unwrapped_positions(eng, model):
    a = si_diamond(3, seed=7)
    pos, lat = a.get_positions(), a.get_cell()
    pos[::3] += lat[0] - 2 * lat[2]
    pos[1::5] -= lat[1]
    E, F, S = check_vs_oracle(eng, model, SimpleAtoms(a.get_chemical_symbols(), pos, lat))
    E0, F0, S0 = run_engine(eng, model, a)
    assert abs(E - E0) < 1e-5 and np.abs(F - F0).max() < 2e-6  # wrapping is a symmetry


def test_cell_smaller_than_cutoff(eng, model):
    """8-atom conventional cell (5.43 A < 2 r_cut): several periodic images per pair, stencil reach > 1."""
    atoms = si_diamond(1, seed=9)
    og = oracle_graph(atoms)
    run_engine(eng, model, atoms)
    assert eng.counts()["n_edges"] == len(og["i1"])
    check_vs_oracle(eng, model, atoms)


def test_non_periodic_cluster(eng, model):
    a = si_diamond(2, seed=2)
    atoms = SimpleAtoms(a.get_chemical_symbols(), a.get_positions(), a.get_cell(), pbc=(False, False, False))
    check_vs_oracle(eng, model, atoms)
    slab = SimpleAtoms(a.get_chemical_symbols(), a.get_positions(), a.get_cell(), pbc=(True, True, False))
    check_vs_oracle(eng, model, slab)


def test_mixed_species_and_rough_structure(eng, model):
    atoms = rough_cell(300, seed=3)
    sym = ["Si" if i % 3 else "Ge" for i in range(len(atoms))]
    atoms = SimpleAtoms(sym, atoms.get_positions(), atoms.get_cell())
    check_vs_oracle(eng, model, atoms)


def test_errors_are_reported_not_fatal(eng, model):
    from distmlip_b200._lib import B2MError

    lone = SimpleAtoms(["Si", "Si"], np.array([[0.0, 0, 0], [20.0, 20, 20]]), np.eye(3) * 40.0)
    with pytest.raises(B2MError) as ei:
        run_engine(eng, model, lone)
    assert "No neighbors" in str(ei.value)  # the reference exit()s here (fpis.c:634-635)
    # the handle stays usable
    check_vs_oracle(eng, model, si_diamond(2))
    eng.set_partition(0, 2)
    with pytest.raises(B2MError) as ei:
        run_engine(eng, model, si_diamond(4))  # 10.9 A slabs <= 2 (r_cut + r_bond)
    assert ei.value.code == -4 and "too close" in str(ei.value)
    eng.set_partition(0, 1)


# ------------------------------------------------------------------ partitioner: bit-exact vs golden / oracle
@pytest.mark.parametrize("name", sorted(GOLD))
def test_partition_matches_reference_golden(eng, model, name):
    atoms, P = golden_cases()[name]
    g = GOLD[name]
    try:
        for p in range(P):
            eng.set_partition(p, P)
            eng.set_structure(atoms.get_positions(), atoms.get_cell(), np.zeros(len(atoms), dtype=np.int32),
                              atoms.get_pbc().astype(np.int32))
            mine = engine_partition_digests(eng, P)
            assert mine == g["parts"][p], (name, p)
    finally:
        eng.set_partition(0, 1)


def test_single_partition_graph_equals_oracle(eng, model):
    atoms = si_diamond(5, seed=13)
    og = oracle_graph(atoms)
    eng.set_structure(atoms.get_positions(), atoms.get_cell(), np.zeros(len(atoms), dtype=np.int32),
                      atoms.get_pbc().astype(np.int32))
    assert digest(eng.partition_info(3)) == digest(np.column_stack([og["i1"], og["i2"], og["off"]]))
    be = og["bond_edges"]
    assert digest(eng.partition_info(4)) == digest(np.column_stack([og["i1"][be], og["i2"][be], og["off"][be]]))
    assert eng.counts()["n_angles"] == len(og["la"])


def test_halo_bond_sections_agree_between_ranks(eng, model):
    """what rank q sends (to-list bond rows) is exactly what rank p expects (halo bond rows), in order."""
    atoms = si_diamond(4, nz=8, seed=21)
    info = []
    try:
        for p in range(2):
            eng.set_partition(p, 2)
            eng.set_structure(atoms.get_positions(), atoms.get_cell(), np.zeros(len(atoms), dtype=np.int32),
                              atoms.get_pbc().astype(np.int32))
            c = eng.counts()
            bonds = eng.partition_info(4)
            info.append(dict(c=c, bonds=bonds, halo=eng.partition_info(1), to=eng.partition_info(6)))
    finally:
        eng.set_partition(0, 1)
    for p in range(2):
        q = 1 - p
        to_q = info[q]["to"]
        assert np.array_equal(to_q[to_q[:, 0] == p, 1], info[p]["halo"])  # atoms: same order (gid ascending)
        sent_atoms = set(info[p]["halo"].tolist())
        owned_q = info[q]["bonds"][: info[q]["c"]["n_bond_own"]]
        sent = owned_q[np.isin(owned_q[:, 1], list(sent_atoms))]
        halo_p = info[p]["bonds"][info[p]["c"]["n_bond_own"]:]
        assert digest(sent) == digest(halo_p) and len(halo_p) > 0


# ------------------------------------------------------------------ Potential / calculator surface
def test_potential_and_calculator_surface(model):
    from distmlip_b200.implementations.matgl import CHGNet_Dist, PESCalculator_Dist, Potential_Dist

    atoms = si_diamond(2)
    dm = CHGNet_Dist.from_existing(make_model())
    dm.enable_distributed_mode([0])
    with pytest.raises(Exception):
        dm.enable_distributed_mode([0])  # chgnet.py:457-458
    refs = np.zeros(len(dm.element_types))
    refs[dm.element_types.index("Si")] = -0.25
    pot = Potential_Dist(model=dm, data_mean=1.5, data_std=2.0, element_refs=refs, calc_site_wise=True)
    E, F, S, H, site = pot(atoms)
    Eo, Fo, So, siteo = potential_ref(model, atoms, data_mean=1.5, data_std=2.0, element_refs=refs)
    assert H is None and abs(E.item() - Eo.item()) / len(atoms) < 1e-6
    assert (F - Fo).abs().max().item() < 4e-6 and (S - So).abs().max().item() < 4e-6
    assert (site - siteo).abs().max().item() < 5e-6
    calc = PESCalculator_Dist(potential=pot, stress_unit="eV/A3", use_voigt=True)
    calc.calculate(atoms, ["energy", "forces", "stress"])
    assert set(calc.results) >= {"energy", "free_energy", "forces", "stress", "magmoms"}
    assert calc.results["stress"].shape == (6,) and calc.results["forces"].shape == (len(atoms), 3)
    pot2 = Potential_Dist(model=dm, calc_forces=False, calc_stresses=False)
    out = pot2(atoms)
    assert out[1] is None and out[2] is None


# ------------------------------------------------------------------ size-independent properties at bench size
@pytest.fixture(scope="module")
def big(eng, model):
    atoms = si_diamond(23)  # 97 336 atoms: the workload bench.py times (BASELINE config[1])
    E, F, S = run_engine(eng, model, atoms)
    return atoms, E, F, S


def test_fullsize_net_force_and_counts(eng, big):
    atoms, E, F, S = big
    c = eng.counts()
    assert abs(c["n_edges"] / len(atoms) - 28.0) < 0.1 and abs(c["n_bond_own"] / len(atoms) - 4.0) < 0.05
    assert np.isfinite(F).all() and np.abs(F.sum(0)).max() < 5e-3  # Newton's third law (fp32 atomics)
    assert np.abs(S - S.T).max() < 1e-5  # symmetric virial


def test_fullsize_translation_and_permutation(eng, model, big):
    atoms, E, F, S = big
    pos = atoms.get_positions()
    shifted = SimpleAtoms(atoms.get_chemical_symbols(), pos + np.array([1.234, -0.77, 3.1]), atoms.get_cell())
    E2, F2, _ = run_engine(eng, model, shifted)
    assert abs(E2 - E) / len(atoms) < 1e-7 and np.abs(F2 - F).max() < 5e-6
    perm = np.random.default_rng(0).permutation(len(atoms))
    permuted = SimpleAtoms(atoms.get_chemical_symbols(), pos[perm], atoms.get_cell())
    E3, F3, _ = run_engine(eng, model, permuted)
    assert abs(E3 - E) / len(atoms) < 1e-7 and np.abs(F3 - F[perm]).max() < 5e-6


def test_supercell_extensivity(eng, model):
    """a periodic cell repeated twice along z has exactly twice the energy and the same forces."""
    a = si_diamond(6, seed=17)
    lat = a.get_cell()
    pos = a.get_positions()
    lat2 = lat.copy()
    lat2[2] *= 2
    b = SimpleAtoms(a.get_chemical_symbols() * 2, np.vstack([pos, pos + lat[2]]), lat2)
    E1, F1, S1 = run_engine(eng, model, a)
    E2, F2, S2 = run_engine(eng, model, b)
    assert abs(E2 - 2 * E1) / len(b) < 2e-7
    assert np.abs(F2[: len(a)] - F1).max() < 2e-6 and np.abs(F2[len(a):] - F1).max() < 2e-6
    assert np.abs(S2 - S1).max() < 2e-6


def test_axis_permutation_symmetry(eng, model):
    """cubic cell: cyclic permutation of the Cartesian axes permutes forces/stress, energy unchanged."""
    a = si_diamond(4, seed=19)
    E1, F1, S1 = run_engine(eng, model, a)
    P = [1, 2, 0]
    b = SimpleAtoms(a.get_chemical_symbols(), a.get_positions()[:, P], a.get_cell()[P][:, P])
    E2, F2, S2 = run_engine(eng, model, b)
    assert abs(E2 - E1) / len(a) < 2e-7
    assert np.abs(F2 - F1[:, P]).max() < 2e-6 and np.abs(S2 - S1[P][:, P]).max() < 2e-6


# ------------------------------------------------------------------ the two kernel generations agree on the device
def test_tcgen05_and_ffma_paths_agree(model):
    """default path = tcgen05 (3xTF32 in TMEM); B2M_LEGACY_FFMA=1 selects the FP32-FFMA tile kernels."""
    atoms = si_diamond(4, seed=23)
    old = os.environ.get("B2M_LEGACY_FFMA")
    try:
        os.environ["B2M_LEGACY_FFMA"] = "1"
        e_ffma = engine_from_model(model)
        os.environ["B2M_LEGACY_FFMA"] = "0"
        e_tc = engine_from_model(model)
    finally:
        if old is None:
            os.environ.pop("B2M_LEGACY_FFMA", None)
        else:
            os.environ["B2M_LEGACY_FFMA"] = old
    E1, F1, S1 = run_engine(e_ffma, model, atoms)
    E2, F2, S2 = run_engine(e_tc, model, atoms)
    assert abs(E1 - E2) / len(atoms) < 1e-7 and np.abs(F1 - F2).max() < 1e-6 and np.abs(S1 - S2).max() < 1e-6
    e_ffma.close()
    e_tc.close()


@pytest.mark.parametrize("nb", [2, 3])
def test_other_block_counts(nb):
    """n_blocks is read from the model (chgnet.py:298): the layer loops, the dead last angle update and the
    saved-tensor bookkeeping must hold for depths other than the default 4."""
    m = make_model(seed=5, num_blocks=nb)
    e = engine_from_model(m)
    atoms = si_diamond(3, seed=29)
    E, F, S = run_engine(e, m, atoms)
    Eo, Fo, So, _ = potential_ref(make_model(seed=5, num_blocks=nb).double(), atoms, dtype=torch.float64)
    assert abs(E - Eo.item()) / len(atoms) < TOL_E and np.abs(F - Fo.numpy()).max() < TOL_F
    assert np.abs(S - So.numpy()).max() < TOL_S
    e.close()


def test_release_workspace_then_reuse(model):
    """b2m_release_workspace frees the resident graph and every per-structure buffer; the handle keeps its weights and the
    next set_structure allocates again (bench.py uses it to make room for its single-partition check)"""
    import torch as _t

    e = engine_from_model(model)
    atoms = si_diamond(6, seed=41)
    E1, F1, S1 = run_engine(e, model, atoms)
    used = _t.cuda.mem_get_info()[0]
    e.release_workspace()
    assert _t.cuda.mem_get_info()[0] > used  # memory came back
    from distmlip_b200._lib import B2MError

    with pytest.raises(B2MError):
        e.compute(True, True)  # no structure any more: an error, not a crash
    E2, F2, S2 = run_engine(e, model, atoms)
    assert abs(E1 - E2) / len(atoms) < 2e-8 and np.abs(F1 - F2).max() < 1e-6
    e.close()


def test_empty_and_degenerate_inputs(eng, model):
    from distmlip_b200._lib import B2MError

    with pytest.raises(B2MError):
        eng.set_structure(np.zeros((0, 3)), np.eye(3) * 10, np.zeros(0, dtype=np.int32), np.ones(3, dtype=np.int32))
    with pytest.raises(B2MError) as ei:
        eng.set_structure(np.zeros((2, 3)) + [[0, 0, 0], [1, 1, 1]], np.zeros((3, 3)), np.zeros(2, dtype=np.int32),
                          np.ones(3, dtype=np.int32))
    assert "singular" in str(ei.value)
    check_vs_oracle(eng, model, si_diamond(2))  # handle still usable
