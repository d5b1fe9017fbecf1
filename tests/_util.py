"""Shared helpers for parity tests: build the oracle graph, map engine-local orderings to it."""
from __future__ import annotations

import numpy as np
import torch

from distmlip_b200.structures import si_diamond
from oracle import graph_ref as G
from oracle import manual_ref as M
from oracle.chgnet_ref import CHGNetRef, build_line_graph, potential_ref


def make_model(seed=0, scale=1.0, num_blocks=4):
    m = CHGNetRef(seed=seed, num_blocks=num_blocks)
    if scale != 1.0:
        with torch.no_grad():
            for n, p in m.named_parameters():
                if "weight" in n and p.ndim == 2 and "embedding" not in n:
                    p.mul_(scale)
    return m


def engine_from_model(model, device=0, data_mean=0.0, data_std=1.0, element_refs=None):
    from distmlip_b200 import _lib

    sd = model.state_dict()
    eng = _lib.Engine(n_elem=sd["atom_embedding.weight"].shape[0], dim=64, max_n=9, max_f=4,
                      n_blocks=model.n_blocks, cutoff=float(model.cutoff),
                      three_body_cutoff=float(model.three_body_cutoff), cutoff_exponent=int(model.cutoff_exponent),
                      device=device)
    eng.load_state_dict({k: v.float() for k, v in sd.items()})
    eng.set_scaling(data_mean, data_std)
    if element_refs is not None:
        eng.set_element_refs(element_refs)
    eng.finalize()
    return eng


def oracle_graph(atoms, rc=5.0, rb=3.0):
    cart = atoms.get_positions()
    lat = atoms.get_cell()
    pbc = atoms.get_pbc().astype(np.int64)
    i1, i2, off, d2, bond = G.neighbor_list(cart, lat, pbc, rc, rb)
    bond_edges, la, lb, ce = build_line_graph(i1, i2, bond)
    vec = cart[i2] + off @ lat - cart[i1]
    return dict(i1=i1, i2=i2, off=off, d2=d2, bond=bond, bond_edges=bond_edges, la=la, lb=lb, ce=ce, vec=vec)


def key5(a):
    """rows (src, dst, ox, oy, oz) -> list of tuples"""
    return [tuple(int(v) for v in r) for r in a]


def maps_to_oracle(eng, og):
    """Returns (edge_perm, bond_perm, angle_perm): engine-local index -> oracle index."""
    ekey = {k: i for i, k in enumerate(key5(np.column_stack([og["i1"], og["i2"], og["off"]])))}
    ee = eng.partition_info(3)
    edge_perm = np.array([ekey[k] for k in key5(ee)], dtype=np.int64)
    be = og["bond_edges"]
    bkey = {k: i for i, k in enumerate(key5(np.column_stack([og["i1"][be], og["i2"][be], og["off"][be]])))}
    bb = eng.partition_info(4)
    bond_perm = np.array([bkey[k] for k in key5(bb)], dtype=np.int64)
    akey = {(int(a), int(b)): i for i, (a, b) in enumerate(zip(og["la"], og["lb"]))}
    aa = eng.partition_info(5)
    angle_perm = np.array([akey[(int(bond_perm[a]), int(bond_perm[b]))] for a, b, _c in aa], dtype=np.int64)
    return edge_perm, bond_perm, angle_perm


def manual_run(model, atoms, og, dtype=torch.float64, data_std=1.0):
    types = np.array([model.element_types.index(s) for s in atoms.get_chemical_symbols()])
    return M.run(model, types, og["vec"], og["i1"], og["i2"], og["bond_edges"], og["la"], og["lb"], og["ce"],
                 data_std=data_std, dtype=dtype)


# ---- order-independent digests shared with tests/golden/make_golden.py ----
MOD = (1 << 61) - 1


def digest(rows):
    rows = np.asarray(rows, dtype=np.int64)
    if rows.size == 0:
        return [0, 0, 0]
    rows = rows.reshape(len(rows), -1)
    acc = np.zeros(len(rows), dtype=object)
    for c in range(rows.shape[1]):
        acc = (acc * 1000003 + (rows[:, c].astype(object) + 7919)) % MOD
    return [int(len(rows)), int(sum(acc) % MOD), int(sum((a * a) % MOD for a in acc) % MOD)]


def golden_cases():
    import importlib.util
    import os

    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "golden", "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.cases()


def oracle_partition_digests(o, p):
    """same fields as make_golden.describe()['parts'][p], from a GraphOracle"""
    P = o.P
    s, d, of = o.edges_of(p)
    bs, bd, bo = o.bonds_owned(p)
    hs, _hd, _ho = o.bonds_halo(p)
    ang = o.angles_of(p)
    return {
        "n_owned": int(len(o.owned(p))),
        "owned": digest(o.owned(p)[:, None]),
        "to": [digest(o.to_list(p, q)[:, None]) if q != p else [0, 0, 0] for q in range(P)],
        "from": [digest(o.from_list(p, q)[:, None]) if q != p else [0, 0, 0] for q in range(P)],
        "edges": digest(np.column_stack([s, d, of])),
        "bonds_owned": digest(np.column_stack([bs, bd, bo])),
        "n_bond_halo": int(len(hs)),
        "n_angles": int(len(ang)),
        "angle_dst_center": digest(ang[:, [1, 5, 6, 7, 8, 9]]) if len(ang) else [0, 0, 0],
    }


def engine_partition_digests(eng, P):
    c = eng.counts()
    own = np.sort(eng.partition_info(0))
    halo, howner = eng.partition_info(1), eng.partition_info(2)
    tl = eng.partition_info(6)
    edges = eng.partition_info(3)
    bonds = eng.partition_info(4)
    ang = eng.partition_info(5)
    nbo = c["n_bond_own"]
    outb = bonds[ang[:, 1]] if len(ang) else np.zeros((0, 5), dtype=np.int64)
    return {
        "n_owned": int(c["n_own"]),
        "owned": digest(own[:, None]),
        "to": [digest(np.sort(tl[tl[:, 0] == q, 1])[:, None]) if len(tl) else [0, 0, 0] for q in range(P)],
        "from": [digest(halo[howner == q][:, None]) for q in range(P)],
        "edges": digest(edges),
        "bonds_owned": digest(bonds[:nbo]),
        "n_bond_halo": int(c["n_bond_halo"]),
        "n_angles": int(c["n_angles"]),
        "angle_dst_center": digest(np.column_stack([outb, ang[:, 2]])) if len(ang) else [0, 0, 0],
    }
