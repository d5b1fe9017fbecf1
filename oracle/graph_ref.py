"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

CPU restatement (numpy/scipy, float64 / int64) of the reference's graph builder:

  * periodic neighbour list  -- DistMLIP/distributed/fpis.c:418-901
  * slab partitioner + halo ("to"/"from") sections + bond (line) graph
                             -- DistMLIP/distributed/subgraph_creation_utils.c:26-931,
                                1102-1154, 1189-1322, 1370-1456, 1512-1529

and a loader for the reference's own C extension compiled by oracle/Makefile into
oracle/_ref/ (the live oracle that pins this restatement; see tests/test_oracle_graph.py).

Everything is returned in *canonical* form (sorted tuples / sets) so that orderings that
are an accident of OpenMP thread splits in the reference do not matter.
"""
from __future__ import annotations

import glob
import importlib.util
import os

import numpy as np

EPSILON = 1e-10  # subgraph_creation_utils.c:9


# --------------------------------------------------------------------------------------
# the reference's own compiled C extension (oracle/_ref)
# --------------------------------------------------------------------------------------
def load_ref_extension():
    """Import oracle/_ref/subgraph_creation_fast*.so; returns module or None."""
    here = os.path.dirname(os.path.abspath(__file__))
    cands = glob.glob(os.path.join(here, "_ref", "subgraph_creation_fast*.so"))
    if not cands:
        return None
    spec = importlib.util.spec_from_file_location("subgraph_creation_fast", cands[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ref_get_subgraphs(cart, frac_wrapped, lattice, pbc, num_partitions, cutoff, bond_cutoff,
                      use_bond_graph=True, tol=1e-8, num_threads=8):
    """Call the reference C extension exactly as dist.py:196-232 does. Returns the 19-tuple."""
    mod = load_ref_extension()
    if mod is None:
        raise RuntimeError("oracle/_ref not built (run `make -C oracle`)")
    cart = np.ascontiguousarray(cart, dtype=np.float64)
    frac = np.ascontiguousarray(frac_wrapped, dtype=np.float64)
    lattice = np.ascontiguousarray(lattice, dtype=np.float64)
    pbc = np.ascontiguousarray(pbc, dtype=np.int64)
    return mod.get_subgraphs_fast(cart, float(cutoff), pbc, lattice, int(num_partitions),
                                  float(bond_cutoff), float(tol), int(num_threads),
                                  bool(use_bond_graph), frac)


# --------------------------------------------------------------------------------------
# restatement: neighbour list (fpis.c)
# --------------------------------------------------------------------------------------
def wrap_frac(cart, lattice, pbc):
    """fpis.c:492-506: frac = cart @ inv(lattice); fmod to [0,1) on periodic axes.
    Returns (wrapped frac, integer correction) with frac_unwrapped = wrapped + corr."""
    frac = cart @ np.linalg.inv(lattice)
    wrapped = frac.copy()
    corr = np.zeros_like(frac)
    for j in range(3):
        if pbc[j]:
            w = np.fmod(frac[:, j], 1.0)
            w[w < 0] += 1.0
            wrapped[:, j] = w
            corr[:, j] = frac[:, j] - w
    return wrapped, corr


def neighbor_list(cart, lattice, pbc, r, bond_r, tol=1e-8):
    """Restates intra_parallel_find_points_in_spheres_c (fpis.c:418-901).

    Edge (i -> j, image) exists iff  tol < |x_j + off.L - x_i|^2 < r^2 + tol  and i != j
    (fpis.c:767, 834): periodic *self* images are never neighbours.
    Returns index_1 (centre i), index_2 (neighbour j), offsets[E,3] (integer-valued, relative
    to the *unwrapped* input coordinates, fpis.c:839-841), d2[E], bond mask (d2 < bond_r^2+tol,
    fpis.c:844), all sorted by (i, j, off).
    """
    from scipy.spatial import cKDTree

    cart = np.asarray(cart, dtype=np.float64)
    lattice = np.asarray(lattice, dtype=np.float64)
    n = len(cart)
    wrapped, corr = wrap_frac(cart, lattice, pbc)
    wc = wrapped @ lattice
    # image range: enough images that every point within r of any centre is present
    recip = np.linalg.inv(lattice).T  # rows = reciprocal vectors / 2pi
    heights = 1.0 / np.linalg.norm(recip, axis=1)
    # image range must reach the (possibly unwrapped) centres: fpis.c get_bounds(:489-491)
    frac_c = wrapped + corr
    ilo = [int(np.floor(frac_c[:, k].min() - (r + 1e-6) / h)) - 1 if pbc[k] else 0 for k, h in enumerate(heights)]
    ihi = [int(np.ceil(frac_c[:, k].max() + (r + 1e-6) / h)) + 1 if pbc[k] else 0 for k, h in enumerate(heights)]
    lo = cart.min(axis=0) - r - 1e-6
    hi = cart.max(axis=0) + r + 1e-6
    pts, idx, img = [], [], []
    for a in range(ilo[0], ihi[0] + 1):
        for b in range(ilo[1], ihi[1] + 1):
            for c in range(ilo[2], ihi[2] + 1):
                shift = a * lattice[0] + b * lattice[1] + c * lattice[2]
                p = wc + shift
                m = np.all((p > lo) & (p < hi), axis=1)
                if m.any():
                    pts.append(p[m])
                    idx.append(np.nonzero(m)[0])
                    img.append(np.broadcast_to(np.array([a, b, c], dtype=np.float64), (m.sum(), 3)))
    pts = np.concatenate(pts)
    idx = np.concatenate(idx)
    img = np.concatenate(img)
    tree = cKDTree(pts)
    ctree = cKDTree(cart)
    pairs = ctree.query_ball_tree(tree, np.sqrt(r * r + tol) + 1e-9)
    i1, i2 = [], []
    for i, lst in enumerate(pairs):
        if lst:
            i1.append(np.full(len(lst), i, dtype=np.int64))
            i2.append(np.array(lst, dtype=np.int64))
    i1 = np.concatenate(i1)
    k2 = np.concatenate(i2)
    dv = pts[k2] - cart[i1]
    d2 = np.einsum("ij,ij->i", dv, dv)
    j = idx[k2]
    keep = (d2 < r * r + tol) & (d2 > tol) & (j != i1)
    i1, k2, d2, j = i1[keep], k2[keep], d2[keep], j[keep]
    off = img[k2] - corr[j]
    off = np.rint(off).astype(np.int64)
    order = np.lexsort((off[:, 2], off[:, 1], off[:, 0], j, i1))
    i1, j, off, d2 = i1[order], j[order], off[order], d2[order]
    bond = d2 < bond_r * bond_r + tol
    return i1, j, off, d2, bond


# --------------------------------------------------------------------------------------
# restatement: partitioner (subgraph_creation_utils.c)
# --------------------------------------------------------------------------------------
def partition_rule(wrapped_cart, frac_wrapped, num_partitions):
    """create_partition (subgraph_creation_utils.c:1370-1456): longest *Cartesian* extent picks
    the axis, walls equally spaced in *fractional* coordinate between min and max, +EPSILON."""
    ext = wrapped_cart.max(axis=0) - wrapped_cart.min(axis=0)
    dim = 0
    for i in (1, 2):
        if ext[i] > ext[dim]:
            dim = i
    fmin = frac_wrapped[:, dim].min()
    fmax = frac_wrapped[:, dim].max()
    length = fmax - fmin
    walls = np.array([(i * (length / num_partitions)) + EPSILON + fmin for i in range(1, num_partitions)])
    # collision rule (:1436-1453)
    moved = True
    while moved:
        moved = False
        for w in range(len(walls)):
            if np.any(frac_wrapped[:, dim] == walls[w]):
                walls[w] += EPSILON
                moved = True
    return dim, walls


def which_partition(frac_dim_values, walls):
    """which_partition (:1312-1322): first wall strictly greater than the coordinate."""
    return np.searchsorted(walls, frac_dim_values, side="right").astype(np.int64)


def check_partition_size(dim, walls, lattice, atom_cutoff, bond_cutoff, use_bond_graph):
    """check_partition_size (:1512-1529). Returns True when the reference accepts."""
    col = np.array([lattice[0, dim], lattice[1, dim], lattice[2, dim]])
    width = walls[0] * np.linalg.norm(col)
    if use_bond_graph:
        return not (width <= 2 * (atom_cutoff + bond_cutoff))
    return not (width <= 2 * atom_cutoff)


class GraphOracle:
    """Canonical description of the partitioned graph for `num_partitions` slabs."""

    def __init__(self, cart, lattice, pbc, num_partitions, cutoff, bond_cutoff,
                 use_bond_graph=True, tol=1e-8, frac_wrapped=None):
        cart = np.asarray(cart, dtype=np.float64)
        lattice = np.asarray(lattice, dtype=np.float64)
        self.n = len(cart)
        self.P = num_partitions
        self.i1, self.i2, self.off, self.d2, self.bond = neighbor_list(
            cart, lattice, pbc, cutoff, bond_cutoff if use_bond_graph else 0.0, tol)
        if frac_wrapped is None:
            frac_wrapped, _ = wrap_frac(cart, lattice, pbc)
        self.frac = frac_wrapped
        wcart = frac_wrapped @ lattice
        if num_partitions > 1:
            self.dim, self.walls = partition_rule(wcart, frac_wrapped, num_partitions)
            self.owner = which_partition(frac_wrapped[:, self.dim], self.walls)
            self.accepts = check_partition_size(self.dim, self.walls, lattice, cutoff, bond_cutoff,
                                                use_bond_graph)
        else:
            self.dim, self.walls = 0, np.zeros(0)
            self.owner = np.zeros(self.n, dtype=np.int64)
            self.accepts = True
        # "to" partition of every atom: src atom of an edge whose dst lives elsewhere
        # (assign_to_partitions_test_2, :1189-1306). -1 = pure.
        src, dst = self.i1, self.i2
        cross = self.owner[src] != self.owner[dst]
        self.to_part = np.full(self.n, -1, dtype=np.int64)
        self.to_part[src[cross]] = self.owner[dst[cross]]
        # the reference asserts uniqueness (:1243-1248)
        chk = {}
        self.unique_to = True
        for s, q in zip(src[cross], self.owner[dst[cross]]):
            if chk.setdefault(s, q) != q:
                self.unique_to = False
                break

    # ---- atoms -------------------------------------------------------------------
    def owned(self, p):
        return np.nonzero(self.owner == p)[0]

    def to_list(self, p, q):
        """atoms owned by p exported to q, ascending global id (create_global_id_array order)."""
        return np.nonzero((self.owner == p) & (self.to_part == q))[0]

    def from_list(self, p, q):
        """halo atoms of p owned by q (== to_list(q, p))."""
        return self.to_list(q, p)

    # ---- edges -------------------------------------------------------------------
    def edges_of(self, p):
        """edges owned by p = edges whose dst (index_2) is owned by p (:178-250).
        Returns (src gid, dst gid, off[.,3]) sorted canonically by (dst, src, off)."""
        m = self.owner[self.i2] == p
        s, d, o = self.i1[m], self.i2[m], self.off[m]
        order = np.lexsort((o[:, 2], o[:, 1], o[:, 0], s, d))
        return s[order], d[order], o[order]

    # ---- bonds -------------------------------------------------------------------
    def bonds_owned(self, p):
        m = self.bond & (self.owner[self.i2] == p)
        s, d, o = self.i1[m], self.i2[m], self.off[m]
        order = np.lexsort((o[:, 2], o[:, 1], o[:, 0], s, d))
        return s[order], d[order], o[order]

    def bonds_halo(self, p):
        """'from' bond nodes of p: every bond whose dst atom is a halo atom of p
        (nodes_to_partition[dst] == p, :516-545)."""
        m = self.bond & (self.to_part[self.i2] == p) & (self.owner[self.i2] != p)
        s, d, o = self.i1[m], self.i2[m], self.off[m]
        order = np.lexsort((o[:, 2], o[:, 1], o[:, 0], s, d))
        return s[order], d[order], o[order]

    def angles_of(self, p):
        """line-graph edges of p: (s->d) -> (d->x), (d->x) owned by p, x != s (atom index
        compare, :716-718), centre d.  Returns array [A, 10]:
        (s, d, off_a[3], x, off_b[3], centre) sorted canonically."""
        bs, bd, bo = self.i1[self.bond], self.i2[self.bond], self.off[self.bond]
        # in-bonds grouped by dst
        order = np.argsort(bd, kind="stable")
        bs_d, bd_d, bo_d = bs[order], bd[order], bo[order]
        start = np.searchsorted(bd_d, np.arange(self.n), side="left")
        end = np.searchsorted(bd_d, np.arange(self.n), side="right")
        out = []
        own = self.owner[bd] == p
        for k in np.nonzero(own)[0]:
            d, x, ob = bs[k], bd[k], bo[k]
            for t in range(start[d], end[d]):
                s = bs_d[t]
                if s == x:
                    continue
                out.append((s, d, *bo_d[t], x, *ob, d))
        if not out:
            return np.zeros((0, 10), dtype=np.int64)
        arr = np.array(out, dtype=np.int64)
        order = np.lexsort(tuple(arr[:, c] for c in range(9, -1, -1)))
        return arr[order]


# --------------------------------------------------------------------------------------
# canonicalisation of the reference's 19-tuple so it can be compared with GraphOracle
# --------------------------------------------------------------------------------------
def canon_from_ref_tuple(t, P):
    """Turn get_subgraphs_fast's 19-tuple (subgraph_creation_fast.c:403-422) into canonical sets."""
    (src_nodes, dst_nodes, markers, _lc, global_ids, i1, i2, offs, dists, lsrc, ldst, within,
     lmarkers, nude, bmap_de, bmap_ude, l2g, _g2l, centers) = t
    out = {"i1": i1, "i2": i2, "off": np.rint(offs).astype(np.int64), "dist": dists, "within": within,
           "parts": []}
    for p in range(P):
        mk = np.append(markers[p], len(global_ids[p]))
        gid = global_ids[p]
        part = {
            "pure": np.sort(gid[mk[0]:mk[1]]),
            "to": [gid[mk[1 + q]:mk[2 + q]] for q in range(P)],
            "from": [gid[mk[1 + P + q]:mk[2 + P + q]] for q in range(P)],
            "n_owned": int(mk[1 + P]),
        }
        s = gid[src_nodes[p]]
        d = gid[dst_nodes[p]]
        o = out["off"][l2g[p]]
        order = np.lexsort((o[:, 2], o[:, 1], o[:, 0], s, d))
        part["edges"] = (s[order], d[order], o[order])
        assert np.array_equal(i1[l2g[p]], s) and np.array_equal(i2[l2g[p]], d)
        if len(lmarkers):
            lm = np.append(lmarkers[p], nude[p])
            part["n_bond_owned"] = int(lm[1 + P])
            part["n_bond_total"] = int(lm[-1])
            # owned bond nodes -> global edge via DE mapping
            ude2edge = np.full(int(lm[-1]), -1, dtype=np.int64)
            ude2edge[bmap_ude[p]] = l2g[p][bmap_de[p]]
            part["ude2edge"] = ude2edge
            part["line_src"] = lsrc[p]
            part["line_dst"] = ldst[p]
            part["center"] = gid[centers[p]]
        out["parts"].append(part)
    return out
