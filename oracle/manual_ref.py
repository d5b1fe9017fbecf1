"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Stage-by-stage CPU mirror of the dataflow the CUDA engine uses (factorised first layers,
hand-derived backward, no autograd).  It exists to (1) prove the factorisation and the manual
backward equal oracle/chgnet_ref.py + autograd, and (2) give per-stage tensors that GPU debug
taps can be diffed against.  Same provenance caveat as chgnet_ref.py: PARITY UNPINNED w.r.t. matgl.

Stage names match distmlip_b200/csrc/*.cu kernels.  Formulas: SURVEY.md §8 a5-a15, §9.
"""
from __future__ import annotations

import math

import numpy as np
import torch


def silu(x):
    return x * torch.sigmoid(x)


def dsilu(x):
    s = torch.sigmoid(x)
    return s * (1 + x * (1 - s))


def _cat(a, b):
    return torch.cat([a, b], 0)


class Weights:
    """Pre-composed weight views (what b2m_load_weights + finalize computes on the host)."""

    def __init__(self, sd, n_blocks, cutoff, cutoff3, p, dtype):
        g = lambda k: sd[k].detach().to(dtype)
        self.rc, self.rc3, self.p = cutoff, cutoff3, p
        self.f2 = g("bond_expansion.frequencies")
        self.f3 = g("threebody_bond_expansion.frequencies")
        self.fa = g("angle_expansion.frequencies")
        self.emb = g("atom_embedding.weight")
        self.Wbe = g("bond_embedding.layers.0.weight")  # [64,9]
        self.Wae = g("angle_embedding.layers.0.weight")
        self.Wabw = g("atom_bond_weights.weight")
        self.W3bw = g("threebody_bond_weights.weight")
        self.atom = []
        for l in range(n_blocks):
            pre = f"atom_graph_layers.{l}.conv_layer."
            W1 = _cat(g(pre + "node_update_func.layers.layers.0.weight"),
                      g(pre + "node_update_func.gates.layers.0.weight"))  # [128,192]
            b1 = _cat(g(pre + "node_update_func.layers.layers.0.bias"),
                      g(pre + "node_update_func.gates.layers.0.bias"))
            D = W1.shape[1] // 3
            self.atom.append(dict(
                W1s=W1[:, :D], W1e=W1[:, D:2 * D], W1t=W1[:, 2 * D:], b1=b1,
                M=W1[:, D:2 * D] @ self.Wbe,  # [128,9]
                W2L=g(pre + "node_update_func.layers.layers.1.weight"),
                b2L=g(pre + "node_update_func.layers.layers.1.bias"),
                W2G=g(pre + "node_update_func.gates.layers.1.weight"),
                b2G=g(pre + "node_update_func.gates.layers.1.bias"),
                Wout=g(pre + "node_out_func.weight")))
        self.bond = []
        for l in range(n_blocks - 1):
            pre = f"bond_graph_layers.{l}.conv_layer."
            W1 = _cat(g(pre + "node_update_func.layers.layers.0.weight"),
                      g(pre + "node_update_func.gates.layers.0.weight"))  # [128,256]
            b1 = _cat(g(pre + "node_update_func.layers.layers.0.bias"),
                      g(pre + "node_update_func.gates.layers.0.bias"))
            D = W1.shape[1] // 4
            WA = _cat(g(pre + "edge_update_func.layers.layers.0.weight"),
                      g(pre + "edge_update_func.gates.layers.0.weight"))
            bA = _cat(g(pre + "edge_update_func.layers.layers.0.bias"),
                      g(pre + "edge_update_func.gates.layers.0.bias"))
            self.bond.append(dict(
                W1a=W1[:, :D], W1g=W1[:, D:2 * D], W1c=W1[:, 2 * D:3 * D], W1b=W1[:, 3 * D:], b1=b1,
                W2L=g(pre + "node_update_func.layers.layers.1.weight"),
                b2L=g(pre + "node_update_func.layers.layers.1.bias"),
                W2G=g(pre + "node_update_func.gates.layers.1.weight"),
                b2G=g(pre + "node_update_func.gates.layers.1.bias"),
                Wout=g(pre + "node_out_func.weight"),
                WAa=WA[:, :D], WAg=WA[:, D:2 * D], WAc=WA[:, 2 * D:3 * D], WAb=WA[:, 3 * D:], bA=bA))
        self.Ws = g("sitewise_readout.weight")
        self.bs = g("sitewise_readout.bias")
        self.F = [(g(f"final_layer.layers.{i}.weight"), g(f"final_layer.layers.{i}.bias")) for i in range(3)]


def rbf_env(d, freq, rc, p):
    """be_k = env(rbf_k) * rbf_k and d(be_k)/dd.  (chgnet.py:116-124)."""
    c = math.sqrt(2.0 / rc)
    arg = d[:, None] * freq[None, :] / rc
    s, co = torch.sin(arg), torch.cos(arg)
    rbf = c * s / d[:, None]
    drbf = c * (freq[None, :] / rc * co / d[:, None] - s / (d[:, None] ** 2))
    c1 = -(p + 1) * (p + 2) / 2
    c2 = p * (p + 2)
    c3 = -p * (p + 1) / 2
    rho = rbf / rc
    env = 1 + c1 * rho**p + c2 * rho ** (p + 1) + c3 * rho ** (p + 2)
    denv = (c1 * p * rho ** (p - 1) + c2 * (p + 1) * rho**p + c3 * (p + 2) * rho ** (p + 1)) / rc
    ok = rbf <= rc
    be = torch.where(ok, env * rbf, torch.zeros_like(rbf))
    dbe_drbf = torch.where(ok, env + rbf * denv, torch.zeros_like(rbf))
    return be, dbe_drbf * drbf


def gated2(pre, W, hidden=True):
    """second half of a GatedMLP given first-layer pre-activations [.,128]."""
    D = pre.shape[1] // 2
    if hidden:
        hid = silu(pre)
        u = hid[:, :D] @ W["W2L"].T + W["b2L"]
        v = hid[:, D:] @ W["W2G"].T + W["b2G"]
    else:
        u, v = pre[:, :D], pre[:, D:]
    oL, oG = silu(u), torch.sigmoid(v)
    return u, v, oL, oG


def gated2_bwd(pre, u, oL, oG, goL, goG, W, hidden=True):
    gu = goL * dsilu(u)
    gv = goG * oG * (1 - oG)
    if hidden:
        ghid = torch.cat([gu @ W["W2L"], gv @ W["W2G"]], 1)
        return ghid * dsilu(pre)
    return torch.cat([gu, gv], 1)


def run(model, node_types, vec, i_src, i_dst, bond_edges, la, lb, ce, data_std=1.0, dtype=torch.float64):
    """Forward + manual backward over the global graph.
    Returns dict with energy, gvec [E,3] (dE/dvec_e), taps."""
    sd = model.state_dict()
    W = Weights(sd, model.n_blocks, float(model.cutoff), float(model.three_body_cutoff),
                int(model.cutoff_exponent), dtype)
    nb = model.n_blocks
    t = lambda a: torch.as_tensor(a, dtype=torch.int64)
    i_src, i_dst, bond_edges, la, lb, ce, node_types = map(t, (i_src, i_dst, bond_edges, la, lb, ce, node_types))
    vec = torch.as_tensor(vec, dtype=dtype)
    n = len(node_types)
    E = len(i_src)
    B = len(bond_edges)
    A = len(la)
    taps = {}
    # ---- geometry / expansions (K_edge_geom, K_bond_init, K_angle_init) ----
    d = vec.norm(dim=1)
    be, dbe = rbf_env(d, W.f2, W.rc, W.p)
    bvec, bd = vec[bond_edges], d[bond_edges]
    tbe, dtbe = rbf_env(bd, W.f3, W.rc3, W.p)
    va, vb = bvec[la], bvec[lb]
    na, nbn = va.norm(dim=1), vb.norm(dim=1)
    cos_raw = -(va * vb).sum(1) / (na * nbn)
    lo, hi = -1 + 1e-7, 1 - 1e-7
    cc = cos_raw.clamp(lo, hi)
    theta = torch.acos(cc)
    karg = theta[:, None] * W.fa[None, :]
    four = torch.zeros(A, 2 * len(W.fa) - 1, dtype=dtype)
    four[:, ::2] = torch.cos(karg) / math.pi
    four[:, 1::2] = torch.sin(karg[:, 1:]) / math.pi
    x = [W.emb[node_types]]
    h = [be[bond_edges] @ W.Wbe.T]
    ang = [four @ W.Wae.T]
    wab = be @ W.Wabw.T
    w3b = tbe @ W.W3bw.T
    edge_bond = torch.full((E,), -1, dtype=torch.int64)
    edge_bond[bond_edges] = torch.arange(B)
    isb = edge_bond >= 0
    saved = []

    def atom_fwd(l, xin, hin):
        w = W.atom[l]
        Aa = xin @ w["W1s"].T
        Cc = xin @ w["W1t"].T + w["b1"]
        T = be @ w["M"].T
        if l > 0:
            Q = hin @ w["W1e"].T
            T = torch.where(isb[:, None], Q[edge_bond.clamp(min=0)], T)
        pre = Aa[i_src] + Cc[i_dst] + T
        u, v, oL, oG = gated2(pre, w)
        msg = oL * oG * wab
        agg = torch.zeros(n, msg.shape[1], dtype=dtype).index_add_(0, i_dst, msg)
        return xin + agg @ w["Wout"].T, (pre, u, oL, oG)

    for l in range(nb - 1):
        xn, sa = atom_fwd(l, x[l], h[l])
        x.append(xn)
        w = W.bond[l]
        Ha = h[l] @ w["W1a"].T
        Hb = h[l] @ w["W1b"].T + w["b1"]
        Xc = xn @ w["W1c"].T
        pre = Ha[la] + ang[l] @ w["W1g"].T + Xc[ce] + Hb[lb]
        u, v, oL, oG = gated2(pre, w)
        aggB = torch.zeros(B, 64, dtype=dtype).index_add_(0, lb, oL * oG) if A else torch.zeros(B, 64, dtype=dtype)
        upd = aggB @ w["Wout"].T
        hn = h[l] + upd * w3b
        h.append(hn)
        Ha2 = hn @ w["WAa"].T
        Hb2 = hn @ w["WAb"].T + w["bA"]
        Xc2 = xn @ w["WAc"].T
        pre2 = Ha2[la] + ang[l] @ w["WAg"].T + Xc2[ce] + Hb2[lb]
        u2, v2, oL2, oG2 = gated2(pre2, w, hidden=False)
        ang.append(ang[l] + oL2 * oG2)
        saved.append(dict(atom=sa, bpre=pre, bu=u, boL=oL, boG=oG, upd=upd, apre=pre2, au=u2, aoL=oL2, aoG=oG2))
        taps[f"x{l + 1}"], taps[f"h{l + 1}"], taps[f"ang{l + 1}"] = xn, hn, ang[-1]
    site = x[nb - 1] @ W.Ws.T + W.bs
    xn, sa_last = atom_fwd(nb - 1, x[nb - 1], h[nb - 1])
    x.append(xn)
    taps[f"x{nb}"] = xn
    y1p = xn @ W.F[0][0].T + W.F[0][1]
    y1 = silu(y1p)
    y2p = y1 @ W.F[1][0].T + W.F[1][1]
    y2 = silu(y2p)
    ea = y2 @ W.F[2][0].T + W.F[2][1]
    energy = data_std * ea.sum()
    # ---------------- backward ----------------
    gy2 = data_std * W.F[2][0].expand(n, -1)
    gy1 = (gy2 * dsilu(y2p)) @ W.F[1][0]
    gx = (gy1 * dsilu(y1p)) @ W.F[0][0]
    gbe = torch.zeros(E, be.shape[1], dtype=dtype)
    gtbe = torch.zeros(B, tbe.shape[1], dtype=dtype)
    gh = torch.zeros(B, 64, dtype=dtype)
    gang = torch.zeros(A, 64, dtype=dtype)

    def atom_bwd(l, gxn, sa, xin, hin):
        nonlocal gbe
        w = W.atom[l]
        pre, u, oL, oG = sa
        gagg = gxn @ w["Wout"]
        gm = gagg[i_dst]
        gwab = gm * oL * oG
        gpre = gated2_bwd(pre, u, oL, oG, gm * oG * wab, gm * oL * wab, w)
        gA = torch.zeros(n, 128, dtype=dtype).index_add_(0, i_src, gpre)
        gC = torch.zeros(n, 128, dtype=dtype).index_add_(0, i_dst, gpre)
        gbe = gbe + gwab @ W.Wabw
        ghin = None
        if l > 0:
            gbe = gbe + torch.where(isb[:, None], torch.zeros_like(gpre), gpre) @ w["M"]
            gQ = gpre[bond_edges]
            ghin = gQ @ w["W1e"]
        else:
            gbe = gbe + gpre @ w["M"]
        return gxn + gA @ w["W1s"] + gC @ w["W1t"], ghin

    gx, ghl = atom_bwd(nb - 1, gx, sa_last, x[nb - 1], h[nb - 1])
    gh = gh + ghl
    taps[f"gx{nb - 1}"] = gx
    for l in range(nb - 2, -1, -1):
        s = saved[l]
        w = W.bond[l]
        # angle update backward (dead for the last block: gang == 0 there)
        gpre2 = gated2_bwd(s["apre"], s["au"], s["aoL"], s["aoG"], gang * s["aoG"], gang * s["aoL"], w, hidden=False)
        gang = gang + gpre2 @ w["WAg"]
        gHa2 = torch.zeros(B, 128, dtype=dtype).index_add_(0, la, gpre2)
        gHb2 = torch.zeros(B, 128, dtype=dtype).index_add_(0, lb, gpre2)
        gXc2 = torch.zeros(n, 128, dtype=dtype).index_add_(0, ce, gpre2)
        gh = gh + gHa2 @ w["WAa"] + gHb2 @ w["WAb"]
        gx = gx + gXc2 @ w["WAc"]
        # bond conv backward
        gtbe = gtbe + (gh * s["upd"]) @ W.W3bw
        gaggB = (gh * w3b) @ w["Wout"]
        gm = gaggB[lb]
        gpre = gated2_bwd(s["bpre"], s["bu"], s["boL"], s["boG"], gm * s["boG"], gm * s["boL"], w)
        gang = gang + gpre @ w["W1g"]
        gHa = torch.zeros(B, 128, dtype=dtype).index_add_(0, la, gpre)
        gHb = torch.zeros(B, 128, dtype=dtype).index_add_(0, lb, gpre)
        gXc = torch.zeros(n, 128, dtype=dtype).index_add_(0, ce, gpre)
        gh = gh + gHa @ w["W1a"] + gHb @ w["W1b"]
        gx = gx + gXc @ w["W1c"]
        # atom conv backward
        gx, ghl = atom_bwd(l, gx, s["atom"], x[l], h[l])
        if ghl is not None:
            gh = gh + ghl
        taps[f"gx{l}"], taps[f"gh{l}"], taps[f"gang{l}"] = gx, gh, gang
    # ---- init backward ----
    gbe_b = gh @ W.Wbe  # h0 = be[bond] @ Wbe^T
    gd = (gbe * dbe).sum(1)
    gd_b = (gbe_b * dbe[bond_edges]).sum(1) + (gtbe * dtbe).sum(1)
    gd = gd.index_add(0, bond_edges, gd_b)
    gfour = gang @ W.Wae
    gth = ((-gfour[:, ::2] * W.fa[None, :] * torch.sin(karg)).sum(1)
           + (gfour[:, 1::2] * W.fa[None, 1:] * torch.cos(karg[:, 1:])).sum(1)) / math.pi
    inside = (cos_raw >= lo) & (cos_raw <= hi)
    gcos = torch.where(inside, -gth / torch.sqrt(1 - cc * cc), torch.zeros_like(gth))
    dva = -vb / (na * nbn)[:, None] - cos_raw[:, None] * va / (na**2)[:, None]
    dvb = -va / (na * nbn)[:, None] - cos_raw[:, None] * vb / (nbn**2)[:, None]
    gbvec = torch.zeros(B, 3, dtype=dtype).index_add_(0, la, gcos[:, None] * dva).index_add_(0, lb, gcos[:, None] * dvb)
    gvec = gd[:, None] * vec / d[:, None]
    gvec = gvec.index_add(0, bond_edges, gbvec)
    taps.update(be=be, tbe=tbe, theta=theta, x0=x[0], h0=h[0], ang0=ang[0], gd=gd, gbvec=gbvec, e_atom=ea)
    return dict(energy=energy, gvec=gvec, site=site, taps=taps)


def forces_from_gvec(gvec, vec, i_src, i_dst, n, volume):
    """pos_bar[dst] += g ; pos_bar[src] -= g ; F = -pos_bar ; strain_bar = sum vec (x) g (pes.py:122-145)."""
    t = lambda a: torch.as_tensor(a, dtype=torch.int64)
    pb = torch.zeros(n, 3, dtype=gvec.dtype).index_add_(0, t(i_dst), gvec).index_add_(0, t(i_src), -gvec)
    virial = torch.as_tensor(vec, dtype=gvec.dtype).T @ gvec
    return -pb, virial / volume * 160.21766208
