"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Stage-by-stage mirror of the engine's TensorNet forward and hand-derived reverse pass (csrc/kernels_tn.cu,
csrc/engine_tn.cuh), written on the same decomposed storage the kernels use, so that every intermediate the engine can
export through `b2m_debug_tensor` has a tensor of the same name and layout here.  tests/test_oracle_tensornet.py checks
this mirror against autograd through oracle/tensornet_ref.py; the GPU tests check the engine against both.

Storage ("decomposed form"): a per-atom, per-channel 3x3 tensor M = I*eye + skew(a) + S is held as 10 numbers
  k = 0: I | 1..3: a_x, a_y, a_z (skew(a) = [[0,-az,ay],[az,0,-ax],[-ay,ax,0]]) | 4..9: S_xx, S_xy, S_xz, S_yy, S_yz, S_zz
in an array [n, 10, C].  Channel mixing (the `linears_tensor` of matgl's TensorNet) acts on each of the 10 rows
independently with the weight of its part, so the form is closed under everything but the 3x3 products, which go
through `full` / `dec`.  Adjoints are kept in the same parameter space.
"""
from __future__ import annotations

import math

import numpy as np
import torch

PART = [0, 1, 1, 1, 2, 2, 2, 2, 2, 2]


def silu(x):
    return x * torch.sigmoid(x)


def dsilu(x):
    s = torch.sigmoid(x)
    return s * (1 + x * (1 - s))


def full(T):
    """[n,10,C] -> [n,C,3,3]"""
    I, ax, ay, az, xx, xy, xz, yy, yz, zz = [T[:, k, :] for k in range(10)]
    rows = [torch.stack([I + xx, xy - az, xz + ay], -1), torch.stack([xy + az, I + yy, yz - ax], -1),
            torch.stack([xz - ay, yz + ax, I + zz], -1)]
    return torch.stack(rows, -2)


def full_adj(G):
    """adjoint of `full`: dE/dM [n,C,3,3] -> parameter-space adjoint [n,10,C]"""
    g = lambda i, j: G[..., i, j]
    return torch.stack([g(0, 0) + g(1, 1) + g(2, 2), g(2, 1) - g(1, 2), g(0, 2) - g(2, 0), g(1, 0) - g(0, 1),
                        g(0, 0), g(0, 1) + g(1, 0), g(0, 2) + g(2, 0), g(1, 1), g(1, 2) + g(2, 1), g(2, 2)], 1)


def dec(M):
    """decompose_tensor in parameter form: [n,C,3,3] -> [n,10,C]"""
    m = lambda i, j: M[..., i, j]
    I = (m(0, 0) + m(1, 1) + m(2, 2)) / 3
    return torch.stack([I, 0.5 * (m(2, 1) - m(1, 2)), 0.5 * (m(0, 2) - m(2, 0)), 0.5 * (m(1, 0) - m(0, 1)),
                        m(0, 0) - I, 0.5 * (m(0, 1) + m(1, 0)), 0.5 * (m(0, 2) + m(2, 0)), m(1, 1) - I,
                        0.5 * (m(1, 2) + m(2, 1)), m(2, 2) - I], 1)


def dec_adj(g):
    """adjoint of `dec`: [n,10,C] -> dE/dM [n,C,3,3]"""
    gI, gax, gay, gaz, gxx, gxy, gxz, gyy, gyz, gzz = [g[:, k, :] for k in range(10)]
    t = (gI - gxx - gyy - gzz) / 3
    rows = [torch.stack([t + gxx, 0.5 * (gxy - gaz), 0.5 * (gxz + gay)], -1),
            torch.stack([0.5 * (gxy + gaz), t + gyy, 0.5 * (gyz - gax)], -1),
            torch.stack([0.5 * (gxz - gay), 0.5 * (gyz + gax), t + gzz], -1)]
    return torch.stack(rows, -2)


NW = torch.tensor([3.0, 2.0, 2.0, 2.0, 1.0, 2.0, 2.0, 1.0, 2.0, 1.0])  # tensor_norm = sum_k NW[k] T_k^2


def nrm(T):
    return (NW.to(T.dtype)[None, :, None] * T * T).sum(1)


def scale_fwd(T):
    """X / (tensor_norm(X) + 1)"""
    q = nrm(T) + 1
    return T / q[:, None, :], q


def scale_bwd(T, q, gout):
    dot = (gout * T).sum(1)
    return gout / q[:, None, :] - (dot / (q * q))[:, None, :] * (2 * NW.to(T.dtype)[None, :, None] * T)


def mix(T, Ws):
    """Ws: three [C,C] nn.Linear weights (out, in), one per part"""
    return torch.stack([T[:, k, :] @ Ws[PART[k]].T for k in range(10)], 1)


def mix_adj(g, Ws):
    return torch.stack([g[:, k, :] @ Ws[PART[k]] for k in range(10)], 1)


def layer_norm(x, w, b, eps=1e-5):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    rstd = 1 / torch.sqrt(var + eps)
    xh = (x - mu) * rstd
    return xh * w + b, xh, rstd


def layer_norm_bwd(gy, xh, rstd, w):
    gh = gy * w
    return rstd * (gh - gh.mean(-1, keepdim=True) - xh * (gh * xh).mean(-1, keepdim=True))


def sym6(v):
    n2 = (v * v).sum(1)
    x, y, z = v[:, 0], v[:, 1], v[:, 2]
    return torch.stack([x * x - n2 / 3, x * y, x * z, y * y - n2 / 3, y * z, z * z - n2 / 3], 1)


def run(model, node_types, vec, i_src, i_dst, data_std=1.0, dtype=torch.float64):
    """Forward + reverse pass for one graph.  Returns dict(energy, gvec [E,3] = dE/dvec, taps)."""
    sd = {k: v.detach().to(dtype) for k, v in model.state_dict().items()}
    t = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.int64)
    src, dst, types = t(i_src), t(i_dst), t(node_types)
    vec = torch.as_tensor(np.asarray(vec), dtype=dtype)
    n, E, C = len(types), len(src), model.units
    rc = float(model.cutoff)
    so3 = model.equivariance_invariance_group != "O(3)"
    width = float(model.bond_expansion.rbf.width)
    mu = sd["bond_expansion.rbf.centers"]
    taps = {}

    # ---------------- geometry ----------------
    d = torch.linalg.norm(vec, dim=1)
    vh = vec / d[:, None]
    diff = d[:, None] - mu[None, :]
    rbf = torch.exp(-width * diff * diff)
    drbf = rbf * (-2 * width * diff)
    inside = d <= rc
    Cc = torch.where(inside, 0.5 * (torch.cos(math.pi * d / rc) + 1), torch.zeros_like(d))
    dCc = torch.where(inside, -0.5 * math.pi / rc * torch.sin(math.pi * d / rc), torch.zeros_like(d))
    taps.update(rbf=rbf, cut=Cc)

    # ---------------- embedding ----------------
    te = "tensor_embedding."
    Wd = torch.cat([sd[te + f"distance_proj{k}.weight"] for k in (1, 2, 3)], 0)  # [3C, nrbf]
    bd = torch.cat([sd[te + f"distance_proj{k}.bias"] for k in (1, 2, 3)], 0)
    P = rbf @ Wd.T + bd
    z = sd[te + "emb.weight"]
    W2 = sd[te + "emb2.weight"]
    U = z @ W2[:, :C].T  # per-species halves of emb2
    V = z @ W2[:, C:].T + sd[te + "emb2.bias"]
    Zij = U[types[src]] + V[types[dst]]
    cz = Cc[:, None] * Zij
    s6 = sym6(vh)
    edge_rows = [P[:, :C] * cz] + [P[:, C:2 * C] * cz * vh[:, a:a + 1] for a in range(3)] + \
                [P[:, 2 * C:] * cz * s6[:, s:s + 1] for s in range(6)]
    T0 = torch.zeros(n, 10, C, dtype=dtype).index_add_(0, dst, torch.stack(edge_rows, 1))
    nr0 = nrm(T0)
    ln0, xh0, rstd0 = layer_norm(nr0, sd[te + "init_norm.weight"], sd[te + "init_norm.bias"])
    s1p = ln0 @ sd[te + "linears_scalar.0.weight"].T + sd[te + "linears_scalar.0.bias"]
    s1 = silu(s1p)
    s2p = s1 @ sd[te + "linears_scalar.1.weight"].T + sd[te + "linears_scalar.1.bias"]
    sc = silu(s2p).reshape(n, C, 3)
    Wt_e = [sd[te + f"linears_tensor.{k}.weight"] for k in range(3)]
    T0m = mix(T0, Wt_e)
    X = torch.stack([T0m[:, k, :] * sc[:, :, PART[k]] for k in range(10)], 1)
    taps.update(P=P, T0=T0, ln0=ln0, s2p=s2p, T0m=T0m, X0=X)

    # ---------------- interaction layers ----------------
    saved = []
    for l in range(model.nblocks):
        p = f"layers.{l}."
        Ws = [sd[p + f"linears_scalar.{k}.weight"] for k in range(3)]
        bs = [sd[p + f"linears_scalar.{k}.bias"] for k in range(3)]
        Wt = [sd[p + f"linears_tensor.{k}.weight"] for k in range(6)]
        f1p = rbf @ Ws[0].T + bs[0]
        f2p = silu(f1p) @ Ws[1].T + bs[1]
        f3p = silu(f2p) @ Ws[2].T + bs[2]
        f3 = silu(f3p)
        fe = (f3 * Cc[:, None]).reshape(E, C, 3)
        Xh, q = scale_fwd(X)
        Y = mix(Xh, Wt[:3])
        msg = torch.zeros(n, 10, C, dtype=dtype).index_add_(
            0, dst, torch.stack([fe[:, :, PART[k]] * Y[src, k, :] for k in range(10)], 1))
        Mf, Yf = full(msg), full(Y)
        Pf = 2 * Yf @ Mf if so3 else Mf @ Yf + Yf @ Mf
        Pd = dec(Pf)
        Pn, q2 = scale_fwd(Pd)
        dX = mix(Pn, Wt[3:])
        Df = full(dX)
        Xn = Xh + dX + dec(Df @ Df)
        saved.append(dict(X=X, q=q, Xh=Xh, Y=Y, msg=msg, Pd=Pd, q2=q2, Pn=Pn, dX=dX, f1p=f1p, f2p=f2p, f3p=f3p, f3=f3,
                          fe=fe, Ws=Ws, Wt=Wt))
        taps.update({f"f3p{l}": f3p, f"Xh{l}": Xh, f"Y{l}": Y, f"msg{l}": msg, f"Pn{l}": Pn, f"dX{l}": dX, f"X{l + 1}": Xn})
        X = Xn

    # ---------------- readout ----------------
    inv = torch.cat([3 * X[:, 0] ** 2, 2 * (X[:, 1:4] ** 2).sum(1),
                     (NW[4:].to(dtype)[None, :, None] * X[:, 4:] ** 2).sum(1)], -1)
    r, xhr, rstdr = layer_norm(inv, sd["out_norm.weight"], sd["out_norm.bias"])
    x = r @ sd["linear.weight"].T + sd["linear.bias"]
    idx = sorted({int(k.split(".")[3]) for k in sd if k.startswith("final_layer.gated.layers.")})
    chain = {"layers": [], "gates": []}
    outs = {}
    for br in ("layers", "gates"):
        hcur = x
        for j, i in enumerate(idx):
            W, b = sd[f"final_layer.gated.{br}.{i}.weight"], sd[f"final_layer.gated.{br}.{i}.bias"]
            pre = hcur @ W.T + b
            chain[br].append((hcur, pre, W))
            hcur = silu(pre) if j < len(idx) - 1 else pre
        outs[br] = hcur
    gate = torch.sigmoid(outs["gates"])
    e_atom = outs["layers"] * gate
    energy = e_atom.sum()
    taps.update(inv=inv, xr=x, e_atom=e_atom)

    # ================= reverse pass =================
    gl = torch.full_like(e_atom, data_std) * gate
    gg = torch.full_like(e_atom, data_std) * outs["layers"] * gate * (1 - gate)
    gx = torch.zeros_like(x)
    for br, gcur in (("layers", gl), ("gates", gg)):
        for j in reversed(range(len(idx))):
            hin, pre, W = chain[br][j]
            if j < len(idx) - 1:
                gcur = gcur * dsilu(pre)
            gcur = gcur @ W
        gx = gx + gcur
    gr = gx @ sd["linear.weight"]
    ginv = layer_norm_bwd(gr, xhr, rstdr, sd["out_norm.weight"])
    gX = torch.zeros_like(X)
    gX[:, 0] = ginv[:, :C] * 6 * X[:, 0]
    gX[:, 1:4] = ginv[:, None, C:2 * C] * 4 * X[:, 1:4]
    gX[:, 4:] = ginv[:, None, 2 * C:] * 2 * NW[4:].to(dtype)[None, :, None] * X[:, 4:]
    taps[f"gX{model.nblocks}"] = gX

    g_rbf = torch.zeros_like(rbf)
    gC = torch.zeros_like(d)
    for l in reversed(range(model.nblocks)):
        s = saved[l]
        Wt = s["Wt"]
        # Xn = Xh + dX + dec(D D)
        Df = full(s["dX"])
        Gsq = dec_adj(gX)
        gdX = gX + full_adj(Gsq @ Df.transpose(-1, -2) + Df.transpose(-1, -2) @ Gsq)
        gXh = gX.clone()
        gPn = mix_adj(gdX, Wt[3:])
        gPd = scale_bwd(s["Pd"], s["q2"], gPn)
        G = dec_adj(gPd)
        Mf, Yf = full(s["msg"]), full(s["Y"])
        if so3:
            gMf = 2 * Yf.transpose(-1, -2) @ G
            gYf = 2 * G @ Mf.transpose(-1, -2)
        else:
            gMf = G @ Yf.transpose(-1, -2) + Yf.transpose(-1, -2) @ G
            gYf = Mf.transpose(-1, -2) @ G + G @ Mf.transpose(-1, -2)
        gmsg, gY = full_adj(gMf), full_adj(gYf)
        # message: msg[t] = sum_e fe[e,:,part] * Y[src]
        gm_e = gmsg[dst]  # [E,10,C]
        Ysrc = s["Y"][src]
        gfe = torch.stack([(gm_e[:, [k for k in range(10) if PART[k] == p], :] *
                            Ysrc[:, [k for k in range(10) if PART[k] == p], :]).sum(1) for p in range(3)], -1)  # [E,C,3]
        gY = gY.index_add(0, src, torch.stack([s["fe"][:, :, PART[k]] * gm_e[:, k, :] for k in range(10)], 1))
        gXh = gXh + mix_adj(gY, Wt[:3])
        gX = scale_bwd(s["X"], s["q"], gXh)
        # edge MLP
        gfe = gfe.reshape(E, 3 * C)
        gC = gC + (gfe * s["f3"]).sum(1)
        g3 = gfe * Cc[:, None] * dsilu(s["f3p"])
        g2 = (g3 @ s["Ws"][2]) * dsilu(s["f2p"])
        g1 = (g2 @ s["Ws"][1]) * dsilu(s["f1p"])
        g_rbf = g_rbf + g1 @ s["Ws"][0]
        taps.update({f"gX{l}": gX, f"gY{l}": gY, f"gmsg{l}": gmsg, f"gdX{l}": gdX, f"gf{l}": gfe})

    # embedding:  X0_k = T0m_k * sc[:, :, part(k)]
    gT0m = torch.stack([gX[:, k, :] * sc[:, :, PART[k]] for k in range(10)], 1)
    gsc = torch.stack([sum(gX[:, k, :] * T0m[:, k, :] for k in range(10) if PART[k] == p) for p in range(3)], -1)
    gs2p = gsc.reshape(n, 3 * C) * dsilu(s2p)
    gs1p = (gs2p @ sd[te + "linears_scalar.1.weight"]) * dsilu(s1p)
    gln0 = gs1p @ sd[te + "linears_scalar.0.weight"]
    gnr0 = layer_norm_bwd(gln0, xh0, rstd0, sd[te + "init_norm.weight"])
    gT0 = mix_adj(gT0m, Wt_e) + gnr0[:, None, :] * 2 * NW.to(dtype)[None, :, None] * T0
    taps["gT0"] = gT0
    ge = gT0[dst]  # [E,10,C]
    wI = ge[:, 0, :] * cz
    wA = ge[:, 1:4, :] * cz[:, None, :]
    wS = ge[:, 4:, :] * cz[:, None, :]
    sA = (wA * vh[:, :, None]).sum(1)  # [E,C]
    sS = (wS * s6[:, :, None]).sum(1)
    gP = torch.cat([wI, sA, sS], 1)
    # d/dCc
    gC = gC + (Zij * (ge[:, 0, :] * P[:, :C] + P[:, C:2 * C] * (ge[:, 1:4, :] * vh[:, :, None]).sum(1)
                      + P[:, 2 * C:] * (ge[:, 4:, :] * s6[:, :, None]).sum(1))).sum(1)
    # d/dvh
    a3 = (wA * P[:, None, C:2 * C]).sum(2)  # [E,3]
    w6 = (wS * P[:, None, 2 * C:]).sum(2)   # [E,6]: xx,xy,xz,yy,yz,zz
    tr = w6[:, 0] + w6[:, 3] + w6[:, 5]
    x_, y_, z_ = vh[:, 0], vh[:, 1], vh[:, 2]
    gvh = a3 + torch.stack([2 * w6[:, 0] * x_ + w6[:, 1] * y_ + w6[:, 2] * z_ - 2 * x_ / 3 * tr,
                            w6[:, 1] * x_ + 2 * w6[:, 3] * y_ + w6[:, 4] * z_ - 2 * y_ / 3 * tr,
                            w6[:, 2] * x_ + w6[:, 4] * y_ + 2 * w6[:, 5] * z_ - 2 * z_ / 3 * tr], 1)
    g_rbf = g_rbf + gP @ Wd
    gd = (g_rbf * drbf).sum(1) + gC * dCc
    gvec = gd[:, None] * vh + (gvh - (gvh * vh).sum(1, keepdim=True) * vh) / d[:, None]
    taps.update(gP=gP, g_rbf=g_rbf, gd=gd, gvh=gvh, gC=gC)
    return dict(energy=energy, gvec=gvec, taps=taps)
