// engine_tn.inl -- TensorNet schedule (included by engine.cu after the halo-exchange helpers).
// Stage by stage the same as oracle/tensornet_manual.py; reference control flow replaced:
// DistMLIP/implementations/matgl/models/tensornet.py:10-147 (forward) and pes.py:109-145 (scaling, reverse pass, forces,
// stress).  One deviation, on purpose: the embedded tensors are exchanged before the first interaction layer (the
// reference's first atom_transfer follows layer 0, tensornet.py:119-127, which leaves the halo sources of layer 0 at
// zero and makes its result depend on the partition count).

static constexpr int TNC = 64;         // channels (units)
static constexpr int TNW = 10 * TNC;   // floats per atom, decomposed form

// B2M_TN_TRACE=1: synchronise and name every stage on stderr (locating a faulting or hanging kernel)
static void tn_stage(b2m_engine* e, const char* name) {
  static const bool on = [] {
    const char* v = getenv("B2M_TN_TRACE");
    return v && v[0] == '1';
  }();
  if (!on) return;
  fprintf(stderr, "[tn] %s ...", name);
  fflush(stderr);
  cudaError_t r = cudaStreamSynchronize(e->st);
  fprintf(stderr, " %s\n", cudaGetErrorString(r));
  fflush(stderr);
}

static const std::vector<float>& TWt(b2m_engine* e, const std::string& k, std::vector<int64_t> shape) {
  auto it = e->host_w.find(k);
  B2M_REQUIRE(it != e->host_w.end(), B2M_ERR_INVALID, "missing weight: " + k);
  e->consumed.insert(k);
  B2M_REQUIRE(e->host_shape[k] == shape, B2M_ERR_INVALID,
              "weight '" + k + "' has an unsupported shape (TensorNet engine: units = 64, Gaussian expansion <= 64 centres)");
  return it->second;
}
// [rows][cols] -> [rows][cols_pad] (zero columns appended)
static std::vector<float> pad_cols(const std::vector<float>& m, int rows, int ncol, int pad) {
  std::vector<float> o((size_t)rows * pad, 0.f);
  for (int r = 0; r < rows; r++)
    for (int c = 0; c < ncol; c++) o[(size_t)r * pad + c] = m[(size_t)r * ncol + c];
  return o;
}

static void tn_finalize_weights(b2m_engine* e) {
  TnState& t = *e->tn;
  const int C = TNC, nr = t.num_rbf, nrp = t.rp.nrp, nb = t.nblocks, ne = e->desc.n_elem;
  e->consumed.clear();
  Packer P;
  std::map<std::string, size_t> off;
  auto put = [&](const std::string& name, const std::vector<float>& v) { off[name] = P.add(v); };
  // forward operand of a Linear [out][in]: its transpose [in][out]; K (= in) padded with zero rows to kpad
  auto fwd = [&](const std::vector<float>& w, int out, int in, int kpad) {
    std::vector<float> tr = transpose(w, out, in);  // [in][out]
    tr.resize((size_t)kpad * out, 0.f);
    return tr;
  };
  const auto& mu = TWt(e, "bond_expansion.rbf.centers", {nr});
  for (int k = 0; k < 64; k++) t.rp.mu[k] = k < nr ? mu[k] : 0.f;
  const std::string te = "tensor_embedding.";
  {
    std::vector<float> Wd, bd;
    for (int k = 1; k <= 3; k++) {
      const auto& w = TWt(e, te + "distance_proj" + std::to_string(k) + ".weight", {C, nr});
      const auto& b = TWt(e, te + "distance_proj" + std::to_string(k) + ".bias", {C});
      Wd.insert(Wd.end(), w.begin(), w.end());
      bd.insert(bd.end(), b.begin(), b.end());
    }
    put("Wd_t", fwd(Wd, 3 * C, nr, nrp));
    put("Wd_r", pad_cols(Wd, 3 * C, nr, nrp));
    {
      const std::vector<float> Wp = pad_cols(Wd, 3 * C, nr, nrp);  // [3C][64]
      auto rows = [&](int r0, int n) { return std::vector<float>(Wp.begin() + (size_t)r0 * nrp, Wp.begin() + (size_t)(r0 + n) * nrp); };
      put("Wdc_a", canon_split(rows(0, 2 * C), 2 * C, nrp, nrp));
      put("Wdc_b", canon_split(rows(2 * C, C), C, nrp, nrp));
      for (int j = 0; j < 3; j++) put("Wdrc" + std::to_string(j), canon_split(transpose(rows(j * C, C), C, nrp), nrp, C, C));
    }
    put("bd", bd);
    const auto& emb = TWt(e, te + "emb.weight", {ne, C});
    const auto& W2 = TWt(e, te + "emb2.weight", {C, 2 * C});
    const auto& b2 = TWt(e, te + "emb2.bias", {C});
    std::vector<float> U((size_t)ne * C), V((size_t)ne * C);
    for (int z = 0; z < ne; z++)
      for (int c = 0; c < C; c++) {
        double u = 0, v = b2[c];
        for (int k = 0; k < C; k++) {
          u += (double)emb[(size_t)z * C + k] * W2[(size_t)c * 2 * C + k];
          v += (double)emb[(size_t)z * C + k] * W2[(size_t)c * 2 * C + C + k];
        }
        U[(size_t)z * C + c] = (float)u, V[(size_t)z * C + c] = (float)v;
      }
    put("U", U), put("V", V);
    for (int k = 0; k < 3; k++) {
      const auto& w = TWt(e, te + "linears_tensor." + std::to_string(k) + ".weight", {C, C});
      put("Wte_t" + std::to_string(k), transpose(w, C, C));
      put("Wte_r" + std::to_string(k), w);
    }
    put("ln0_g", TWt(e, te + "init_norm.weight", {C})), put("ln0_b", TWt(e, te + "init_norm.bias", {C}));
    const auto& s0 = TWt(e, te + "linears_scalar.0.weight", {2 * C, C});
    const auto& s1 = TWt(e, te + "linears_scalar.1.weight", {3 * C, 2 * C});
    put("Ws0_t", transpose(s0, 2 * C, C)), put("Ws0_r", s0), put("bs0", TWt(e, te + "linears_scalar.0.bias", {2 * C}));
    put("Ws1_t", transpose(s1, 3 * C, 2 * C)), put("Ws1_r", s1), put("bs1", TWt(e, te + "linears_scalar.1.bias", {3 * C}));
  }
  for (int l = 0; l < nb; l++) {
    const std::string p = "layers." + std::to_string(l) + ".", q = "L" + std::to_string(l) + ".";
    const auto& w0 = TWt(e, p + "linears_scalar.0.weight", {C, nr});
    const auto& w1 = TWt(e, p + "linears_scalar.1.weight", {2 * C, C});
    const auto& w2 = TWt(e, p + "linears_scalar.2.weight", {3 * C, 2 * C});
    put(q + "W0t", fwd(w0, C, nr, nrp)), put(q + "W0r", pad_cols(w0, C, nr, nrp));
    put(q + "W1t", transpose(w1, 2 * C, C)), put(q + "W1r", w1);
    put(q + "W2t", transpose(w2, 3 * C, 2 * C)), put(q + "W2r", w2);
    {
      const std::vector<float> w0p = pad_cols(w0, C, nr, nrp);  // [C][64]
      put(q + "W0c", canon_split(w0p, C, nrp, nrp));
      put(q + "W0rc", canon_split(transpose(w0p, C, nrp), nrp, C, C));
      put(q + "W1c", canon_split(w1, 2 * C, C, C));
      put(q + "W1rc", canon_split(transpose(w1, 2 * C, C), C, 2 * C, 2 * C));
      for (int j = 0; j < 3; j++) {
        const std::vector<float> blk(w2.begin() + (size_t)j * C * 2 * C, w2.begin() + (size_t)(j + 1) * C * 2 * C);  // [C][2C]
        put(q + "W2c" + std::to_string(j), canon_split(blk, C, 2 * C, 2 * C));
        put(q + "W2rc" + std::to_string(j), canon_split(transpose(blk, C, 2 * C), 2 * C, C, C));
      }
    }
    put(q + "b0", TWt(e, p + "linears_scalar.0.bias", {C}));
    put(q + "b1", TWt(e, p + "linears_scalar.1.bias", {2 * C}));
    put(q + "b2", TWt(e, p + "linears_scalar.2.bias", {3 * C}));
    for (int k = 0; k < 6; k++) {
      const auto& w = TWt(e, p + "linears_tensor." + std::to_string(k) + ".weight", {C, C});
      put(q + "Wt_t" + std::to_string(k), transpose(w, C, C));
      put(q + "Wt_r" + std::to_string(k), w);
    }
  }
  put("lnr_g", TWt(e, "out_norm.weight", {3 * C})), put("lnr_b", TWt(e, "out_norm.bias", {3 * C}));
  {
    const auto& wl = TWt(e, "linear.weight", {C, 3 * C});
    put("Wl_t", transpose(wl, C, 3 * C)), put("Wl_r", wl), put("bl", TWt(e, "linear.bias", {C}));
  }
  // final_layer.gated.{layers,gates}.{i}: the Linear modules of the two chains (activations occupy the odd indices)
  std::vector<int> idx;
  for (auto& kv : e->host_w) {
    const std::string pre = "final_layer.gated.layers.";
    if (kv.first.rfind(pre, 0) == 0 && kv.first.size() > 7 && kv.first.substr(kv.first.size() - 7) == ".weight")
      idx.push_back(atoi(kv.first.c_str() + pre.size()));
  }
  std::sort(idx.begin(), idx.end());
  B2M_REQUIRE(idx.size() >= 2, B2M_ERR_INVALID, "final_layer.gated needs at least one hidden layer");
  const char* brn[2] = {"layers", "gates"};
  std::vector<std::pair<int, int>> dims;  // (in, out) of the hidden layers
  for (int br = 0; br < 2; br++) {
    int in = C;
    for (size_t j = 0; j < idx.size(); j++) {
      const std::string k = std::string("final_layer.gated.") + brn[br] + "." + std::to_string(idx[j]);
      auto it = e->host_shape.find(k + ".weight");
      B2M_REQUIRE(it != e->host_shape.end() && it->second.size() == 2 && it->second[1] == in, B2M_ERR_INVALID,
                  "unexpected readout layer " + k);
      const int out = (int)it->second[0];
      const auto& w = TWt(e, k + ".weight", {out, in});
      const auto& b = TWt(e, k + ".bias", {out});
      const std::string q = std::string("R") + brn[br] + std::to_string(j) + ".";
      if (j + 1 < idx.size()) {
        B2M_REQUIRE(out % 64 == 0 && in % 32 == 0, B2M_ERR_INVALID, "readout hidden widths must be multiples of 64");
        put(q + "Wt", transpose(w, out, in)), put(q + "Wr", w), put(q + "b", b);
        if (br == 0) dims.push_back({in, out});
      } else {
        B2M_REQUIRE(out == 1, B2M_ERR_INVALID, "ntargets must be 1");
        put(q + "w", w);
        t.blast[br] = b[0];
        t.wlast_in = in;
      }
      in = out;
    }
  }
  for (auto& kv : e->host_w)
    if (!e->consumed.count(kv.first))
      throw Error(B2M_ERR_INVALID, "state_dict tensor '" + kv.first +
                                       "' is not used by this engine (unsupported TensorNet configuration; refusing to ignore it)");
  if (!e->elem_refs.empty()) {
    e->erefbuf.ensure(e->elem_refs.size());
    B2M_CK(cudaMemcpyAsync(e->erefbuf.p, e->elem_refs.data(), e->elem_refs.size() * sizeof(double), cudaMemcpyHostToDevice,
                           e->st));
  }
  e->wbuf.ensure(P.host.size() + 64);
  B2M_CK(cudaMemcpyAsync(e->wbuf.p, P.host.data(), P.host.size() * sizeof(float), cudaMemcpyHostToDevice, e->st));
  B2M_CK(cudaStreamSynchronize(e->st));
  auto dp = [&](const std::string& n) { return (const float*)(e->wbuf.p + off.at(n)); };
  t.Wd_t = dp("Wd_t"), t.Wd_r = dp("Wd_r"), t.bd = dp("bd"), t.U = dp("U"), t.V = dp("V");
  t.Wdc_a = dp("Wdc_a"), t.Wdc_b = dp("Wdc_b");
  for (int j = 0; j < 3; j++) t.Wdrc[j] = dp("Wdrc" + std::to_string(j));
  {
    const char* ff = getenv("B2M_TN_FFMA");
    t.tc = !(ff && ff[0] == '1');
  }
  for (int k = 0; k < 3; k++) t.Wte_t[k] = dp("Wte_t" + std::to_string(k)), t.Wte_r[k] = dp("Wte_r" + std::to_string(k));
  t.ln0_g = dp("ln0_g"), t.ln0_b = dp("ln0_b");
  t.Ws0_t = dp("Ws0_t"), t.Ws0_r = dp("Ws0_r"), t.bs0 = dp("bs0");
  t.Ws1_t = dp("Ws1_t"), t.Ws1_r = dp("Ws1_r"), t.bs1 = dp("bs1");
  t.L.resize(nb);
  for (int l = 0; l < nb; l++) {
    const std::string q = "L" + std::to_string(l) + ".";
    TnLayerW& w = t.L[l];
    w.W0t = dp(q + "W0t"), w.W0r = dp(q + "W0r"), w.W1t = dp(q + "W1t"), w.W1r = dp(q + "W1r");
    w.W2t = dp(q + "W2t"), w.W2r = dp(q + "W2r"), w.b0 = dp(q + "b0"), w.b1 = dp(q + "b1"), w.b2 = dp(q + "b2");
    for (int k = 0; k < 6; k++)
      w.Wt_t[k] = dp(q + "Wt_t" + std::to_string(k)), w.Wt_r[k] = dp(q + "Wt_r" + std::to_string(k));
    w.W0c = dp(q + "W0c"), w.W0rc = dp(q + "W0rc"), w.W1c = dp(q + "W1c"), w.W1rc = dp(q + "W1rc");
    for (int j = 0; j < 3; j++) w.W2c[j] = dp(q + "W2c" + std::to_string(j)), w.W2rc[j] = dp(q + "W2rc" + std::to_string(j));
  }
  t.lnr_g = dp("lnr_g"), t.lnr_b = dp("lnr_b"), t.Wl_t = dp("Wl_t"), t.Wl_r = dp("Wl_r"), t.bl = dp("bl");
  for (int br = 0; br < 2; br++) {
    t.chain[br].clear();
    for (size_t j = 0; j + 1 < idx.size(); j++) {
      const std::string q = std::string("R") + brn[br] + std::to_string(j) + ".";
      t.chain[br].push_back(TnChainW{dp(q + "Wt"), dp(q + "Wr"), dp(q + "b"), dims[j].first, dims[j].second});
    }
    t.wlast[br] = dp(std::string("R") + brn[br] + std::to_string(idx.size() - 1) + ".w");
  }
  e->d_eref = e->elem_refs.empty() ? nullptr : e->erefbuf.p;
  e->finalized = true;
}

static void tn_alloc_workspace(b2m_engine* e) {
  TnState& t = *e->tn;
  Graph& g = e->g;
  const size_t nl = g.n_loc, no = g.n_own, E = (size_t)g.E, C = TNC, nrp = t.rp.nrp;
  const int nb = t.nblocks;
  t.rbf.ensure(E * nrp + 64), t.cut.ensure(E + 64), t.P.ensure(E * 3 * C + 64);
  t.T0.ensure(no * TNW + 64), t.nr0.ensure(no * C + 64), t.ln0.ensure(no * C + 64), t.st0.ensure(no * 2 + 64);
  t.s1p.ensure(no * 2 * C + 64), t.s1.ensure(no * 2 * C + 64), t.s2p.ensure(no * 3 * C + 64), t.T0m.ensure(no * TNW + 64);
  t.f1.ensure(E * C + 64), t.f2.ensure(E * 2 * C + 64);
  t.X.resize(nb + 1);
  for (auto& b : t.X) b.ensure(nl * TNW + 64);
  for (auto* v : {&t.f1p, &t.f2p, &t.f3p, &t.q, &t.Xh, &t.Y, &t.msg, &t.Pn, &t.dX}) v->resize(nb);
  for (int l = 0; l < nb; l++) {
    t.f1p[l].ensure(E * C + 64), t.f2p[l].ensure(E * 2 * C + 64), t.f3p[l].ensure(E * 3 * C + 64);
    t.q[l].ensure(nl * C + 64), t.Xh[l].ensure(nl * TNW + 64), t.Y[l].ensure(nl * TNW + 64);
    t.msg[l].ensure(no * TNW + 64), t.Pn[l].ensure(no * TNW + 64), t.dX[l].ensure(no * TNW + 64);
  }
  t.inv.ensure(no * 3 * C + 64), t.str.ensure(no * 2 + 64), t.r.ensure(no * 3 * C + 64), t.xr.ensure(no * C + 64);
  t.lout.ensure(no + 64), t.gout.ensure(no + 64), t.e_atom.ensure(no + 64);
  size_t wmax = 64;
  for (int br = 0; br < 2; br++) {
    t.cpre[br].resize(t.chain[br].size()), t.cact[br].resize(t.chain[br].size());
    for (size_t j = 0; j < t.chain[br].size(); j++) {
      t.cpre[br][j].ensure(no * t.chain[br][j].out + 64), t.cact[br][j].ensure(no * t.chain[br][j].out + 64);
      wmax = std::max<size_t>(wmax, t.chain[br][j].out);
    }
  }
  {
    t.gX.ensure(nl * TNW + 64), t.gY.ensure(nl * TNW + 64), t.gmsg.ensure(no * TNW + 64), t.gdX.ensure(no * TNW + 64);
    t.gPn.ensure(no * TNW + 64), t.gf.ensure(E * 3 * C + 64);  // the reverse edge MLP reuses f2 / f1 (forward scratch) for its adjoints
    t.g_rbf.ensure(E * nrp + 64), t.gC.ensure(E + 64), t.gvh.ensure(E * 3 + 64), t.gd.ensure(E + 64);
    t.gT0m.ensure(no * TNW + 64), t.gT0.ensure(no * TNW + 64), t.gs2p.ensure(no * 3 * C + 64);
    t.gs1p.ensure(no * 2 * C + 64), t.gln0.ensure(no * C + 64), t.gnr0.ensure(no * C + 64);
    t.gr.ensure(no * std::max<size_t>(3 * C, wmax) + 64), t.ginv.ensure(no * 3 * C + 64), t.gxr.ensure(no * C + 64);
    t.gca.ensure(no * wmax + 64), t.gcb.ensure(no * wmax + 64);
  }
  e->forces.ensure((size_t)g.N * 3 + 64);
  e->scal.ensure(16);
  if (e->world > 1) {
    size_t tot_to = 0;
    for (int q = 0; q < e->world; q++) tot_to += g.n_to[q];
    const size_t m = tot_to * TNW;
    if (e->leader != nullptr) {
      e->precv[0].ensure(m + 64), e->precv[1].ensure(m + 64);
    } else {
      e->sendbuf.ensure(m + 64), e->recvbuf.ensure(m + 64);
    }
  }
}

// plain / SiLU / reverse row GEMM on the engine's stream
static void tn_gemm(b2m_engine* e, const float* A, int lda, const float* B, float* C, int ldc, int M, int N, int K,
                    const float* bias, int epi = 0, float* Cpre = nullptr, const float* Pre = nullptr, int ldp = 0,
                    bool accum = false) {
  TnGemm g;
  g.A = A, g.lda = lda, g.B0 = B, g.C = C, g.ldc = ldc, g.M = M, g.N = N, g.K = K, g.bias = bias, g.epi = epi;
  g.Cpre = Cpre, g.Pre = Pre, g.ldp = ldp, g.accum = accum ? 1 : 0;
  launch_tn_gemm(e->st, g, 1);
}
// channel mix of a decomposed tensor: out[:, k, :] (+)= in[:, k, :] @ B[part(k)]
static void tn_mix(b2m_engine* e, const float* in, float* out, int rows, const float* B0, const float* B1,
                   const float* B2, bool accum = false) {
  TnGemm g;
  g.A = in, g.lda = TNW, g.zA = TNC, g.B0 = B0, g.B1 = B1, g.B2 = B2, g.bsel = 1;
  g.C = out, g.ldc = TNW, g.zC = TNC, g.M = rows, g.N = TNC, g.K = TNC, g.accum = accum ? 1 : 0;
  launch_tn_gemm(e->st, g, 10);
}

static void tn_edge_mlp(b2m_engine* e, int l) {
  TnState& t = *e->tn;
  const TnLayerW& w = t.L[l];
  const int E = (int)e->g.E, C = TNC;
  if (t.tc) {  // tcgen05 3xTF32 row GEMM (kernels_tc.cu) with the SiLU epilogue; the 128 -> 192 layer as three column blocks
    launch_gemm_tc_epi(e->st, t.rbf.p, t.rp.nrp, w.W0c, t.f1.p, C, E, C, C, w.b0, false, 1, t.f1p[l].p, nullptr, 0, e->num_sms);
    launch_gemm_tc_epi(e->st, t.f1.p, C, w.W1c, t.f2.p, 2 * C, E, 2 * C, C, w.b1, false, 1, t.f2p[l].p, nullptr, 0, e->num_sms);
    tn_stage(e, "tn_edge_mlp:launch_gemm_tc_epi");
    for (int j = 0; j < 3; j++)
      launch_gemm_tc_epi(e->st, t.f2.p, 2 * C, w.W2c[j], t.f3p[l].p + j * C, 3 * C, E, C, 2 * C, w.b2 + j * C, false, 0,
                         nullptr, nullptr, 0, e->num_sms);
    return;
  }
  tn_gemm(e, t.rbf.p, t.rp.nrp, w.W0t, t.f1.p, C, E, C, t.rp.nrp, w.b0, 1, t.f1p[l].p);
  tn_stage(e, "tn_edge_mlp:tn_gemm");
  tn_gemm(e, t.f1.p, C, w.W1t, t.f2.p, 2 * C, E, 2 * C, C, w.b1, 1, t.f2p[l].p);
  tn_stage(e, "tn_edge_mlp:tn_gemm");
  tn_gemm(e, t.f2.p, 2 * C, w.W2t, t.f3p[l].p, 3 * C, E, 3 * C, 2 * C, w.b2);
  tn_stage(e, "tn_edge_mlp:tn_gemm");
}
// reverse of the edge MLP: g3 [E,3C] (adjoint of the last pre-activation) -> g_rbf += ...
static void tn_edge_mlp_bwd(b2m_engine* e, int l) {
  TnState& t = *e->tn;
  const TnLayerW& w = t.L[l];
  const int E = (int)e->g.E, C = TNC;
  if (t.tc) {  // the 192 -> 128 product as three K-chunks accumulated in place; the last one carries the SiLU' factor
    for (int j = 0; j < 3; j++)
      launch_gemm_tc_epi(e->st, t.gf.p + j * C, 3 * C, w.W2rc[j], t.f2.p, 2 * C, E, 2 * C, C, nullptr, j > 0, j == 2 ? 2 : 0,
                         nullptr, j == 2 ? t.f2p[l].p : nullptr, 2 * C, e->num_sms);
    launch_gemm_tc_epi(e->st, t.f2.p, 2 * C, w.W1rc, t.f1.p, C, E, C, 2 * C, nullptr, false, 2, nullptr, t.f1p[l].p, C, e->num_sms);
    tn_stage(e, "tn_edge_mlp_bwd:launch_gemm_tc_epi");
    launch_gemm_tc_epi(e->st, t.f1.p, C, w.W0rc, t.g_rbf.p, t.rp.nrp, E, t.rp.nrp, C, nullptr, true, 0, nullptr, nullptr, 0,
                       e->num_sms);
    return;
  }
  tn_gemm(e, t.gf.p, 3 * C, w.W2r, t.f2.p, 2 * C, E, 2 * C, 3 * C, nullptr, 2, nullptr, t.f2p[l].p, 2 * C);
  tn_stage(e, "tn_edge_mlp_bwd:tn_gemm");
  tn_gemm(e, t.f2.p, 2 * C, w.W1r, t.f1.p, C, E, C, 2 * C, nullptr, 2, nullptr, t.f1p[l].p, C);
  tn_stage(e, "tn_edge_mlp_bwd:tn_gemm");
  tn_gemm(e, t.f1.p, C, w.W0r, t.g_rbf.p, t.rp.nrp, E, t.rp.nrp, C, nullptr, 0, nullptr, nullptr, 0, true);
  tn_stage(e, "tn_edge_mlp_bwd:tn_gemm");
}

static void tn_forward(b2m_engine* e) {
  TnState& t = *e->tn;
  Graph& g = e->g;
  const int C = TNC, nb = t.nblocks, no = g.n_own, nl = g.n_loc;
  B2M_REQUIRE(g.E < (1LL << 31) / 8, B2M_ERR_INVALID, "too many edges for one TensorNet partition");
  const int E = (int)g.E;
  B2M_CK(cudaMemsetAsync(e->scal.p, 0, 16 * sizeof(double), e->st));
  launch_tn_edge_geom(e->st, g.E, g.e_vec.p, t.rp, t.rbf.p, t.cut.p);
  tn_stage(e, "tn_forward:launch_tn_edge_geom");
  // ---- embedding (tensor_embedding_dist, tensornet.py:104-112) ----
  if (t.tc) {
    launch_gemm_tc_epi(e->st, t.rbf.p, t.rp.nrp, t.Wdc_a, t.P.p, 3 * C, E, 2 * C, C, t.bd, false, 0, nullptr, nullptr, 0, e->num_sms);
    launch_gemm_tc_epi(e->st, t.rbf.p, t.rp.nrp, t.Wdc_b, t.P.p + 2 * C, 3 * C, E, C, C, t.bd + 2 * C, false, 0, nullptr, nullptr, 0,
                       e->num_sms);
  } else {
    tn_gemm(e, t.rbf.p, t.rp.nrp, t.Wd_t, t.P.p, 3 * C, E, 3 * C, t.rp.nrp, t.bd);
    tn_stage(e, "tn_forward:tn_gemm");
  }
  launch_tn_embed_agg(e->st, no, g.row_ptr.p, g.e_src.p, g.type.p, t.U, t.V, t.P.p, t.cut.p, g.e_vec.p, t.T0.p, t.nr0.p);
  tn_stage(e, "tn_forward:launch_tn_embed_agg");
  launch_tn_layernorm(e->st, no, C, t.nr0.p, t.ln0_g, t.ln0_b, t.ln0.p, t.st0.p);
  tn_stage(e, "tn_forward:launch_tn_layernorm");
  tn_gemm(e, t.ln0.p, C, t.Ws0_t, t.s1.p, 2 * C, no, 2 * C, C, t.bs0, 1, t.s1p.p);
  tn_stage(e, "tn_forward:tn_gemm");
  tn_gemm(e, t.s1.p, 2 * C, t.Ws1_t, t.s2p.p, 3 * C, no, 3 * C, 2 * C, t.bs1);
  tn_stage(e, "tn_forward:tn_gemm");
  tn_mix(e, t.T0.p, t.T0m.p, no, t.Wte_t[0], t.Wte_t[1], t.Wte_t[2]);
  tn_stage(e, "tn_forward:tn_mix");
  launch_tn_embed_out(e->st, no, t.T0m.p, t.s2p.p, t.X[0].p);
  tn_stage(e, "tn_forward:launch_tn_embed_out");
  // ---- interaction layers (dist_forward, tensornet.py:119-127) ----
  for (int l = 0; l < nb; l++) {
    const TnLayerW& w = t.L[l];
    halo_forward_begin(e, 2, l);  // halo rows of X[l] travel while the edge MLP of this layer runs
    tn_edge_mlp(e, l);
    tn_stage(e, "tn_forward:tn_edge_mlp");
    halo_forward_end(e);
    tn_stage(e, "tn_forward:halo_forward_end");
    launch_tn_scale(e->st, nl, t.X[l].p, t.Xh[l].p, t.q[l].p);
    tn_stage(e, "tn_forward:launch_tn_scale");
    tn_mix(e, t.Xh[l].p, t.Y[l].p, nl, w.Wt_t[0], w.Wt_t[1], w.Wt_t[2]);
    tn_stage(e, "tn_forward:tn_mix");
    launch_tn_msg(e->st, no, g.row_ptr.p, g.e_src.p, t.f3p[l].p, t.cut.p, t.Y[l].p, t.msg[l].p);
    tn_stage(e, "tn_forward:launch_tn_msg");
    launch_tn_prod(e->st, no, t.msg[l].p, t.Y[l].p, t.so3, t.Pn[l].p);
    tn_stage(e, "tn_forward:launch_tn_prod");
    tn_mix(e, t.Pn[l].p, t.dX[l].p, no, w.Wt_t[3], w.Wt_t[4], w.Wt_t[5]);
    tn_stage(e, "tn_forward:tn_mix");
    launch_tn_update(e->st, no, t.Xh[l].p, t.dX[l].p, t.X[l + 1].p);
    tn_stage(e, "tn_forward:launch_tn_update");
  }
  // ---- readout (tensornet.py:129-147; the transfer after the last layer feeds nothing and is skipped) ----
  launch_tn_invariants(e->st, no, t.X[nb].p, t.inv.p);
  tn_stage(e, "tn_forward:launch_tn_invariants");
  launch_tn_layernorm(e->st, no, 3 * C, t.inv.p, t.lnr_g, t.lnr_b, t.r.p, t.str.p);
  tn_stage(e, "tn_forward:launch_tn_layernorm");
  tn_gemm(e, t.r.p, 3 * C, t.Wl_t, t.xr.p, C, no, C, 3 * C, t.bl);
  tn_stage(e, "tn_forward:tn_gemm");
  const float* hlast[2];
  for (int br = 0; br < 2; br++) {
    const float* h = t.xr.p;
    for (size_t j = 0; j < t.chain[br].size(); j++) {
      const TnChainW& cw = t.chain[br][j];
      tn_gemm(e, h, cw.in, cw.Wt, t.cact[br][j].p, cw.out, no, cw.out, cw.in, cw.b, 1, t.cpre[br][j].p);
      tn_stage(e, "tn_forward:tn_gemm");
      h = t.cact[br][j].p;
    }
    hlast[br] = h;
  }
  launch_tn_readout_final(e->st, no, t.wlast_in, hlast[0], t.wlast[0], t.blast[0], hlast[1], t.wlast[1], t.blast[1],
                          g.type.p, e->d_eref, (float)e->desc.data_std, t.lout.p, t.gout.p, t.e_atom.p, e->scal.p);
}

static void tn_backward(b2m_engine* e) {
  TnState& t = *e->tn;
  Graph& g = e->g;
  const int C = TNC, nb = t.nblocks, no = g.n_own, nl = g.n_loc, E = (int)g.E;
  launch_zero_rows(e->st, t.gC.p, g.E);
  tn_stage(e, "tn_backward:launch_zero_rows");
  launch_zero_rows(e->st, t.g_rbf.p, g.E * t.rp.nrp);
  tn_stage(e, "tn_backward:launch_zero_rows");
  launch_zero_rows(e->st, t.gX.p, (int64_t)nl * TNW);
  tn_stage(e, "tn_backward:launch_zero_rows");
  launch_zero_rows(e->st, e->forces.p, g.N * 3);
  tn_stage(e, "tn_backward:launch_zero_rows");
  // ---- readout ----
  const int nh = (int)t.chain[0].size();
  launch_tn_readout_seed(e->st, no, t.wlast_in, t.lout.p, t.gout.p, (float)e->desc.data_std, t.wlast[0], t.wlast[1],
                         t.cpre[0][nh - 1].p, t.cpre[1][nh - 1].p, t.gca.p, t.gcb.p);
  for (int br = 0; br < 2; br++) {
    float* cur = br == 0 ? t.gca.p : t.gcb.p;  // adjoint of the pre-activation of hidden layer j, in place along the chain
    for (int j = nh - 1; j >= 1; j--) {
      const TnChainW& cw = t.chain[br][j];
      // (g_pre_j @ W_j) * SiLU'(pre_{j-1}); the widths of a chain are equal in matgl's readout, so in place is not
      // possible (the GEMM reads its input rows while writing): ping-pong through gr (free until the chains are done)
      tn_gemm(e, cur, cw.out, cw.Wr, t.gr.p, cw.in, no, cw.in, cw.out, nullptr, 2, nullptr, t.cpre[br][j - 1].p, cw.in);
      tn_stage(e, "tn_backward:tn_gemm");
      B2M_CK(cudaMemcpyAsync(cur, t.gr.p, (size_t)no * cw.in * sizeof(float), cudaMemcpyDeviceToDevice, e->st));
    }
    const TnChainW& c0 = t.chain[br][0];
    tn_gemm(e, cur, c0.out, c0.Wr, t.gxr.p, c0.in, no, c0.in, c0.out, nullptr, 0, nullptr, nullptr, 0, br == 1);
    tn_stage(e, "tn_backward:tn_gemm");
  }
  tn_gemm(e, t.gxr.p, C, t.Wl_r, t.gr.p, 3 * C, no, 3 * C, C, nullptr);
  tn_stage(e, "tn_backward:tn_gemm");
  launch_tn_layernorm_bwd(e->st, no, 3 * C, t.inv.p, t.str.p, t.lnr_g, t.gr.p, t.ginv.p);
  tn_stage(e, "tn_backward:launch_tn_layernorm_bwd");
  launch_tn_invariants_bwd(e->st, no, t.X[nb].p, t.ginv.p, t.gX.p);
  tn_stage(e, "tn_backward:launch_tn_invariants_bwd");
  // ---- interaction layers ----
  for (int l = nb - 1; l >= 0; l--) {
    const TnLayerW& w = t.L[l];
    launch_tn_update_bwd(e->st, no, t.dX[l].p, t.gX.p, t.gdX.p);  // gX stays: it is also the adjoint of Xh (residual)
    tn_mix(e, t.gdX.p, t.gPn.p, no, w.Wt_r[3], w.Wt_r[4], w.Wt_r[5]);
    tn_stage(e, "tn_backward:tn_mix");
    launch_zero_rows(e->st, t.gY.p + (size_t)no * TNW, (int64_t)(nl - no) * TNW);
    tn_stage(e, "tn_backward:launch_zero_rows");
    launch_tn_prod_bwd(e->st, no, t.msg[l].p, t.Y[l].p, t.so3, t.gPn.p, t.gmsg.p, t.gY.p);
    tn_stage(e, "tn_backward:launch_tn_prod_bwd");
    launch_tn_msg_bwd(e->st, no, g.row_ptr.p, g.e_src.p, t.f3p[l].p, t.cut.p, t.Y[l].p, t.gmsg.p, t.gf.p, t.gC.p, t.gY.p);
    tn_stage(e, "tn_backward:launch_tn_msg_bwd");
    tn_edge_mlp_bwd(e, l);
    tn_stage(e, "tn_backward:tn_edge_mlp_bwd");
    tn_mix(e, t.gY.p, t.gX.p, nl, w.Wt_r[0], w.Wt_r[1], w.Wt_r[2], true);
    tn_stage(e, "tn_backward:tn_mix");
    launch_tn_scale_bwd(e->st, nl, t.X[l].p, t.q[l].p, t.gX.p);
    tn_stage(e, "tn_backward:launch_tn_scale_bwd");
    halo_backward(e, t.gX.p, false, TNW);
    tn_stage(e, "tn_backward:halo_backward");
  }
  // ---- embedding ----
  launch_tn_embed_out_bwd(e->st, no, t.T0m.p, t.s2p.p, t.gX.p, t.gT0m.p, t.gs2p.p);
  tn_stage(e, "tn_backward:launch_tn_embed_out_bwd");
  tn_gemm(e, t.gs2p.p, 3 * C, t.Ws1_r, t.gs1p.p, 2 * C, no, 2 * C, 3 * C, nullptr, 2, nullptr, t.s1p.p, 2 * C);
  tn_stage(e, "tn_backward:tn_gemm");
  tn_gemm(e, t.gs1p.p, 2 * C, t.Ws0_r, t.gln0.p, C, no, C, 2 * C, nullptr);
  tn_stage(e, "tn_backward:tn_gemm");
  launch_tn_layernorm_bwd(e->st, no, C, t.nr0.p, t.st0.p, t.ln0_g, t.gln0.p, t.gnr0.p);
  tn_stage(e, "tn_backward:launch_tn_layernorm_bwd");
  tn_mix(e, t.gT0m.p, t.gT0.p, no, t.Wte_r[0], t.Wte_r[1], t.Wte_r[2]);
  tn_stage(e, "tn_backward:tn_mix");
  launch_tn_norm_bwd_add(e->st, no, t.T0.p, t.gnr0.p, t.gT0.p);
  tn_stage(e, "tn_backward:launch_tn_norm_bwd_add");
  launch_tn_embed_agg_bwd(e->st, g.E, g.e_src.p, g.e_dst.p, g.type.p, t.U, t.V, t.P.p, t.cut.p, g.e_vec.p, t.gT0.p,
                          t.gf.p, t.gC.p, t.gvh.p);
  if (t.tc) {
    for (int j = 0; j < 3; j++)
      launch_gemm_tc_epi(e->st, t.gf.p + j * C, 3 * C, t.Wdrc[j], t.g_rbf.p, t.rp.nrp, E, t.rp.nrp, C, nullptr, true, 0, nullptr,
                         nullptr, 0, e->num_sms);
  } else {
    tn_gemm(e, t.gf.p, 3 * C, t.Wd_r, t.g_rbf.p, t.rp.nrp, E, t.rp.nrp, 3 * C, nullptr, 0, nullptr, nullptr, 0, true);
    tn_stage(e, "tn_backward:tn_gemm");
  }
  launch_tn_edge_final(e->st, g.E, g.e_src.p, g.e_dst.p, g.e_vec.p, g.gid.p, t.rp, t.g_rbf.p, t.gC.p, t.gvh.p, t.gd.p,
                       e->forces.p, e->scal.p + 1);
}

// name -> (pointer, rows, cols) for b2m_debug_tensor
static bool tn_debug_lookup(b2m_engine* e, const std::string& n, const float*& src, int64_t& r, int64_t& c) {
  TnState& t = *e->tn;
  Graph& g = e->g;
  const int nb = t.nblocks;
  auto lay = [&](const std::string& pre, int& l) {
    if (n.rfind(pre, 0) != 0 || n.size() <= pre.size() || !isdigit(n[pre.size()])) return false;
    for (size_t i = pre.size(); i < n.size(); i++)
      if (!isdigit(n[i])) return false;
    l = atoi(n.c_str() + pre.size());
    return true;
  };
  int l = 0;
  const int64_t no = g.n_own, nl = g.n_loc, E = g.E;
  if (n == "rbf") src = t.rbf.p, r = E, c = t.rp.nrp;
  else if (n == "cut") src = t.cut.p, r = E, c = 1;
  else if (n == "P") src = t.P.p, r = E, c = 3 * TNC;
  else if (n == "T0") src = t.T0.p, r = no, c = TNW;
  else if (n == "ln0") src = t.ln0.p, r = no, c = TNC;
  else if (n == "s2p") src = t.s2p.p, r = no, c = 3 * TNC;
  else if (n == "T0m") src = t.T0m.p, r = no, c = TNW;
  else if (lay("X", l) && l <= nb) src = t.X[l].p, r = l == nb ? no : nl, c = TNW;
  else if (lay("f3p", l) && l < nb) src = t.f3p[l].p, r = E, c = 3 * TNC;
  else if (lay("Xh", l) && l < nb) src = t.Xh[l].p, r = nl, c = TNW;
  else if (lay("Y", l) && l < nb) src = t.Y[l].p, r = nl, c = TNW;
  else if (lay("msg", l) && l < nb) src = t.msg[l].p, r = no, c = TNW;
  else if (lay("Pn", l) && l < nb) src = t.Pn[l].p, r = no, c = TNW;
  else if (lay("dX", l) && l < nb) src = t.dX[l].p, r = no, c = TNW;
  else if (n == "inv") src = t.inv.p, r = no, c = 3 * TNC;
  else if (n == "xr") src = t.xr.p, r = no, c = TNC;
  else if (n == "e_atom") src = t.e_atom.p, r = no, c = 1;
  // reverse pass: the buffers hold what the last stage that used them wrote (layer 0 / the embedding)
  else if (n == "gX0") src = t.gX.p, r = nl, c = TNW;
  else if (n == "gY0") src = t.gY.p, r = nl, c = TNW;
  else if (n == "gmsg0") src = t.gmsg.p, r = no, c = TNW;
  else if (n == "gdX0") src = t.gdX.p, r = no, c = TNW;
  else if (n == "gT0") src = t.gT0.p, r = no, c = TNW;
  else if (n == "gP") src = t.gf.p, r = E, c = 3 * TNC;
  else if (n == "g_rbf") src = t.g_rbf.p, r = E, c = t.rp.nrp;
  else if (n == "gC") src = t.gC.p, r = E, c = 1;
  else if (n == "gvh") src = t.gvh.p, r = E, c = 3;
  else if (n == "gd") src = t.gd.p, r = E, c = 1;
  else if (n == "e_vec") src = reinterpret_cast<const float*>(g.e_vec.p), r = E, c = 4;
  else return false;
  return true;
}

static void tn_release(b2m_engine* e) {
  if (!e->tn) return;
  TnState& t = *e->tn;
  auto drop = [](DBuf<float>& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr, b.cap = 0;
  };
  for (auto* b : {&t.rbf, &t.cut, &t.P, &t.T0, &t.nr0, &t.ln0, &t.st0, &t.s1p, &t.s1, &t.s2p, &t.T0m, &t.f1, &t.f2, &t.inv,
                  &t.str, &t.r, &t.xr, &t.lout, &t.gout, &t.e_atom, &t.gX, &t.gY, &t.gmsg, &t.gdX, &t.gPn, &t.gf,
                  &t.g_rbf, &t.gC, &t.gvh, &t.gd, &t.gT0m, &t.gT0, &t.gs2p, &t.gs1p, &t.gln0, &t.gnr0, &t.gr,
                  &t.ginv, &t.gxr, &t.gca, &t.gcb})
    drop(*b);
  for (auto* v : {&t.X, &t.f1p, &t.f2p, &t.f3p, &t.q, &t.Xh, &t.Y, &t.msg, &t.Pn, &t.dX, &t.cpre[0], &t.cpre[1],
                  &t.cact[0], &t.cact[1]})
    for (auto& b : *v) drop(b);
}
