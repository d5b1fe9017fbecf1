// tn_state.cuh -- weights and workspace of the TensorNet path (engine_tn.inl drives kernels_tn.cu with it).
#pragma once
#include <vector>

#include "kernels.cuh"

namespace b2m {

struct TnLayerW {
  const float *W0t, *b0, *W1t, *b1, *W2t, *b2;  // edge MLP num_rbf -> C -> 2C -> 3C, forward operands [K][N]
  const float *W0r, *W1r, *W2r;                 // raw nn.Linear weights [out][in] (reverse pass; W0r padded to nrp columns)
  const float *Wt_t[6], *Wt_r[6];               // linears_tensor 0..5: transposed (forward) and raw (reverse)
  // tcgen05 operands of the edge MLP (canonical hi/lo planes, see engine.cu canon_split): forward W0 (64->64, rbf columns
  // padded), W1 (64->128), W2 as three 128->64 column blocks; reverse W2 as three 64->128 K-chunks, W1 (128->64), W0 (64->64)
  const float *W0c, *W1c, *W2c[3], *W2rc[3], *W1rc, *W0rc;
};
struct TnChainW {  // one hidden Linear of a readout chain
  const float *Wt, *Wr, *b;
  int in, out;
};

struct TnState {
  int units = 64, num_rbf = 32, nblocks = 2, so3 = 0;
  TnRadial rp;
  // ---- weights (device pointers into the engine's weight buffer) ----
  const float *Wd_t = nullptr, *bd = nullptr, *Wd_r = nullptr;  // three distance projections stacked: [nrp][3C], [3C], [3C][nrp]
  const float *Wdc_a = nullptr, *Wdc_b = nullptr, *Wdrc[3] = {nullptr, nullptr, nullptr};  // tcgen05: 64->128 | 64->64 ; reverse chunks
  bool tc = true;  // edge-level GEMMs on the tcgen05 row GEMM (B2M_TN_FFMA=1: the FP32-FFMA tiles, A/B checks)
  const float *U = nullptr, *V = nullptr;                       // emb2 halves applied to the embedding table: [n_elem][C]
  const float *Wte_t[3] = {nullptr, nullptr, nullptr}, *Wte_r[3] = {nullptr, nullptr, nullptr};
  const float *ln0_g = nullptr, *ln0_b = nullptr;
  const float *Ws0_t = nullptr, *bs0 = nullptr, *Ws0_r = nullptr, *Ws1_t = nullptr, *bs1 = nullptr, *Ws1_r = nullptr;
  std::vector<TnLayerW> L;
  const float *lnr_g = nullptr, *lnr_b = nullptr, *Wl_t = nullptr, *bl = nullptr, *Wl_r = nullptr;
  std::vector<TnChainW> chain[2];  // hidden layers of final_layer.gated.{layers, gates}
  const float* wlast[2] = {nullptr, nullptr};
  float blast[2] = {0.f, 0.f};
  int wlast_in = 64;
  // ---- workspace ----
  DBuf<float> rbf, cut, P, T0, nr0, ln0, st0, s1p, s1, s2p, T0m;
  DBuf<float> f1, f2;  // [E][C], [E][2C]: activations of the edge MLP (forward), adjoints of its hidden layers (reverse)
  std::vector<DBuf<float>> X;                                       // nblocks + 1 : [n_loc][10][C]
  std::vector<DBuf<float>> f1p, f2p, f3p, q, Xh, Y, msg, Pn, dX;    // per layer
  DBuf<float> inv, str, r, xr, lout, gout, e_atom;
  std::vector<DBuf<float>> cpre[2], cact[2];
  // reverse pass
  DBuf<float> gX, gY, gmsg, gdX, gPn, gf, g_rbf, gC, gvh, gd, gT0m, gT0, gs2p, gs1p, gln0, gnr0, gr, ginv, gxr,
      gca, gcb;
};

}  // namespace b2m
