// Plan for the mlp-grad conditional AR-DAE (reference: models/graddae/mlp.py:341-483).
//   train : loss = mean((sigma*g + eps)^2), g = -dE/dx~, plus every parameter gradient of that
//           double-backprop loss, computed with the primal / score / tangent / adjoint sweeps of
//           SURVEY.md 8a-3 as a chain of tcgen05 GEMMs with fused epilogues.
//   score : glogprob only (sweeps 1-2).
// Precision: the PRIMAL forward runs "3xTF32" (operands split into tf32 hi/lo pairs, fp32-accurate
// products).  With std_scale = 1e4 the inputs reach |x| ~ 1e4..1e5, pre-activations ~1e4, and the
// loss gradient lives on the units within O(10) of their softplus kink: a plain tf32 forward
// (relative 5e-4 -> absolute +-10) makes those sigmoids garbage (measured: 47 % gradient error on
// the reference-generated fixture).  The score / tangent / adjoint / weight-gradient sweeps are
// linear in their operands and run plain tf32 (measured gradient error ~1e-3).
// Parameter tensors arrive as per-tensor device pointers in state_dict order:
//   ctx_encode (W,b) x L ; inp_encode (W,b) x L ; neglogprob (W,b) x (L+1).
#pragma once
#include <cstring>
#include <memory>

#include "kernels.cuh"
#include "plan.cuh"

namespace ardae {

struct CdaeConfig {
  int d = 0, c = 0, H = 0, L = 0;  // input_dim, context_dim, h_dim, num_hidden_layers
  int B = 0, S = 0;                // data rows, samples per data row (N = B*S)
  int train = 1;                   // 0: score-only plan
  int kind = 0;                    // 0: mlp-grad (energy network, score by back-prop; models/graddae/mlp.py)
                                   // 1: mlp-res  (network outputs the score; models/resdae/mlp.py:286-413)
};

struct CdaeBindings {  // per-call user pointers (plain device memory, no TMA)
  const float* x = nullptr;      // [N, d]
  const float* ctx = nullptr;    // [B, c]
  const float* sigma = nullptr;  // [N]
  float* eps = nullptr;          // [N, d]  (input, or output when gen_eps)
  int gen_eps = 0;
  uint64_t seed = 0;
  float inv_count = 0.0f;        // 1 / (N_global * d)
  float* loss_out = nullptr;     // device scalar
  float* score_out = nullptr;    // [N, d] or null
};

// Collects DeriveItems and emits the single weight-derivation launch.
struct DeriveList {
  std::vector<DeriveItem> host;
  int blocks = 0;
  // W[rows(out), cols(in)] at src (pitch src_ld).  want3: forward 3xTF32 operand; wantT: transpose.
  W3 add(Workspace& ws, const float* src, int rows, int cols, int src_ld, bool want3, bool wantT, bool want16 = false) {
    W3 w;
    w.in = cols; w.out = rows; w.kp = round_up(cols, 32);
    DeriveItem it;
    std::memset(&it, 0, sizeof(it));
    it.src = src; it.rows = rows; it.cols = cols; it.src_ld = src_ld;
    if (want3) {
      w.b3 = Mat(ws.floats(static_cast<size_t>(rows) * 3 * w.kp), rows, 3 * w.kp, 3 * w.kp);
      it.dst3 = w.b3.p; it.kp = w.kp; it.ld3 = 3 * w.kp;
    }
    if (wantT) {
      w.T = ws.mat(cols, rows);
      it.dstT = w.T.p; it.ldT = w.T.ld;
    }
    if (want16) {
      w.k16 = round_up(cols, 64);
      w.h16 = ws.mat16(rows, 2 * w.k16);
      it.dst16 = w.h16.p; it.k16 = w.k16;
    }
    it.first_block = blocks;
    it.tiles_x = (cols + 31) / 32;
    blocks += it.tiles_x * ((rows + 31) / 32);
    host.push_back(it);
    return w;
  }
  int emit(Workspace& ws, Plan& plan) {
    float* table = ws.floats((host.size() * sizeof(DeriveItem) + 3) / 4 + 1);
    if (ws.dry) {
      plan.add(nullptr);
      return 0;
    }
    DeriveItem* dev = reinterpret_cast<DeriveItem*>(table);
    ARDAE_CUDA_OK(cudaMemcpy(dev, host.data(), host.size() * sizeof(DeriveItem), cudaMemcpyHostToDevice));
    const int nitems = static_cast<int>(host.size());
    const int nb = blocks;
    plan.add([dev, nitems, nb](cudaStream_t s) {
      derive_weights_kernel<<<nb, 256, 0, s>>>(dev, nitems);
      return static_cast<int>(cudaGetLastError());
    });
    return 0;
  }
};

struct CdaePlan {
  CdaeConfig cfg;
  Plan plan;
  Workspace ws;
  CdaeBindings bind;  // ops read this at launch time
  DeriveList derive;
  size_t tn_need = 0;  // split-K partial workspace, measured by the dry build

  int ntensors() const { return 6 * cfg.L + 2; }

  // 16-bit spill plan (chain16_sm100.cuh + gemm_tn16.cuh): mlp-grad plans (training update, or score only: sweeps 1-2)
  // whose H -> H layers fit the fused chain kernel.  ARDAE_SPILL16=0 keeps the fp32-spill plan (A/B measurements, parity triage).
  bool spill16_plan() const {
    static int env = -1;
    if (env < 0) {
      const char* e = std::getenv("ARDAE_SPILL16");
      env = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    const int kp = round_up(cfg.d, 32);
    return env != 0 && cfg.kind == 0 && chain_supported(cfg.H, 1) && 2 * cfg.L <= kChainMaxLayers &&
           kp <= cfg.H / 2 && cfg.H % 64 == 0;
  }

  // The training update of the mlp-grad CDAE as FOUR chain launches (primal 3xTF32 / score / tangent / adjoint, each
  // carrying its 128-row tile through all 2L layers incl. the d-wide ends) + 2L bf16 weight-gradient contractions.
  // Every [N, H] array that crosses HBM is bfloat16; [N, d] arrays (x~, g, r, eps) stay fp32.
  int build16(float* const* params, float* const* grads) {
    const int d = cfg.d, c = cfg.c, H = cfg.H, L = cfg.L, B = cfg.B, S = cfg.S;
    const int N = B * S;
    const bool dry = ws.dry;
    const bool train = cfg.train != 0;
    auto P = [&](int i) -> float* { return dry ? nullptr : params[i]; };
    auto G = [&](int i) -> float* { return (dry || !train) ? nullptr : grads[i]; };
    auto iC = [&](int l) { return 2 * l; };
    auto iA = [&](int l) { return 2 * L + 2 * l; };
    auto iW = [&](int l) { return 4 * L + 2 * l; };
    const int kp = round_up(d, 32), ny = round_up(d, 64);

    derive = DeriveList();
    std::vector<W3> Cw(L), Aw(L), Ww(L);
    for (int l = 0; l < L; ++l) Cw[l] = derive.add(ws, P(iC(l)), H, l == 0 ? c : H, l == 0 ? c : H, true, true);
    const bool s3h = s3h_supported(H);  // primal sweep on the fp16 pipe
    for (int l = 0; l < L; ++l) Aw[l] = derive.add(ws, P(iA(l)), H, l == 0 ? d : H, l == 0 ? d : H, true, true, s3h);
    for (int l = 1; l < L; ++l) Ww[l] = derive.add(ws, P(iW(l)), H, H, H, true, true, s3h);
    const int ld1 = 2 * H + 1;
    W3 W1u = derive.add(ws, P(iW(0)), H, H, ld1, true, true, s3h);
    W3 W1c = derive.add(ws, P(iW(0)) ? P(iW(0)) + H : nullptr, H, H, ld1, true, true);
    Ww[0] = W1u;
    int rc = derive.emit(ws, plan);
    if (rc) return rc;
    float* wsig = ws.floats(H);
    const float* wo = P(iW(L));

    // ---- activations
    Pair xt = make_pair(ws, N, d), ctxp = make_pair(ws, B, c);
    Mat rowbias = ws.mat(B, H);
    Mat gmat(ws.floats(static_cast<size_t>(N) * kp), N, d, kp), rmat(ws.floats(static_cast<size_t>(N) * kp), N, d, kp);
    Mat16 x16h = ws.mat16(N, ny), x16l = ws.mat16(N, ny), r16h = ws.mat16(N, ny), r16l = ws.mat16(N, ny);
    float* sig = ws.floats(N);
    std::vector<Pair> Cc(L);
    for (int l = 0; l < L; ++l) Cc[l] = make_pair(ws, B, H);
    std::vector<Mat16> U(L), V(L), DA(L), DP(L), UD(L), VD(L), TA(L), TP(L);
    for (auto* arr : {&U, &V, &DA, &DP})
      for (int l = 0; l < L; ++l) (*arr)[l] = ws.mat16(N, H);
    if (train)
      for (auto* arr : {&UD, &VD, &TA, &TP})
        for (int l = 0; l < L; ++l) (*arr)[l] = ws.mat16(N, H);
    Mat gsum = ws.mat(B, H);
    std::vector<Mat> DC(L);
    for (int l = 0; l < L; ++l) DC[l] = ws.mat(B, H);
    if (dry) {
      tn_need = tn16_workspace_bytes(H, H, N);
      const size_t b2 = tn16_workspace_bytes(H, d, N);
      if (b2 > tn_need) tn_need = b2;
    }
    const size_t tn_ws_bytes = tn_need;
    float* tn_ws = ws.floats(tn_ws_bytes / 4);
    size_t ctx_tn_bytes = tn_workspace_bytes(H, H, B);
    {
      const size_t b2 = tn_workspace_bytes(H, c, B);
      if (b2 > ctx_tn_bytes) ctx_tn_bytes = b2;
    }
    float* ctx_tn_ws = ws.floats(ctx_tn_bytes / 4);
    CdaeBindings* bd = &bind;

    // ---- prologue: sigma copy, x~ = x + sigma*eps (tf32 pair + bf16 pair), context pair, exact w_sigma
    {
      const float* w1 = P(iW(0));
      plan.add([=](cudaStream_t s) {
        ARDAE_CUDA_OK(cudaMemcpyAsync(sig, bd->sigma, sizeof(float) * N, cudaMemcpyDeviceToDevice, s));
        if (bd->loss_out) ARDAE_CUDA_OK(cudaMemsetAsync(bd->loss_out, 0, sizeof(float), s));
        split2d_kernel<<<grid_for(static_cast<size_t>(B) * c), 256, 0, s>>>(
            bd->ctx, c, ctxp.buf.p, ctxp.buf.ld, B, c, ctxp.kp, 1.0f, 0.0f);
        copy2d_kernel<<<1, 256, 0, s>>>(w1 + 2 * H, ld1, wsig, 1, H, 1, 1, 1.0f, 0.0f, 0);
        return static_cast<int>(cudaGetLastError());
      });
    }
    // ---- context branch on the B distinct rows (side lane, underneath the perturbation prologue)
    plan.fork();
    for (int l = 0; l < L; ++l) {
      GemmNTDesc g = nt3_desc(l == 0 ? ctxp : Cc[l - 1], Cw[l], Cc[l], EPI_SOFTPLUS);
      g.bias = P(iC(l) + 1);
      plan.nt(g);
    }
    {
      GemmNTDesc g = nt3_desc_plain(Cc[L - 1], W1c, rowbias, EPI_LINEAR);
      g.bias = P(iW(0) + 1);
      plan.nt(g);
    }
    plan.cur_lane = 0;
    plan.add([=](cudaStream_t s) {
      cdae_perturb_kernel<<<grid_for(static_cast<size_t>(N) * d), 256, 0, s>>>(
          bd->x, sig, bd->eps, xt.buf.p, N, d, xt.buf.ld, xt.kp, bd->gen_eps, bd->seed, replay_counter(), x16h.p, x16l.p,
          x16h.ld);
      return static_cast<int>(cudaGetLastError());
    });
    plan.join();
    auto base = [&](int mode) {
      Chain16Desc cd;
      cd.mode = mode; cd.M = N; cd.H = H; cd.row_scale = sig;
      return cd;
    };
    bool delta_fused = false;
    // ---- sweep 1: primal forward (3xTF32), all 2L layers
    {
      Chain16Desc cd = base(CHAIN_SOFTPLUS3);
      cd.A0_32 = xt.hi().p; cd.lda0_32 = xt.buf.ld; cd.A0lo = xt.lo().p; cd.lda0lo = xt.buf.ld;
      auto s3 = [&](const W3& w, const float* bias, const Mat16& out, int kin) {
        Chain16LayerDesc q;
        q.W = w.b3.p; q.ldw = w.b3.ld; q.kin = kin; q.bias = bias; q.out = out.p; q.ldo = out.ld;
        q.W16 = w.h16.p; q.ldw16 = w.h16.ld; q.kin16 = w.k16;
        return q;
      };
      cd.layers.push_back(s3(Aw[0], P(iA(0) + 1), U[0], kp));
      for (int l = 1; l < L; ++l) cd.layers.push_back(s3(Aw[l], P(iA(l) + 1), U[l], H));
      {
        Chain16LayerDesc q = s3(W1u, nullptr, V[0], H);
        q.group_bias = rowbias.p; q.group = S; q.ldg = rowbias.ld; q.col_vec = wsig;
        cd.layers.push_back(q);
      }
      for (int l = 1; l < L; ++l) cd.layers.push_back(s3(Ww[l], P(iW(l) + 1), V[l], H));
      // the fp16-pipe kernel also emits delta_L from its last epilogue (wo 16-byte aligned: arena tensors are)
      const bool fuse_delta = s3h && (reinterpret_cast<uintptr_t>(wo) & 15) == 0;
      if (fuse_delta) {
        cd.wo = wo; cd.delta16 = DP[L - 1].p; cd.ld_delta16 = DP[L - 1].ld;
      }
      delta_fused = fuse_delta;
      if (s3h) plan.chain_s3h(cd); else plan.chain16(cd);
    }
    // ---- sweep 2: score backward, ending in g = delta a_1 . A_1 (fp32 [N, kp])
    if (!delta_fused) {
      const Mat16 vl = V[L - 1], dpl = DP[L - 1];
      plan.add([=](cudaStream_t s) {
        cdae_init_delta16_kernel<<<grid_for(static_cast<size_t>(N) * H / 8), 256, 0, s>>>(vl.p, vl.ld, wo, dpl.p, dpl.ld, N, H);
        return static_cast<int>(cudaGetLastError());
      });
    }
    {
      Chain16Desc cd = base(CHAIN_MUL_SIG);
      cd.A0_16 = DP[L - 1].p; cd.lda0_16 = DP[L - 1].ld;
      auto bw = [&](const Mat& wT, const Mat16& aux1, const Mat16& out) {
        Chain16LayerDesc q;
        q.W = wT.p; q.ldw = wT.ld; q.aux1 = aux1.p; q.ld1 = aux1.ld; q.out = out.p; q.ldo = out.ld;
        return q;
      };
      for (int l = L - 1; l >= 1; --l) cd.layers.push_back(bw(Ww[l].T, V[l - 1], DP[l - 1]));
      cd.layers.push_back(bw(W1u.T, U[L - 1], DA[L - 1]));
      for (int l = L - 1; l >= 1; --l) cd.layers.push_back(bw(Aw[l].T, U[l - 1], DA[l - 1]));
      {
        Chain16LayerDesc q;
        q.W = Aw[0].T.p; q.ldw = Aw[0].T.ld; q.nout = kp; q.w_rows = d; q.out32 = gmat.p; q.ld_out32 = gmat.ld;
        cd.layers.push_back(q);
      }
      plan.chain16(cd);
    }
    if (!train) {  // glogprob: sweeps 1-2 only
      plan.add([=](cudaStream_t s) {
        unpad_kernel<<<grid_for(static_cast<size_t>(N) * d), 256, 0, s>>>(gmat.p, gmat.ld, bd->score_out, N, d, 1.0f);
        return static_cast<int>(cudaGetLastError());
      });
      return plan.error;
    }
    // ---- loss + residual direction r (fp32 + bf16 pair)
    plan.add([=](cudaStream_t s) {
      cdae_loss_kernel<<<grid_for(static_cast<size_t>(N) * gmat.ld), 256, 0, s>>>(
          gmat.p, gmat.ld, sig, bd->eps, rmat.p, N, d, bd->inv_count, bd->loss_out, bd->score_out, r16h.p, r16l.p, r16h.ld);
      return static_cast<int>(cudaGetLastError());
    });
    // ---- sweep 3: tangent forward from r (also emits t = delta * tangent_pre * (1 - sig))
    {
      Chain16Desc cd = base(CHAIN_TANGENT);
      cd.A0_32 = rmat.p; cd.lda0_32 = rmat.ld;
      auto tg = [&](const Mat& w, int kin, const Mat16& aux1, const Mat16& aux2, const Mat16& out, const Mat16& out2) {
        Chain16LayerDesc q;
        q.W = w.p; q.ldw = w.ld; q.kin = kin; q.aux1 = aux1.p; q.ld1 = aux1.ld; q.aux2 = aux2.p; q.ld2 = aux2.ld;
        q.out = out.p; q.ldo = out.ld; q.out2 = out2.p; q.ldo2 = out2.ld;
        return q;
      };
      cd.layers.push_back(tg(Aw[0].hi(), kp, U[0], DA[0], UD[0], TA[0]));
      for (int l = 1; l < L; ++l) cd.layers.push_back(tg(Aw[l].hi(), H, U[l], DA[l], UD[l], TA[l]));
      for (int l = 0; l < L; ++l) cd.layers.push_back(tg(Ww[l].hi(), H, V[l], DP[l], VD[l], TP[l]));
      cd.layers.back().colsum = G(iW(L)); cd.layers.back().colsum_scale = -1.0f;  // d w_o = -sum_n vdot_L
      cd.layers.back().colsum2 = G(iW(L - 1) + 1);                                // d beta_L = sum_n adj p_L (= t_L)
      plan.chain16(cd);
    }
    // ---- sweep 4: adjoint backward (in place over the t buffers)
    {
      Chain16Desc cd = base(CHAIN_ADJOINT);
      cd.A0_16 = TP[L - 1].p; cd.lda0_16 = TP[L - 1].ld;
      auto adj = [&](const Mat& wT, const Mat16& aux1, const Mat16& t, float* colsum) {
        Chain16LayerDesc q;
        q.W = wT.p; q.ldw = wT.ld; q.aux1 = aux1.p; q.ld1 = aux1.ld; q.aux2 = t.p; q.ld2 = t.ld;
        q.out = t.p; q.ldo = t.ld; q.colsum = colsum;
        return q;
      };
      for (int l = L - 1; l >= 1; --l) {
        Chain16LayerDesc q = adj(Ww[l].T, V[l - 1], TP[l - 1], G(iW(l - 1) + 1));
        if (l - 1 == 0) {  // d w_1sigma = sum_n sigma_n * adj p_1
          q.colsum_w = G(iW(0)) ? G(iW(0)) + 2 * H : nullptr;
          q.colsum_w_stride = ld1;
        }
        cd.layers.push_back(q);
      }
      cd.layers.push_back(adj(W1u.T, U[L - 1], TA[L - 1], G(iA(L - 1) + 1)));
      for (int l = L - 1; l >= 1; --l) cd.layers.push_back(adj(Aw[l].T, U[l - 1], TA[l - 1], G(iA(l - 1) + 1)));
      plan.chain16(cd);
    }
    // ---- context branch backward (B rows, fp32 / tf32) on the side lane underneath the weight-gradient contractions
    plan.fork();
    {
      const Mat16 ap1 = TP[0];
      plan.add([=](cudaStream_t s) {
        group_sum16_kernel<<<B, 256, 0, s>>>(ap1.p, ap1.ld, gsum.p, gsum.ld, B, S, H, 1);
        return static_cast<int>(cudaGetLastError());
      });
    }
    auto ctx_tn = [&](const Mat& X0, const Mat& Y0, float* dst, int ldo) {
      GemmTNDesc t;
      t.X0 = X0.p; t.ldx0 = X0.ld; t.Y0 = Y0.p; t.ldy0 = Y0.ld;
      t.M = X0.cols; t.N = Y0.cols; t.K = X0.rows; t.out = dst; t.ldo = ldo;
      t.scale = 1.0f; t.beta = 1.0f;
      t.workspace = ctx_tn_ws; t.workspace_bytes = ctx_tn_bytes;
      plan.tn(t);
    };
    ctx_tn(gsum, Cc[L - 1].hi(), G(iW(0)) ? G(iW(0)) + H : nullptr, ld1);
    {
      GemmNTDesc g = nt_desc(gsum, W1c.T, DC[L - 1], EPI_MUL_SIG);
      set_aux1(g, Cc[L - 1].hi());
      g.colsum = G(iC(L - 1) + 1);
      plan.nt(g);
    }
    for (int l = L - 1; l >= 1; --l) {
      GemmNTDesc g = nt_desc(DC[l], Cw[l].T, DC[l - 1], EPI_MUL_SIG);
      set_aux1(g, Cc[l - 1].hi());
      g.colsum = G(iC(l - 1) + 1);
      plan.nt(g);
    }
    for (int l = L - 1; l >= 1; --l) ctx_tn(DC[l], Cc[l - 1].hi(), G(iC(l)), H);
    ctx_tn(DC[0], ctxp.hi(), G(iC(0)), c);
    plan.cur_lane = 0;
    // ---- weight gradients: dW = adj^T . act + delta^T . tangent   (accumulated into .grad)
    auto tn16_desc = [&](std::initializer_list<std::pair<Mat16, Mat16>> pairs, int Mo, int No, int Ny, float* dst, int ldo) {
      GemmTN16Desc t;
      for (const auto& pr : pairs) {
        t.X[t.npairs] = pr.first.p; t.ldx[t.npairs] = pr.first.ld;
        t.Y[t.npairs] = pr.second.p; t.ldy[t.npairs] = pr.second.ld;
        ++t.npairs;
      }
      t.M = Mo; t.N = No; t.Ny = Ny; t.K = N; t.out = dst; t.ldo = ldo;
      t.workspace = tn_ws; t.workspace_bytes = tn_ws_bytes;
      return t;
    };
    plan.tn16(tn16_desc({{TP[0], U[L - 1]}, {DP[0], UD[L - 1]}}, H, H, H, G(iW(0)), ld1));
    {  // the 2(L-1) [H,H] layers: one launch walking them in order
      std::vector<GemmTN16Desc> batch;
      for (int l = L - 1; l >= 1; --l) batch.push_back(tn16_desc({{TP[l], V[l - 1]}, {DP[l], VD[l - 1]}}, H, H, H, G(iW(l)), H));
      for (int l = L - 1; l >= 1; --l) batch.push_back(tn16_desc({{TA[l], U[l - 1]}, {DA[l], UD[l - 1]}}, H, H, H, G(iA(l)), H));
      plan.tn16_multi(batch);
    }
    plan.tn16(tn16_desc({{TA[0], x16h}, {TA[0], x16l}, {DA[0], r16h}, {DA[0], r16l}}, H, d, ny, G(iA(0)), d));
    plan.join();
    return plan.error;
  }

  // returns 0 or error; when ws.dry only measures
  int build(float* const* params, float* const* grads) {
    const int d = cfg.d, c = cfg.c, H = cfg.H, L = cfg.L, B = cfg.B, S = cfg.S;
    const int N = B * S;
    if (d <= 0 || c <= 0 || H <= 0 || L < 2 || B <= 0 || S <= 0)
      return fail(-2, "cdae: bad config (need d,c,H,B,S > 0 and num_hidden_layers >= 2)");
    if (H % 4 != 0) return fail(-2, "cdae: h_dim must be a multiple of 4");
    plan.dry = ws.dry;
    if (spill16_plan()) return build16(params, grads);
    const bool dry = ws.dry;
    const bool train = cfg.train != 0;
    auto P = [&](int i) -> float* { return dry ? nullptr : params[i]; };
    auto G = [&](int i) -> float* { return (dry || !train) ? nullptr : grads[i]; };
    auto iC = [&](int l) { return 2 * l; };          // ctx_encode weight l (bias +1)
    auto iA = [&](int l) { return 2 * L + 2 * l; };  // inp_encode
    auto iW = [&](int l) { return 4 * L + 2 * l; };  // neglogprob (l == L: fc)

    // ---- derived weights
    derive = DeriveList();
    std::vector<W3> Cw(L), Aw(L), Ww(L);
    for (int l = 0; l < L; ++l) Cw[l] = derive.add(ws, P(iC(l)), H, l == 0 ? c : H, l == 0 ? c : H, true, train);
    for (int l = 0; l < L; ++l) Aw[l] = derive.add(ws, P(iA(l)), H, l == 0 ? d : H, l == 0 ? d : H, true, true);
    for (int l = 1; l < L; ++l) Ww[l] = derive.add(ws, P(iW(l)), H, H, H, true, true);
    const int ld1 = 2 * H + 1;
    W3 W1u = derive.add(ws, P(iW(0)), H, H, ld1, true, true);
    W3 W1c = derive.add(ws, P(iW(0)) ? P(iW(0)) + H : nullptr, H, H, ld1, true, train);
    Ww[0] = W1u;
    const bool res = cfg.kind == 1;
    if (cfg.kind != 0 && cfg.kind != 1) return fail(-2, "cdae: kind must be 0 (mlp-grad) or 1 (mlp-res)");
    W3 Wo;  // mlp-res: dae.fc [d, H]
    if (res) Wo = derive.add(ws, P(iW(L)), d, H, H, true, train);
    int rc = derive.emit(ws, plan);
    if (rc) return rc;
    float* wsig = ws.floats(H);  // exact fp32 copy of W_1[:, 2H]
    const float* wo = P(iW(L));

    // ---- activations
    Pair xt = make_pair(ws, N, d), ctxp = make_pair(ws, B, c);
    Mat gmat = ws.mat(N, d), rowbias = ws.mat(B, H);
    float* sig = ws.floats(N);
    std::vector<Pair> Cc(L), U(L), V(L);
    std::vector<Mat> DA(L), DP(L);
    for (int l = 0; l < L; ++l) Cc[l] = make_pair(ws, B, H);
    // Fused path: the H -> H layers of each sweep run as ONE launch per chain (chain_sm100.cuh) with the
    // activation operand resident on chip; only the d -> H first layer / H -> d last layer stay per-layer GEMMs.
    const bool use_chain = chain_supported(H, 1) && 2 * L - 1 <= kChainMaxLayers;
    // primal forward as ONE chain when the N-row first layer is long enough to cover the context branch on the side
    // lane; small plans (score on B rows) keep two chains so the context branch overlaps the first of them
    const bool merge_s3 = N >= 16384;
    for (int l = 0; l < L; ++l) {
      // chain plans: dense hi arrays; a lo part only where the chain starts from the (hi, lo) pair (U[0])
      if (use_chain) {
        U[l] = make_pair_split(ws, N, H, l == 0 || (!merge_s3 && l == L - 1));
        V[l] = (cfg.kind == 1 && l == L - 1) ? make_pair(ws, N, H) : make_pair_split(ws, N, H, false);
      } else {
        U[l] = make_pair(ws, N, H);
        V[l] = make_pair(ws, N, H);
      }
    }
    for (int l = 0; l < L; ++l) DA[l] = ws.mat(N, H);
    for (int l = 0; l < L; ++l) DP[l] = ws.mat(N, H);
    std::vector<Mat> UD(L), VD(L), TA(L), TP(L), DC(L);
    Mat rmat, gsum;
    float* tn_ws = nullptr;
    size_t tn_ws_bytes = 0;
    float* ctx_tn_ws = nullptr;
    size_t ctx_tn_bytes = 0;
    if (train) {
      rmat = ws.mat(N, d);
      gsum = ws.mat(B, H);
      if (!res) {
        for (int l = 0; l < L; ++l) UD[l] = ws.mat(N, H);
        for (int l = 0; l < L; ++l) VD[l] = ws.mat(N, H);
        for (int l = 0; l < L; ++l) TA[l] = ws.mat(N, H);
        for (int l = 0; l < L; ++l) TP[l] = ws.mat(N, H);
      }
      for (int l = 0; l < L; ++l) DC[l] = ws.mat(B, H);
      if (dry) {
        const int shapes[5][3] = {{H, H, N}, {H, d, N}, {H, H, B}, {H, c, B}, {d, H, N}};
        tn_need = 0;
        for (auto& sh : shapes) {
          const size_t b = tn_workspace_bytes(sh[0], sh[1], sh[2]);
          if (b > tn_need) tn_need = b;
        }
      }
      tn_ws_bytes = tn_need;
      tn_ws = ws.floats(tn_ws_bytes / 4);
      ctx_tn_bytes = tn_workspace_bytes(H, H, B);
      const size_t b2 = tn_workspace_bytes(H, c, B);
      if (b2 > ctx_tn_bytes) ctx_tn_bytes = b2;
      ctx_tn_ws = ws.floats(ctx_tn_bytes / 4);
    }
    CdaeBindings* bd = &bind;
    auto tn_desc = [&](const Mat& X0, const Mat& Y0, const Mat* X1, const Mat* Y1, float* dst, int ldo) {
      GemmTNDesc t;
      t.X0 = X0.p; t.ldx0 = X0.ld; t.Y0 = Y0.p; t.ldy0 = Y0.ld;
      if (X1) { t.X1 = X1->p; t.ldx1 = X1->ld; t.Y1 = Y1->p; t.ldy1 = Y1->ld; }
      t.M = X0.cols; t.N = Y0.cols; t.K = X0.rows; t.out = dst; t.ldo = ldo;
      t.scale = 1.0f; t.beta = 1.0f;
      // the side lane (context branch) has its own split-K scratch: both lanes run concurrently
      t.workspace = plan.cur_lane == 1 ? ctx_tn_ws : tn_ws;
      t.workspace_bytes = plan.cur_lane == 1 ? ctx_tn_bytes : tn_ws_bytes;
      return t;
    };
    auto tn2 = [&](const Mat& X0, const Mat& Y0, const Mat* X1, const Mat* Y1, float* dst, int ldo) {
      plan.tn(tn_desc(X0, Y0, X1, Y1, dst, ldo));
    };

    // ---- prologue: sigma copy, x~ = x + sigma*eps (tf32 pair), context pair, exact w_sigma
    {
      const float* w1 = P(iW(0));
      plan.add([=](cudaStream_t s) {
        ARDAE_CUDA_OK(cudaMemcpyAsync(sig, bd->sigma, sizeof(float) * N, cudaMemcpyDeviceToDevice, s));
        if (bd->loss_out) ARDAE_CUDA_OK(cudaMemsetAsync(bd->loss_out, 0, sizeof(float), s));
        cdae_perturb_kernel<<<grid_for(static_cast<size_t>(N) * d), 256, 0, s>>>(
            bd->x, sig, bd->eps, xt.buf.p, N, d, xt.buf.ld, xt.kp, bd->gen_eps, bd->seed, replay_counter());
        split2d_kernel<<<grid_for(static_cast<size_t>(B) * c), 256, 0, s>>>(
            bd->ctx, c, ctxp.buf.p, ctxp.buf.ld, B, c, ctxp.kp, 1.0f, 0.0f);
        copy2d_kernel<<<1, 256, 0, s>>>(w1 + 2 * H, ld1, wsig, 1, H, 1, 1, 1.0f, 0.0f, 0);
        return static_cast<int>(cudaGetLastError());
      });
    }
    // ---- context branch on the B distinct rows (reference runs it on all N: graddae/mlp.py:415,426);
    //      independent of the N-row input branch until layer p_1: side lane
    plan.fork();
    for (int l = 0; l < L; ++l) {
      GemmNTDesc g = nt3_desc(l == 0 ? ctxp : Cc[l - 1], Cw[l], Cc[l], EPI_SOFTPLUS);
      g.bias = P(iC(l) + 1);
      plan.nt(g);
    }
    {
      GemmNTDesc g = nt3_desc_plain(Cc[L - 1], W1c, rowbias, EPI_LINEAR);
      g.bias = P(iW(0) + 1);
      plan.nt(g);
    }
    plan.cur_lane = 0;  // (still forked: joined right before p_1 needs the per-row bias)
    // ---- sweep 1: primal forward (3xTF32)
    auto chain_of = [&](int mode, const Mat& a0) {
      ChainDesc cd;
      cd.mode = mode; cd.M = N; cd.H = H; cd.A0 = a0.p; cd.lda0 = a0.ld; cd.row_scale = sig;
      return cd;
    };
    auto s3_layer = [&](const W3& w, const float* bias, const Pair& out) {
      ChainLayerDesc q;
      q.W = w.b3.p; q.ldw = w.b3.ld; q.bias = bias; q.out = out.hi().p; q.ldo = out.hi().ld;
      return q;
    };
    if (use_chain) {
      {
        GemmNTDesc g = nt3_desc(xt, Aw[0], U[0], EPI_SOFTPLUS);
        g.bias = P(iA(0) + 1);
        plan.nt(g);
      }
      // ONE chain for the whole primal forward (inp_encode layers 1..L-1, then p_1 .. p_L) when merge_s3: the context
      // branch only has to be finished before the chain starts (it runs on the side lane underneath the first-layer
      // GEMM above), and nothing restarts from an (hi, lo) pair in the middle (no lo spill, no second
      // initial-activation load).  Otherwise an inp chain, the join, and a p chain restarting from U[L-1].
      if (!merge_s3 && L > 1) {
        ChainDesc cu = chain_of(CHAIN_SOFTPLUS3, U[0].hi());
        cu.A0lo = U[0].lo().p; cu.lda0lo = U[0].lo().ld;
        for (int l = 1; l < L; ++l) cu.layers.push_back(s3_layer(Aw[l], P(iA(l) + 1), U[l]));
        cu.layers.back().out_lo = U[L - 1].lo().p;
        cu.layers.back().ld_out_lo = U[L - 1].lo().ld;
        plan.chain(cu);
      }
      plan.join();
      const Pair& entry = merge_s3 ? U[0] : U[L - 1];
      ChainDesc cd = chain_of(CHAIN_SOFTPLUS3, entry.hi());
      cd.A0lo = entry.lo().p; cd.lda0lo = entry.lo().ld;
      if (merge_s3)
        for (int l = 1; l < L; ++l) cd.layers.push_back(s3_layer(Aw[l], P(iA(l) + 1), U[l]));
      {
        ChainLayerDesc q = s3_layer(W1u, nullptr, V[0]);
        q.group_bias = rowbias.p; q.group = S; q.ldg = rowbias.ld; q.col_vec = wsig;
        cd.layers.push_back(q);
      }
      for (int l = 1; l < L; ++l) cd.layers.push_back(s3_layer(Ww[l], P(iW(l) + 1), V[l]));
      if (res) {  // the output layer f = v_L Wo^T + b_o runs 3xTF32 on the (hi, lo) pair
        cd.layers.back().out_lo = V[L - 1].lo().p;
        cd.layers.back().ld_out_lo = V[L - 1].lo().ld;
      }
      plan.chain(cd);
    } else {
    for (int l = 0; l < L; ++l) {
      GemmNTDesc g = nt3_desc(l == 0 ? xt : U[l - 1], Aw[l], U[l], EPI_SOFTPLUS);
      g.bias = P(iA(l) + 1);
      plan.nt(g);
    }
    plan.join();
    {
      GemmNTDesc g = nt3_desc(U[L - 1], W1u, V[0], EPI_SOFTPLUS);
      g.group_bias = rowbias.p; g.group = S; g.ldg = rowbias.ld;
      g.row_scale = sig; g.col_vec = wsig;
      plan.nt(g);
    }
    for (int l = 1; l < L; ++l) {
      GemmNTDesc g = nt3_desc(V[l - 1], Ww[l], V[l], EPI_SOFTPLUS);
      g.bias = P(iW(l) + 1);
      plan.nt(g);
    }
    }
    if (res) {
      // =============================================================== mlp-res: plain forward / backward
      // f = dae.fc(v_L)  [N, d]  (the score estimate itself: resdae/mlp.py:382,411)
      {
        GemmNTDesc g = nt3_desc_plain(V[L - 1], Wo, gmat, EPI_LINEAR);
        g.bias = P(iW(L) + 1);
        plan.nt(g);
      }
      if (!train) {
        plan.add([=](cudaStream_t s) {
          unpad_kernel<<<grid_for(static_cast<size_t>(N) * d), 256, 0, s>>>(gmat.p, gmat.ld, bd->score_out, N, d, 1.0f);
          return static_cast<int>(cudaGetLastError());
        });
        return plan.error;
      }
      // loss = mean((sigma f + eps)^2) (:386) and r = d loss / d f
      {
        float* gbo = G(iW(L) + 1);
        plan.add([=](cudaStream_t s) {
          cdae_loss_kernel<<<grid_for(static_cast<size_t>(N) * gmat.ld), 256, 0, s>>>(
              gmat.p, gmat.ld, sig, bd->eps, rmat.p, N, d, bd->inv_count, bd->loss_out, bd->score_out);
          dim3 grid((d + 31) / 32, N >= 2048 ? 64 : (N + 63) / 64);
          colsum_kernel<<<grid, 256, 0, s>>>(rmat.p, rmat.ld, N, d, gbo, 1.0f);  // d b_o = sum_n r
          return static_cast<int>(cudaGetLastError());
        });
      }
      // delta p_L = (r Wo) * sig(p_L)
      {
        GemmNTDesc g = nt_desc(rmat, Wo.T, DP[L - 1], EPI_MUL_SIG);
        set_aux1(g, V[L - 1].hi());
        g.colsum = G(iW(L - 1) + 1);
        plan.nt(g);
      }
      auto bwl = [&](const Mat& wT, const Mat& aux1, const Mat& out, float* colsum) {
        ChainLayerDesc q;
        q.W = wT.p; q.ldw = wT.ld; q.aux1 = aux1.p; q.ld1 = aux1.ld; q.out = out.p; q.ldo = out.ld; q.colsum = colsum;
        return q;
      };
      if (use_chain) {
        ChainDesc cd = chain_of(CHAIN_MUL_SIG, DP[L - 1]);
        for (int l = L - 1; l >= 1; --l) {
          ChainLayerDesc q = bwl(Ww[l].T, V[l - 1].hi(), DP[l - 1], G(iW(l - 1) + 1));
          if (l - 1 == 0) {  // d w_1sigma = sum_n sigma_n * delta p_1
            q.colsum_w = G(iW(0)) ? G(iW(0)) + 2 * H : nullptr;
            q.colsum_w_stride = ld1;
          }
          cd.layers.push_back(q);
        }
        cd.layers.push_back(bwl(W1u.T, U[L - 1].hi(), DA[L - 1], G(iA(L - 1) + 1)));
        for (int l = L - 1; l >= 1; --l) cd.layers.push_back(bwl(Aw[l].T, U[l - 1].hi(), DA[l - 1], G(iA(l - 1) + 1)));
        plan.chain(cd);
      } else {
        for (int l = L - 1; l >= 1; --l) {
          GemmNTDesc g = nt_desc(DP[l], Ww[l].T, DP[l - 1], EPI_MUL_SIG);
          set_aux1(g, V[l - 1].hi());
          g.colsum = G(iW(l - 1) + 1);
          if (l - 1 == 0) {
            g.colsum_w = G(iW(0)) ? G(iW(0)) + 2 * H : nullptr;
            g.colsum_w_stride = ld1;
            g.row_w = sig;
          }
          plan.nt(g);
        }
        {
          GemmNTDesc g = nt_desc(DP[0], W1u.T, DA[L - 1], EPI_MUL_SIG);
          set_aux1(g, U[L - 1].hi());
          g.colsum = G(iA(L - 1) + 1);
          plan.nt(g);
        }
        for (int l = L - 1; l >= 1; --l) {
          GemmNTDesc g = nt_desc(DA[l], Aw[l].T, DA[l - 1], EPI_MUL_SIG);
          set_aux1(g, U[l - 1].hi());
          g.colsum = G(iA(l - 1) + 1);
          plan.nt(g);
        }
      }
      // context branch backward (B rows) on the side lane, from the per-data-row sums of delta p_1
      plan.fork();
      {
        const Mat dp1 = DP[0];
        plan.add([=](cudaStream_t s) {
          group_sum_kernel<<<B, 256, 0, s>>>(dp1.p, dp1.ld, gsum.p, gsum.ld, B, S, H, 1);
          return static_cast<int>(cudaGetLastError());
        });
      }
      {
        const Mat y = Cc[L - 1].hi();
        tn2(gsum, y, nullptr, nullptr, G(iW(0)) ? G(iW(0)) + H : nullptr, ld1);
      }
      {
        GemmNTDesc g = nt_desc(gsum, W1c.T, DC[L - 1], EPI_MUL_SIG);
        set_aux1(g, Cc[L - 1].hi());
        g.colsum = G(iC(L - 1) + 1);
        plan.nt(g);
      }
      for (int l = L - 1; l >= 1; --l) {
        GemmNTDesc g = nt_desc(DC[l], Cw[l].T, DC[l - 1], EPI_MUL_SIG);
        set_aux1(g, Cc[l - 1].hi());
        g.colsum = G(iC(l - 1) + 1);
        plan.nt(g);
      }
      for (int l = L - 1; l >= 1; --l) {
        const Mat y = Cc[l - 1].hi();
        tn2(DC[l], y, nullptr, nullptr, G(iC(l)), H);
      }
      {
        const Mat y = ctxp.hi();
        tn2(DC[0], y, nullptr, nullptr, G(iC(0)), c);
      }
      plan.cur_lane = 0;
      // weight gradients: dW = delta^T . activation
      {
        const Mat y = V[L - 1].hi();
        tn2(rmat, y, nullptr, nullptr, G(iW(L)), H);
      }
      {
        const Mat y = U[L - 1].hi();
        tn2(DP[0], y, nullptr, nullptr, G(iW(0)), ld1);
      }
      {  // the 2(L-1) [H, H] contractions as one launch
        std::vector<GemmTNDesc> batch;
        for (int l = L - 1; l >= 1; --l) batch.push_back(tn_desc(DP[l], V[l - 1].hi(), nullptr, nullptr, G(iW(l)), H));
        for (int l = L - 1; l >= 1; --l) batch.push_back(tn_desc(DA[l], U[l - 1].hi(), nullptr, nullptr, G(iA(l)), H));
        plan.tn_batch(batch);
      }
      {
        const Mat xh = xt.hi();
        tn2(DA[0], xh, nullptr, nullptr, G(iA(0)), d);
      }
      plan.join();
      return plan.error;
    }
    // ---- sweep 2: score backward (tf32)
    {
      const Mat vl = V[L - 1].hi(), dpl = DP[L - 1];
      plan.add([=](cudaStream_t s) {
        cdae_init_delta_kernel<<<grid_for(static_cast<size_t>(N) * H / 4), 256, 0, s>>>(vl.p, vl.ld, wo, dpl.p, dpl.ld, N, H);
        return static_cast<int>(cudaGetLastError());
      });
    }
    auto bw_layer = [&](const Mat& wT, const Mat& aux1, const Mat& out) {
      ChainLayerDesc q;
      q.W = wT.p; q.ldw = wT.ld; q.aux1 = aux1.p; q.ld1 = aux1.ld; q.out = out.p; q.ldo = out.ld;
      return q;
    };
    if (use_chain) {
      ChainDesc cd = chain_of(CHAIN_MUL_SIG, DP[L - 1]);
      for (int l = L - 1; l >= 1; --l) cd.layers.push_back(bw_layer(Ww[l].T, V[l - 1].hi(), DP[l - 1]));
      cd.layers.push_back(bw_layer(W1u.T, U[L - 1].hi(), DA[L - 1]));
      for (int l = L - 1; l >= 1; --l) cd.layers.push_back(bw_layer(Aw[l].T, U[l - 1].hi(), DA[l - 1]));
      plan.chain(cd);
    } else {
    for (int l = L - 1; l >= 1; --l) {
      GemmNTDesc g = nt_desc(DP[l], Ww[l].T, DP[l - 1], EPI_MUL_SIG);
      set_aux1(g, V[l - 1].hi());
      plan.nt(g);
    }
    {
      GemmNTDesc g = nt_desc(DP[0], W1u.T, DA[L - 1], EPI_MUL_SIG);
      set_aux1(g, U[L - 1].hi());
      plan.nt(g);
    }
    for (int l = L - 1; l >= 1; --l) {
      GemmNTDesc g = nt_desc(DA[l], Aw[l].T, DA[l - 1], EPI_MUL_SIG);
      set_aux1(g, U[l - 1].hi());
      plan.nt(g);
    }
    }
    {
      GemmNTDesc g = nt_desc(DA[0], Aw[0].T, gmat, EPI_LINEAR);
      g.round_out = 0;
      plan.nt(g);
    }
    if (!train) {
      plan.add([=](cudaStream_t s) {
        unpad_kernel<<<grid_for(static_cast<size_t>(N) * d), 256, 0, s>>>(gmat.p, gmat.ld, bd->score_out, N, d, 1.0f);
        return static_cast<int>(cudaGetLastError());
      });
      return plan.error;
    }
    // ---- loss + residual direction r
    plan.add([=](cudaStream_t s) {
      cdae_loss_kernel<<<grid_for(static_cast<size_t>(N) * gmat.ld), 256, 0, s>>>(
          gmat.p, gmat.ld, sig, bd->eps, rmat.p, N, d, bd->inv_count, bd->loss_out, bd->score_out);
      return static_cast<int>(cudaGetLastError());
    });
    // ---- sweep 3: tangent forward (also emits t = delta * tangent_pre * (1 - sig))
    if (use_chain) {
      {
        GemmNTDesc g = nt_desc(rmat, Aw[0].hi(), UD[0], EPI_TANGENT);
        set_aux1(g, U[0].hi()); set_aux2(g, DA[0]); set_out2(g, TA[0]);
        plan.nt(g);
      }
      auto tg_layer = [&](const Mat& w, const Mat& aux1, const Mat& aux2, const Mat& out, const Mat& out2) {
        ChainLayerDesc q;
        q.W = w.p; q.ldw = w.ld; q.aux1 = aux1.p; q.ld1 = aux1.ld; q.aux2 = aux2.p; q.ld2 = aux2.ld;
        q.out = out.p; q.ldo = out.ld; q.out2 = out2.p; q.ldo2 = out2.ld;
        return q;
      };
      ChainDesc cd = chain_of(CHAIN_TANGENT, UD[0]);
      for (int l = 1; l < L; ++l) cd.layers.push_back(tg_layer(Aw[l].hi(), U[l].hi(), DA[l], UD[l], TA[l]));
      for (int l = 0; l < L; ++l) cd.layers.push_back(tg_layer(Ww[l].hi(), V[l].hi(), DP[l], VD[l], TP[l]));
      cd.layers.back().colsum = G(iW(L)); cd.layers.back().colsum_scale = -1.0f;  // d w_o = -sum_n vdot_L
      cd.layers.back().colsum2 = G(iW(L - 1) + 1);                                // d beta_L = sum_n adj p_L (= t_L)
      plan.chain(cd);
    } else {
    for (int l = 0; l < L; ++l) {
      GemmNTDesc g = nt_desc(l == 0 ? rmat : UD[l - 1], Aw[l].hi(), UD[l], EPI_TANGENT);
      set_aux1(g, U[l].hi()); set_aux2(g, DA[l]); set_out2(g, TA[l]);
      plan.nt(g);
    }
    for (int l = 0; l < L; ++l) {
      GemmNTDesc g = nt_desc(l == 0 ? UD[L - 1] : VD[l - 1], Ww[l].hi(), VD[l], EPI_TANGENT);
      set_aux1(g, V[l].hi()); set_aux2(g, DP[l]); set_out2(g, TP[l]);
      if (l == L - 1) {
        g.colsum = G(iW(L)); g.colsum_scale = -1.0f;  // d w_o = -sum_n vdot_L
        g.colsum2 = G(iW(L - 1) + 1);                 // d beta_L = sum_n adj p_L (= t_L)
      }
      plan.nt(g);
    }
    }
    // ---- sweep 4: adjoint backward (in place over the t buffers)
    // context-branch backward (B rows): side lane underneath the rest of the N-row work
    auto ctx_backward = [&]() {
    plan.fork();
    // ---- context branch backward (B rows)
    {
      const Mat ap1 = TP[0];
      plan.add([=](cudaStream_t s) {
        group_sum_kernel<<<B, 256, 0, s>>>(ap1.p, ap1.ld, gsum.p, gsum.ld, B, S, H, 1);
        return static_cast<int>(cudaGetLastError());
      });
    }
    {
      const Mat y = Cc[L - 1].hi();
      tn2(gsum, y, nullptr, nullptr, G(iW(0)) ? G(iW(0)) + H : nullptr, ld1);
    }
    {
      GemmNTDesc g = nt_desc(gsum, W1c.T, DC[L - 1], EPI_MUL_SIG);
      set_aux1(g, Cc[L - 1].hi());
      g.colsum = G(iC(L - 1) + 1);
      plan.nt(g);
    }
    for (int l = L - 1; l >= 1; --l) {
      GemmNTDesc g = nt_desc(DC[l], Cw[l].T, DC[l - 1], EPI_MUL_SIG);
      set_aux1(g, Cc[l - 1].hi());
      g.colsum = G(iC(l - 1) + 1);
      plan.nt(g);
    }
    for (int l = L - 1; l >= 1; --l) {
      const Mat y = Cc[l - 1].hi();
      tn2(DC[l], y, nullptr, nullptr, G(iC(l)), H);
    }
    {
      const Mat y = ctxp.hi();
      tn2(DC[0], y, nullptr, nullptr, G(iC(0)), c);
    }
    plan.cur_lane = 0;
    };
    if (use_chain) {
      auto adj_layer = [&](const Mat& wT, const Mat& aux1, const Mat& t, float* colsum) {
        ChainLayerDesc q;
        q.W = wT.p; q.ldw = wT.ld; q.aux1 = aux1.p; q.ld1 = aux1.ld; q.aux2 = t.p; q.ld2 = t.ld;
        q.out = t.p; q.ldo = t.ld; q.colsum = colsum;
        return q;
      };
      ChainDesc cd = chain_of(CHAIN_ADJOINT, TP[L - 1]);
      for (int l = L - 1; l >= 1; --l) {
        ChainLayerDesc q = adj_layer(Ww[l].T, V[l - 1].hi(), TP[l - 1], G(iW(l - 1) + 1));
        if (l - 1 == 0) {  // d w_1sigma = sum_n sigma_n * adj p_1
          q.colsum_w = G(iW(0)) ? G(iW(0)) + 2 * H : nullptr;
          q.colsum_w_stride = ld1;
        }
        cd.layers.push_back(q);
      }
      cd.layers.push_back(adj_layer(W1u.T, U[L - 1].hi(), TA[L - 1], G(iA(L - 1) + 1)));
      for (int l = L - 1; l >= 1; --l) cd.layers.push_back(adj_layer(Aw[l].T, U[l - 1].hi(), TA[l - 1], G(iA(l - 1) + 1)));
      plan.chain(cd);
      ctx_backward();  // adj p_1 (TP[0]) is final only once the chain has finished
    } else {
    for (int l = L - 1; l >= 1; --l) {
      GemmNTDesc g = nt_desc(TP[l], Ww[l].T, TP[l - 1], EPI_ADJOINT);
      set_aux1(g, V[l - 1].hi()); set_aux2(g, TP[l - 1]);
      g.colsum = G(iW(l - 1) + 1);
      if (l - 1 == 0) {  // d w_1sigma = sum_n sigma_n * adj p_1
        g.colsum_w = G(iW(0)) ? G(iW(0)) + 2 * H : nullptr;
        g.colsum_w_stride = ld1;
        g.row_w = sig;
      }
      plan.nt(g);
    }
    // adj p_1 (TP[0]) is final: the context-branch backward runs underneath the rest of sweep 4
    ctx_backward();
    {
      GemmNTDesc g = nt_desc(TP[0], W1u.T, TA[L - 1], EPI_ADJOINT);
      set_aux1(g, U[L - 1].hi()); set_aux2(g, TA[L - 1]);
      g.colsum = G(iA(L - 1) + 1);
      plan.nt(g);
    }
    for (int l = L - 1; l >= 1; --l) {
      GemmNTDesc g = nt_desc(TA[l], Aw[l].T, TA[l - 1], EPI_ADJOINT);
      set_aux1(g, U[l - 1].hi()); set_aux2(g, TA[l - 1]);
      g.colsum = G(iA(l - 1) + 1);
      plan.nt(g);
    }
    }
    // ---- weight gradients: dW = adj^T . act + delta^T . tangent   (accumulated into .grad)
    const Mat xt_hi = xt.hi();
    {
      const Mat y = U[L - 1].hi();
      tn2(TP[0], y, &DP[0], &UD[L - 1], G(iW(0)), ld1);
    }
    {  // the 2(L-1) two-pair [H, H] contractions as one launch (grid.z = layer)
      std::vector<GemmTNDesc> batch;
      for (int l = L - 1; l >= 1; --l) batch.push_back(tn_desc(TP[l], V[l - 1].hi(), &DP[l], &VD[l - 1], G(iW(l)), H));
      for (int l = L - 1; l >= 1; --l) batch.push_back(tn_desc(TA[l], U[l - 1].hi(), &DA[l], &UD[l - 1], G(iA(l)), H));
      plan.tn_batch(batch);
    }
    tn2(TA[0], xt_hi, &DA[0], &rmat, G(iA(0)), d);
    plan.join();
    return plan.error;
  }
};

}  // namespace ardae
