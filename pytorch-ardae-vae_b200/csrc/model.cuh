// Plan for the implicit-posterior VAE with the noise-concat MLP encoder
// (reference: models/ivae/toy.py `ToyIPVAE`, models/ivae/mnist.py `MNISTIPVAE`, enc_type 'concat').
//   encode   : z = f(x, eps) for B data rows x nz noise samples each     (forward only)
//   forward  : encode + decoder + ELBO terms, keeps the tape in the workspace
//   backward : parameter gradients from the tape for upstream (loss_scale * dloss + gz on z)
// Forward GEMMs run 3xTF32 (fp32-accurate: z is later scaled by std_scale = 1e4 and differenced
// against zbar, so a plain tf32 forward would leave an absolute error of ~S*5e-4*|z| on the CDAE
// input); backward GEMMs run tf32.
// Parameter tensors (state_dict order):
//   toy  : inp_encode (n_inp linears), fc.layers (n_fc), fc.fc, decode.main (n_dec), mean_fn, logvar_fn
//   mnist: inp_encode (n_inp linears), fc.layers (n_fc), fc.fc, decode.main (n_dec), logit_fn
#pragma once
#include <array>
#include "cdae.cuh"
#include "conv_gemm.cuh"

namespace ardae {

struct ModelConfig {
  int kind = 0;   // 0 toy (gaussian decoder, noise concatenated at every fc layer), 1 mnist, 2 conv (ConvIPVAE)
  int img_h = 0, img_c = 0;  // kind 2: image height (= width) and channels; D = img_c*img_h*img_h
  int D = 0, n = 0, h = 0, zd = 0;
  int n_inp = 0, n_fc = 0, n_dec = 0;  // number of linears in inp_encode, hidden fc layers, decode.main
  int act = 1;    // 0 relu, 1 softplus
  int B = 0, nz = 1;
  int mode = 0;   // 0 encode only, 1 forward + backward, 2 IWS log-likelihood (forward only),
                  // 3 decode(z) only, 4 encode._forward_inp(x) only, 5 encode._forward_all(inp, nos) only (nz = 1)
};

struct ModelBindings {
  const float* x = nullptr;      // [B, D]
  const float* noise = nullptr;  // [R, n] or null (= zeros: encode(std=0))
  float* z_out = nullptr;        // [R, zd]
  float* zbar_out = nullptr;     // [B, zd] mean code encode(x, std=0) from the same pass (fwd_mean plan), or null
  float* hidden_out = nullptr;   // kind 3: [B, 2h] cat(h0, h) at std = 0 (Encoder.forward_hidden, the 'hidden1a' context), or null
  float* heads_out = nullptr;    // [R, D] logits | [R, 2D] mu,logvar  (optional)
  float* sums = nullptr;         // [3] loss, recon, prior (device; overwritten)
  float beta = 1.0f;
  const float* beta_dev = nullptr;  // non-null: beta is read from this device scalar at kernel time (graph replay + annealing)
  float inv_rows = 0.0f;         // 1 / R_global
  // backward
  float loss_scale = 0.0f;
  const float* gz = nullptr;     // [R, zd] upstream gradient on z, or null
  float gz_scale = 1.0f;
  // IWS (mode 2)
  const float* eta = nullptr;    // [R, zd] standard normal behind MVN.rsample, or null (Philox)
  uint64_t seed = 0;
  float* iws_out = nullptr;      // [B] per-image log p_hat(x)
  float* iws_total = nullptr;    // device scalar, += sum over images
  // sub-module calls (modes 3-5)
  const float* z_in = nullptr;   // [R, zd]   decode(z)
  const float* inp_in = nullptr; // [R, feat] encode._forward_all(inp, nos)
  float* inp_out = nullptr;      // [B, feat] encode._forward_inp(x)
  int* status = nullptr;         // set to 1+image if a covariance is not positive definite
};

struct ModelPlan {
  ModelConfig cfg;
  Plan fwd, bwd_dec, bwd_enc;
  Plan fwd_mean;  // mode 0: z-bar = encode(x, std=0) from the per-data-row bias of the sampling pass (B rows)
  Workspace ws;
  ModelBindings bind;
  DeriveList derive;
  size_t tn_need = 0;

  int n_heads() const { return cfg.kind == 0 ? 2 : (cfg.kind == 2 ? 3 : 1); }  // kind 2: deconv1, deconv2, logit_fn
  int ntensors() const {
    if (cfg.kind == 3) return 2 * (cfg.n_inp + 2 + cfg.n_fc + 2 + cfg.n_dec + 1);  // aux main + 2 heads, fc + 2 heads, decoder + logits
    return 2 * (cfg.n_inp + cfg.n_fc + 1 + cfg.n_dec + n_heads());
  }

  // ------------------------------------------------------------------------------------------------------------------
  // kind 3: hierarchical implicit-posterior VAE `net.MNISTAuxIPVAE` (models/ivae/auxmnist.py:47-300, encoders
  // models/vae/auxmnist.py:31-68,147-191, decoder models/vae/mnist.py): z0 = mu0(x) + exp(lv0(x)/2) eps0 on the noise
  // width, z = mu(x, z0) + exp(lv(x, z0)/2) eps.  `noise` packs the two draws as [R, n + zd] (eps0 | eps); a null noise
  // pointer is `std = 0`.  Tensor order = state_dict order: aux_encode.main (n_inp linears), aux_encode.reparam
  // (mean_fn, logvar_fn), encode.fc (n_fc linears, first one [h, D + n]), encode.reparam (mean_fn, logvar_fn),
  // decode.main (n_dec linears), decode.reparam.logit_fn.  The x-half of encode.fc layer 0 enters once per data row
  // (rowbias0), as in the concat models.
  int build_aux(float* const* params, float* const* grads) {
    const ModelConfig& c = cfg;
    const int D = c.D, n = c.n, h = c.h, zd = c.zd, B = c.B, nz = c.nz, R = B * nz;
    if (c.mode < 0 || c.mode > 3) return fail(-2, "model: the auxmnist kind supports modes 0 (encode), 1 (train), 2 (IWS), 3 (decode)");
    if (c.mode == 3 && nz != 1) return fail(-2, "model: sub-module plans (modes 3-5) take nz = 1");
    if (c.act != 1) return fail(-2, "model: the auxmnist kind is softplus only");
    if (c.mode == 2 && (zd > 64 || nz < 2 * zd)) return fail(-2, "iws: need z_dim <= 64 and sample_size >= 2*z_dim");
    const bool dry = ws.dry, train = c.mode == 1, dec = c.mode >= 1;
    fwd.dry = bwd_dec.dry = bwd_enc.dry = fwd_mean.dry = dry;
    const int ACT = EPI_SOFTPLUS, DACT = EPI_MUL_SIG;
    auto P = [&](int i) -> float* { return dry ? nullptr : params[i]; };
    auto G = [&](int i) -> float* { return (dry || !train) ? nullptr : grads[i]; };
    auto iA = [&](int l) { return 2 * l; };
    auto iAH = [&](int k) { return 2 * (c.n_inp + k); };
    auto iF = [&](int l) { return 2 * (c.n_inp + 2 + l); };
    auto iFH = [&](int k) { return 2 * (c.n_inp + 2 + c.n_fc + k); };
    auto iD = [&](int l) { return 2 * (c.n_inp + 2 + c.n_fc + 2 + l); };
    const int iH = 2 * (c.n_inp + 2 + c.n_fc + 2 + c.n_dec);
    const int np = round_up(n, 4), zdp = round_up(zd, 4), Dp = round_up(D, 4), ne = n + zd;

    // ---- derived weights
    derive = DeriveList();
    std::vector<W3> Aw(c.n_inp), Fw(c.n_fc), Dw(c.n_dec);
    for (int l = 0; l < c.n_inp; ++l) Aw[l] = derive.add(ws, P(iA(l)), h, l == 0 ? D : h, l == 0 ? D : h, true, train && l > 0);
    const int ld0 = D + n;
    W3 F0i = derive.add(ws, P(iF(0)), h, D, ld0, true, false);
    W3 F0n = derive.add(ws, P(iF(0)) ? P(iF(0)) + D : nullptr, h, n, ld0, true, train);
    for (int l = 1; l < c.n_fc; ++l) Fw[l] = derive.add(ws, P(iF(l)), h, h, h, true, train);
    for (int l = 0; l < c.n_dec && dec; ++l) Dw[l] = derive.add(ws, P(iD(l)), h, l == 0 ? zd : h, l == 0 ? zd : h, true, train);
    // two-head operands: forward [2*out, 3*kp] (heads stacked), transposed [h, 2*outp] for the backward
    auto heads_w = [&](int i0, int out, int outp, bool want) {
      W3 w;
      w.in = h; w.out = 2 * out; w.kp = round_up(h, 32);
      if (!want) return w;
      w.b3 = Mat(ws.floats(static_cast<size_t>(2) * out * 3 * w.kp), 2 * out, 3 * w.kp, 3 * w.kp);
      if (train) w.T = ws.mat(h, 2 * outp);
      for (int k = 0; k < 2; ++k) {
        DeriveItem it;
        std::memset(&it, 0, sizeof(it));
        it.src = P(i0 + 2 * k); it.rows = out; it.cols = h; it.src_ld = h;
        it.dst3 = dry ? nullptr : w.b3.p + static_cast<size_t>(k) * out * w.b3.ld;
        it.kp = w.kp; it.ld3 = w.b3.ld;
        it.dstT = (dry || !train) ? nullptr : w.T.p + k * outp; it.ldT = train ? w.T.ld : 0;
        it.first_block = derive.blocks;
        it.tiles_x = (h + 31) / 32;
        derive.blocks += it.tiles_x * ((out + 31) / 32);
        derive.host.push_back(it);
      }
      return w;
    };
    W3 AH = heads_w(iAH(0), n, np, true), FH = heads_w(iFH(0), zd, zdp, true);
    W3 Hw;  // decoder logits
    Hw.in = h; Hw.out = D; Hw.kp = round_up(h, 32);
    if (dec) {
      Hw.b3 = Mat(ws.floats(static_cast<size_t>(D) * 3 * Hw.kp), D, 3 * Hw.kp, 3 * Hw.kp);
      if (train) Hw.T = ws.mat(h, Dp);
      DeriveItem it;
      std::memset(&it, 0, sizeof(it));
      it.src = P(iH); it.rows = D; it.cols = h; it.src_ld = h;
      it.dst3 = dry ? nullptr : Hw.b3.p; it.kp = Hw.kp; it.ld3 = Hw.b3.ld;
      it.dstT = (dry || !train) ? nullptr : Hw.T.p; it.ldT = train ? Hw.T.ld : 0;
      it.first_block = derive.blocks;
      it.tiles_x = (h + 31) / 32;
      derive.blocks += it.tiles_x * ((D + 31) / 32);
      derive.host.push_back(it);
    }
    int rc = derive.emit(ws, fwd);
    if (rc) return rc;

    // ---- buffers
    Pair xin = make_pair(ws, B, D);
    std::vector<Pair> A(c.n_inp), Fh(c.n_fc), Dh(c.n_dec);
    for (int l = 0; l < c.n_inp; ++l) A[l] = make_pair(ws, B, h);
    Mat M0 = ws.mat(B, 2 * np);  // [mu0 | lv0]
    Pair z0p = make_pair(ws, R, n);
    Mat rowbias0 = ws.mat(B, h);
    for (int l = 0; l < c.n_fc; ++l) Fh[l] = make_pair(ws, R, h);
    Mat M = ws.mat(R, 2 * zdp);  // [mu | lv]
    Pair zp = make_pair(ws, R, zd);
    Mat zbuf = ws.mat(R, zd);
    Mat heads, dheads, dzdec, dzt, dM, dz0, dM0, gsum0;
    std::vector<Mat> dD(c.n_dec), dF(c.n_fc), dA(c.n_inp);
    float *lw0 = nullptr, *wbuf = nullptr, *tn_ws = nullptr;
    size_t tn_bytes = 0;
    if (dec) {
      for (int l = 0; l < c.n_dec; ++l) Dh[l] = make_pair(ws, R, h);
      heads = ws.mat(R, Dp);
    }
    if (c.mode == 2) {
      lw0 = ws.floats(R);
      wbuf = ws.floats(R);
    }
    if (train) {
      dheads = ws.mat(R, Dp);
      dzdec = ws.mat(R, zd);
      dzt = ws.mat(R, zd);
      dM = ws.mat(R, 2 * zdp);
      dz0 = ws.mat(R, n);
      dM0 = ws.mat(B, 2 * np);
      gsum0 = ws.mat(B, h);
      for (int l = 0; l < c.n_dec; ++l) dD[l] = ws.mat(R, h);
      for (int l = 0; l < c.n_fc; ++l) dF[l] = ws.mat(R, h);
      for (int l = 0; l < c.n_inp; ++l) dA[l] = ws.mat(B, h);
      if (dry) {
        const int shapes[8][3] = {{h, h, R}, {h, n, R}, {2 * zd, h, R}, {D, h, R}, {h, zd, R}, {h, D, B}, {h, h, B}, {2 * n, h, B}};
        tn_need = 0;
        for (auto& sh : shapes) {
          const size_t b = tn_workspace_bytes(sh[0], sh[1], sh[2]);
          if (b > tn_need) tn_need = b;
        }
      }
      tn_bytes = tn_need;
      tn_ws = ws.floats(tn_bytes / 4);
    }
    ModelBindings* bd = &bind;

    // ---- decoder (models/vae/mnist.py Decoder: n_dec softplus layers + logits)
    auto decoder_fwd = [&]() {
      for (int l = 0; l < c.n_dec; ++l) {
        GemmNTDesc g = nt3_desc(l == 0 ? zp : Dh[l - 1], Dw[l], Dh[l], ACT);
        g.bias = P(iD(l) + 1);
        fwd.nt(g);
      }
      GemmNTDesc g = nt3_desc_plain(Dh[c.n_dec - 1], Hw, heads.cols_from(0, D), EPI_LINEAR);
      g.bias = P(iH + 1);
      fwd.nt(g);
    };
    if (c.mode == 3) {  // Decoder.forward(z) minus the sampler
      fwd.add([=](cudaStream_t s) {
        split2d_kernel<<<grid_for(static_cast<size_t>(R) * zd), 256, 0, s>>>(bd->z_in, zd, zp.buf.p, zp.buf.ld, R, zd, zp.kp, 1.0f, 0.0f);
        return static_cast<int>(cudaGetLastError());
      });
      decoder_fwd();
      fwd.add([=](cudaStream_t s) {
        unpad_kernel<<<grid_for(static_cast<size_t>(R) * D), 256, 0, s>>>(heads.p, heads.ld, bd->heads_out, R, D, 1.0f);
        return static_cast<int>(cudaGetLastError());
      });
      return fwd.error;
    }

    // ================================================================= forward
    fwd.add([=](cudaStream_t s) {
      split2d_kernel<<<grid_for(static_cast<size_t>(B) * D), 256, 0, s>>>(bd->x, D, xin.buf.p, xin.buf.ld, B, D, xin.kp, 2.0f, -1.0f);
      if (bd->sums) ARDAE_CUDA_OK(cudaMemsetAsync(bd->sums, 0, 3 * sizeof(float), s));
      return static_cast<int>(cudaGetLastError());
    });
    for (int l = 0; l < c.n_inp; ++l) {  // aux_encode.main on the data rows
      GemmNTDesc g = nt3_desc(l == 0 ? xin : A[l - 1], Aw[l], A[l], ACT);
      g.bias = P(iA(l) + 1);
      fwd.nt(g);
    }
    auto two_heads = [&](Plan& pl, const Pair& in, const W3& w, int out, int outp, const Mat& dst, int ib0) {
      for (int k = 0; k < 2; ++k) {
        W3 hk = w;
        hk.out = out;
        hk.b3 = w.b3.rows_from(k * out, out);
        GemmNTDesc g = nt3_desc_plain(in, hk, dst.cols_from(k * outp, out), EPI_LINEAR);
        g.bias = P(ib0 + 2 * k + 1);
        pl.nt(g);
      }
    };
    two_heads(fwd, A[c.n_inp - 1], AH, n, np, M0, iAH(0));
    fwd.add([=](cudaStream_t s) {  // z0 = mu0[b] + exp(lv0[b] / 2) eps0[r]
      aux_reparam_kernel<<<grid_for(static_cast<size_t>(R) * n), 256, 0, s>>>(
          M0.p, M0.ld, np, bd->noise, ne, R, n, nz, z0p.buf.p, z0p.buf.ld, z0p.kp, nullptr, 0, nullptr);
      return static_cast<int>(cudaGetLastError());
    });
    {
      GemmNTDesc g = nt3_desc_plain(xin, F0i, rowbias0, EPI_LINEAR);
      g.bias = P(iF(0) + 1);
      fwd.nt(g);
    }
    {
      GemmNTDesc g = nt3_desc(z0p, F0n, Fh[0], ACT);
      g.group_bias = rowbias0.p; g.group = nz; g.ldg = rowbias0.ld;
      fwd.nt(g);
    }
    for (int l = 1; l < c.n_fc; ++l) {
      GemmNTDesc g = nt3_desc(Fh[l - 1], Fw[l], Fh[l], ACT);
      g.bias = P(iF(l) + 1);
      fwd.nt(g);
    }
    two_heads(fwd, Fh[c.n_fc - 1], FH, zd, zdp, M, iFH(0));
    fwd.add([=](cudaStream_t s) {  // z = mu + exp(lv / 2) eps
      aux_reparam_kernel<<<grid_for(static_cast<size_t>(R) * zd), 256, 0, s>>>(
          M.p, M.ld, zdp, bd->noise != nullptr ? bd->noise + n : nullptr, ne, R, zd, 1, zp.buf.p, zp.buf.ld, zp.kp, zbuf.p,
          zbuf.ld, bd->z_out);
      return static_cast<int>(cudaGetLastError());
    });
    if (c.mode == 0) {
      // ---- std = 0 pass on the data rows: z-bar = mu(x, mu0(x)) and the 'hidden1a' context cat(h0, h)
      std::vector<Pair> Fb(c.n_fc);
      for (int l = 0; l < c.n_fc; ++l) Fb[l] = make_pair(ws, B, h);
      Pair z0b = make_pair(ws, B, n), zbp = make_pair(ws, B, zd);
      Mat Mb = ws.mat(B, 2 * zdp), hbuf = ws.mat(B, 2 * h);
      fwd_mean.add([=](cudaStream_t s) {
        aux_reparam_kernel<<<grid_for(static_cast<size_t>(B) * n), 256, 0, s>>>(M0.p, M0.ld, np, nullptr, 0, B, n, 1, z0b.buf.p,
                                                                              z0b.buf.ld, z0b.kp, nullptr, 0, nullptr);
        return static_cast<int>(cudaGetLastError());
      });
      {
        GemmNTDesc g = nt3_desc(z0b, F0n, Fb[0], ACT);
        g.group_bias = rowbias0.p; g.group = 1; g.ldg = rowbias0.ld;
        fwd_mean.nt(g);
      }
      for (int l = 1; l < c.n_fc; ++l) {
        GemmNTDesc g = nt3_desc(Fb[l - 1], Fw[l], Fb[l], ACT);
        g.bias = P(iF(l) + 1);
        fwd_mean.nt(g);
      }
      two_heads(fwd_mean, Fb[c.n_fc - 1], FH, zd, zdp, Mb, iFH(0));
      {
        const Pair a_last = A[c.n_inp - 1], f_last = Fb[c.n_fc - 1];
        fwd_mean.add([=](cudaStream_t s) {
          aux_reparam_kernel<<<grid_for(static_cast<size_t>(B) * zd), 256, 0, s>>>(Mb.p, Mb.ld, zdp, nullptr, 0, B, zd, 1, zbp.buf.p,
                                                                                 zbp.buf.ld, zbp.kp, nullptr, 0, bd->zbar_out);
          if (bd->hidden_out != nullptr) {  // [B, 2h] = [h0 | h], contiguous rows
            pair_sum_kernel<<<grid_for(static_cast<size_t>(B) * h), 256, 0, s>>>(a_last.hi().p, a_last.lo().p, a_last.buf.ld,
                                                                               hbuf.p, hbuf.ld, nullptr, B, h);
            pair_sum_kernel<<<grid_for(static_cast<size_t>(B) * h), 256, 0, s>>>(f_last.hi().p, f_last.lo().p, f_last.buf.ld,
                                                                               hbuf.p + h, hbuf.ld, nullptr, B, h);
            unpad_kernel<<<grid_for(static_cast<size_t>(B) * 2 * h), 256, 0, s>>>(hbuf.p, hbuf.ld, bd->hidden_out, B, 2 * h, 1.0f);
          }
          return static_cast<int>(cudaGetLastError());
        });
      }
      return fwd.error ? fwd.error : fwd_mean.error;
    }
    if (c.mode == 2) {
      const size_t smem = sizeof(float) * (static_cast<size_t>(zd) * zd + 2 * zd + 64 * zd);
      fwd.add([=](cudaStream_t s) {
        iws_moments_kernel<<<B, 256, smem, s>>>(zbuf.p, zbuf.ld, nz, zd, bd->eta, bd->seed, zp.buf.p, zp.buf.ld, zp.kp, lw0,
                                                bd->status, 1e-5f);
        return static_cast<int>(cudaGetLastError());
      });
    }
    decoder_fwd();
    if (c.mode == 2) {
      fwd.add([=](cudaStream_t s) {
        iws_loglik_kernel<<<R, 128, 0, s>>>(heads.p, heads.ld, Dp, bd->x, D, nz, 1, lw0, wbuf);
        iws_logmeanexp_kernel<<<B, 256, 0, s>>>(wbuf, nz, bd->iws_out, bd->iws_total);
        return static_cast<int>(cudaGetLastError());
      });
      return fwd.error;
    }
    fwd.add([=](cudaStream_t s) {
      bern_elbo_kernel<<<R, 256, 0, s>>>(heads.p, heads.ld, bd->x, D, zbuf.p, zbuf.ld, zd, nz, bd->beta, bd->inv_rows, bd->sums,
                                         dheads.p, dheads.ld, bd->beta_dev);
      if (bd->heads_out)
        unpad_kernel<<<grid_for(static_cast<size_t>(R) * D), 256, 0, s>>>(heads.p, heads.ld, bd->heads_out, R, D, 1.0f);
      return static_cast<int>(cudaGetLastError());
    });

    // ================================================================= backward
    auto tn1 = [&](Plan& pl, const Mat& X, const Mat& Y, float* dst, int ldo) {
      GemmTNDesc t;
      t.X0 = X.p; t.ldx0 = X.ld; t.Y0 = Y.p; t.ldy0 = Y.ld;
      t.M = X.cols; t.N = Y.cols; t.K = X.rows; t.out = dst; t.ldo = ldo;
      t.scale = 1.0f; t.beta = 1.0f; t.workspace = tn_ws; t.workspace_bytes = tn_bytes;
      pl.tn(t);
    };
    auto colsum_op = [&](Plan& pl, const Mat& X, float* dst) {
      pl.add([=](cudaStream_t s) {
        // tall-skinny operands (pixel-major conv gradients: 1e5 rows x 16..32 columns) need many row slabs
        const int col_blocks = (X.cols + 31) / 32;
        int slabs = X.rows >= 2048 ? 32 : (X.rows + 63) / 64;
        if (X.rows >= 16384 && col_blocks * slabs < 592) slabs = std::min(X.rows / 256, 592 / col_blocks);
        dim3 grid(col_blocks, slabs);
        colsum_kernel<<<grid, 256, 0, s>>>(X.p, X.ld, X.rows, X.cols, dst, 1.0f);
        return static_cast<int>(cudaGetLastError());
      });
    };
    // ---- decoder part
    {
      const Mat dh = dheads.cols_from(0, D);
      tn1(bwd_dec, dh, Dh[c.n_dec - 1].hi(), G(iH), h);
      colsum_op(bwd_dec, dh, G(iH + 1));
      GemmNTDesc g = nt_desc(dheads, Hw.T, dD[c.n_dec - 1], DACT);
      set_aux1(g, Dh[c.n_dec - 1].hi());
      g.colsum = G(iD(c.n_dec - 1) + 1);
      bwd_dec.nt(g);
    }
    for (int l = c.n_dec - 1; l >= 1; --l) {
      GemmNTDesc g = nt_desc(dD[l], Dw[l].T, dD[l - 1], DACT);
      set_aux1(g, Dh[l - 1].hi());
      g.colsum = G(iD(l - 1) + 1);
      bwd_dec.nt(g);
    }
    for (int l = c.n_dec - 1; l >= 1; --l) tn1(bwd_dec, dD[l], Dh[l - 1].hi(), G(iD(l)), h);
    tn1(bwd_dec, dD[0], zp.hi(), G(iD(0)), zd);
    {
      GemmNTDesc g = nt_desc(dD[0], Dw[0].T, dzdec, EPI_LINEAR);
      g.round_out = 0;
      bwd_dec.nt(g);
    }
    // ---- encoder part: z = mu + exp(lv/2) eps  ->  d mu = dz, d lv = dz * (z - mu) / 2
    bwd_enc.add([=](cudaStream_t s) {
      dz_total_kernel<<<grid_for(static_cast<size_t>(R) * zd), 256, 0, s>>>(
          dzdec.p, dzdec.ld, zbuf.p, zbuf.ld, bd->gz, bd->gz_scale, bd->loss_scale, bd->beta * bd->inv_rows, dzt.p, dzt.ld, R, zd,
          bd->beta_dev, bd->inv_rows);
      aux_reparam_bwd_kernel<<<grid_for(static_cast<size_t>(R) * zd), 256, 0, s>>>(
          dzt.p, dzt.ld, M.p, M.ld, zdp, bd->noise != nullptr ? bd->noise + n : nullptr, ne, R, zd, dM.p, dM.ld);
      return static_cast<int>(cudaGetLastError());
    });
    for (int k = 0; k < 2; ++k) {
      const Mat dk = dM.cols_from(k * zdp, zd);
      tn1(bwd_enc, dk, Fh[c.n_fc - 1].hi(), G(iFH(k)), h);
      colsum_op(bwd_enc, dk, G(iFH(k) + 1));
    }
    {
      GemmNTDesc g = nt_desc(dM, FH.T, dF[c.n_fc - 1], DACT);
      set_aux1(g, Fh[c.n_fc - 1].hi());
      g.colsum = G(iF(c.n_fc - 1) + 1);
      bwd_enc.nt(g);
    }
    for (int l = c.n_fc - 1; l >= 1; --l) {
      GemmNTDesc g = nt_desc(dF[l], Fw[l].T, dF[l - 1], DACT);
      set_aux1(g, Fh[l - 1].hi());
      g.colsum = G(iF(l - 1) + 1);
      bwd_enc.nt(g);
    }
    for (int l = c.n_fc - 1; l >= 1; --l) tn1(bwd_enc, dF[l], Fh[l - 1].hi(), G(iF(l)), h);
    // fc layer 0: z0 half over the R rows, x half through the per-data-row sum
    tn1(bwd_enc, dF[0], z0p.hi(), G(iF(0)) ? G(iF(0)) + D : nullptr, ld0);
    {
      const Mat d0 = dF[0];
      bwd_enc.add([=](cudaStream_t s) {
        group_sum_kernel<<<B, 256, 0, s>>>(d0.p, d0.ld, gsum0.p, gsum0.ld, B, nz, h, 1);
        return static_cast<int>(cudaGetLastError());
      });
    }
    tn1(bwd_enc, gsum0, xin.hi(), G(iF(0)), ld0);
    {  // d z0 = d hid_0 . W_z0
      GemmNTDesc g = nt_desc(dF[0], F0n.T, dz0, EPI_LINEAR);
      g.round_out = 0;
      bwd_enc.nt(g);
    }
    bwd_enc.add([=](cudaStream_t s) {  // through z0 = mu0[b] + exp(lv0[b]/2) eps0[r]
      aux_reparam_bwd_group_kernel<<<B, 128, 0, s>>>(dz0.p, dz0.ld, M0.p, M0.ld, np, bd->noise, ne, B, nz, n, dM0.p, dM0.ld);
      return static_cast<int>(cudaGetLastError());
    });
    for (int k = 0; k < 2; ++k) {
      const Mat dk = dM0.cols_from(k * np, n);
      tn1(bwd_enc, dk, A[c.n_inp - 1].hi(), G(iAH(k)), h);
      colsum_op(bwd_enc, dk, G(iAH(k) + 1));
    }
    {
      GemmNTDesc g = nt_desc(dM0, AH.T, dA[c.n_inp - 1], DACT);
      set_aux1(g, A[c.n_inp - 1].hi());
      g.colsum = G(iA(c.n_inp - 1) + 1);
      bwd_enc.nt(g);
    }
    for (int l = c.n_inp - 1; l >= 1; --l) {
      GemmNTDesc g = nt_desc(dA[l], Aw[l].T, dA[l - 1], DACT);
      set_aux1(g, A[l - 1].hi());
      g.colsum = G(iA(l - 1) + 1);
      bwd_enc.nt(g);
    }
    for (int l = c.n_inp - 1; l >= 1; --l) tn1(bwd_enc, dA[l], A[l - 1].hi(), G(iA(l)), h);
    tn1(bwd_enc, dA[0], xin.hi(), G(iA(0)), D);
    return fwd.error ? fwd.error : (bwd_dec.error ? bwd_dec.error : bwd_enc.error);
  }

  int build(float* const* params, float* const* grads) {
    const ModelConfig& c = cfg;
    const int D = c.D, n = c.n, h = c.h, zd = c.zd, B = c.B, nz = c.nz, R = B * nz;
    if (D <= 0 || n <= 0 || h <= 0 || zd <= 0 || B <= 0 || nz <= 0 || c.n_inp < 1 || c.n_fc < 1 || c.n_dec < 1)
      return fail(-2, "model: bad config");
    if (c.kind < 0 || c.kind > 3) return fail(-2, "model: kind must be 0 (toy), 1 (mnist), 2 (conv) or 3 (auxmnist)");
    if (c.kind == 3) return build_aux(params, grads);
    const bool conv = c.kind == 2;
    // conv geometry (models/ivae/conv.py:64-67): three 5x5 stride-2 pad-2 convs
    auto cos_ = [](int hin) { return (hin + 4 - 5) / 2 + 1; };
    const int s1 = conv ? c.img_h : 0, s2 = conv ? cos_(s1) : 0, s4 = conv ? cos_(s2) : 0, s8 = conv ? cos_(s4) : 0;
    const int d7 = conv ? (s8 - 1) * 2 + 1 : 0;         // deconv1 output (7 for 28x28)
    const int d8 = d7 + 1;                               // after ZeroPad2d((0,1,0,1))
    const int d15 = conv ? (d8 - 1) * 2 + 1 : 0;        // deconv2 output
    if (conv) {
      if (c.img_h <= 0 || c.img_c <= 0 || D != c.img_c * c.img_h * c.img_h) return fail(-2, "model: conv needs D = img_c*img_h^2");
      if ((d15 - 1) * 2 + 1 - 1 != c.img_h) return fail(-2, "model: conv decoder geometry does not reproduce img_h (28, 12, 20, ... work)");
      if (c.n_inp != 3 || c.n_fc != 1 || c.n_dec != 2) return fail(-2, "model: conv expects n_inp=3, n_fc=1, n_dec=2");
    }
    const int feat = conv ? s8 * s8 * 32 : h;             // width of the per-data-row features fed to fc layer 0
    auto dwid = [&](int l) { return conv ? (l == 0 ? 300 : feat) : h; };  // decoder fc widths (vae/conv.py:109)
    if (c.mode < 0 || c.mode > 5) return fail(-2, "model: bad mode");
    if (c.mode >= 3 && nz != 1) return fail(-2, "model: sub-module plans (modes 3-5) take nz = 1");
    const bool dry = ws.dry, train = c.mode == 1, dec = c.mode >= 1 && c.mode <= 3;
    const bool enc_inp = c.mode != 3 && c.mode != 5;  // the input stack runs
    const bool enc_fc = c.mode != 3 && c.mode != 4;   // the fc stack runs
    if (c.mode == 2 && (zd > 64 || nz < 2 * zd)) return fail(-2, "iws: need z_dim <= 64 and sample_size >= 2*z_dim (ivae/mnist.py:382)");
    fwd.dry = bwd_dec.dry = bwd_enc.dry = dry;
    const int ACT = c.act ? EPI_SOFTPLUS : EPI_RELU;
    const int DACT = c.act ? EPI_MUL_SIG : EPI_MUL_STEP;
    const bool toy = c.kind == 0;
    auto P = [&](int i) -> float* { return dry ? nullptr : params[i]; };
    auto G = [&](int i) -> float* { return (dry || !train) ? nullptr : grads[i]; };
    auto iI = [&](int l) { return 2 * l; };
    auto iF = [&](int l) { return 2 * (c.n_inp + l); };                 // l == n_fc: fc.fc
    auto iD = [&](int l) { return 2 * (c.n_inp + c.n_fc + 1 + l); };
    auto iH = [&](int k) { return 2 * (c.n_inp + c.n_fc + 1 + c.n_dec + k); };
    const int nH = n_heads();
    const int Dp = round_up(D, 4);  // each head starts on a 16-byte column boundary

    // ---- derived weights
    derive = DeriveList();
    std::vector<W3> Iw(c.n_inp), Dw(c.n_dec);
    for (int l = 0; l < c.n_inp && !conv; ++l) {
      const int in = l == 0 ? D : h;
      Iw[l] = derive.add(ws, P(iI(l)), h, in, in, true, train && l > 0);
    }
    // fc layer 0: [h, feat+n] = [W_inp | W_noise]
    const int ld0 = feat + n;
    W3 F0i = derive.add(ws, P(iF(0)), h, feat, ld0, true, train);
    // fused N-row sampling pass (enc_sample_sm100.cuh): MNIST-kind encode plans
    const bool enc_fused = c.kind == 1 && (c.mode == 0 || (c.mode == 2 && zd % 4 == 0)) && c.n_fc == 1 &&
                           enc_sample_supported(n, h, zd);
    // fused decoder + BCE row sums (dec_iws_sm100.cuh): MNIST-kind IWS plans
    const bool dec_fused = c.kind == 1 && c.mode == 2 && dec_iws_supported(zd, h, c.n_dec);
    W3 F0n = derive.add(ws, P(iF(0)) ? P(iF(0)) + feat : nullptr, h, n, ld0, true, false, enc_fused);
    // later fc layers (toy: input is [hid | eps]; mnist has none) and the final fc.fc
    std::vector<W3> Fw(c.n_fc + 1);
    for (int l = 1; l <= c.n_fc; ++l) {
      const int out = l == c.n_fc ? zd : h;
      const int in = toy ? h + n : h;
      Fw[l] = derive.add(ws, P(iF(l)), out, in, in, true, train, enc_fused && l == c.n_fc);
    }
    for (int l = 0; l < c.n_dec; ++l) {
      const int in = l == 0 ? zd : dwid(l - 1);
      if (dec) Dw[l] = derive.add(ws, P(iD(l)), dwid(l), in, in, true, train, dec_fused);
    }
    // heads: combined [nH*D, h] forward operand and [h, nH*D] transpose
    W3 Hw;
    Hw.in = h; Hw.out = nH * D; Hw.kp = round_up(h, 32);
    if (dec_fused) {
      Hw.k16 = round_up(h, 64);
      Hw.h16 = ws.mat16(D, 2 * Hw.k16);
    }
    if (dec && !conv) {
      Hw.b3 = Mat(ws.floats(static_cast<size_t>(nH) * D * 3 * Hw.kp), nH * D, 3 * Hw.kp, 3 * Hw.kp);
      if (train) Hw.T = ws.mat(h, nH * Dp);
      for (int k = 0; k < nH; ++k) {
        DeriveItem it;
        std::memset(&it, 0, sizeof(it));
        it.src = P(iH(k)); it.rows = D; it.cols = h; it.src_ld = h;
        it.dst3 = dry ? nullptr : Hw.b3.p + static_cast<size_t>(k) * D * Hw.b3.ld;
        it.kp = Hw.kp; it.ld3 = Hw.b3.ld;
        it.dstT = (dry || !train) ? nullptr : Hw.T.p + k * Dp; it.ldT = Hw.T.ld;
        if (dec_fused && k == 0 && !dry) { it.dst16 = Hw.h16.p; it.k16 = Hw.k16; }
        it.first_block = derive.blocks;
        it.tiles_x = (h + 31) / 32;
        derive.blocks += it.tiles_x * ((D + 31) / 32);
        derive.host.push_back(it);
      }
    }
    // conv layers as GEMMs (conv_gemm.cuh): encoder convs in the standard layouts ([Co, Ci*25] forward operand and
    // its transpose); the deconvs need the transposed 3xTF32 operand, derived by derive_t3_kernel below
    const bool cg = conv && conv_gemm_enabled();
    const int IC_ = c.img_c;
    std::vector<W3> Cw(3);
    Mat dT3[3], dWr[3];                       // deconv1, deconv2, logit deconv: [Co*25, 3*kp] and [Ci, Co*25]
    const int dci[3] = {32, 32, 16};          // deconv input channels
    const int dck[3] = {32 * 25, 16 * 25, IC_ * 25};
    if (cg) {
      const int cci[3] = {IC_, 16, 32}, cco[3] = {16, 32, 32};
      for (int l = 0; l < 3 && enc_inp; ++l)
        Cw[l] = derive.add(ws, P(iI(l)), cco[l], cci[l] * 25, cci[l] * 25, true, train && l > 0);
      for (int k = 0; k < 3 && dec; ++k) {
        dT3[k] = Mat(ws.floats(static_cast<size_t>(dck[k]) * 96), dck[k], 96, 96);  // kp = 32 >= Ci
        if (train) dWr[k] = ws.mat(dci[k], dck[k]);
      }
    }
    int rc = derive.emit(ws, fwd);
    if (rc) return rc;
    if (cg && dec) {
      const float* wsrc[3] = {P(iH(0)), P(iH(1)), P(iH(2))};
      for (int k = 0; k < 3; ++k) {
        const float* w = wsrc[k];
        const Mat t3 = dT3[k], wr = dWr[k];
        const int rows = dci[k], cols = dck[k];
        fwd.add([=](cudaStream_t s) {
          derive_t3_kernel<<<grid_for(static_cast<size_t>(rows) * cols), 256, 0, s>>>(w, rows, cols, t3.p, 32, wr.p, wr.ld);
          return static_cast<int>(cudaGetLastError());
        });
      }
    }

    // ---- buffers
    Pair xin;
    if (!conv) xin = make_pair(ws, B, D);
    std::vector<Pair> I(c.n_inp), Fh(c.n_fc + 1), Dh(c.n_dec);
    for (int l = 0; l < c.n_inp && !conv; ++l) I[l] = make_pair(ws, B, h);
    // conv front end: NCHW fp32 feature maps on the B data rows, then the flattened conv3 output as a pair
    float *c1 = nullptr, *c2 = nullptr, *c3 = nullptr;
    // GEMM form: im2col pairs col_l [B*Ho^2, Ci*25] and pixel-major activated outputs a_l [B*Ho^2, Co]
    Pair colp[3];
    Mat act_[3];
    const int eho[3] = {s2, s4, s8}, eci[3] = {c.img_c, 16, 32}, eco[3] = {16, 32, 32};
    if (conv && !cg) {
      c1 = ws.floats(static_cast<size_t>(B) * 16 * s2 * s2);
      c2 = ws.floats(static_cast<size_t>(B) * 32 * s4 * s4);
      c3 = ws.floats(static_cast<size_t>(B) * feat);
    }
    if (cg && enc_inp) {
      for (int l = 0; l < 3; ++l) {
        colp[l] = make_pair(ws, B * eho[l] * eho[l], eci[l] * 25);
        act_[l] = ws.mat(B * eho[l] * eho[l], eco[l]);
      }
    }
    if (conv) I[c.n_inp - 1] = make_pair(ws, B, feat);
    const Pair inp_pair = I[c.n_inp - 1];
    Pair epsp = make_pair(ws, R, n);
    Mat rowbias0 = ws.mat(B, h);
    // fc hidden outputs; for toy they live inside the concat buffer [hid | eps] of the next layer
    for (int l = 0; l < c.n_fc; ++l) Fh[l] = make_pair(ws, R, toy ? h + n : h);
    Pair zp = make_pair(ws, R, zd);
    Mat zbuf = ws.mat(R, zd);  // plain fp32 z (exact sum hi+lo is not needed: LINEAR plain output)
    Mat heads, dheads, dzdec, dzt;
    std::vector<Mat> dD(c.n_dec), dF(c.n_fc), dI(c.n_inp);
    Mat gsum0;
    float* tn_ws = nullptr;
    size_t tn_bytes = 0;
    float *lw0 = nullptr, *wbuf = nullptr;
    // conv back end (vae/conv.py:126-131): h1 [R,32,s8,s8] -> deconv1 -> pad -> deconv2 -> logit deconv -> crop
    float *h1 = nullptr, *h2p = nullptr, *h3 = nullptr, *dh3 = nullptr, *dh2 = nullptr, *dh1 = nullptr;
    Pair dp_[3];
    float* dcols_fwd = nullptr;
    Mat dcolb[3], ddp[3];   // backward: im2col'd upstream gradients [rows, Co*25] and d pre-activations (pixel-major)
    Mat ecol[3], eda[3];    // encoder backward: dcol_l [B*Ho^2, Ci*25] (l = 1, 2) and dA_l [B*Ho^2, Co]
    if (dec) {
      for (int l = 0; l < c.n_dec; ++l) Dh[l] = make_pair(ws, R, dwid(l));
      if (conv) {
        heads = Mat(ws.floats(static_cast<size_t>(R) * D), R, D, D);
        if (!cg) {
          h1 = ws.floats(static_cast<size_t>(R) * feat);
          h2p = ws.floats(static_cast<size_t>(R) * 32 * d8 * d8);   // zero-padded (pad row/col never written)
          h3 = ws.floats(static_cast<size_t>(R) * 16 * d15 * d15);
        } else {
          // pixel-major pairs: p1 = decode.fc output [R*s8^2, 32], p2 = padded deconv1 output [R*d8^2, 32] (pad row /
          // column stay zero), p3 = deconv2 output [R*d15^2, 16]; one scratch for the three column matrices
          dp_[0] = make_pair(ws, R * s8 * s8, 32);
          dp_[1] = make_pair(ws, R * d8 * d8, 32);
          dp_[2] = make_pair(ws, R * d15 * d15, 16);
          size_t need = static_cast<size_t>(R) * s8 * s8 * dck[0];
          need = std::max(need, static_cast<size_t>(R) * d8 * d8 * dck[1]);
          need = std::max(need, static_cast<size_t>(R) * d15 * d15 * round_up(dck[2], 4));
          dcols_fwd = ws.floats(need);
        }
      } else {
        heads = ws.mat(R, nH * Dp);
      }
    }
    if (c.mode == 2) {
      lw0 = ws.floats(R);
      wbuf = ws.floats(R);
    }
    if (train) {
      dheads = conv ? Mat(ws.floats(static_cast<size_t>(R) * D), R, D, D) : ws.mat(R, nH * Dp);
      dzdec = ws.mat(R, zd);
      dzt = ws.mat(R, zd);
      for (int l = 0; l < c.n_dec; ++l) dD[l] = ws.mat(R, dwid(l));
      for (int l = 0; l < c.n_fc; ++l) dF[l] = ws.mat(R, h);
      for (int l = 0; l < c.n_inp; ++l) dI[l] = ws.mat(B, (conv && l == c.n_inp - 1) ? feat : (conv ? 4 : h));
      gsum0 = ws.mat(B, h);
      if (conv && !cg) {
        dh3 = ws.floats(static_cast<size_t>(R) * 16 * d15 * d15);
        dh2 = ws.floats(static_cast<size_t>(R) * 32 * d8 * d8);
        dh1 = ws.floats(static_cast<size_t>(R) * feat);
      }
      if (cg) {
        const int drows[3] = {R * s8 * s8, R * d8 * d8, R * d15 * d15};
        for (int k = 0; k < 3; ++k) {
          dcolb[k] = ws.mat(drows[k], dck[k]);
          ddp[k] = ws.mat(drows[k], dci[k]);
        }
        for (int l = 0; l < 3; ++l) {
          eda[l] = ws.mat(B * eho[l] * eho[l], eco[l]);
          if (l > 0) ecol[l] = ws.mat(B * eho[l] * eho[l], eci[l] * 25);
        }
      }
      if (dry) {
        std::vector<std::array<int, 3>> shapes = {{h, h + n, R}, {h, h, R}, {zd, h + n, R}, {nH * D, h, R}, {h, zd, R},
                                                  {h, D, B}, {h, h, B}, {h, n, R}, {h, feat, B}, {feat, 300, R}};
        if (cg) {
          for (int l = 0; l < 3; ++l) shapes.push_back({eco[l], eci[l] * 25, B * eho[l] * eho[l]});
          shapes.push_back({dci[0], dck[0], R * s8 * s8});
          shapes.push_back({dci[1], dck[1], R * d8 * d8});
          shapes.push_back({dci[2], dck[2], R * d15 * d15});
        }
        tn_need = 0;
        for (auto& sh : shapes) {
          const size_t b = tn_workspace_bytes(sh[0], sh[1], sh[2]);
          if (b > tn_need) tn_need = b;
        }
      }
      tn_bytes = tn_need;
      tn_ws = ws.floats(tn_bytes / 4);
    }
    ModelBindings* bd = &bind;

    // ================================================================= forward
    const int CACT = c.act ? CONV_ACT_SOFTPLUS : CONV_ACT_RELU;
    const int IH = c.img_h, IC = c.img_c;
    fwd.add([=](cudaStream_t s) {
      if (!conv && enc_inp)
        split2d_kernel<<<grid_for(static_cast<size_t>(B) * D), 256, 0, s>>>(
            bd->x, D, xin.buf.p, xin.buf.ld, B, D, xin.kp, toy ? 1.0f : 2.0f, toy ? 0.0f : -1.0f);
      if (bd->sums) ARDAE_CUDA_OK(cudaMemsetAsync(bd->sums, 0, 3 * sizeof(float), s));
      return static_cast<int>(cudaGetLastError());
    });
    // the (HBM-bound, R-row) noise split runs on the plan's side lane underneath the latency-bound B-row input stack;
    // joined right before the first layer that consumes the noise pair
    if (enc_fc && !enc_fused) fwd.fork();
    if (enc_fc && !enc_fused) fwd.add([=](cudaStream_t s) {
      if (bd->noise != nullptr) {
        split2d_kernel<<<grid_for(static_cast<size_t>(R) * n), 256, 0, s>>>(
            bd->noise, n, epsp.buf.p, epsp.buf.ld, R, n, epsp.kp, 1.0f, 0.0f);
      } else {
        ARDAE_CUDA_OK(cudaMemsetAsync(epsp.buf.p, 0, sizeof(float) * R * epsp.buf.ld, s));
      }
      return static_cast<int>(cudaGetLastError());
    });
    fwd.cur_lane = 0;
    if (c.mode == 5) {  // encode._forward_all(inp, nos): the per-row features are given
      fwd.add([=](cudaStream_t s) {
        split2d_kernel<<<grid_for(static_cast<size_t>(B) * feat), 256, 0, s>>>(
            bd->inp_in, feat, inp_pair.buf.p, inp_pair.buf.ld, B, feat, inp_pair.kp, 1.0f, 0.0f);
        return static_cast<int>(cudaGetLastError());
      });
    }
    // inp_encode on the B data rows
    for (int l = 0; l < c.n_inp && !conv && enc_inp; ++l) {
      GemmNTDesc g = nt3_desc(l == 0 ? xin : I[l - 1], Iw[l], I[l], ACT);
      g.bias = P(iI(l) + 1);
      fwd.nt(g);
    }
    if (cg && enc_inp) {
      // x <- 2x-1, conv(1->16) -> conv(16->32) -> conv(32->32), 5x5 s2 p2 + activation (ivae/conv.py:84-96):
      // im2col (pair) -> 3xTF32 GEMM with the bias + activation epilogue, pixel-major outputs
      const int hin[3] = {c.img_h, s2, s4};
      for (int l = 0; l < 3; ++l) {
        const Pair cp = colp[l];
        const Mat prev = l > 0 ? act_[l - 1] : Mat();
        const int Ho = eho[l], Hi = hin[l], Ci = eci[l];
        fwd.add([=](cudaStream_t s) {
          const ImgView in = l == 0 ? img_nchw(const_cast<float*>(bd->x), Ci, Hi) : img_pix(prev.p, prev.ld, Ci, Hi, Hi);
          launch_im2col5s2(in, Ho, l == 0 ? 2.0f : 1.0f, l == 0 ? -1.0f : 0.0f, cp.buf.p, cp.buf.ld, cp.kp, B, s);
          return static_cast<int>(cudaGetLastError());
        });
        GemmNTDesc g = nt3_desc_plain(colp[l], Cw[l], act_[l], ACT);
        g.bias = P(iI(l) + 1);
        fwd.nt(g);
      }
      {
        const Mat a3 = act_[2];
        fwd.add([=](cudaStream_t s) {  // flatten in the reference's NCHW order (feature c*P + p), as a pair
          chw_pix_permute_kernel<false><<<grid_for(static_cast<size_t>(B) * feat), 256, 0, s>>>(
              a3.p, a3.ld, 0, inp_pair.buf.p, inp_pair.buf.ld, inp_pair.kp, B, 32, s8 * s8, 0);
          return static_cast<int>(cudaGetLastError());
        });
      }
    }
    if (conv && !cg && enc_inp) {
      // x <- 2x-1, conv(1->16) -> conv(16->32) -> conv(32->32), all 5x5 s2 p2 + activation (ivae/conv.py:84-96)
      const float *w1 = P(iI(0)), *b1 = P(iI(0) + 1), *w2 = P(iI(1)), *b2 = P(iI(1) + 1), *w3 = P(iI(2)), *b3 = P(iI(2) + 1);
      fwd.add([=](cudaStream_t s) {
        conv5s2_fwd_kernel<<<grid_for(static_cast<size_t>(B) * 16 * s2 * s2), 256, 0, s>>>(
            bd->x, IC, IH, IH, IH, w1, b1, c1, 16, s2, s2, s2, B, 2.0f, -1.0f, CACT, nullptr);
        conv5s2_fwd_kernel<<<grid_for(static_cast<size_t>(B) * 32 * s4 * s4), 256, 0, s>>>(
            c1, 16, s2, s2, s2, w2, b2, c2, 32, s4, s4, s4, B, 1.0f, 0.0f, CACT, nullptr);
        conv5s2_fwd_kernel<<<grid_for(static_cast<size_t>(B) * 32 * s8 * s8), 256, 0, s>>>(
            c2, 32, s4, s4, s4, w3, b3, c3, 32, s8, s8, s8, B, 1.0f, 0.0f, CACT, nullptr);
        split2d_kernel<<<grid_for(static_cast<size_t>(B) * feat), 256, 0, s>>>(
            c3, feat, inp_pair.buf.p, inp_pair.buf.ld, B, feat, inp_pair.kp, 1.0f, 0.0f);
        return static_cast<int>(cudaGetLastError());
      });
    }
    if (c.mode == 4) {  // encode._forward_inp(x): the features of the data rows
      const Mat ih = inp_pair.hi(), il = inp_pair.lo();
      Mat ibuf = ws.mat(B, feat);
      fwd.add([=](cudaStream_t s) {
        pair_sum_kernel<<<grid_for(static_cast<size_t>(B) * feat), 256, 0, s>>>(ih.p, il.p, ih.ld, ibuf.p, ibuf.ld,
                                                                              bd->inp_out, B, feat);
        return static_cast<int>(cudaGetLastError());
      });
      return fwd.error;
    }
    // fc layer 0: input half once per data row (rowbias0), noise half over the R rows
    if (enc_fc) {
      GemmNTDesc g = nt3_desc_plain(inp_pair, F0i, rowbias0, EPI_LINEAR);
      g.bias = P(iF(0) + 1);
      fwd.nt(g);
    }
    if (enc_fc && !enc_fused) fwd.join();
    if (enc_fused) {
      EncSampleDesc ed;
      ed.W1 = F0n.h16.p; ed.ldw1 = F0n.h16.ld; ed.W2 = Fw[c.n_fc].h16.p; ed.ldw2 = Fw[c.n_fc].h16.ld;
      ed.rowbias = rowbias0.p; ed.ldb = rowbias0.ld; ed.bias2 = P(iF(c.n_fc) + 1); ed.z_out = zbuf.p;
      ed.R = R; ed.nz = nz; ed.n = n; ed.h = h; ed.zd = zd;
      if (!dry) {
        PreparedEncSample pe;
        int rc2 = prepare_enc_sample(ed, &pe);
        if (rc2) return rc2;
        auto sp = std::make_shared<PreparedEncSample>(pe);
        float* zdst = c.mode == 2 ? zbuf.p : nullptr;  // IWS plans keep z in the workspace (moment matching reads it)
        float* znull = ws.floats(static_cast<size_t>(n) * 128 + 4);  // zero noise rows for encode(std=0) (workspace is zeroed)
        fwd.add([=](cudaStream_t s) {
          // noise == NULL (encode(x, std=0)): every row reads the same zero row through a zero row pitch
          PreparedEncSample q = *sp;
          if (bd->noise == nullptr) q.params.n = 0;
          return launch_prepared_enc_sample(q, bd->noise != nullptr ? bd->noise : znull, zdst != nullptr ? zdst : bd->z_out, s);
        });
      } else {
        ws.floats(static_cast<size_t>(n) * 128 + 4);
        fwd.add(nullptr);
      }
      fwd.tag_last("enc_sample");
    }
    if (enc_fc && !enc_fused) {
      Pair out = Fh[0];
      out.w = h;  // write only the hid columns of the (possibly wider) concat buffer
      GemmNTDesc g = nt3_desc(epsp, F0n, out, ACT);
      g.group_bias = rowbias0.p; g.group = nz; g.ldg = rowbias0.ld;
      // N-row sampling with 256 < h <= 512: two 256-wide tiles (A read twice) instead of three 128-wide ones
      if (h > 256 && h <= 512 && R >= 16384) g.force_block_n = 256;
      fwd.nt(g);
    }
    if (toy && enc_fc && !enc_fused) {
      // append eps to every concat buffer: columns [h, h+n) of hi and lo halves
      fwd.add([=](cudaStream_t s) {
        for (int l = 0; l < c.n_fc; ++l) {
          const Pair& f = Fh[l];
          if (bd->noise != nullptr)
            split2d_kernel<<<grid_for(static_cast<size_t>(R) * n), 256, 0, s>>>(
                bd->noise, n, f.buf.p + h, f.buf.ld, R, n, f.kp, 1.0f, 0.0f);
          else {
            ARDAE_CUDA_OK(cudaMemset2DAsync(f.buf.p + h, sizeof(float) * f.buf.ld, 0, sizeof(float) * n, R, s));
            ARDAE_CUDA_OK(cudaMemset2DAsync(f.buf.p + f.kp + h, sizeof(float) * f.buf.ld, 0, sizeof(float) * n, R, s));
          }
        }
        return static_cast<int>(cudaGetLastError());
      });
    }
    for (int l = 1; l < c.n_fc && enc_fc && !enc_fused; ++l) {
      Pair out = Fh[l];
      out.w = h;
      GemmNTDesc g = nt3_desc(Fh[l - 1], Fw[l], out, ACT);
      g.bias = P(iF(l) + 1);
      fwd.nt(g);
    }
    if (c.mode == 3) {  // decode(z): the latent rows are given
      fwd.add([=](cudaStream_t s) {
        split2d_kernel<<<grid_for(static_cast<size_t>(R) * zd), 256, 0, s>>>(bd->z_in, zd, zp.buf.p, zp.buf.ld, R, zd,
                                                                           zp.kp, 1.0f, 0.0f);
        return static_cast<int>(cudaGetLastError());
      });
    }
    if (enc_fc && !enc_fused) {  // z = fc.fc([hid | eps])  (plain fp32 to the user buffer layout, and as a tf32 pair)
      GemmNTDesc g = nt3_desc(Fh[c.n_fc - 1], Fw[c.n_fc], zp, EPI_LINEAR);
      g.bias = P(iF(c.n_fc) + 1);
      fwd.nt(g);
      const Mat zh = zp.hi(), zl = zp.lo();
      fwd.add([=](cudaStream_t s) {
        // z = hi + lo (exact to ~2^-22 relative)
        pair_sum_kernel<<<grid_for(static_cast<size_t>(R) * zd), 256, 0, s>>>(zh.p, zl.p, zh.ld, zbuf.p, zbuf.ld,
                                                                            bd->z_out, R, zd);
        return static_cast<int>(cudaGetLastError());
      });
    }
    if (c.mode == 0) {
      // ---- mean code: for eps = 0 the noise half of fc layer 0 contributes nothing, so hid_0 = act(rowbias0) on the
      // B data rows; the remaining fc layers see [hid | 0].  Reuses the input stack of the sampling pass
      // (ivae_ardae.py:735,748 recompute it: encode(x, std=0) twice per update).
      fwd_mean.dry = dry;
      std::vector<Pair> Fb(c.n_fc);
      for (int l = 0; l < c.n_fc; ++l) Fb[l] = make_pair(ws, B, toy ? h + n : h);  // eps columns stay zero
      Pair zbp = make_pair(ws, B, zd);
      Mat zbb = ws.mat(B, zd);
      {
        const Pair f0 = Fb[0];
        const int ACTI = c.act;
        fwd_mean.add([=](cudaStream_t s) {
          act_split_kernel<<<grid_for(static_cast<size_t>(B) * h), 256, 0, s>>>(rowbias0.p, rowbias0.ld, f0.buf.p,
                                                                               f0.buf.ld, B, h, f0.kp, ACTI);
          return static_cast<int>(cudaGetLastError());
        });
      }
      for (int l = 1; l < c.n_fc; ++l) {
        Pair out = Fb[l];
        out.w = h;
        GemmNTDesc g = nt3_desc(Fb[l - 1], Fw[l], out, ACT);
        g.bias = P(iF(l) + 1);
        fwd_mean.nt(g);
      }
      {
        GemmNTDesc g = nt3_desc(Fb[c.n_fc - 1], Fw[c.n_fc], zbp, EPI_LINEAR);
        g.bias = P(iF(c.n_fc) + 1);
        fwd_mean.nt(g);
        const Mat zh = zbp.hi(), zl = zbp.lo();
        fwd_mean.add([=](cudaStream_t s) {
          pair_sum_kernel<<<grid_for(static_cast<size_t>(B) * zd), 256, 0, s>>>(zh.p, zl.p, zh.ld, zbb.p, zbb.ld,
                                                                              bd->zbar_out, B, zd);
          return static_cast<int>(cudaGetLastError());
        });
      }
      return fwd.error ? fwd.error : fwd_mean.error;
    }
    if (!dec) return fwd.error;
    if (c.mode == 2) {
      // moment-matched Gaussian proposal per image; overwrites the z pair with newz = mu + L eta
      const size_t smem = sizeof(float) * (static_cast<size_t>(zd) * zd + 2 * zd + 64 * zd);
      fwd.add([=](cudaStream_t s) {
        iws_moments_kernel<<<B, 256, smem, s>>>(zbuf.p, zbuf.ld, nz, zd, bd->eta, bd->seed, zp.buf.p, zp.buf.ld,
                                                zp.kp, lw0, bd->status);
        return static_cast<int>(cudaGetLastError());
      });
    }
    if (dec_fused) {
      DecIwsDesc dd;
      for (int l = 0; l < c.n_dec; ++l) {
        dd.W[l] = Dw[l].h16.p; dd.ldw[l] = Dw[l].h16.ld; dd.bias[l] = P(iD(l) + 1);
      }
      dd.Wlogit = Hw.h16.p; dd.ldwl = Hw.h16.ld; dd.bias_logit = P(iH(0) + 1);
      dd.z_hi = zp.hi().p; dd.z_lo = zp.lo().p; dd.ldz = zp.buf.ld; dd.lw0 = lw0; dd.w = wbuf;
      dd.R = R; dd.S = nz; dd.zd = zd; dd.h = h; dd.D = D; dd.nhid = c.n_dec;
      if (!dry) {
        PreparedDecIws pd;
        int rc2 = prepare_dec_iws(dd, &pd);
        if (rc2) return rc2;
        auto sp = std::make_shared<PreparedDecIws>(pd);
        fwd.add([=](cudaStream_t s) {
          int rc3 = launch_prepared_dec_iws(*sp, bd->x, s);
          if (rc3) return rc3;
          iws_logmeanexp_kernel<<<B, 256, 0, s>>>(wbuf, nz, bd->iws_out, bd->iws_total);
          return static_cast<int>(cudaGetLastError());
        });
      } else {
        fwd.add(nullptr);
      }
      fwd.tag_last("dec_iws");
      return fwd.error;
    }
    // decoder
    for (int l = 0; l < c.n_dec; ++l) {
      GemmNTDesc g = nt3_desc(l == 0 ? zp : Dh[l - 1], Dw[l], Dh[l], ACT);
      g.bias = P(iD(l) + 1);
      fwd.nt(g);
    }
    for (int k = 0; k < nH && !conv; ++k) {
      W3 hk = Hw;
      hk.out = D;
      hk.b3 = Hw.b3.rows_from(k * D, D);
      GemmNTDesc g = nt3_desc_plain(Dh[c.n_dec - 1], hk, heads.cols_from(k * Dp, D), EPI_LINEAR);
      g.bias = P(iH(k) + 1);
      fwd.nt(g);
    }
    if (cg) {
      // deconv = (in . W) scattered by col2im (+ bias, activation); the reference pads deconv1's activated output to
      // d8 x d8 and crops the last row / column of the logits (vae/conv.py:128-131)
      const Pair hl = Dh[c.n_dec - 1];
      {
        const Pair p1 = dp_[0];
        fwd.add([=](cudaStream_t s) {
          chw_pix_permute_kernel<true><<<grid_for(static_cast<size_t>(R) * feat), 256, 0, s>>>(
              hl.buf.p, hl.buf.ld, hl.kp, p1.buf.p, p1.buf.ld, p1.kp, R, 32, s8 * s8, 0);
          return static_cast<int>(cudaGetLastError());
        });
      }
      const int gin[3] = {s8, d8, d15};            // grid of the GEMM rows (input pixels)
      const int gout[3] = {d7, d15, c.img_h};      // logical output size (the logit deconv's 29th row / column is cropped)
      const int gpitch[3] = {d8, d15, c.img_h};
      const int gco[3] = {32, 16, c.img_c};
      for (int k = 0; k < 3; ++k) {
        const Mat cols = Mat(dcols_fwd, R * gin[k] * gin[k], dck[k], round_up(dck[k], 4));
        W3 wk;
        wk.in = dci[k]; wk.out = dck[k]; wk.kp = 32; wk.b3 = dT3[k];
        GemmNTDesc g = nt3_desc_plain(dp_[k], wk, cols, EPI_LINEAR);
        fwd.nt(g);
        const float* bias = P(iH(k) + 1);
        const Pair nxt = k < 2 ? dp_[k + 1] : Pair();
        const int Ho = gin[k], Hout = gout[k], pitch = gpitch[k], Co = gco[k];
        fwd.add([=](cudaStream_t s) {
          const ImgView out = k < 2 ? img_pix(nxt.buf.p, nxt.buf.ld, Co, Hout, pitch) : img_nchw(heads.p, Co, Hout);
          col2im5s2_kernel<<<grid_for(static_cast<size_t>(R) * Hout * Hout * Co), 256, 0, s>>>(
              cols.p, cols.ld, Ho, out, bias, CACT, k < 2 ? C2I_ACT_PAIR : C2I_PLAIN, k < 2 ? nxt.kp : 0, nullptr, R);
          return static_cast<int>(cudaGetLastError());
        });
      }
    }
    if (conv && !cg) {
      // deconvs as the adjoint (backward-data) form of the 5x5 s2 p2 conv; the reference pads deconv1's
      // activated output to d8 x d8 and crops the last row/column of the logits (vae/conv.py:128-131)
      const Pair hl = Dh[c.n_dec - 1];
      const float *wd1 = P(iH(0)), *bd1 = P(iH(0) + 1), *wd2 = P(iH(1)), *bd2 = P(iH(1) + 1), *wl = P(iH(2)), *bl = P(iH(2) + 1);
      fwd.add([=](cudaStream_t s) {
        pair_sum_kernel<<<grid_for(static_cast<size_t>(R) * feat), 256, 0, s>>>(hl.hi().p, hl.lo().p, hl.buf.ld, h1, feat,
                                                                              nullptr, R, feat);
        conv5s2_bwd_data_kernel<<<grid_for(static_cast<size_t>(R) * 32 * d7 * d7), 256, 0, s>>>(
            h1, 32, s8, s8, s8, wd1, bd1, h2p, 32, d7, d8, d8, R, CACT, nullptr, 0);
        conv5s2_bwd_data_kernel<<<grid_for(static_cast<size_t>(R) * 16 * d15 * d15), 256, 0, s>>>(
            h2p, 32, d8, d8, d8, wd2, bd2, h3, 16, d15, d15, d15, R, CACT, nullptr, 0);
        conv5s2_bwd_data_kernel<<<grid_for(static_cast<size_t>(R) * IC * IH * IH), 256, 0, s>>>(
            h3, 16, d15, d15, d15, wl, bl, heads.p, IC, IH, IH, IH, R, CONV_ACT_NONE, nullptr, 0);
        return static_cast<int>(cudaGetLastError());
      });
    }
    if (c.mode == 3) {  // head-major [nH][R][D] logits / (mu, logvar) to the caller
      fwd.add([=](cudaStream_t s) {
        for (int k = 0; k < (conv ? 1 : nH); ++k)
          unpad_kernel<<<grid_for(static_cast<size_t>(R) * D), 256, 0, s>>>(
              heads.p + k * Dp, heads.ld, bd->heads_out + static_cast<size_t>(k) * R * D, R, D, 1.0f);
        return static_cast<int>(cudaGetLastError());
      });
      return fwd.error;
    }
    if (c.mode == 2) {
      fwd.add([=](cudaStream_t s) {
        iws_loglik_kernel<<<R, 128, 0, s>>>(heads.p, heads.ld, Dp, bd->x, D, nz, toy ? 0 : 1, lw0, wbuf);
        iws_logmeanexp_kernel<<<B, 256, 0, s>>>(wbuf, nz, bd->iws_out, bd->iws_total);
        return static_cast<int>(cudaGetLastError());
      });
      return fwd.error;
    }
    fwd.add([=](cudaStream_t s) {
      if (toy)
        gauss_elbo_kernel<<<R, 128, 0, s>>>(heads.p, heads.ld, Dp, bd->x, D, zbuf.p, zbuf.ld, zd, nz, bd->beta,
                                            bd->inv_rows, bd->sums, dheads.p, dheads.ld, bd->beta_dev);
      else
        bern_elbo_kernel<<<R, 256, 0, s>>>(heads.p, heads.ld, bd->x, D, zbuf.p, zbuf.ld, zd, nz, bd->beta,
                                           bd->inv_rows, bd->sums, dheads.p, dheads.ld, bd->beta_dev);
      if (bd->heads_out)  // head-major [nH][R][D]
        for (int k = 0; k < nH; ++k)
          unpad_kernel<<<grid_for(static_cast<size_t>(R) * D), 256, 0, s>>>(
              heads.p + k * Dp, heads.ld, bd->heads_out + static_cast<size_t>(k) * R * D, R, D, 1.0f);
      return static_cast<int>(cudaGetLastError());
    });

    // ================================================================= backward
    auto tn1 = [&](Plan& pl, const Mat& X, const Mat& Y, float* dst, int ldo) {
      GemmTNDesc t;
      t.X0 = X.p; t.ldx0 = X.ld; t.Y0 = Y.p; t.ldy0 = Y.ld;
      t.M = X.cols; t.N = Y.cols; t.K = X.rows; t.out = dst; t.ldo = ldo;
      t.scale = 1.0f; t.beta = 1.0f; t.workspace = tn_ws; t.workspace_bytes = tn_bytes;
      pl.tn(t);
    };
    auto colsum_op = [&](Plan& pl, const Mat& X, float* dst) {
      pl.add([=](cudaStream_t s) {
        // tall-skinny operands (pixel-major conv gradients: 1e5 rows x 16..32 columns) need many row slabs
        const int col_blocks = (X.cols + 31) / 32;
        int slabs = X.rows >= 2048 ? 32 : (X.rows + 63) / 64;
        if (X.rows >= 16384 && col_blocks * slabs < 592) slabs = std::min(X.rows / 256, 592 / col_blocks);
        dim3 grid(col_blocks, slabs);
        colsum_kernel<<<grid, 256, 0, s>>>(X.p, X.ld, X.rows, X.cols, dst, 1.0f);
        return static_cast<int>(cudaGetLastError());
      });
    };
    // ---- decoder part (skipped when loss_scale == 0)
    for (int k = 0; k < nH && !conv; ++k) {
      const Mat dh = dheads.cols_from(k * Dp, D);
      tn1(bwd_dec, dh, Dh[c.n_dec - 1].hi(), G(iH(k)), h);
      colsum_op(bwd_dec, dh, G(iH(k) + 1));
    }
    if (!conv) {
      GemmNTDesc g = nt_desc(dheads, Hw.T, dD[c.n_dec - 1], DACT);
      set_aux1(g, Dh[c.n_dec - 1].hi());
      g.colsum = G(iD(c.n_dec - 1) + 1);
      bwd_dec.nt(g);
    } else if (cg) {
      // logit deconv <- deconv2 <- deconv1: dcols = im2col(d out) ; dW += in^T . dcols ; d in_pre = (dcols . W^T) * act'(in)
      // (act'(0) = 0 on the pad row / column of p2); the bias gradients are the column sums of the d pre-activations
      const int gin[3] = {s8, d8, d15}, gout[3] = {d7, d15, c.img_h}, gpitch[3] = {d8, d15, c.img_h};
      const int gco[3] = {32, 16, c.img_c};
      const Mat dlast = dD[c.n_dec - 1];
      float* gb_last = G(iD(c.n_dec - 1) + 1);
      {
        float* gbl = G(iH(2) + 1);
        const int P2 = c.img_h * c.img_h, ICc = c.img_c;
        bwd_dec.add([=](cudaStream_t s) {
          chan_sum_nchw_kernel<<<dim3(ICc, 128), 256, 0, s>>>(dheads.p, dheads.ld, R, ICc, P2, gbl);
          return static_cast<int>(cudaGetLastError());
        });
      }
      for (int k = 2; k >= 0; --k) {
        const Mat dc = dcolb[k], din = ddp[k];
        const Mat up = k < 2 ? ddp[k + 1] : Mat();
        const int Ho = gin[k], Hout = gout[k], pitch = gpitch[k], Co = gco[k];
        bwd_dec.add([=](cudaStream_t s) {
          const ImgView src = k < 2 ? img_pix(up.p, up.ld, Co, Hout, pitch) : img_nchw(dheads.p, Co, Hout);
          launch_im2col5s2(src, Ho, 1.0f, 0.0f, dc.p, dc.ld, 0, R, s);
          return static_cast<int>(cudaGetLastError());
        });
        tn1(bwd_dec, dp_[k].hi(), dc, G(iH(k)), dck[k]);
        GemmNTDesc g = nt_desc(dc, dWr[k], din, DACT);
        set_aux1(g, dp_[k].hi());
        if (k > 0) g.colsum = G(iH(k - 1) + 1);  // bias of the deconv below = sum of this d pre-activation
        bwd_dec.nt(g);
      }
      {
        const Mat d1 = ddp[0];
        bwd_dec.add([=](cudaStream_t s) {  // back to the NCHW-flattened rows of decode.fc's output
          chw_pix_permute_kernel<false><<<grid_for(static_cast<size_t>(R) * feat), 256, 0, s>>>(
              d1.p, d1.ld, 0, dlast.p, dlast.ld, 0, R, 32, s8 * s8, 1);
          dim3 grid((feat + 31) / 32, R >= 2048 ? 32 : (R + 63) / 64);
          colsum_kernel<<<grid, 256, 0, s>>>(dlast.p, dlast.ld, R, feat, gb_last, 1.0f);
          return static_cast<int>(cudaGetLastError());
        });
      }
    } else {
      float *gwd1 = G(iH(0)), *gbd1 = G(iH(0) + 1), *gwd2 = G(iH(1)), *gbd2 = G(iH(1) + 1), *gwl = G(iH(2)), *gbl = G(iH(2) + 1);
      const float *wd1 = P(iH(0)), *wd2 = P(iH(1)), *wl = P(iH(2));
      const Mat dlast = dD[c.n_dec - 1];
      float* gb_last = G(iD(c.n_dec - 1) + 1);
      bwd_dec.add([=](cudaStream_t s) {
        // logit deconv: d h3pre = conv(dlogit) * act'(h3) ; dW = <dlogit (conv input role), h3 (conv output role)>
        conv5s2_fwd_kernel<<<grid_for(static_cast<size_t>(R) * 16 * d15 * d15), 256, 0, s>>>(
            dheads.p, IC, IH, IH, IH, wl, nullptr, dh3, 16, d15, d15, d15, R, 1.0f, 0.0f, CACT, h3);
        conv5s2_bwd_weight_kernel<<<16 * IC, 256, 0, s>>>(dheads.p, IC, IH, IH, IH, h3, 16, d15, d15, d15, gwl, nullptr, R, 1.0f, 0.0f);
        chan_sum_kernel<<<IC, 256, 0, s>>>(dheads.p, IC, IH, IH, IH, R, gbl);
        // deconv2
        conv5s2_fwd_kernel<<<grid_for(static_cast<size_t>(R) * 32 * d8 * d8), 256, 0, s>>>(
            dh3, 16, d15, d15, d15, wd2, nullptr, dh2, 32, d8, d8, d8, R, 1.0f, 0.0f, CACT, h2p);
        conv5s2_bwd_weight_kernel<<<32 * 16, 256, 0, s>>>(dh3, 16, d15, d15, d15, h2p, 32, d8, d8, d8, gwd2, nullptr, R, 1.0f, 0.0f);
        chan_sum_kernel<<<16, 256, 0, s>>>(dh3, 16, d15, d15, d15, R, gbd2);
        // deconv1 (its input h1 is the activated output of decode.fc: multiply by act' there)
        conv5s2_fwd_kernel<<<grid_for(static_cast<size_t>(R) * 32 * s8 * s8), 256, 0, s>>>(
            dh2, 32, d7, d8, d8, wd1, nullptr, dh1, 32, s8, s8, s8, R, 1.0f, 0.0f, CONV_ACT_NONE, nullptr);
        conv5s2_bwd_weight_kernel<<<32 * 32, 256, 0, s>>>(dh2, 32, d7, d8, d8, h1, 32, s8, s8, s8, gwd1, nullptr, R, 1.0f, 0.0f);
        chan_sum_kernel<<<32, 256, 0, s>>>(dh2, 32, d7, d8, d8, R, gbd1);
        mul_dact_kernel<<<grid_for(static_cast<size_t>(R) * feat), 256, 0, s>>>(dh1, h1, dlast.p, static_cast<size_t>(R) * feat, CACT, 1);
        dim3 grid((feat + 31) / 32, R >= 2048 ? 32 : (R + 63) / 64);
        colsum_kernel<<<grid, 256, 0, s>>>(dlast.p, dlast.ld, R, feat, gb_last, 1.0f);
        return static_cast<int>(cudaGetLastError());
      });
    }
    for (int l = c.n_dec - 1; l >= 1; --l) {
      GemmNTDesc g = nt_desc(dD[l], Dw[l].T, dD[l - 1], DACT);
      set_aux1(g, Dh[l - 1].hi());
      g.colsum = G(iD(l - 1) + 1);
      bwd_dec.nt(g);
    }
    for (int l = c.n_dec - 1; l >= 1; --l) tn1(bwd_dec, dD[l], Dh[l - 1].hi(), G(iD(l)), dwid(l - 1));
    tn1(bwd_dec, dD[0], zp.hi(), G(iD(0)), zd);
    {
      GemmNTDesc g = nt_desc(dD[0], Dw[0].T, dzdec, EPI_LINEAR);
      g.round_out = 0;
      bwd_dec.nt(g);
    }
    // ---- encoder part
    bwd_enc.add([=](cudaStream_t s) {
      dz_total_kernel<<<grid_for(static_cast<size_t>(R) * zd), 256, 0, s>>>(
          dzdec.p, dzdec.ld, zbuf.p, zbuf.ld, bd->gz, bd->gz_scale, bd->loss_scale, bd->beta * bd->inv_rows, dzt.p, dzt.ld, R, zd, bd->beta_dev,
          bd->inv_rows);
      return static_cast<int>(cudaGetLastError());
    });
    {
      const Mat lastin = toy ? Fh[c.n_fc - 1].hi() : Fh[c.n_fc - 1].hi();
      tn1(bwd_enc, dzt, lastin, G(iF(c.n_fc)), toy ? h + n : h);
      colsum_op(bwd_enc, dzt, G(iF(c.n_fc) + 1));
    }
    {
      // d hid_last = (dz . W_z[:, :h]) * act'(hid_last)
      GemmNTDesc g = nt_desc(dzt, Fw[c.n_fc].T.rows_from(0, h), dF[c.n_fc - 1], DACT);
      Pair hv = Fh[c.n_fc - 1];
      hv.w = h;
      set_aux1(g, hv.hi());
      g.colsum = G(iF(c.n_fc - 1) + 1);
      bwd_enc.nt(g);
    }
    for (int l = c.n_fc - 1; l >= 1; --l) {
      GemmNTDesc g = nt_desc(dF[l], Fw[l].T.rows_from(0, h), dF[l - 1], DACT);
      Pair hv = Fh[l - 1];
      hv.w = h;
      set_aux1(g, hv.hi());
      g.colsum = G(iF(l - 1) + 1);
      bwd_enc.nt(g);
    }
    for (int l = c.n_fc - 1; l >= 1; --l) tn1(bwd_enc, dF[l], Fh[l - 1].hi(), G(iF(l)), toy ? h + n : h);
    // layer 0: noise half over R rows, input half through the per-data-row sum
    tn1(bwd_enc, dF[0], epsp.hi(), G(iF(0)) ? G(iF(0)) + feat : nullptr, ld0);
    {
      const Mat d0 = dF[0];
      bwd_enc.add([=](cudaStream_t s) {
        group_sum_kernel<<<B, 256, 0, s>>>(d0.p, d0.ld, gsum0.p, gsum0.ld, B, nz, h, 1);
        return static_cast<int>(cudaGetLastError());
      });
    }
    tn1(bwd_enc, gsum0, inp_pair.hi(), G(iF(0)), ld0);
    {
      GemmNTDesc g = nt_desc(gsum0, F0i.T, dI[c.n_inp - 1], DACT);
      set_aux1(g, inp_pair.hi());
      if (!conv) g.colsum = G(iI(c.n_inp - 1) + 1);  // conv: the bias is per channel, summed below
      if (conv) g.round_out = 0;
      bwd_enc.nt(g);
    }
    if (cg) {
      // back through conv3, conv2, conv1 (dI[last] = d pre-activation of conv3's output in NCHW order):
      // dW_l += dA_l^T . col_l ; db_l = column sums ; dA_{l-1} = col2im(dA_l . W_l) * act'(a_{l-1})
      {
        const Mat dl = dI[c.n_inp - 1], d3 = eda[2];
        bwd_enc.add([=](cudaStream_t s) {
          chw_pix_permute_kernel<true><<<grid_for(static_cast<size_t>(B) * feat), 256, 0, s>>>(
              dl.p, dl.ld, 0, d3.p, d3.ld, 0, B, 32, s8 * s8, 1);
          return static_cast<int>(cudaGetLastError());
        });
      }
      for (int l = 2; l >= 0; --l) {
        tn1(bwd_enc, eda[l], colp[l].hi(), G(iI(l)), eci[l] * 25);
        colsum_op(bwd_enc, eda[l], G(iI(l) + 1));
        if (l == 0) break;
        GemmNTDesc g = nt_desc(eda[l], Cw[l].T, ecol[l], EPI_LINEAR);
        bwd_enc.nt(g);
        const Mat dc = ecol[l], dprev = eda[l - 1], uprev = act_[l - 1];
        const int Ho = eho[l], Hi = eho[l - 1], Ci = eci[l];
        bwd_enc.add([=](cudaStream_t s) {
          col2im5s2_kernel<<<grid_for(static_cast<size_t>(B) * Hi * Hi * Ci), 256, 0, s>>>(
              dc.p, dc.ld, Ho, img_pix(dprev.p, dprev.ld, Ci, Hi, Hi), nullptr, CACT, C2I_MUL_DACT, 0, uprev.p, B);
          return static_cast<int>(cudaGetLastError());
        });
      }
      return fwd.error ? fwd.error : (bwd_dec.error ? bwd_dec.error : bwd_enc.error);
    }
    if (conv) {
      // back through conv3, conv2, conv1 (dI[last] = d pre-activation of conv3's output, [B, 32*s8*s8])
      const float* dc3 = dI[c.n_inp - 1].p;
      const int ldc3 = dI[c.n_inp - 1].ld;
      float *gw1 = G(iI(0)), *gb1 = G(iI(0) + 1), *gw2 = G(iI(1)), *gb2 = G(iI(1) + 1), *gw3 = G(iI(2)), *gb3 = G(iI(2) + 1);
      const float *w2 = P(iI(1)), *w3 = P(iI(2));
      float* dc2 = ws.floats(static_cast<size_t>(B) * 32 * s4 * s4);
      float* dc1 = ws.floats(static_cast<size_t>(B) * 16 * s2 * s2);
      float* dc3c = ws.floats(static_cast<size_t>(B) * feat);  // contiguous copy (dI pitch may be padded)
      bwd_enc.add([=](cudaStream_t s) {
        unpad_kernel<<<grid_for(static_cast<size_t>(B) * feat), 256, 0, s>>>(dc3, ldc3, dc3c, B, feat, 1.0f);
        conv5s2_bwd_weight_kernel<<<32 * 32, 256, 0, s>>>(c2, 32, s4, s4, s4, dc3c, 32, s8, s8, s8, gw3, gb3, B, 1.0f, 0.0f);
        conv5s2_bwd_data_kernel<<<grid_for(static_cast<size_t>(B) * 32 * s4 * s4), 256, 0, s>>>(
            dc3c, 32, s8, s8, s8, w3, nullptr, dc2, 32, s4, s4, s4, B, CACT, c2, 1);
        conv5s2_bwd_weight_kernel<<<32 * 16, 256, 0, s>>>(c1, 16, s2, s2, s2, dc2, 32, s4, s4, s4, gw2, gb2, B, 1.0f, 0.0f);
        conv5s2_bwd_data_kernel<<<grid_for(static_cast<size_t>(B) * 16 * s2 * s2), 256, 0, s>>>(
            dc2, 32, s4, s4, s4, w2, nullptr, dc1, 16, s2, s2, s2, B, CACT, c1, 1);
        conv5s2_bwd_weight_kernel<<<16 * IC, 256, 0, s>>>(bd->x, IC, IH, IH, IH, dc1, 16, s2, s2, s2, gw1, gb1, B, 2.0f, -1.0f);
        return static_cast<int>(cudaGetLastError());
      });
      return fwd.error ? fwd.error : (bwd_dec.error ? bwd_dec.error : bwd_enc.error);
    }
    for (int l = c.n_inp - 1; l >= 1; --l) {
      GemmNTDesc g = nt_desc(dI[l], Iw[l].T, dI[l - 1], DACT);
      set_aux1(g, I[l - 1].hi());
      g.colsum = G(iI(l - 1) + 1);
      bwd_enc.nt(g);
    }
    for (int l = c.n_inp - 1; l >= 1; --l) tn1(bwd_enc, dI[l], I[l - 1].hi(), G(iI(l)), h);
    tn1(bwd_enc, dI[0], xin.hi(), G(iI(0)), D);
    return fwd.error ? fwd.error : (bwd_dec.error ? bwd_dec.error : bwd_enc.error);
  }
};

}  // namespace ardae
