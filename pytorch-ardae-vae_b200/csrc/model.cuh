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
#include "cdae.cuh"

namespace ardae {

struct ModelConfig {
  int kind = 0;   // 0 toy (gaussian decoder, noise concatenated at every fc layer), 1 mnist
  int D = 0, n = 0, h = 0, zd = 0;
  int n_inp = 0, n_fc = 0, n_dec = 0;  // number of linears in inp_encode, hidden fc layers, decode.main
  int act = 1;    // 0 relu, 1 softplus
  int B = 0, nz = 1;
  int mode = 0;   // 0 encode only, 1 forward + backward, 2 IWS log-likelihood (forward only)
};

struct ModelBindings {
  const float* x = nullptr;      // [B, D]
  const float* noise = nullptr;  // [R, n] or null (= zeros: encode(std=0))
  float* z_out = nullptr;        // [R, zd]
  float* heads_out = nullptr;    // [R, D] logits | [R, 2D] mu,logvar  (optional)
  float* sums = nullptr;         // [3] loss, recon, prior (device; overwritten)
  float beta = 1.0f;
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
  int* status = nullptr;         // set to 1+image if a covariance is not positive definite
};

struct ModelPlan {
  ModelConfig cfg;
  Plan fwd, bwd_dec, bwd_enc;
  Workspace ws;
  ModelBindings bind;
  DeriveList derive;
  size_t tn_need = 0;

  int n_heads() const { return cfg.kind == 0 ? 2 : 1; }
  int ntensors() const { return 2 * (cfg.n_inp + cfg.n_fc + 1 + cfg.n_dec + n_heads()); }

  int build(float* const* params, float* const* grads) {
    const ModelConfig& c = cfg;
    const int D = c.D, n = c.n, h = c.h, zd = c.zd, B = c.B, nz = c.nz, R = B * nz;
    if (D <= 0 || n <= 0 || h <= 0 || zd <= 0 || B <= 0 || nz <= 0 || c.n_inp < 1 || c.n_fc < 1 || c.n_dec < 1)
      return fail(-2, "model: bad config");
    if (c.kind != 0 && c.kind != 1) return fail(-2, "model: kind must be 0 (toy) or 1 (mnist)");
    const bool dry = ws.dry, train = c.mode == 1, dec = c.mode != 0;
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
    for (int l = 0; l < c.n_inp; ++l) {
      const int in = l == 0 ? D : h;
      Iw[l] = derive.add(ws, P(iI(l)), h, in, in, true, train && l > 0);
    }
    // fc layer 0: [h, h+n] = [W_inp | W_noise]
    const int ld0 = h + n;
    W3 F0i = derive.add(ws, P(iF(0)), h, h, ld0, true, train);
    W3 F0n = derive.add(ws, P(iF(0)) ? P(iF(0)) + h : nullptr, h, n, ld0, true, false);
    // later fc layers (toy: input is [hid | eps]; mnist has none) and the final fc.fc
    std::vector<W3> Fw(c.n_fc + 1);
    for (int l = 1; l <= c.n_fc; ++l) {
      const int out = l == c.n_fc ? zd : h;
      const int in = toy ? h + n : h;
      Fw[l] = derive.add(ws, P(iF(l)), out, in, in, true, train);
    }
    for (int l = 0; l < c.n_dec; ++l) {
      const int in = l == 0 ? zd : h;
      if (dec) Dw[l] = derive.add(ws, P(iD(l)), h, in, in, true, train);
    }
    // heads: combined [nH*D, h] forward operand and [h, nH*D] transpose
    W3 Hw;
    Hw.in = h; Hw.out = nH * D; Hw.kp = round_up(h, 32);
    if (dec) {
      Hw.b3 = Mat(ws.floats(static_cast<size_t>(nH) * D * 3 * Hw.kp), nH * D, 3 * Hw.kp, 3 * Hw.kp);
      if (train) Hw.T = ws.mat(h, nH * Dp);
      for (int k = 0; k < nH; ++k) {
        DeriveItem it;
        std::memset(&it, 0, sizeof(it));
        it.src = P(iH(k)); it.rows = D; it.cols = h; it.src_ld = h;
        it.dst3 = dry ? nullptr : Hw.b3.p + static_cast<size_t>(k) * D * Hw.b3.ld;
        it.kp = Hw.kp; it.ld3 = Hw.b3.ld;
        it.dstT = (dry || !train) ? nullptr : Hw.T.p + k * Dp; it.ldT = Hw.T.ld;
        it.first_block = derive.blocks;
        it.tiles_x = (h + 31) / 32;
        derive.blocks += it.tiles_x * ((D + 31) / 32);
        derive.host.push_back(it);
      }
    }
    int rc = derive.emit(ws, fwd);
    if (rc) return rc;

    // ---- buffers
    Pair xin = make_pair(ws, B, D);
    std::vector<Pair> I(c.n_inp), Fh(c.n_fc + 1), Dh(c.n_dec);
    for (int l = 0; l < c.n_inp; ++l) I[l] = make_pair(ws, B, h);
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
    if (dec) {
      for (int l = 0; l < c.n_dec; ++l) Dh[l] = make_pair(ws, R, h);
      heads = ws.mat(R, nH * Dp);
    }
    if (c.mode == 2) {
      lw0 = ws.floats(R);
      wbuf = ws.floats(R);
    }
    if (train) {
      dheads = ws.mat(R, nH * Dp);
      dzdec = ws.mat(R, zd);
      dzt = ws.mat(R, zd);
      for (int l = 0; l < c.n_dec; ++l) dD[l] = ws.mat(R, h);
      for (int l = 0; l < c.n_fc; ++l) dF[l] = ws.mat(R, h);
      for (int l = 0; l < c.n_inp; ++l) dI[l] = ws.mat(B, h);
      gsum0 = ws.mat(B, h);
      if (dry) {
        const int shapes[8][3] = {{h, h + n, R}, {h, h, R}, {zd, h + n, R}, {nH * D, h, R}, {h, zd, R},
                                  {h, D, B}, {h, h, B}, {h, n, R}};
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
    fwd.add([=](cudaStream_t s) {
      split2d_kernel<<<grid_for(static_cast<size_t>(B) * D), 256, 0, s>>>(
          bd->x, D, xin.buf.p, xin.buf.ld, B, D, xin.kp, toy ? 1.0f : 2.0f, toy ? 0.0f : -1.0f);
      if (bd->noise != nullptr) {
        split2d_kernel<<<grid_for(static_cast<size_t>(R) * n), 256, 0, s>>>(
            bd->noise, n, epsp.buf.p, epsp.buf.ld, R, n, epsp.kp, 1.0f, 0.0f);
      } else {
        ARDAE_CUDA_OK(cudaMemsetAsync(epsp.buf.p, 0, sizeof(float) * R * epsp.buf.ld, s));
      }
      if (bd->sums) ARDAE_CUDA_OK(cudaMemsetAsync(bd->sums, 0, 3 * sizeof(float), s));
      return static_cast<int>(cudaGetLastError());
    });
    // inp_encode on the B data rows
    for (int l = 0; l < c.n_inp; ++l) {
      GemmNTDesc g = nt3_desc(l == 0 ? xin : I[l - 1], Iw[l], I[l], ACT);
      g.bias = P(iI(l) + 1);
      fwd.nt(g);
    }
    // fc layer 0: input half once per data row (rowbias0), noise half over the R rows
    {
      GemmNTDesc g = nt3_desc_plain(I[c.n_inp - 1], F0i, rowbias0, EPI_LINEAR);
      g.bias = P(iF(0) + 1);
      fwd.nt(g);
    }
    {
      Pair out = Fh[0];
      out.w = h;  // write only the hid columns of the (possibly wider) concat buffer
      GemmNTDesc g = nt3_desc(epsp, F0n, out, ACT);
      g.group_bias = rowbias0.p; g.group = nz; g.ldg = rowbias0.ld;
      fwd.nt(g);
    }
    if (toy) {
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
    for (int l = 1; l < c.n_fc; ++l) {
      Pair out = Fh[l];
      out.w = h;
      GemmNTDesc g = nt3_desc(Fh[l - 1], Fw[l], out, ACT);
      g.bias = P(iF(l) + 1);
      fwd.nt(g);
    }
    {  // z = fc.fc([hid | eps])  (plain fp32 to the user buffer layout, and as a tf32 pair)
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
    // decoder
    for (int l = 0; l < c.n_dec; ++l) {
      GemmNTDesc g = nt3_desc(l == 0 ? zp : Dh[l - 1], Dw[l], Dh[l], ACT);
      g.bias = P(iD(l) + 1);
      fwd.nt(g);
    }
    for (int k = 0; k < nH; ++k) {
      W3 hk = Hw;
      hk.out = D;
      hk.b3 = Hw.b3.rows_from(k * D, D);
      GemmNTDesc g = nt3_desc_plain(Dh[c.n_dec - 1], hk, heads.cols_from(k * Dp, D), EPI_LINEAR);
      g.bias = P(iH(k) + 1);
      fwd.nt(g);
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
                                            bd->inv_rows, bd->sums, dheads.p, dheads.ld);
      else
        bern_elbo_kernel<<<R, 256, 0, s>>>(heads.p, heads.ld, bd->x, D, zbuf.p, zbuf.ld, zd, nz, bd->beta,
                                           bd->inv_rows, bd->sums, dheads.p, dheads.ld);
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
        dim3 grid((X.cols + 31) / 32, X.rows >= 2048 ? 32 : (X.rows + 63) / 64);
        colsum_kernel<<<grid, 256, 0, s>>>(X.p, X.ld, X.rows, X.cols, dst, 1.0f);
        return static_cast<int>(cudaGetLastError());
      });
    };
    // ---- decoder part (skipped when loss_scale == 0)
    for (int k = 0; k < nH; ++k) {
      const Mat dh = dheads.cols_from(k * Dp, D);
      tn1(bwd_dec, dh, Dh[c.n_dec - 1].hi(), G(iH(k)), h);
      colsum_op(bwd_dec, dh, G(iH(k) + 1));
    }
    {
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
    // ---- encoder part
    bwd_enc.add([=](cudaStream_t s) {
      dz_total_kernel<<<grid_for(static_cast<size_t>(R) * zd), 256, 0, s>>>(
          dzdec.p, dzdec.ld, zbuf.p, zbuf.ld, bd->gz, bd->gz_scale, bd->loss_scale, bd->beta * bd->inv_rows, dzt.p, dzt.ld, R, zd);
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
    tn1(bwd_enc, dF[0], epsp.hi(), G(iF(0)) ? G(iF(0)) + h : nullptr, ld0);
    {
      const Mat d0 = dF[0];
      bwd_enc.add([=](cudaStream_t s) {
        group_sum_kernel<<<B, 256, 0, s>>>(d0.p, d0.ld, gsum0.p, gsum0.ld, B, nz, h, 1);
        return static_cast<int>(cudaGetLastError());
      });
    }
    tn1(bwd_enc, gsum0, I[c.n_inp - 1].hi(), G(iF(0)), ld0);
    {
      GemmNTDesc g = nt_desc(gsum0, F0i.T, dI[c.n_inp - 1], DACT);
      set_aux1(g, I[c.n_inp - 1].hi());
      g.colsum = G(iI(c.n_inp - 1) + 1);
      bwd_enc.nt(g);
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
