// Fused Bernoulli decoder + BCE row sum of the IWS evaluator (MNIST-kind decoder: z -> h -> ... -> h -> D logits):
//     w[r] = lw0[r] - sum_px [ softplus(l[r, px]) - x[r / S, px] * l[r, px] ],     l = W_logit . dec_main(z'[r]) + b
// (models/ivae/mnist.py:421-430 inside logprob_w_cov_gaussian_posterior; utils/vae.py:21-30).
// One CTA carries a 128-row tile through the hidden layers with the activation in shared memory, then walks the D
// logits in blocks of <= 160 columns: every block is reduced into per-row BCE partial sums in the epilogue, so neither
// the hidden activations nor the logits (15.7 MB per image at 5000 samples) ever reach HBM -- the per-layer path wrote
// and re-read both.  The logit accumulator is double-buffered in TMEM: the MMAs of block j+1 run under the epilogue
// of block j.
//
// Arithmetic: fp32-accurate products on the fp16 tensor pipe (three-product scheme of chain_s3h_sm100.cuh), exact
// power-of-two scaling: z' * 2^2, hidden * 2^-2, weights * 2^4.
// Shapes: z_dim <= 64, h <= 320, up to 4 hidden layers, any D.  Warps as in enc_sample_sm100.cuh.
#pragma once
#include "enc_sample_sm100.cuh"

namespace ardae {

constexpr int kDecMaxHidden = 4;

struct alignas(64) DecIwsParams {
  CUtensorMap tmW[kDecMaxHidden];  // fp16 [h rows, 2*kin16(l)] = [W hi | lo] * 2^4, box {64, NH}
  CUtensorMap tmWlogit;            // fp16 [D rows, 2*k2], box {64, NL}
  const float* bias[kDecMaxHidden];
  const float* bias_logit;         // [D]
  const float* z_hi;               // z' as a tf32 (hi, lo) pair [R, ldz] (iws_moments_kernel output)
  const float* z_lo;
  const float* x;                  // [B, D] in {0,1} (any float)
  const float* lw0;                // [R] log prior - log q of the sample
  float* w;                        // [R] log importance weight
  int ldz;
  int R, S, zd, h, D, nhid;
  int k0;                          // z_dim rounded up to 64
  int k2;                          // h rounded up to 64
  int NH;                          // columns per N-half of a hidden layer
  int NL;                          // logit block width (multiple of 32, <= 160)
};

struct DecIwsConfig {
  static constexpr int kTile = kBlockM * 64 * 2;
  static constexpr int kWStage = 160 * 64 * 2;         // 20 KB
  static constexpr int kNumWStages = 3;
  static constexpr int kOffA = kNumWStages * kWStage;
  static constexpr int kOffPart = kOffA + 10 * kTile;  // [4][128] per-group BCE partial sums
  static constexpr int kDataBytes = kOffPart + 4 * 128 * 4;
  static constexpr int kSmemBytes = kDataBytes + 1024 + 256;
  static constexpr int kGroups = 4;
  static constexpr int kThreads = 128 + kGroups * 128;
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

constexpr float kDecZScale = 4.0f;          // 2^2
constexpr float kDecHidScale = 0.25f;       // 2^-2
constexpr float kDecAcc0 = 1.0f / 64.0f;    // 1 / (z scale * weight scale)
constexpr float kDecAcc = 0.25f;            // 1 / (hidden scale * weight scale)

__global__ void __launch_bounds__(DecIwsConfig::kThreads, 1)
dec_iws_kernel(const __grid_constant__ DecIwsParams p) {
  using Cfg = DecIwsConfig;
  constexpr int G = Cfg::kGroups, NW = Cfg::kNumWStages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* w_full = reinterpret_cast<uint64_t*>(smem + Cfg::kDataBytes);
  uint64_t* w_empty = w_full + NW;
  uint64_t* a_ready = w_empty + NW;     // [1] the A operand of the next layer is in shared memory (phase = layer)
  uint64_t* acc_full = a_ready + 1;     // [1] hidden-layer accumulator complete (phase = layer)
  uint64_t* lg_full = acc_full + 1;     // [2] logit block buffer complete
  uint64_t* lg_empty = lg_full + 2;     // [2] logit block buffer drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lg_empty + 2);
  float* part = reinterpret_cast<float*>(smem + Cfg::kOffPart);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kBlockM;
  const int NH = p.NH, NL = p.NL;
  const int nkb2 = p.k2 >> 6;
  const int nblk = (p.D + NL - 1) / NL;  // logit blocks
  uint8_t* hi_tiles = smem + Cfg::kOffA;
  uint8_t* lo_tiles = hi_tiles + 5 * Cfg::kTile;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.tmW[0]);
    ptx::prefetch_tmap(&p.tmWlogit);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < NW; ++s) {
        ptx::mbar_init(&w_full[s], 1);
        ptx::mbar_init(&w_empty[s], 1);
      }
      ptx::mbar_init(a_ready, 4 * G);
      ptx::mbar_init(acc_full, 1);
      for (int b = 0; b < 2; ++b) {
        ptx::mbar_init(&lg_full[b], 1);
        ptx::mbar_init(&lg_empty[b], 4 * G);
      }
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t acc = *tmem_slot;  // hidden layers: columns [0, 2*NH); logit blocks: [0, NL) and [256, 256 + NL)

  if (warp == 0) {
    // ------------------------------------------------------------ weight producer
    if (ptx::elect_one()) {
      int it = 0;
      auto load = [&](const CUtensorMap* tm, int kc, int row, uint32_t bytes) {
        const int s = it % NW;
        ptx::mbar_wait(&w_empty[s], ((it / NW) & 1) ^ 1);
        ptx::mbar_expect_tx(&w_full[s], bytes);
        ptx::tma_load_2d(smem + s * Cfg::kWStage, tm, &w_full[s], kc, row);
        ++it;
      };
      for (int l = 0; l < p.nhid; ++l) {
        const int kin = l == 0 ? p.k0 : p.k2;
        for (int h = 0; h < 2; ++h)
          for (int j = 0; j < 2 * (kin >> 6); ++j)  // k-block kb of W hi, then of W lo
            load(&p.tmW[l], ((j & 1) ? kin : 0) + (j >> 1) * 64, h * NH, static_cast<uint32_t>(NH) * 64 * 2);
      }
      for (int b = 0; b < nblk; ++b)
        for (int j = 0; j < 2 * nkb2; ++j)
          load(&p.tmWlogit, ((j & 1) ? p.k2 : 0) + (j >> 1) * 64, b * NL, static_cast<uint32_t>(NL) * 64 * 2);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (ptx::elect_one()) {
      int it = 0;
      auto mma_block = [&](uint32_t d_t, uint32_t idesc, int kb, bool first, bool wlo) {
        const int s = it % NW;
        ptx::mbar_wait(&w_full[s], (it / NW) & 1);
        ptx::tc_fence_after();
        const uint32_t b_addr = ptx::smem_u32(smem + s * Cfg::kWStage);
        const uint32_t h_addr = ptx::smem_u32(hi_tiles + kb * Cfg::kTile);
        const uint32_t l_addr = ptx::smem_u32(lo_tiles + kb * Cfg::kTile);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t adesc = ptx::make_smem_desc_sw128(h_addr + k * 32, 0, 1024);
          const uint64_t bdesc = ptx::make_smem_desc_sw128(b_addr + k * 32, 0, 1024);
          ptx::umma_f16(d_t, adesc, bdesc, idesc, (first && k == 0) ? 0u : 1u);
        }
        if (!wlo) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adesc = ptx::make_smem_desc_sw128(l_addr + k * 32, 0, 1024);
            const uint64_t bdesc = ptx::make_smem_desc_sw128(b_addr + k * 32, 0, 1024);
            ptx::umma_f16(d_t, adesc, bdesc, idesc, 1u);
          }
        }
        ptx::umma_commit(&w_empty[s]);
        ++it;
      };
      const uint32_t idesc_h = ptx::make_idesc_f16(kBlockM, NH);
      for (int l = 0; l < p.nhid; ++l) {
        const int nkb = (l == 0 ? p.k0 : p.k2) >> 6;
        ptx::mbar_wait(a_ready, l & 1);
        ptx::tc_fence_after();
        for (int h = 0; h < 2; ++h)
          for (int kb = 0; kb < nkb; ++kb) {
            mma_block(acc + h * NH, idesc_h, kb, kb == 0, false);
            mma_block(acc + h * NH, idesc_h, kb, false, true);
          }
        ptx::umma_commit(acc_full);
      }
      // logit blocks: double-buffered accumulator
      ptx::mbar_wait(a_ready, p.nhid & 1);
      ptx::tc_fence_after();
      const uint32_t idesc_l = ptx::make_idesc_f16(kBlockM, NL);
      for (int b = 0; b < nblk; ++b) {
        const int buf = b & 1;
        ptx::mbar_wait(&lg_empty[buf], ((b >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_t = acc + buf * 256;
        for (int kb = 0; kb < nkb2; ++kb) {
          mma_block(d_t, idesc_l, kb, kb == 0, false);
          mma_block(d_t, idesc_l, kb, false, true);
        }
        ptx::umma_commit(&lg_full[buf]);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue warps
    const int g = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int m = m0 + r;
    const bool row_ok = m < p.R;
    const int swz = r & 7;
    const uint32_t row_off = static_cast<uint32_t>(r) * 128u;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;

    auto store_sub = [&](int c, int sub, const float (&sv)[16]) {
      const uint32_t base = static_cast<uint32_t>((c >> 1) * Cfg::kTile) + row_off;
      const uint32_t hb = ptx::smem_u32(hi_tiles) + base, lb = ptx::smem_u32(lo_tiles) + base;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) split_f16x2<true>(sv[q * 8 + 2 * i], sv[q * 8 + 2 * i + 1], hw[i], lw[i]);
        const uint32_t off = static_cast<uint32_t>((((c & 1) * 4 + sub * 2 + q) ^ swz) << 4);
        sts128u(hb + off, hw[0], hw[1], hw[2], hw[3]);
        sts128u(lb + off, lw[0], lw[1], lw[2], lw[3]);
      }
    };

    // ---- z' = hi + lo -> scaled fp16 pair (columns >= z_dim of the k-block(s) are zero)
    for (int c = g; c < 2 * (p.k0 >> 6); c += G) {
#pragma unroll 1
      for (int sub = 0; sub < 2; ++sub) {
        const int c0 = c * 32 + sub * 16;
        float sv[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) sv[j] = 0.0f;
        if (row_ok) {
          const float* sh = p.z_hi + static_cast<size_t>(m) * p.ldz + c0;
          const float* sl = p.z_lo + static_cast<size_t>(m) * p.ldz + c0;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (c0 + j < p.zd) sv[j] = (__ldg(sh + j) + __ldg(sl + j)) * kDecZScale;
        }
        store_sub(c, sub, sv);
      }
    }
    ptx::fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(a_ready);

    // ---- hidden layers: h_l = softplus(acc * scale + b_l) -> A operand of the next layer
#pragma unroll 1
    for (int l = 0; l < p.nhid; ++l) {
      ptx::mbar_wait(acc_full, l & 1);  // all MMAs of the layer have retired: the A tiles may be overwritten
      ptx::tc_fence_after();
      const float as = l == 0 ? kDecAcc0 : kDecAcc;
      const float* bl = p.bias[l];
#pragma unroll 1
      for (int c = g; c < 2 * nkb2; c += G) {
#pragma unroll 1
        for (int sub = 0; sub < 2; ++sub) {
          const int ns = c * 32 + sub * 16;
          float sv[16];
          if (ns < 2 * NH) {
            uint32_t accu[16];
            tmem_ld_32x16(acc + lane_addr + ns, accu);
            float add[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) add[j] = (ns + j < p.h) ? __ldg(bl + ns + j) : 0.0f;
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j)
              sv[j] = (ns + j < p.h) ? softplus_fast(fmaf(__uint_as_float(accu[j]), as, add[j])) * kDecHidScale : 0.0f;
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) sv[j] = 0.0f;
          }
          store_sub(c, sub, sv);
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(a_ready);
    }

    // ---- logit blocks: BCE partial sums of this thread's row over the chunks its group owns
    float bce = 0.0f;
    const float* xr = p.x + static_cast<size_t>((row_ok ? m : 0) / p.S) * p.D;
#pragma unroll 1
    for (int b = 0; b < nblk; ++b) {
      const int buf = b & 1;
      ptx::mbar_wait(&lg_full[buf], (b >> 1) & 1);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int c = g; c * 32 < NL; c += G) {
#pragma unroll 1
        for (int sub = 0; sub < 2; ++sub) {
          const int px0 = b * NL + c * 32 + sub * 16;
          if (px0 >= p.D) break;
          uint32_t accu[16];
          tmem_ld_32x16(acc + buf * 256 + lane_addr + c * 32 + sub * 16, accu);
          float bb[16], xv[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const bool ok = px0 + j < p.D;
            bb[j] = ok ? __ldg(p.bias_logit + px0 + j) : 0.0f;
            xv[j] = ok ? __ldg(xr + px0 + j) : 0.0f;
          }
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float lg = fmaf(__uint_as_float(accu[j]), kDecAcc, bb[j]);
            const float t = softplus_fast(lg) - xv[j] * lg;
            bce += (px0 + j < p.D) ? t : 0.0f;
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&lg_empty[buf]);
    }
    part[g * 128 + r] = bce;
    ptx::named_bar_sync(1, G * 128);
    if (g == 0 && row_ok) p.w[m] = p.lw0[m] - (part[r] + part[128 + r] + part[256 + r] + part[384 + r]);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(acc, 512);
  }
}

// ------------------------------------------------------------------ host side
struct DecIwsDesc {
  const uint16_t* W[kDecMaxHidden] = {nullptr, nullptr, nullptr, nullptr}; int ldw[kDecMaxHidden] = {0, 0, 0, 0};
  const float* bias[kDecMaxHidden] = {nullptr, nullptr, nullptr, nullptr};
  const uint16_t* Wlogit = nullptr; int ldwl = 0;
  const float* bias_logit = nullptr;
  const float* z_hi = nullptr; const float* z_lo = nullptr; int ldz = 0;
  const float* lw0 = nullptr;
  float* w = nullptr;
  int R = 0, S = 1, zd = 0, h = 0, D = 0, nhid = 0;
};

inline bool dec_iws_supported(int zd, int h, int nhid) {
  static int env = -1;
  if (env < 0) {
    const char* e = std::getenv("ARDAE_IWS_FUSED");
    env = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return env != 0 && zd >= 1 && zd <= 64 && h >= 1 && h <= 320 && nhid >= 1 && nhid <= kDecMaxHidden;
}

struct PreparedDecIws {
  DecIwsParams params;
  dim3 grid;
};

inline int prepare_dec_iws(const DecIwsDesc& d, PreparedDecIws* out) {
  if (!dec_iws_supported(d.zd, d.h, d.nhid) || d.R <= 0 || d.S <= 0 || d.D <= 0) return fail(-2, "dec_iws: unsupported shape");
  if (!d.Wlogit || !d.bias_logit || !d.z_hi || !d.z_lo || !d.lw0 || !d.w) return fail(-2, "dec_iws: missing pointer");
  PreparedDecIws pr;
  std::memset(&pr.params, 0, sizeof(pr.params));
  DecIwsParams& p = pr.params;
  p.k0 = (d.zd + 63) / 64 * 64;
  p.k2 = (d.h + 63) / 64 * 64;
  p.NH = (((d.h + 31) / 32 * 32) / 2 + 15) / 16 * 16;
  p.NL = d.D >= 160 ? 160 : (d.D + 31) / 32 * 32;
  int rc;
  for (int l = 0; l < d.nhid; ++l) {
    if (!d.W[l] || !d.bias[l]) return fail(-2, "dec_iws: missing layer operand");
    const int kin = l == 0 ? p.k0 : p.k2;
    if ((rc = encode_tmap_2d_bf16(&p.tmW[l], d.W[l], 2 * kin, d.h, d.ldw[l], 64, p.NH, CU_TENSOR_MAP_SWIZZLE_128B,
                                  CU_TENSOR_MAP_DATA_TYPE_FLOAT16)))
      return rc;
    p.bias[l] = d.bias[l];
  }
  if ((rc = encode_tmap_2d_bf16(&p.tmWlogit, d.Wlogit, 2 * p.k2, d.D, d.ldwl, 64, p.NL, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_DATA_TYPE_FLOAT16)))
    return rc;
  p.bias_logit = d.bias_logit; p.z_hi = d.z_hi; p.z_lo = d.z_lo; p.ldz = d.ldz; p.lw0 = d.lw0; p.w = d.w;
  p.R = d.R; p.S = d.S; p.zd = d.zd; p.h = d.h; p.D = d.D; p.nhid = d.nhid;
  pr.grid = dim3((d.R + kBlockM - 1) / kBlockM, 1, 1);
  ARDAE_CUDA_OK(cudaFuncSetAttribute(reinterpret_cast<const void*>(&dec_iws_kernel),
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, DecIwsConfig::kSmemBytes));
  *out = pr;
  return 0;
}

inline int launch_prepared_dec_iws(PreparedDecIws pr, const float* x, cudaStream_t stream) {
  pr.params.x = x;
  void* args[1] = {&pr.params};
  ARDAE_CUDA_OK(cudaLaunchKernel(reinterpret_cast<const void*>(&dec_iws_kernel), pr.grid, dim3(DecIwsConfig::kThreads),
                                 args, DecIwsConfig::kSmemBytes, stream));
  return 0;
}

}  // namespace ardae
