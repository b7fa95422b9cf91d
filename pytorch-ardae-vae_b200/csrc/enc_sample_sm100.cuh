// Fused N-row sampling pass of the noise-concat MLP encoder (MNIST-kind: one hidden fc layer):
//     z[r, :] = W_z . softplus( W_eps . eps[r, :] + rowbias[r / nz, :] ) + b_z            r < R = B * nz
// (models/ivae/mnist.py:161-165 `_forward_all` on the nz samples of every data row).  The
// input half of fc layer 0 enters as the per-data-row bias `rowbias` = W_inp . inp(x_b) + b (computed once per data
// row by the B-row plan), so the only per-sample work is the noise half.  One CTA carries a 128-row tile through both
// layers; the [128, h] hidden activation lives in shared memory only (the per-layer path wrote and re-read it as a
// fp32 (hi, lo) pair: ~1 GB of HBM traffic per update at config 2).
//
// Arithmetic: fp32-accurate products on the fp16 tensor pipe (three-product scheme of chain_s3h_sm100.cuh) with
// exact power-of-two scaling: eps * 2^4, hidden * 2^-2, weights * 2^4.
// Shapes: noise_dim <= 128, h <= 320, z_dim <= 32.  Warps: 0 weight TMA producer, 1 MMA issuer + TMEM owner,
// 4..19 sixteen epilogue warps (four groups; group g owns the 32-column chunks c % 4 == g).
#pragma once
#include "chain_s3h_sm100.cuh"

namespace ardae {

struct alignas(64) EncSampleParams {
  CUtensorMap tmW1;        // fp16 [h rows, 2*k1] = [W_eps hi | lo] * 2^4, K-major, box {64, NH}
  CUtensorMap tmW2;        // fp16 [z_dim rows, 2*k2] = [W_z hi | lo] * 2^4, box {64, 32}
  const float* noise;      // [R, n] fp32
  const float* rowbias;    // [B, ldb]: input half + bias of fc layer 0
  const float* bias2;      // [z_dim]
  float* z_out;            // [R, z_dim]
  int ldb;
  int R, nz, n, h, zd;
  int k1;                  // noise_dim rounded up to 64 (64 or 128)
  int k2;                  // h rounded up to 64
  int NH;                  // columns per N-half of layer 1: (h rounded up to 32) / 2 rounded up to 16
};

struct EncSampleConfig {
  static constexpr int kTile = kBlockM * 64 * 2;       // 16 KB A tile (128 rows x 64 fp16)
  static constexpr int kWStage = 160 * 64 * 2;         // 20 KB: one k-block of one N-half of layer 1 (<= 160 rows)
  static constexpr int kNumWStages = 3;
  static constexpr int kOffA = kNumWStages * kWStage;  // hi tiles [5] then lo tiles [5] (layer 1 uses the first 2 of each)
  static constexpr int kDataBytes = kOffA + 10 * kTile;
  static constexpr int kSmemBytes = kDataBytes + 1024 + 256;
  static constexpr int kGroups = 4;
  static constexpr int kThreads = 128 + kGroups * 128;
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

constexpr float kEncEpsScale = 16.0f;     // 2^4
constexpr float kEncHidScale = 0.25f;     // 2^-2
constexpr float kEncAcc1 = 1.0f / 256.0f; // 1 / (eps scale * weight scale)
constexpr float kEncAcc2 = 0.25f;         // 1 / (hidden scale * weight scale)

__global__ void __launch_bounds__(EncSampleConfig::kThreads, 1)
enc_sample_kernel(const __grid_constant__ EncSampleParams p) {
  using Cfg = EncSampleConfig;
  constexpr int G = Cfg::kGroups, NW = Cfg::kNumWStages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* w_full = reinterpret_cast<uint64_t*>(smem + Cfg::kDataBytes);
  uint64_t* w_empty = w_full + NW;
  uint64_t* a_ready = w_empty + NW;    // [2] A operand of layer 1 / layer 2 is in shared memory
  uint64_t* acc_full = a_ready + 2;    // [2] accumulator of layer 1 / layer 2 complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kBlockM;
  const int NH = p.NH;
  const int nkb1 = p.k1 >> 6, nkb2 = p.k2 >> 6;
  uint8_t* hi_tiles = smem + Cfg::kOffA;
  uint8_t* lo_tiles = hi_tiles + 5 * Cfg::kTile;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.tmW1);
    ptx::prefetch_tmap(&p.tmW2);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < NW; ++s) {
        ptx::mbar_init(&w_full[s], 1);
        ptx::mbar_init(&w_empty[s], 1);
      }
      for (int i = 0; i < 2; ++i) {
        ptx::mbar_init(&a_ready[i], 4 * G);
        ptx::mbar_init(&acc_full[i], 1);
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
  const uint32_t acc1 = *tmem_slot;        // columns [0, 2*NH)
  const uint32_t acc2 = acc1 + 384;        // columns [384, 416)

  if (warp == 0) {
    // ------------------------------------------------------------ weight producer
    if (ptx::elect_one()) {
      int it = 0;
      // layer 1: per N-half, k-block kb of W_eps hi, then of W_eps lo (columns [0,k1) and [k1,2*k1) of the operand)
      for (int h = 0; h < 2; ++h)
        for (int j = 0; j < 2 * nkb1; ++j, ++it) {
          const int s = it % NW;
          ptx::mbar_wait(&w_empty[s], ((it / NW) & 1) ^ 1);
          ptx::mbar_expect_tx(&w_full[s], static_cast<uint32_t>(NH) * 64 * 2);
          ptx::tma_load_2d(smem + s * Cfg::kWStage, &p.tmW1, &w_full[s], ((j & 1) ? p.k1 : 0) + (j >> 1) * 64, h * NH);
        }
      // layer 2: k-block kb of W_z hi, then of W_z lo
      for (int j = 0; j < 2 * nkb2; ++j, ++it) {
        const int s = it % NW;
        ptx::mbar_wait(&w_empty[s], ((it / NW) & 1) ^ 1);
        ptx::mbar_expect_tx(&w_full[s], 32u * 64 * 2);
        ptx::tma_load_2d(smem + s * Cfg::kWStage, &p.tmW2, &w_full[s], ((j & 1) ? p.k2 : 0) + (j >> 1) * 64, 0);
      }
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
      // ---- layer 1: acc1[:, h*NH : (h+1)*NH] = eps . W_eps^T (three products)
      ptx::mbar_wait(&a_ready[0], 0);
      ptx::tc_fence_after();
      const uint32_t idesc1 = ptx::make_idesc_f16(kBlockM, NH);
      for (int h = 0; h < 2; ++h)
        for (int kb = 0; kb < nkb1; ++kb) {
          mma_block(acc1 + h * NH, idesc1, kb, kb == 0, false);  // hi . Whi + lo . Whi
          mma_block(acc1 + h * NH, idesc1, kb, false, true);     // hi . Wlo
        }
      ptx::umma_commit(&acc_full[0]);
      // ---- layer 2: acc2 = hidden . W_z^T
      ptx::mbar_wait(&a_ready[1], 0);
      ptx::tc_fence_after();
      const uint32_t idesc2 = ptx::make_idesc_f16(kBlockM, 32);
      for (int kb = 0; kb < nkb2; ++kb) {
        mma_block(acc2, idesc2, kb, kb == 0, false);
        mma_block(acc2, idesc2, kb, false, true);
      }
      ptx::umma_commit(&acc_full[1]);
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

    // ---- phase A: noise tile -> scaled fp16 (hi, lo) A operand of layer 1 (columns >= n are zero)
    if (g < 2 * nkb1) {
      const int c = g;  // 32-column chunks of the (64-padded) noise width
      const float* src = p.noise + static_cast<size_t>(row_ok ? m : 0) * p.n;
      const bool vec = (p.n % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.noise) & 15) == 0);
#pragma unroll 1
      for (int sub = 0; sub < 2; ++sub) {
        const int c0 = c * 32 + sub * 16;
        float sv[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) sv[j] = 0.0f;
        if (row_ok) {
          if (vec && c0 + 16 <= p.n) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(src + c0) + q);
              sv[q * 4 + 0] = t.x * kEncEpsScale; sv[q * 4 + 1] = t.y * kEncEpsScale;
              sv[q * 4 + 2] = t.z * kEncEpsScale; sv[q * 4 + 3] = t.w * kEncEpsScale;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c0 + j < p.n) sv[j] = __ldg(src + c0 + j) * kEncEpsScale;
          }
        }
        store_sub(c, sub, sv);
      }
    }
    ptx::fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&a_ready[0]);

    // ---- phase C: hidden = softplus(acc1 / 256 + rowbias[b]) -> A operand of layer 2
    ptx::mbar_wait(&acc_full[0], 0);  // (all layer-1 MMAs have retired: their A tiles may be overwritten)
    ptx::tc_fence_after();
    {
      const float* brow = p.rowbias + static_cast<size_t>((row_ok ? m : 0) / p.nz) * p.ldb;
      const int nch = 2 * nkb2;  // 32-column chunks of the (64-padded) hidden width
#pragma unroll 1
      for (int c = g; c < nch; c += G) {
#pragma unroll 1
        for (int sub = 0; sub < 2; ++sub) {
          const int ns = c * 32 + sub * 16;
          float sv[16];
          if (ns < 2 * NH) {  // a real accumulator column range (NH is a multiple of 16)
            uint32_t accu[16];
            tmem_ld_32x16(acc1 + lane_addr + ns, accu);
            float add[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) add[j] = (ns + j < p.h) ? __ldg(brow + ns + j) : 0.0f;
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j)
              sv[j] = (ns + j < p.h) ? softplus_fast(fmaf(__uint_as_float(accu[j]), kEncAcc1, add[j])) * kEncHidScale : 0.0f;
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
      if (lane == 0) ptx::mbar_arrive(&a_ready[1]);
    }

    // ---- phase E: z = acc2 / 4 + b_z (fp32 rows straight to global); one group suffices
    if (g == 0) {
      ptx::mbar_wait(&acc_full[1], 0);
      ptx::tc_fence_after();
      uint32_t accu[32];
      ptx::tmem_ld_32x32(acc2 + lane_addr, accu);
      ptx::tmem_ld_wait();
      if (row_ok) {
        float* dst = p.z_out + static_cast<size_t>(m) * p.zd;
        if (p.zd == 32 && (reinterpret_cast<uintptr_t>(p.z_out) & 15) == 0) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            reinterpret_cast<float4*>(dst)[q] =
                make_float4(fmaf(__uint_as_float(accu[q * 4 + 0]), kEncAcc2, __ldg(p.bias2 + q * 4 + 0)),
                            fmaf(__uint_as_float(accu[q * 4 + 1]), kEncAcc2, __ldg(p.bias2 + q * 4 + 1)),
                            fmaf(__uint_as_float(accu[q * 4 + 2]), kEncAcc2, __ldg(p.bias2 + q * 4 + 2)),
                            fmaf(__uint_as_float(accu[q * 4 + 3]), kEncAcc2, __ldg(p.bias2 + q * 4 + 3)));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < p.zd) dst[j] = fmaf(__uint_as_float(accu[j]), kEncAcc2, __ldg(p.bias2 + j));
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(acc1, 512);
  }
}

// ------------------------------------------------------------------ host side
struct EncSampleDesc {
  const uint16_t* W1 = nullptr; int ldw1 = 0;  // fp16 [h, 2*k1], k1 = noise_dim rounded up to 64
  const uint16_t* W2 = nullptr; int ldw2 = 0;  // fp16 [zd, 2*k2]
  const float* noise = nullptr;                // [R, n]
  const float* rowbias = nullptr; int ldb = 0; // [B, ldb]
  const float* bias2 = nullptr;
  float* z_out = nullptr;
  int R = 0, nz = 1, n = 0, h = 0, zd = 0;
};

inline bool enc_sample_supported(int n, int h, int zd) {
  static int env = -1;
  if (env < 0) {
    const char* e = std::getenv("ARDAE_ENC_FUSED");
    env = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return env != 0 && n >= 1 && n <= 128 && h >= 1 && h <= 320 && zd >= 1 && zd <= 32;
}

struct PreparedEncSample {
  EncSampleParams params;
  dim3 grid;
};

inline int prepare_enc_sample(const EncSampleDesc& d, PreparedEncSample* out) {
  if (!enc_sample_supported(d.n, d.h, d.zd) || d.R <= 0 || d.nz <= 0) return fail(-2, "enc_sample: unsupported shape");
  if (!d.W1 || !d.W2 || !d.rowbias || !d.bias2 || !d.z_out) return fail(-2, "enc_sample: missing pointer");
  PreparedEncSample pr;
  std::memset(&pr.params, 0, sizeof(pr.params));
  EncSampleParams& p = pr.params;
  p.k1 = (d.n + 63) / 64 * 64;
  p.k2 = (d.h + 63) / 64 * 64;
  p.NH = (((d.h + 31) / 32 * 32) / 2 + 15) / 16 * 16;
  int rc;
  if ((rc = encode_tmap_2d_bf16(&p.tmW1, d.W1, 2 * p.k1, d.h, d.ldw1, 64, p.NH, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_DATA_TYPE_FLOAT16)))
    return rc;
  if ((rc = encode_tmap_2d_bf16(&p.tmW2, d.W2, 2 * p.k2, d.zd, d.ldw2, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_DATA_TYPE_FLOAT16)))
    return rc;
  p.noise = d.noise; p.rowbias = d.rowbias; p.ldb = d.ldb; p.bias2 = d.bias2; p.z_out = d.z_out;
  p.R = d.R; p.nz = d.nz; p.n = d.n; p.h = d.h; p.zd = d.zd;
  pr.grid = dim3((d.R + kBlockM - 1) / kBlockM, 1, 1);
  ARDAE_CUDA_OK(cudaFuncSetAttribute(reinterpret_cast<const void*>(&enc_sample_kernel),
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, EncSampleConfig::kSmemBytes));
  *out = pr;
  return 0;
}

// `noise` may change per call (the plan binds it at launch time)
inline int launch_prepared_enc_sample(PreparedEncSample pr, const float* noise, float* z_out, cudaStream_t stream) {
  pr.params.noise = noise;
  pr.params.z_out = z_out;
  void* args[1] = {&pr.params};
  ARDAE_CUDA_OK(cudaLaunchKernel(reinterpret_cast<const void*>(&enc_sample_kernel), pr.grid,
                                 dim3(EncSampleConfig::kThreads), args, EncSampleConfig::kSmemBytes, stream));
  return 0;
}

}  // namespace ardae
