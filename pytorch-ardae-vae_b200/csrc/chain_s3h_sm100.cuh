// Primal forward sweep of the CDAE as ONE launch with fp32-accurate products on the fp16 tensor pipe ("3xFP16").
//
// The tf32 variant (chain16_sm100.cuh, CHAIN_SOFTPLUS3) splits activations and weights into tf32 (hi, lo) pairs and
// issues hi.Whi + lo.Whi + hi.Wlo as kind::tf32 MMAs; it is tensor-bound at ~0.9 of the tf32 peak.  fp16 carries the
// same 11 significant bits as tf32, and kind::f16 runs at twice the rate, so the same three-product scheme on fp16
// pairs halves the tensor time.  fp16's narrow exponent is handled by exact power-of-two scaling:
//   activations are stored as (hi, lo) = split(x * 2^-8)   (|x| up to 1.6e7; |x| below 0.016 keeps an absolute
//                                                           accuracy of 8e-6, far below the 3-product error of the
//                                                           1e3..1e5 pre-activations this path exists for)
//   weights     are stored as (hi, lo) = split(w * 2^4)     (|w| up to 4e3)
//   the fp32 accumulator is multiplied by 2^4 in the epilogue.  Values beyond the range saturate (finite).
// Both operand halves live in shared memory (K-major 128 x 64 fp16 tiles, SWIZZLE_128B): 4 + 4 tiles = 128 KB next to
// the 6 x 16 KB weight ring; TMEM holds only the accumulator.  Pipelining, barriers and warp roles are those of
// chain16_kernel (two N-halves per layer, epilogue of half 0 under the MMAs of half 1, next layer starting on
// finished k-blocks).  Output: bf16 spill rows written straight from registers.
// Reference math: models/graddae/mlp.py:400-430 (inp_encode + neglogprob forward), SURVEY.md 8a-3 sweep (1).
#pragma once
#include <cuda_fp16.h>

#include <type_traits>

#include "chain16_sm100.cuh"

namespace ardae {

constexpr float kS3hActScale = 0.00390625f;  // 2^-8
constexpr float kS3hWScale = 16.0f;          // 2^4
constexpr float kS3hAccScale = 16.0f;        // 1 / (act * w scale)
constexpr int kS3hBlockK = 64;               // fp16 elements per k-block (128-byte rows)

struct alignas(64) ChainS3hLayerParams {
  CUtensorMap tmW;          // fp16 [H rows, 2*kin16] = [Whi | Wlo] (scaled by 2^4), K-major, box {64, H/2}
  const float* bias;
  const float* group_bias;  // row m adds group_bias[(m / group) * ldg + n]
  const float* col_vec;     // adds row_scale[m] * col_vec[n]
  uint16_t* out16;          // bf16 spill [M, H]
  int ld_out16;
  int group, ldg;
  int kin16;                // input width rounded up to 64 (<= H; < H only for layer 0)
};

struct alignas(64) ChainS3hParams {
  const float* a0_hi;       // initial activation as a tf32 (hi, lo) pair [M, kin0] fp32 (x~ of the perturbation prologue)
  const float* a0_lo;
  int a0_ld;
  int kin0;                 // true width of the initial activation (multiple of 32)
  const float* row_scale;   // [M] (sigma) or null
  int M, H, nlayers;
  int vec_ok;
  // optional second output of the LAST layer: delta_L[m, n] = -w_o[n] * (1 - exp(-v_L[m, n])) as bf16 -- the start of
  // the score sweep (graddae/mlp.py:426-441: d E / d pre-activation of the last hidden layer), from the bf16-rounded v_L
  const float* wo;
  uint16_t* delta16;
  int ld_delta16;
  ChainS3hLayerParams layer[kChainMaxLayers];
};

struct ChainS3hConfig {
  static constexpr int kWStage = 128 * kS3hBlockK * 2;  // 16 KB: one k-block of one N-half
  static constexpr int kNumWStages = 6;
  static constexpr int kTile = kBlockM * kS3hBlockK * 2;  // 16 KB A tile
  static constexpr int kOffHi = kNumWStages * kWStage;
  static constexpr int kOffLo = kOffHi + 4 * kTile;
  static constexpr int kDataBytes = kOffLo + 4 * kTile;
  static constexpr int kSmemBytes = kDataBytes + 1024 + 256;
  static constexpr int kGroups = 4;  // epilogue groups of four warps: the sweep is paced by epilogue instruction issue
  static constexpr int kThreads = 128 + kGroups * 128;
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

namespace ptx {
// kind::f16 with fp16 operands (a_format = b_format = 0), fp32 accumulate, both K-major
__host__ __device__ constexpr uint32_t make_idesc_f16(int m, int n) {
  return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
}  // namespace ptx

// TMEM -> registers: 32 lanes x 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// add[j] += scale * src[j], j < 16 (16-byte aligned broadcast loads)
__device__ __forceinline__ void add_cols16(const float* __restrict__ src, float scale, float (&add)[16]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(src) + q);
    add[q * 4 + 0] = fmaf(scale, t.x, add[q * 4 + 0]);
    add[q * 4 + 1] = fmaf(scale, t.y, add[q * 4 + 1]);
    add[q * 4 + 2] = fmaf(scale, t.z, add[q * 4 + 2]);
    add[q * 4 + 3] = fmaf(scale, t.w, add[q * 4 + 3]);
  }
}

// (hi, lo) fp16 split of two scaled values, packed as f16x2 words (element 0 in the low half)
template <bool SIGNED>
__device__ __forceinline__ void split_f16x2(float s0, float s1, uint32_t& hi, uint32_t& lo) {
  s0 = fminf(s0, 65000.0f);  // saturate (finite) instead of overflowing to inf
  s1 = fminf(s1, 65000.0f);
  if (SIGNED) {
    s0 = fmaxf(s0, -65000.0f);
    s1 = fmaxf(s1, -65000.0f);
  }
  const __half2 h = __floats2half2_rn(s0, s1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(s0 - hf.x, s1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

__global__ void __launch_bounds__(ChainS3hConfig::kThreads, 1)
chain_s3h_kernel(const __grid_constant__ ChainS3hParams p) {
  using Cfg = ChainS3hConfig;
  constexpr int G = Cfg::kGroups, NW = Cfg::kNumWStages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* w_full = reinterpret_cast<uint64_t*>(smem + Cfg::kDataBytes);
  uint64_t* w_empty = w_full + NW;
  uint64_t* acc_full = w_empty + NW;   // [2]
  uint64_t* a_ready = acc_full + 2;    // [2]
  uint64_t* kfree = a_ready + 2;       // [2] the current layer no longer reads A k-block kb (kb < NBk/2)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kfree + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kBlockM;
  const int H = p.H;
  const int NB = H >> 5;     // 32-column epilogue chunks
  const int NB0 = NB >> 1;   // chunks per N-half
  const int HH = H >> 1;
  const int NBk = H >> 6;    // 64-wide k-blocks of an H-wide activation (H % 128 == 0: 2 or 4)
  const int khalf = NBk >> 1;
  const int nl = p.nlayers;
  uint8_t* hi_tiles = smem + Cfg::kOffHi;
  uint8_t* lo_tiles = smem + Cfg::kOffLo;

  if (warp == 0 && lane == 0) ptx::prefetch_tmap(&p.layer[0].tmW);
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < NW; ++s) {
        ptx::mbar_init(&w_full[s], 1);
        ptx::mbar_init(&w_empty[s], 1);
      }
      for (int h = 0; h < 2; ++h) {
        ptx::mbar_init(&acc_full[h], 1);
        ptx::mbar_init(&a_ready[h], 4 * G);
        ptx::mbar_init(&kfree[h], 1);
      }
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t acc_t = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ weight producer
    if (ptx::elect_one()) {
      const uint32_t wbytes = static_cast<uint32_t>(HH) * kS3hBlockK * 2;
      int it = 0;
      for (int l = 0; l < nl; ++l) {
        const ChainS3hLayerParams& L = p.layer[l];
        const int nkb = L.kin16 >> 6;
        for (int h = 0; h < 2; ++h) {
          for (int j = 0; j < 2 * nkb; ++j, ++it) {  // k-block kb of Whi, then of Wlo
            const int s = it % NW;
            ptx::mbar_wait(&w_empty[s], ((it / NW) & 1) ^ 1);
            const int kc = ((j & 1) ? L.kin16 : 0) + (j >> 1) * kS3hBlockK;
            ptx::mbar_expect_tx(&w_full[s], wbytes);
            ptx::tma_load_2d(smem + s * Cfg::kWStage, &L.tmW, &w_full[s], kc, h * HH);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::make_idesc_f16(kBlockM, HH);
      int it = 0;
      for (int l = 0; l < nl; ++l) {
        const int nkb = p.layer[l].kin16 >> 6;
        for (int h = 0; h < 2; ++h) {
          const uint32_t d_t = acc_t + h * HH;
          for (int kb = 0; kb < nkb; ++kb) {
            if (h == 0 && kb == 0) {
              ptx::mbar_wait(&a_ready[0], l & 1);  // A k-blocks [0, NBk/2) written, accumulator half 0 drained
              ptx::tc_fence_after();
            }
            if (h == 0 && kb == khalf) {
              ptx::mbar_wait(&a_ready[1], l & 1);
              ptx::tc_fence_after();
            }
            const uint32_t h_addr = ptx::smem_u32(hi_tiles + kb * Cfg::kTile);
            const uint32_t l_addr = ptx::smem_u32(lo_tiles + kb * Cfg::kTile);
            {  // Whi: hi . Whi + lo . Whi
              const int s = it % NW;
              ptx::mbar_wait(&w_full[s], (it / NW) & 1);
              ptx::tc_fence_after();
              const uint32_t b_addr = ptx::smem_u32(smem + s * Cfg::kWStage);
#pragma unroll
              for (int k = 0; k < kS3hBlockK / 16; ++k) {
                const uint64_t adesc = ptx::make_smem_desc_sw128(h_addr + k * 32, 0, 1024);
                const uint64_t bdesc = ptx::make_smem_desc_sw128(b_addr + k * 32, 0, 1024);
                ptx::umma_f16(d_t, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
              }
#pragma unroll
              for (int k = 0; k < kS3hBlockK / 16; ++k) {
                const uint64_t adesc = ptx::make_smem_desc_sw128(l_addr + k * 32, 0, 1024);
                const uint64_t bdesc = ptx::make_smem_desc_sw128(b_addr + k * 32, 0, 1024);
                ptx::umma_f16(d_t, adesc, bdesc, idesc, 1u);
              }
              ptx::umma_commit(&w_empty[s]);
              ++it;
            }
            {  // Wlo: hi . Wlo
              const int s = it % NW;
              ptx::mbar_wait(&w_full[s], (it / NW) & 1);
              ptx::tc_fence_after();
              const uint32_t b_addr = ptx::smem_u32(smem + s * Cfg::kWStage);
#pragma unroll
              for (int k = 0; k < kS3hBlockK / 16; ++k) {
                const uint64_t adesc = ptx::make_smem_desc_sw128(h_addr + k * 32, 0, 1024);
                const uint64_t bdesc = ptx::make_smem_desc_sw128(b_addr + k * 32, 0, 1024);
                ptx::umma_f16(d_t, adesc, bdesc, idesc, 1u);
              }
              ptx::umma_commit(&w_empty[s]);
              ++it;
            }
            if (h == 1 && kb < khalf) ptx::umma_commit(&kfree[kb]);  // half-0 epilogue may overwrite k-block kb
          }
          if (h == 0 && nkb <= khalf) {  // short first layer: consume this layer's a_ready[1] phase as well
            ptx::mbar_wait(&a_ready[1], l & 1);
            ptx::tc_fence_after();
          }
          if (h == 1)
            for (int kb = nkb; kb < khalf; ++kb) ptx::umma_commit(&kfree[kb]);
          ptx::umma_commit(&acc_full[h]);
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue warps
    const int g = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int m = m0 + r;
    const bool row_ok = m < p.M;
    const int swz = r & 7;
    const uint32_t row_off = static_cast<uint32_t>(r) * 128u;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const float rs = (p.row_scale != nullptr && row_ok) ? p.row_scale[m] : 0.0f;

    // writes 16 scaled values (columns [c*32 + sub*16, +16) of this thread's row) as fp16 (hi, lo) into the A tiles
    auto store_sub = [&](int c, int sub, const float (&sv)[16], auto sgn) {
      const uint32_t base = static_cast<uint32_t>((c >> 1) * Cfg::kTile) + row_off;
      const uint32_t hb = ptx::smem_u32(hi_tiles) + base, lb = ptx::smem_u32(lo_tiles) + base;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) split_f16x2<decltype(sgn)::value>(sv[q * 8 + 2 * i], sv[q * 8 + 2 * i + 1], hw[i], lw[i]);
        const uint32_t off = static_cast<uint32_t>((((c & 1) * 4 + sub * 2 + q) ^ swz) << 4);
        sts128u(hb + off, hw[0], hw[1], hw[2], hw[3]);
        sts128u(lb + off, lw[0], lw[1], lw[2], lw[3]);
      }
    };

    // ---- pseudo-layer -1: x~ = hi + lo (fp32) -> scaled fp16 pair; pad chunks of the first k-blocks are zero
    {
      const int nch0 = (p.layer[0].kin16 >> 6) * 2;
      for (int c = g; c < nch0; c += G) {
#pragma unroll 1
        for (int sub = 0; sub < 2; ++sub) {
          float sv[16];
          if (c * 32 < p.kin0 && row_ok) {
            const float* sh = p.a0_hi + static_cast<size_t>(m) * p.a0_ld + c * 32 + sub * 16;
            const float* sl = p.a0_lo + static_cast<size_t>(m) * p.a0_ld + c * 32 + sub * 16;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 a = __ldg(reinterpret_cast<const float4*>(sh) + q);
              const float4 b = __ldg(reinterpret_cast<const float4*>(sl) + q);
              sv[q * 4 + 0] = (a.x + b.x) * kS3hActScale; sv[q * 4 + 1] = (a.y + b.y) * kS3hActScale;
              sv[q * 4 + 2] = (a.z + b.z) * kS3hActScale; sv[q * 4 + 3] = (a.w + b.w) * kS3hActScale;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) sv[j] = 0.0f;
          }
          store_sub(c, sub, sv, std::true_type());
        }
      }
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(&a_ready[0]);
        ptx::mbar_arrive(&a_ready[1]);
      }
    }

#pragma unroll 1
    for (int l = 0; l < nl; ++l) {
      const ChainS3hLayerParams& L = p.layer[l];
      const bool last = (l == nl - 1);
      const float* gb_row = (L.group_bias != nullptr)
                                ? L.group_bias + static_cast<size_t>((row_ok ? m : 0) / L.group) * L.ldg
                                : nullptr;
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        ptx::mbar_wait(&acc_full[h], l & 1);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int c = h * NB0 + g; c < (h + 1) * NB0; c += G) {
          const int nc = c * 32;
          if (h == 0) {  // the half-1 MMAs of this layer still read A k-block c/2 until kfree fires
            ptx::mbar_wait(&kfree[c >> 1], l & 1);
            ptx::tc_fence_after();
          }
#pragma unroll 1
          for (int sub = 0; sub < 2; ++sub) {
            const int ns = nc + sub * 16;
            uint32_t accu[16];
            tmem_ld_32x16(acc_t + lane_addr + ns, accu);
            float add[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) add[j] = 0.0f;
            if (L.bias != nullptr) add_cols16(L.bias + ns, 1.0f, add);
            if (gb_row != nullptr) add_cols16(gb_row + ns, 1.0f, add);
            if (L.col_vec != nullptr) add_cols16(L.col_vec + ns, rs, add);
            ptx::tmem_ld_wait();
            float sv[16];
            uint32_t ow[8];
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              const float r0 = softplus_fast(fmaf(__uint_as_float(accu[j]), kS3hAccScale, add[j]));
              const float r1 = softplus_fast(fmaf(__uint_as_float(accu[j + 1]), kS3hAccScale, add[j + 1]));
              sv[j] = r0 * kS3hActScale;
              sv[j + 1] = r1 * kS3hActScale;
              ow[j >> 1] = pack_bf16x2(r0, r1);
            }
            if (!last) store_sub(c, sub, sv, std::false_type());
            if (row_ok) {  // bf16 spill: one 32-byte sector of this thread's row (256-bit store)
              uint16_t* dst = L.out16 + static_cast<size_t>(m) * L.ld_out16 + ns;
              asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(ow[0]), "r"(ow[1]),
                           "r"(ow[2]), "r"(ow[3]), "r"(ow[4]), "r"(ow[5]), "r"(ow[6]), "r"(ow[7])
                           : "memory");
            }
            if (last && p.delta16 != nullptr && row_ok) {  // delta_L from the rounded v_L (what the later sweeps re-read)
              uint32_t dw[8];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 w4 = __ldg(reinterpret_cast<const float4*>(p.wo + ns) + q);
                const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
                float o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const uint32_t pk = ow[q * 2 + (e >> 1)];
                  const float u = __uint_as_float((e & 1) ? (pk & 0xFFFF0000u) : (pk << 16));
                  const float sg = (u < 0.01f) ? u * (1.0f - u * (0.5f - u * (1.0f / 6.0f))) : 1.0f - __expf(-u);
                  o[e] = -wv[e] * sg;
                }
                dw[q * 2] = pack_bf16x2(o[0], o[1]);
                dw[q * 2 + 1] = pack_bf16x2(o[2], o[3]);
              }
              uint16_t* dd = p.delta16 + static_cast<size_t>(m) * p.ld_delta16 + ns;
              asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dd), "r"(dw[0]), "r"(dw[1]),
                           "r"(dw[2]), "r"(dw[3]), "r"(dw[4]), "r"(dw[5]), "r"(dw[6]), "r"(dw[7])
                           : "memory");
            }
          }
        }
        // A chunks of this half written (shared-memory writes fenced for the async proxy), accumulator half drained
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&a_ready[h]);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(acc_t, 256);
  }
}

}  // namespace ardae
