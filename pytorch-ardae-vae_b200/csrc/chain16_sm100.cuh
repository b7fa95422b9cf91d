// Fused layer-chain kernel, 16-bit spill variant: the same on-chip pipeline as chain_sm100.cuh (one launch per
// sweep, the activation operand of every layer resident in TMEM / shared memory, tf32 tensor-core products with
// fp32 accumulation), but every [rows, H] array that leaves or enters the SM -- softplus outputs, deltas, tangents,
// t / adjoints -- is stored as bfloat16.  The round-1 update moved ~20 GB of fp32 spill per step and was HBM-bound
// at 0.30 of the tensor roofline; the chain operands themselves stay tf32 on chip, only the spill narrows
// (scripts/tf32_sensitivity.py: every gradient tensor stays within 4e-3 rel-L2 at H = 256).
//
// Also absorbed here (they were separate [N, H]-array-streaming launches): the d -> H first layer of the forward
// sweeps (per-layer input width `kin`) and the H -> d last layer of the score sweep (`nout` < H: linear fp32 output).
//
// Modes as in chain_sm100.cuh (reference math: SURVEY.md 8a-3, models/graddae/mlp.py:400-444):
//   CHAIN_MUL_SIG   out = acc * sig(aux1)                               A' = tf32(out)
//   CHAIN_TANGENT   out = acc * sig(aux1), out2 = aux2*acc*(1-sig)      A' = tf32(out)
//   CHAIN_ADJOINT   out = acc * sig(aux1) + aux2                        A' = tf32(out)
//   CHAIN_SOFTPLUS3 out = softplus(acc + bias terms) ("3xTF32": hi in shared memory, lo in TMEM); spill = bf16(out)
// Shared memory: 6 x 16 KB weight k-block ring + 128 KB = 16 / 8 aux-out slots of one / two 8 KB bf16 tiles
// (64-byte rows, TMA SWIZZLE_64B), in-place epilogue, TMA store from the same bytes.  SOFTPLUS3 keeps its eight
// 16 KB fp32 hi tiles there and writes the bf16 spill with 64-byte row stores straight from registers.
#pragma once
#include "chain_sm100.cuh"

namespace ardae {

constexpr int kTile16Bytes = kBlockM * 32 * 2;  // 128 rows x 32 bf16

struct alignas(64) Chain16LayerParams {
  CUtensorMap tmW;     // fp32 B operand [nout rows, kin] K-major, box {32, min(nout, H/2)}; SOFTPLUS3: [H, 3*kin]
  CUtensorMap tmAux1;  // bf16 [M, H], box {32, 128}, SWIZZLE_64B
  CUtensorMap tmAux2;
  CUtensorMap tmOut;
  CUtensorMap tmOut2;
  const float* bias;        // SOFTPLUS3
  const float* group_bias;  // SOFTPLUS3: row m adds group_bias[(m / group) * ldg + n]
  const float* col_vec;     // SOFTPLUS3: adds row_scale[m] * col_vec[n]
  uint16_t* out16;          // SOFTPLUS3: bf16 spill [M, H], row pitch ld_out16 elements
  float* out32;             // narrow last layer: fp32 [M, nout] (row pitch ld_out32)
  float* colsum;            // colsum[n]  += colsum_scale * sum_m out[m, n]
  float* colsum2;           // TANGENT: colsum2[n] += sum_m out2[m, n]
  float* colsum_w;          // colsum_w[n * stride] += sum_m out[m, n] * row_scale[m]
  float colsum_scale;
  int colsum_w_stride;
  int group, ldg;
  int ld_out16, ld_out32;
  int kin;                  // input width of this layer (multiple of 32, <= H; < H only for layer 0)
  int nout;                 // H, or (last layer only) a multiple of 32 <= H/2: linear layer, fp32 rows to out32
};

struct alignas(64) Chain16Params {
  CUtensorMap tmA0;        // a0_mode 1: bf16 [M, H] (box {32,128}, SWIZZLE_64B); SOFTPLUS3: fp32 hi part [M, kin0] (SWIZZLE_128B)
  const float* a0_f32;     // a0_mode 0: fp32 [M, kin0] read straight from global (tangent sweep: r); SOFTPLUS3: lo part
  int a0_ld;
  int a0_mode;
  const float* row_scale;  // [M] (sigma) or null
  int M, H, nlayers;
  int vec_ok;
  long long* dbg;          // profiling only (chain16w): clock64 stamps of CTA 0, [layer][half][group][8]
  Chain16LayerParams layer[kChainMaxLayers];
};

template <int MODE>
struct Chain16Config {
  static constexpr bool kS3 = MODE == CHAIN_SOFTPLUS3;
  static constexpr bool kAux2 = MODE == CHAIN_TANGENT || MODE == CHAIN_ADJOINT;
  static constexpr bool kOut2 = MODE == CHAIN_TANGENT;
  static constexpr int kGroups = 2;
  static constexpr int kWStage = 128 * kBlockK * 4;
  // The weight ring is what feeds the MMAs: a 16 KB k-block is consumed every ~500 cycles with 6 stages in flight --
  // stages x 16 KB / (L2 -> SM latency), Little's law, not L2 bandwidth (halving the bytes by multicast changes
  // nothing).  The 224 KB of shared memory are therefore split in favour of the weight ring: 8 stages (128 KB) + 96 KB
  // of aux / out slots (score sweep: 12 slots of 8 KB; tangent / adjoint: 6 slots of two tiles -- every epilogue group
  // holds one slot per N-half and hands it back at the end of the half, so NB/2 slots are in use and the rest is
  // prefetch).  Measured at config 2 (score / tangent / adjoint, ms): 6 stages 0.365 / 0.500 / 0.413, 7 stages
  // 0.350 / 0.466 / 0.389, 8 stages 0.335 / 0.430 / 0.375, 10 stages (4 slots, no prefetch) 0.337 / 0.621 / 0.434.
#ifndef ARDAE_CHAIN16_WSTAGES_1
#define ARDAE_CHAIN16_WSTAGES_1 8
#endif
#ifndef ARDAE_CHAIN16_WSTAGES_2
#define ARDAE_CHAIN16_WSTAGES_2 8
#endif
  static constexpr int kNumWStages = kS3 ? 6 : (kAux2 ? ARDAE_CHAIN16_WSTAGES_2 : ARDAE_CHAIN16_WSTAGES_1);
  static constexpr int kAuxSlot = kS3 ? 0 : (kAux2 ? 2 : 1) * kTile16Bytes;
  static constexpr int kAuxBytes = (14 - kNumWStages) * kTileBytes;   // 14 x 16 KB = 224 KB in all
  static constexpr int kNumAux = kS3 ? 0 : kAuxBytes / kAuxSlot;
  static constexpr int kOffAux = kNumWStages * kWStage;
  static constexpr int kDataBytes = kOffAux + kAuxBytes;
  static constexpr int kSmemBytes = kDataBytes + 1024 + 768;
  static_assert(kSmemBytes <= 232448 && kNumAux <= 16 && (kS3 ? kAuxBytes == 8 * kTileBytes : kNumAux >= 4), "shared memory budget");
  static constexpr int kThreads = 128 + kGroups * 128;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }
__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(Chain16Config<MODE>::kThreads, 1)
chain16_kernel(const __grid_constant__ Chain16Params p) {
  using Cfg = Chain16Config<MODE>;
  constexpr bool S3 = Cfg::kS3, HAS_AUX2 = Cfg::kAux2, HAS_OUT2 = Cfg::kOut2;
  constexpr int G = Cfg::kGroups, NW = Cfg::kNumWStages;
  constexpr int NAUX = Cfg::kNumAux > 0 ? Cfg::kNumAux : 1;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* w_full = reinterpret_cast<uint64_t*>(smem + Cfg::kDataBytes);
  uint64_t* w_empty = w_full + NW;
  uint64_t* aux_full = w_empty + NW;    // [16]
  uint64_t* aux_empty = aux_full + 16;  // [16]
  uint64_t* acc_full = aux_empty + 16;  // [2] accumulator half h of the current layer is complete
  uint64_t* a_ready = acc_full + 2;     // [2] A chunks of half h are written and accumulator half h is drained
  uint64_t* kfree = a_ready + 2;        // [4] the current layer no longer reads A chunk c (c < NB/2)
  uint64_t* a0_full = kfree + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a0_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kBlockM;
  const int H = p.H;
  const int NB = H >> 5;    // 32-column chunks == k-blocks (even: H is a multiple of 64)
  const int NB0 = NB >> 1;  // chunks per N-half
  const int HH = H >> 1;    // columns per N-half
  const int nl = p.nlayers;
  const int NBin0 = p.layer[0].kin >> 5;           // chunks of the initial activation
  const bool a0_ring = !S3 && p.a0_mode == 1;      // initial activation arrives as bf16 tiles through the aux ring
  const int ring_base = a0_ring ? NB : 0;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.layer[0].tmW);
    if (S3 || a0_ring) ptx::prefetch_tmap(&p.tmA0);
    if (!S3) {
      ptx::prefetch_tmap(&p.layer[0].tmAux1);
      ptx::prefetch_tmap(&p.layer[0].tmOut);
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < NW; ++s) {
        ptx::mbar_init(&w_full[s], 1);
        ptx::mbar_init(&w_empty[s], 1);
      }
      for (int a = 0; a < 16; ++a) {
        ptx::mbar_init(&aux_full[a], 1);
        ptx::mbar_init(&aux_empty[a], 1);  // the store-issuing thread of the group that consumed the slot
      }
      for (int h = 0; h < 2; ++h) {
        ptx::mbar_init(&acc_full[h], 1);
        ptx::mbar_init(&a_ready[h], 4 * G);  // every epilogue warp
      }
      for (int c = 0; c < 4; ++c) ptx::mbar_init(&kfree[c], 1);
      ptx::mbar_init(a0_full, 1);
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc_t = tmem_base;       // accumulator columns [0, H)
  const uint32_t a_t = tmem_base + 256;   // A operand (SOFTPLUS3: its lo part) columns [256, 256 + H)
  uint8_t* hi_tiles = smem + Cfg::kOffAux;  // SOFTPLUS3 only

  if (warp == 0) {
    // ------------------------------------------------------------ weight producer
    if (ptx::elect_one()) {
      int it = 0;
      for (int l = 0; l < nl; ++l) {
        const Chain16LayerParams& L = p.layer[l];
        const CUtensorMap* tw = &L.tmW;
        const int NBl = L.kin >> 5;
        const bool narrow = L.nout < H;
        const int nst = S3 ? 2 * NBl : NBl;
        const uint32_t wbytes = static_cast<uint32_t>(narrow ? L.nout : HH) * kBlockK * 4;
        for (int h = 0; h < (narrow ? 1 : 2); ++h) {
          for (int j = 0; j < nst; ++j, ++it) {
            const int s = it % NW;
            const uint32_t ph = (it / NW) & 1;
            ptx::mbar_wait(&w_empty[s], ph ^ 1);
            // SOFTPLUS3: k-block kb of Whi (columns [0,kin)) then of Wlo (columns [2kin,3kin)); both serve hi, Whi also lo
            const int kc = S3 ? (((j & 1) ? 2 * L.kin : 0) + (j >> 1) * kBlockK) : j * kBlockK;
            ptx::mbar_expect_tx(&w_full[s], wbytes);
            ptx::tma_load_2d(smem + s * Cfg::kWStage, tw, &w_full[s], kc, h * HH);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (ptx::elect_one()) {
      int it = 0;
      for (int l = 0; l < nl; ++l) {
        const Chain16LayerParams& L = p.layer[l];
        const int NBl = L.kin >> 5;
        const bool narrow = L.nout < H;
        const uint32_t idesc = ptx::make_idesc_tf32(kBlockM, narrow ? L.nout : HH, 0, 0);
        if (S3 && l == 0) ptx::mbar_wait(a0_full, 0);  // hi tiles of the initial activation have landed (TMA)
        for (int h = 0; h < (narrow ? 1 : 2); ++h) {
          const uint32_t d_t = acc_t + h * HH;
          for (int kb = 0; kb < NBl; ++kb) {
            if (h == 0 && kb == 0) {
              ptx::mbar_wait(&a_ready[0], l & 1);  // A chunks [0, NB/2) written, accumulator half 0 drained
              ptx::tc_fence_after();
            }
            if (h == 0 && kb == NB0) {
              ptx::mbar_wait(&a_ready[1], l & 1);  // A chunks [NB/2, NB) written, accumulator half 1 drained
              ptx::tc_fence_after();
            }
            {
              const int s = it % NW;
              const uint32_t ph = (it / NW) & 1;
              ptx::mbar_wait(&w_full[s], ph);
              ptx::tc_fence_after();
              const uint32_t b_addr = ptx::smem_u32(smem + s * Cfg::kWStage);
              if (S3) {
                const uint32_t h_addr = ptx::smem_u32(hi_tiles + kb * kTileBytes);
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                  const uint64_t adesc = ptx::make_smem_desc_sw128(h_addr + k * kUmmaK * 4, 0, 1024);
                  const uint64_t bdesc = ptx::make_smem_desc_sw128(b_addr + k * kUmmaK * 4, 0, 1024);
                  ptx::umma_tf32(d_t, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
                }
              }
#pragma unroll
              for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                const uint64_t bdesc = ptx::make_smem_desc_sw128(b_addr + k * kUmmaK * 4, 0, 1024);
                ptx::umma_tf32_ts(d_t, a_t + kb * kBlockK + k * kUmmaK, bdesc, idesc, (S3 || (kb | k) != 0) ? 1u : 0u);
              }
              ptx::umma_commit(&w_empty[s]);
              ++it;
            }
            if (S3) {  // hi . Wlo
              const int s = it % NW;
              const uint32_t ph = (it / NW) & 1;
              ptx::mbar_wait(&w_full[s], ph);
              ptx::tc_fence_after();
              const uint32_t b_addr = ptx::smem_u32(smem + s * Cfg::kWStage);
              const uint32_t h_addr = ptx::smem_u32(hi_tiles + kb * kTileBytes);
#pragma unroll
              for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                const uint64_t adesc = ptx::make_smem_desc_sw128(h_addr + k * kUmmaK * 4, 0, 1024);
                const uint64_t bdesc = ptx::make_smem_desc_sw128(b_addr + k * kUmmaK * 4, 0, 1024);
                ptx::umma_tf32(d_t, adesc, bdesc, idesc, 1u);
              }
              ptx::umma_commit(&w_empty[s]);
              ++it;
            }
            // second half: this layer is done with A chunk kb -> the half-0 epilogue may overwrite it
            if (h == 1 && kb < NB0) ptx::umma_commit(&kfree[kb]);
          }
          if (h == 0 && NBl <= NB0 && !narrow) {  // short first layer: consume this layer's a_ready[1] phase as well
            ptx::mbar_wait(&a_ready[1], l & 1);
            ptx::tc_fence_after();
          }
          if (h == 1)
            for (int c = NBl; c < NB0; ++c) ptx::umma_commit(&kfree[c]);  // chunks this layer never read
          ptx::umma_commit(&acc_full[h]);
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------ aux / initial-activation producer
    if (ptx::elect_one()) {
      if (S3) {
        ptx::mbar_expect_tx(a0_full, static_cast<uint32_t>(NBin0) * kTileBytes);
        for (int c = 0; c < NBin0; ++c) ptx::tma_load_2d(hi_tiles + c * kTileBytes, &p.tmA0, a0_full, c * 32, m0);
      } else {
        int it = 0;
        if (a0_ring) {
          for (int c = 0; c < NB; ++c, ++it) {
            const int a = it % NAUX;
            ptx::mbar_wait(&aux_empty[a], ((it / NAUX) & 1) ^ 1);
            ptx::mbar_expect_tx(&aux_full[a], kTile16Bytes);
            ptx::tma_load_2d(smem + Cfg::kOffAux + a * Cfg::kAuxSlot, &p.tmA0, &aux_full[a], c * 32, m0);
          }
        }
        for (int l = 0; l < nl; ++l) {
          const Chain16LayerParams& L = p.layer[l];
          if (L.nout < H) break;  // narrow last layer: linear, no aux
          for (int c = 0; c < NB; ++c, ++it) {
            const int a = it % NAUX;
            ptx::mbar_wait(&aux_empty[a], ((it / NAUX) & 1) ^ 1);
            uint8_t* slot = smem + Cfg::kOffAux + a * Cfg::kAuxSlot;
            ptx::mbar_expect_tx(&aux_full[a], Cfg::kAuxSlot);
            ptx::tma_load_2d(slot, &L.tmAux1, &aux_full[a], c * 32, m0);
            if (HAS_AUX2) ptx::tma_load_2d(slot + kTile16Bytes, &L.tmAux2, &aux_full[a], c * 32, m0);
          }
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
    const bool leader = (quarter == 0 && lane == 0);
    const int swz = r & 7;                       // fp32 tiles (128-byte rows, SWIZZLE_128B)
    const int sw16 = (r >> 1) & 3;               // bf16 tiles (64-byte rows, SWIZZLE_64B)
    const uint32_t row_off = static_cast<uint32_t>(r) * 128u;
    const uint32_t row_off16 = static_cast<uint32_t>(r) * 64u;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t bar_a = 1 + 2 * g, bar_b = 2 + 2 * g;
    const float rs = (p.row_scale != nullptr && row_ok) ? p.row_scale[m] : 0.0f;

    // ---- pseudo-layer -1: bring the initial activation into TMEM
    for (int c = g; c < NBin0; c += G) {
      uint32_t v[32];
      if (S3 || !a0_ring) {
        const float* src = p.a0_f32 + static_cast<size_t>(row_ok ? m : 0) * p.a0_ld + c * 32;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row_ok) t = __ldg(reinterpret_cast<const float4*>(src) + q);
          v[q * 4 + 0] = __float_as_uint(t.x); v[q * 4 + 1] = __float_as_uint(t.y);
          v[q * 4 + 2] = __float_as_uint(t.z); v[q * 4 + 3] = __float_as_uint(t.w);
        }
      } else {
        const int it = c;
        const int a = it % NAUX;
        ptx::mbar_wait(&aux_full[a], (it / NAUX) & 1);
        const uint32_t src = ptx::smem_u32(smem + Cfg::kOffAux + a * Cfg::kAuxSlot) + row_off16;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 t = lds128u(src + ((q ^ sw16) << 4));
          const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            v[q * 8 + 2 * i] = w[i] << 16;             // bf16 -> fp32 bit pattern (exactly tf32-representable)
            v[q * 8 + 2 * i + 1] = w[i] & 0xFFFF0000u;
          }
        }
        ptx::named_bar_sync(bar_a, 128);  // all four warps have read the slot
        if (leader) ptx::mbar_arrive(&aux_empty[a]);
      }
      ptx::tmem_st_32x32(a_t + lane_addr + c * 32, v);
    }
    ptx::tmem_st_wait();
    ptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      ptx::mbar_arrive(&a_ready[0]);
      ptx::mbar_arrive(&a_ready[1]);
    }

    int prev_slot = -1;  // slot whose TMA store may still be reading it (released one chunk later)
#pragma unroll 1
    for (int l = 0; l < nl; ++l) {
      const Chain16LayerParams& L = p.layer[l];
      const bool last = (l == nl - 1);
      if (L.nout < H) {
        // ---- narrow linear last layer (score sweep: g = delta a_1 . A_1): fp32 rows straight to global
        ptx::mbar_wait(&acc_full[0], l & 1);
        ptx::tc_fence_after();
        for (int c = g; c * 32 < L.nout; c += G) {
          uint32_t accu[32];
          ptx::tmem_ld_32x32(acc_t + lane_addr + c * 32, accu);
          ptx::tmem_ld_wait();
          if (row_ok) {
            float4* dst = reinterpret_cast<float4*>(L.out32 + static_cast<size_t>(m) * L.ld_out32 + c * 32);
#pragma unroll
            for (int q = 0; q < 8; ++q)
              dst[q] = make_float4(__uint_as_float(accu[q * 4 + 0]), __uint_as_float(accu[q * 4 + 1]),
                                   __uint_as_float(accu[q * 4 + 2]), __uint_as_float(accu[q * 4 + 3]));
          }
        }
        break;
      }
      const float* gb_row = (S3 && L.group_bias != nullptr)
                                ? L.group_bias + static_cast<size_t>((row_ok ? m : 0) / L.group) * L.ldg
                                : nullptr;
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        ptx::mbar_wait(&acc_full[h], l & 1);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int c = h * NB0 + g; c < (h + 1) * NB0; c += G) {
          const int nc = c * 32;
          if (h == 0) {  // the half-1 MMAs of this layer still read A chunk c until kfree[c] fires
            ptx::mbar_wait(&kfree[c], l & 1);
            ptx::tc_fence_after();
          }
          uint32_t accu[32];
          ptx::tmem_ld_32x32(acc_t + lane_addr + nc, accu);
          const int it = ring_base + NB * l + c;
          const int a = it % NAUX;
          if (!S3) ptx::mbar_wait(&aux_full[a], (it / NAUX) & 1);
          ptx::tmem_ld_wait();
          float v[32];  // A operand of the next layer (tf32): `out`, or the lo part for SOFTPLUS3
          uint8_t* slot = smem + Cfg::kOffAux + a * Cfg::kAuxSlot;  // non-S3: aux tile(s) in, out tile(s) over them
          const uint32_t s1 = ptx::smem_u32(slot) + row_off16;      // aux1 / out row of this thread
          const uint32_t s2 = s1 + kTile16Bytes;                    // aux2 / out2
          if (S3) {
            float add[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) add[j] = 0.0f;
            const bool vec = p.vec_ok != 0;
            if (L.bias != nullptr) add_cols32(L.bias + nc, 1.0f, vec, 32, add);
            if (gb_row != nullptr) add_cols32(gb_row + nc, 1.0f, vec, 32, add);
            if (L.col_vec != nullptr) add_cols32(L.col_vec + nc, rs, vec, 32, add);
            const uint32_t h1 = ptx::smem_u32(hi_tiles + c * kTileBytes) + row_off;
            uint32_t ow[16];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float o[4], full[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float res = softplus_fast(__uint_as_float(accu[q * 4 + j]) + add[q * 4 + j]);
                const float hi = round_tf32_fast(res);
                o[j] = hi;
                full[j] = res;
                v[q * 4 + j] = round_tf32_fast(res - hi);
              }
              sts128(h1 + ((q ^ swz) << 4), o[0], o[1], o[2], o[3]);
              ow[q * 2 + 0] = pack_bf16x2(full[0], full[1]);
              ow[q * 2 + 1] = pack_bf16x2(full[2], full[3]);
            }
            if (row_ok) {  // bf16 spill: 64 contiguous bytes of this thread's row
              uint4* dst = reinterpret_cast<uint4*>(L.out16 + static_cast<size_t>(m) * L.ld_out16 + nc);
#pragma unroll
              for (int q = 0; q < 4; ++q) dst[q] = make_uint4(ow[q * 4 + 0], ow[q * 4 + 1], ow[q * 4 + 2], ow[q * 4 + 3]);
            }
          } else {
            uint4 x1[4], x2[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t soff = static_cast<uint32_t>((q ^ sw16) << 4);
              x1[q] = lds128u(s1 + soff);
              if (HAS_AUX2) x2[q] = lds128u(s2 + soff);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t soff = static_cast<uint32_t>((q ^ sw16) << 4);
              const uint32_t w1[4] = {x1[q].x, x1[q].y, x1[q].z, x1[q].w};
              const uint32_t w2[4] = {HAS_AUX2 ? x2[q].x : 0u, HAS_AUX2 ? x2[q].y : 0u, HAS_AUX2 ? x2[q].z : 0u,
                                      HAS_AUX2 ? x2[q].w : 0u};
              uint32_t ow[4], ow2[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float o[2], ob[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const int j = q * 8 + i * 2 + e;
                  const float u = e ? bf16_hi(w1[i]) : bf16_lo(w1[i]);
                  const float y2 = e ? bf16_hi(w2[i]) : bf16_lo(w2[i]);
                  const float pre = __uint_as_float(accu[j]);
                  float sg, oms;
                  sig_fast(u, sg, oms);
                  float res, res2 = 0.0f;
                  if (MODE == CHAIN_MUL_SIG) {
                    res = pre * sg;
                  } else if (MODE == CHAIN_TANGENT) {
                    res = pre * sg;
                    res2 = y2 * pre * oms;
                  } else {
                    res = fmaf(pre, sg, y2);
                  }
                  o[e] = res;
                  ob[e] = res2;
                  v[j] = round_tf32_fast(res);
                }
                ow[i] = pack_bf16x2(o[0], o[1]);
                ow2[i] = pack_bf16x2(ob[0], ob[1]);
              }
              sts128u(s1 + soff, ow[0], ow[1], ow[2], ow[3]);
              if (HAS_OUT2) sts128u(s2 + soff, ow2[0], ow2[1], ow2[2], ow2[3]);
            }
          }
          if (!last) {
            uint32_t vu[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) vu[j] = __float_as_uint(v[j]);
            ptx::tmem_st_32x32(a_t + lane_addr + nc, vu);
          }
          ptx::fence_proxy_async_smem();
          if (!S3) {
            ptx::named_bar_sync(bar_b, 128);
            if (leader) {
              ptx::tma_store_2d(&L.tmOut, slot, nc, m0);
              if (HAS_OUT2) ptx::tma_store_2d(&L.tmOut2, slot + kTile16Bytes, nc, m0);
              ptx::tma_store_commit();
              if (prev_slot >= 0) {  // the previous store of this group has read its slot: hand it back to the producer
                ptx::tma_store_wait_read<1>();
                ptx::mbar_arrive(&aux_empty[prev_slot]);
              }
              prev_slot = a;
            }
          }
          // ---- fused column sums (bias gradients, d w_sigma, d w_o)
          if (!S3 && (L.colsum != nullptr || L.colsum_w != nullptr || (HAS_OUT2 && L.colsum2 != nullptr))) {
            if (!row_ok) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = 0.0f;
            }
            if (L.colsum_w != nullptr) {
              float w[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) w[i] = v[i] * rs;
              const float t = warp_transpose_reduce32(w, lane);
              atomicAdd(L.colsum_w + static_cast<size_t>(nc + lane) * L.colsum_w_stride, t);
            }
            if (L.colsum != nullptr) {
              const float t = warp_transpose_reduce32(v, lane);
              atomicAdd(L.colsum + nc + lane, L.colsum_scale * t);
            }
            if (HAS_OUT2 && L.colsum2 != nullptr) {
              // re-read this thread's own out2 row from the staging tile (still intact: the next write to it
              // happens after this thread passes the next barrier)
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint4 t4 = lds128u(s2 + ((q ^ sw16) << 4));
                const uint32_t w[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  v[q * 8 + 2 * i] = row_ok ? bf16_lo(w[i]) : 0.0f;
                  v[q * 8 + 2 * i + 1] = row_ok ? bf16_hi(w[i]) : 0.0f;
                }
              }
              const float t = warp_transpose_reduce32(v, lane);
              atomicAdd(L.colsum2 + nc + lane, t);
            }
          }
        }
        // A chunks of this half written (TMEM stores complete, shared-memory writes fenced), accumulator half drained
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&a_ready[h]);
        if (HAS_OUT2 && L.colsum2 != nullptr) ptx::named_bar_sync(bar_a, 128);  // colsum2 re-reads of the slot are done
        if (!S3 && leader && prev_slot >= 0) {  // do not sit on a slot while waiting for the next accumulator half
          ptx::tma_store_wait_read<0>();
          ptx::mbar_arrive(&aux_empty[prev_slot]);
          prev_slot = -1;
        }
      }  // h
    }
    if (!S3 && leader) ptx::tma_store_wait_all<0>();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace ardae
