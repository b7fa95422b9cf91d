// Weight-gradient contraction over bf16 operands (the 16-bit spill of chain16_sm100.cuh):
//   out[M, N] += sum_pairs sum_k X_p[k, M] . Y_p[k, N]        (1..4 operand pairs, fp32 accumulation in TMEM)
// Both operands are "MN-major" (reduction over the slow dimension = the rows of two activation arrays), loaded by
// TMA as boxes of 64 bf16 (128 bytes) x 64 rows, SWIZZLE_128B, and consumed by tcgen05.mma kind::f16 straight from
// shared memory: canonical MN-major layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units, LBO = one TMA box
// (next 64 output columns), SBO = 1 KB (next 8 reduction rows).  Split-K over ~2 CTAs per SM; every split adds
// its tile into the gradient with red.global.add (or writes a partial tile for the fixed-order reduce launch).
// Half the HBM bytes of the tf32 contraction (gemm_sm100.cuh::gemm_tn_kernel) and twice its tensor rate.
#pragma once
#include "chain16_host.cuh"
#include "gemm_host.cuh"

namespace ardae {

constexpr int kTN16BlockK = 64;   // reduction rows per stage
constexpr int kTN16MaxPairs = 4;

struct alignas(64) GemmTN16Params {
  CUtensorMap tmX[kTN16MaxPairs];  // bf16 dims {M, K}, box {64, 64}
  CUtensorMap tmY[kTN16MaxPairs];  // bf16 dims {N, K}, box {64, 64}
  int M, N, K;
  int npairs;
  int kb_per_split;
  float* partial;    // [nsplit][Mpad][Npad]
  float* red_out;    // non-null: red.global.add into out[M, N] (pitch red_ld)
  int red_ld;
  int red_vec;
  int npad;
};

template <int BLOCK_N>
struct GemmTN16Config {
  static constexpr int kBoxBytes = 64 * kTN16BlockK * 2;  // 8 KB: 64 reduction rows x 64 bf16
  static constexpr int kStageA = (kBlockM / 64) * kBoxBytes;
  static constexpr int kStageB = (BLOCK_N / 64) * kBoxBytes;
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kNumStages = BLOCK_N >= 256 ? 2 : (BLOCK_N >= 128 ? 3 : 4);
  static constexpr int kDataBytes = kStage * kNumStages;
  static constexpr int kSmemBytes = kDataBytes + 1024 + 256;
  static constexpr int kTmemCols = BLOCK_N < 32 ? 32 : BLOCK_N;
  static_assert(BLOCK_N % 64 == 0 && BLOCK_N <= 256, "BLOCK_N");
};

namespace ptx {
// kind::f16 instruction descriptor, bf16 operands, fp32 accumulate:
//   [4,6) c_format = 1 (F32)   [7,10) a_format = 1 (BF16)   [10,13) b_format = 1 (BF16)
//   [15] a_major (1 = MN)   [16] b_major   [17,23) N >> 3   [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
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

template <int BLOCK_N>
__global__ void __launch_bounds__(kGemmThreads, 2) gemm_tn16_kernel(const __grid_constant__ GemmTN16Params p) {
  using Cfg = GemmTN16Config<BLOCK_N>;
  constexpr int NSTAGE = Cfg::kNumStages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kDataBytes);
  uint64_t* empty_bar = full_bar + NSTAGE;
  uint64_t* tmem_full_bar = empty_bar + NSTAGE;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int split = blockIdx.x;
  const int m0 = blockIdx.y * kBlockM;
  const int n0 = blockIdx.z * BLOCK_N;
  const int total_kb = (p.K + kTN16BlockK - 1) / kTN16BlockK;
  const int kb_begin = split * p.kb_per_split;
  int kb_end = kb_begin + p.kb_per_split;
  if (kb_end > total_kb) kb_end = total_kb;
  const int nkb = kb_end > kb_begin ? kb_end - kb_begin : 0;
  const int iters = nkb * p.npairs;  // pair-major: all k-blocks of pair 0, then pair 1, ...

  if (warp == 0 && lane == 0) {
    for (int q = 0; q < p.npairs; ++q) {
      ptx::prefetch_tmap(&p.tmX[q]);
      ptx::prefetch_tmap(&p.tmY[q]);
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < NSTAGE; ++s) {
        ptx::mbar_init(&full_bar[s], 1);
        ptx::mbar_init(&empty_bar[s], 1);
      }
      ptx::mbar_init(tmem_full_bar, 1);
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, Cfg::kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (ptx::elect_one()) {
      for (int it = 0; it < iters; ++it) {
        const int s = it % NSTAGE;
        const uint32_t ph = (it / NSTAGE) & 1;
        const int pair = it / nkb;
        const int k0 = (kb_begin + (it - pair * nkb)) * kTN16BlockK;
        ptx::mbar_wait(&empty_bar[s], ph ^ 1);
        ptx::mbar_expect_tx(&full_bar[s], Cfg::kStage);
        uint8_t* sa = smem + s * Cfg::kStage;
        uint8_t* sb = sa + Cfg::kStageA;
#pragma unroll
        for (int b = 0; b < kBlockM / 64; ++b)
          ptx::tma_load_2d(sa + b * Cfg::kBoxBytes, &p.tmX[pair], &full_bar[s], m0 + b * 64, k0);
#pragma unroll
        for (int b = 0; b < BLOCK_N / 64; ++b)
          ptx::tma_load_2d(sb + b * Cfg::kBoxBytes, &p.tmY[pair], &full_bar[s], n0 + b * 64, k0);
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(kBlockM, BLOCK_N, 1, 1);
      for (int it = 0; it < iters; ++it) {
        const int s = it % NSTAGE;
        const uint32_t ph = (it / NSTAGE) & 1;
        ptx::mbar_wait(&full_bar[s], ph);
        ptx::tc_fence_after();
        const uint32_t a_addr = ptx::smem_u32(smem + s * Cfg::kStage);
        const uint32_t b_addr = a_addr + Cfg::kStageA;
#pragma unroll
        for (int k = 0; k < kTN16BlockK / 16; ++k) {  // one MMA = 16 reduction rows = two 8-row swizzle atoms (2 KB)
          const uint64_t adesc = ptx::make_smem_desc_sw128(a_addr + k * 2048, Cfg::kBoxBytes, 1024);
          const uint64_t bdesc = ptx::make_smem_desc_sw128(b_addr + k * 2048, Cfg::kBoxBytes, 1024);
          ptx::umma_bf16(tmem_base, adesc, bdesc, idesc, (it | k) != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&empty_bar[s]);
      }
      ptx::umma_commit(tmem_full_bar);
    }
  } else {
    const int quarter = warp & 3;
    const int Mpad = gridDim.y * kBlockM;
    const int Npad = p.npad;
    if (iters > 0) {
      ptx::mbar_wait(tmem_full_bar, 0);
      ptx::tc_fence_after();
    }
    const int r = quarter * 32 + lane;
    if (p.red_out != nullptr) {
      if (iters > 0) {
        const bool row_ok = m0 + r < p.M;  // tcgen05.ld is warp-collective: only the reductions are predicated
        float* orow = p.red_out + static_cast<size_t>(row_ok ? m0 + r : 0) * p.red_ld + n0;
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          if (n0 + c * 32 >= p.N) break;
          uint32_t accu[32];
          ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + c * 32, accu);
          ptx::tmem_ld_wait();
          if (!row_ok) continue;
          if (p.red_vec && n0 + c * 32 + 32 <= p.N) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(orow + c * 32 + q * 4),
                           "f"(__uint_as_float(accu[q * 4 + 0])), "f"(__uint_as_float(accu[q * 4 + 1])),
                           "f"(__uint_as_float(accu[q * 4 + 2])), "f"(__uint_as_float(accu[q * 4 + 3]))
                           : "memory");
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + c * 32 + j < p.N) atomicAdd(orow + c * 32 + j, __uint_as_float(accu[j]));
          }
        }
      }
    } else {
      float* dst = p.partial + (static_cast<size_t>(split) * Mpad + (m0 + r)) * Npad + n0;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t accu[32];
        if (iters > 0) {
          ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + c * 32, accu);
          ptx::tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) accu[i] = 0u;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<uint4*>(dst + c * 32 + q * 4) =
              make_uint4(accu[q * 4 + 0], accu[q * 4 + 1], accu[q * 4 + 2], accu[q * 4 + 3]);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------
// Several same-shape contractions (the 2(L-1) [H,H] weight gradients of one update) in ONE launch: every CTA keeps its
// (split, m-tile) slot and walks the problems in order.  The TMA producer runs ahead into the next problem while the
// epilogue warps drain the accumulator (and the co-resident CTA keeps streaming), so launch, ramp-up, tail and the
// red.add burst are paid once instead of once per layer (round-1 profile: ~25 us of a 100 us contraction).
constexpr int kTN16MaxProbs = 10;
struct alignas(64) GemmTN16Prob {
  CUtensorMap tmX[2];
  CUtensorMap tmY[2];
  float* red_out;
  int red_ld;
  int npairs;
};
struct alignas(64) GemmTN16MultiParams {
  GemmTN16Prob prob[kTN16MaxProbs];
  int nprob;
  int M, N, K;
  int kb_per_split;
};

// One CTA per SM, four 48 KB stages (192 KB of HBM traffic in flight per SM) and a DOUBLE-BUFFERED accumulator (2 x BLOCK_N
// TMEM columns): the MMAs of problem q+1 fill one buffer while the epilogue warps drain problem q from the other, so the
// operand stream never pauses for the red.add epilogue.
template <int BLOCK_N>
struct GemmTN16MultiConfig {
  using Base = GemmTN16Config<BLOCK_N>;
  static constexpr int kNumStages = BLOCK_N >= 256 ? 4 : 6;
  static constexpr int kDataBytes = Base::kStage * kNumStages;
  static constexpr int kSmemBytes = kDataBytes + 1024 + 256;
  static constexpr int kTmemCols = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

template <int BLOCK_N>
__global__ void __launch_bounds__(kGemmThreads, 1) gemm_tn16_multi_kernel(const __grid_constant__ GemmTN16MultiParams p) {
  using Cfg = GemmTN16Config<BLOCK_N>;
  using MCfg = GemmTN16MultiConfig<BLOCK_N>;
  constexpr int NSTAGE = MCfg::kNumStages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + MCfg::kDataBytes);
  uint64_t* empty_bar = full_bar + NSTAGE;
  uint64_t* tmem_full_bar = empty_bar + NSTAGE;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int split = blockIdx.x;
  const int m0 = blockIdx.y * kBlockM;
  const int n0 = blockIdx.z * BLOCK_N;
  const int total_kb = (p.K + kTN16BlockK - 1) / kTN16BlockK;
  const int kb_begin = split * p.kb_per_split;
  int kb_end = kb_begin + p.kb_per_split;
  if (kb_end > total_kb) kb_end = total_kb;
  const int nkb = kb_end > kb_begin ? kb_end - kb_begin : 0;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.prob[0].tmX[0]);
    ptx::prefetch_tmap(&p.prob[0].tmY[0]);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < NSTAGE; ++s) {
        ptx::mbar_init(&full_bar[s], 1);
        ptx::mbar_init(&empty_bar[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        ptx::mbar_init(&tmem_full_bar[b], 1);
        ptx::mbar_init(&tmem_empty_bar[b], 4);  // one arrival per epilogue warp
      }
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, MCfg::kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (nkb == 0) {  // (only when there are more splits than k-blocks) nothing to add
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, MCfg::kTmemCols);
    return;
  }

  if (warp == 0) {
    if (ptx::elect_one()) {
      int it = 0;
      for (int q = 0; q < p.nprob; ++q) {
        const GemmTN16Prob& pr = p.prob[q];
        if (q + 1 < p.nprob) {
          ptx::prefetch_tmap(&p.prob[q + 1].tmX[0]);
          ptx::prefetch_tmap(&p.prob[q + 1].tmY[0]);
        }
        for (int pair = 0; pair < pr.npairs; ++pair) {
          for (int kb = 0; kb < nkb; ++kb, ++it) {
            const int s = it % NSTAGE;
            const int k0 = (kb_begin + kb) * kTN16BlockK;
            ptx::mbar_wait(&empty_bar[s], ((it / NSTAGE) & 1) ^ 1);
            ptx::mbar_expect_tx(&full_bar[s], Cfg::kStage);
            uint8_t* sa = smem + s * Cfg::kStage;
            uint8_t* sb = sa + Cfg::kStageA;
#pragma unroll
            for (int b = 0; b < kBlockM / 64; ++b)
              ptx::tma_load_2d(sa + b * Cfg::kBoxBytes, &pr.tmX[pair], &full_bar[s], m0 + b * 64, k0);
#pragma unroll
            for (int b = 0; b < BLOCK_N / 64; ++b)
              ptx::tma_load_2d(sb + b * Cfg::kBoxBytes, &pr.tmY[pair], &full_bar[s], n0 + b * 64, k0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(kBlockM, BLOCK_N, 1, 1);
      int it = 0;
      for (int q = 0; q < p.nprob; ++q) {
        const int iters = nkb * p.prob[q].npairs;
        const int buf = q & 1;
        const uint32_t d_t = tmem_base + buf * BLOCK_N;
        ptx::mbar_wait(&tmem_empty_bar[buf], ((q >> 1) & 1) ^ 1);  // the epilogue has drained this buffer (problem q-2)
        ptx::tc_fence_after();
        for (int i = 0; i < iters; ++i, ++it) {
          const int s = it % NSTAGE;
          ptx::mbar_wait(&full_bar[s], (it / NSTAGE) & 1);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem + s * Cfg::kStage);
          const uint32_t b_addr = a_addr + Cfg::kStageA;
#pragma unroll
          for (int k = 0; k < kTN16BlockK / 16; ++k) {
            const uint64_t adesc = ptx::make_smem_desc_sw128(a_addr + k * 2048, Cfg::kBoxBytes, 1024);
            const uint64_t bdesc = ptx::make_smem_desc_sw128(b_addr + k * 2048, Cfg::kBoxBytes, 1024);
            ptx::umma_bf16(d_t, adesc, bdesc, idesc, (i | k) != 0 ? 1u : 0u);
          }
          ptx::umma_commit(&empty_bar[s]);
        }
        ptx::umma_commit(&tmem_full_bar[buf]);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const bool row_ok = m0 + r < p.M;
    for (int q = 0; q < p.nprob; ++q) {
      const GemmTN16Prob& pr = p.prob[q];
      const int buf = q & 1;
      ptx::mbar_wait(&tmem_full_bar[buf], (q >> 1) & 1);
      ptx::tc_fence_after();
      float* orow = pr.red_out + static_cast<size_t>(row_ok ? m0 + r : 0) * pr.red_ld + n0;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        if (n0 + c * 32 >= p.N) break;
        uint32_t accu[32];
        ptx::tmem_ld_32x32(tmem_base + buf * BLOCK_N + (static_cast<uint32_t>(quarter * 32) << 16) + c * 32, accu);
        ptx::tmem_ld_wait();
        if (c == BLOCK_N / 32 - 1 || n0 + (c + 1) * 32 >= p.N) {  // last TMEM read: hand the buffer back
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[buf]);
        }
        if (!row_ok) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(orow + c * 32 + j * 4),
                       "f"(__uint_as_float(accu[j * 4 + 0])), "f"(__uint_as_float(accu[j * 4 + 1])),
                       "f"(__uint_as_float(accu[j * 4 + 2])), "f"(__uint_as_float(accu[j * 4 + 3]))
                       : "memory");
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, MCfg::kTmemCols);
  }
}

// ------------------------------------------------------------------ host side
struct GemmTN16Desc {
  const uint16_t* X[kTN16MaxPairs] = {nullptr, nullptr, nullptr, nullptr}; int ldx[kTN16MaxPairs] = {0, 0, 0, 0};  // [K, M]
  const uint16_t* Y[kTN16MaxPairs] = {nullptr, nullptr, nullptr, nullptr}; int ldy[kTN16MaxPairs] = {0, 0, 0, 0};  // [K, Ny]
  int npairs = 0;
  int M = 0, N = 0, K = 0;   // output M x N (N <= Ny: extra operand columns are computed and dropped)
  int Ny = 0;                // columns that exist in Y (0 = N); must cover the 64-column TMA boxes or be zero padded
  float* out = nullptr; int ldo = 0;  // out += sum (beta = 1, scale = 1)
  float* workspace = nullptr;
  size_t workspace_bytes = 0;
  int target_ctas = 296;
  int atomic = 0;  // 0 auto, 1 red.global.add epilogue, -1 two-pass reduce
};

inline int tn16_block_n(int N) { return N <= 64 ? 64 : (N <= 128 ? 128 : 256); }

inline size_t tn16_workspace_bytes(int M, int N, int K, int target_ctas = 296) {
  target_ctas = tn_target_ctas(target_ctas);
  const int bn = tn16_block_n(N);
  const int mt = (M + kBlockM - 1) / kBlockM, nt = (N + bn - 1) / bn;
  const int total_kb = (K + kTN16BlockK - 1) / kTN16BlockK;
  int nsplit = target_ctas / (mt * nt);
  if (nsplit < 1) nsplit = 1;
  if (nsplit > total_kb) nsplit = total_kb;
  return static_cast<size_t>(nsplit) * mt * kBlockM * nt * bn * sizeof(float);
}

struct PreparedTN16 {
  GemmTN16Params params;
  const void* fn = nullptr;
  dim3 grid;
  int smem = 0;
  int nsplit = 0, Mpad = 0, Npad = 0;
  float* out = nullptr; int M = 0, N = 0, ldo = 0;
  bool atomic = false;
};

inline int prepare_gemm_tn16(const GemmTN16Desc& d_in, PreparedTN16* out) {
  GemmTN16Desc d = d_in;
  d.target_ctas = tn_target_ctas(d.target_ctas);
  if (d.M <= 0 || d.N <= 0 || d.K <= 0 || d.npairs < 1 || d.npairs > kTN16MaxPairs) return fail(-2, "gemm_tn16: bad problem");
  if (!d.out || !d.workspace) return fail(-2, "gemm_tn16: missing pointer");
  const int bn = tn16_block_n(d.N);
  const int mt = (d.M + kBlockM - 1) / kBlockM, nt = (d.N + bn - 1) / bn;
  const int total_kb = (d.K + kTN16BlockK - 1) / kTN16BlockK;
  int nsplit = d.target_ctas / (mt * nt);
  if (nsplit < 1) nsplit = 1;
  if (nsplit > total_kb) nsplit = total_kb;
  const int kb_per_split = (total_kb + nsplit - 1) / nsplit;
  nsplit = (total_kb + kb_per_split - 1) / kb_per_split;
  const size_t need = static_cast<size_t>(nsplit) * mt * kBlockM * nt * bn * sizeof(float);
  if (need > d.workspace_bytes) return fail(-3, "gemm_tn16: workspace too small");
  PreparedTN16 pr;
  std::memset(&pr.params, 0, sizeof(pr.params));
  GemmTN16Params& p = pr.params;
  int rc;
  const int ny = d.Ny > 0 ? d.Ny : d.N;
  for (int q = 0; q < d.npairs; ++q) {
    if (!d.X[q] || !d.Y[q]) return fail(-2, "gemm_tn16: missing operand");
    if ((rc = encode_tmap_2d_bf16(&p.tmX[q], d.X[q], d.M, d.K, d.ldx[q], 64, kTN16BlockK, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = encode_tmap_2d_bf16(&p.tmY[q], d.Y[q], ny, d.K, d.ldy[q], 64, kTN16BlockK, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  p.npairs = d.npairs;
  p.M = d.M; p.N = d.N; p.K = d.K; p.kb_per_split = kb_per_split; p.partial = d.workspace;
  const bool red_vec = (reinterpret_cast<uintptr_t>(d.out) & 15) == 0 && d.ldo % 4 == 0 && d.N % 4 == 0;
  pr.atomic = d.atomic > 0 || (d.atomic == 0 && tn_atomic_default() && red_vec);
  if (pr.atomic) {
    p.red_out = d.out; p.red_ld = d.ldo; p.red_vec = red_vec ? 1 : 0;
  }
  switch (bn) {
    case 64: pr.fn = reinterpret_cast<const void*>(&gemm_tn16_kernel<64>); pr.smem = GemmTN16Config<64>::kSmemBytes; break;
    case 128: pr.fn = reinterpret_cast<const void*>(&gemm_tn16_kernel<128>); pr.smem = GemmTN16Config<128>::kSmemBytes; break;
    default: pr.fn = reinterpret_cast<const void*>(&gemm_tn16_kernel<256>); pr.smem = GemmTN16Config<256>::kSmemBytes; break;
  }
  pr.grid = dim3(nsplit, mt, nt);
  pr.nsplit = nsplit; pr.Mpad = mt * kBlockM; pr.Npad = nt * bn;
  p.npad = pr.Npad;
  pr.out = d.out; pr.M = d.M; pr.N = d.N; pr.ldo = d.ldo;
  ARDAE_CUDA_OK(cudaFuncSetAttribute(pr.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, pr.smem));
  *out = pr;
  return 0;
}

inline int launch_prepared_tn16(const PreparedTN16& pr, cudaStream_t stream) {
  void* args[1] = {const_cast<GemmTN16Params*>(&pr.params)};
  ARDAE_CUDA_OK(cudaLaunchKernel(pr.fn, pr.grid, dim3(kGemmThreads), args, pr.smem, stream));
  if (pr.atomic) return 0;
  const int total = pr.M * pr.N;
  splitk_reduce_kernel<<<(total + 255) / 256, 256, 0, stream>>>(pr.params.partial, pr.nsplit, pr.Mpad, pr.Npad, pr.out,
                                                                 pr.M, pr.N, pr.ldo, 1.0f, 1.0f);
  ARDAE_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace ardae

namespace ardae {

struct PreparedTN16Multi {
  GemmTN16MultiParams params;
  const void* fn = nullptr;
  dim3 grid;
  int smem = 0;
};

// One launch for same-shape, <= 2-pair, vector-reducible contractions (ARDAE_TN_MULTI=0: one launch each).
inline bool tn16_multi_ok(const std::vector<GemmTN16Desc>& ds) {
  static int env = -1;
  if (env < 0) {
    const char* e = std::getenv("ARDAE_TN_MULTI");
    env = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  if (!env || !tn_atomic_default() || ds.size() < 2 || ds.size() > static_cast<size_t>(kTN16MaxProbs)) return false;
  for (const GemmTN16Desc& d : ds) {
    const bool red_vec = (reinterpret_cast<uintptr_t>(d.out) & 15) == 0 && d.ldo % 4 == 0 && d.N % 32 == 0;
    if (!red_vec || d.atomic < 0 || d.npairs < 1 || d.npairs > 2 || d.M != ds[0].M || d.N != ds[0].N || d.K != ds[0].K ||
        (d.Ny > 0 && d.Ny != d.N))
      return false;
  }
  return true;
}

inline int prepare_gemm_tn16_multi(const std::vector<GemmTN16Desc>& ds, PreparedTN16Multi* out) {
  if (!tn16_multi_ok(ds)) return fail(-2, "gemm_tn16_multi: problems are not batchable");
  const GemmTN16Desc& d0 = ds[0];
  const int bn = tn16_block_n(d0.N);
  const int mt = (d0.M + kBlockM - 1) / kBlockM, nt = (d0.N + bn - 1) / bn;
  const int total_kb = (d0.K + kTN16BlockK - 1) / kTN16BlockK;
  int nsplit = num_sms() / (mt * nt);  // one CTA per SM
  if (nsplit < 1) nsplit = 1;
  if (nsplit > total_kb) nsplit = total_kb;
  const int kb_per_split = (total_kb + nsplit - 1) / nsplit;
  nsplit = (total_kb + kb_per_split - 1) / kb_per_split;
  PreparedTN16Multi pr;
  std::memset(&pr.params, 0, sizeof(pr.params));
  GemmTN16MultiParams& p = pr.params;
  int rc;
  for (size_t i = 0; i < ds.size(); ++i) {
    const GemmTN16Desc& d = ds[i];
    GemmTN16Prob& q = p.prob[i];
    for (int k = 0; k < d.npairs; ++k) {
      if (!d.X[k] || !d.Y[k]) return fail(-2, "gemm_tn16_multi: missing operand");
      if ((rc = encode_tmap_2d_bf16(&q.tmX[k], d.X[k], d.M, d.K, d.ldx[k], 64, kTN16BlockK, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
      if ((rc = encode_tmap_2d_bf16(&q.tmY[k], d.Y[k], d.N, d.K, d.ldy[k], 64, kTN16BlockK, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    }
    q.npairs = d.npairs; q.red_out = d.out; q.red_ld = d.ldo;
  }
  p.nprob = static_cast<int>(ds.size());
  p.M = d0.M; p.N = d0.N; p.K = d0.K; p.kb_per_split = kb_per_split;
  switch (bn) {
    case 64: pr.fn = reinterpret_cast<const void*>(&gemm_tn16_multi_kernel<64>); pr.smem = GemmTN16MultiConfig<64>::kSmemBytes; break;
    case 128: pr.fn = reinterpret_cast<const void*>(&gemm_tn16_multi_kernel<128>); pr.smem = GemmTN16MultiConfig<128>::kSmemBytes; break;
    default: pr.fn = reinterpret_cast<const void*>(&gemm_tn16_multi_kernel<256>); pr.smem = GemmTN16MultiConfig<256>::kSmemBytes; break;
  }
  pr.grid = dim3(nsplit, mt, nt);
  ARDAE_CUDA_OK(cudaFuncSetAttribute(pr.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, pr.smem));
  *out = pr;
  return 0;
}

inline int launch_prepared_tn16_multi(const PreparedTN16Multi& pr, cudaStream_t stream) {
  void* args[1] = {const_cast<GemmTN16MultiParams*>(&pr.params)};
  ARDAE_CUDA_OK(cudaLaunchKernel(pr.fn, pr.grid, dim3(kGemmThreads), args, pr.smem, stream));
  return 0;
}

}  // namespace ardae
