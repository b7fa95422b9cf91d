// tcgen05 / TMEM / TMA GEMM kernels (tf32 operands, fp32 accumulate) with the fused
// epilogues the AR-DAE path needs.  Two kernels:
//
//   gemm_nt_kernel : Out[M,N] = epi( alpha * A[M,K] . B[N,K]^T + bias terms ; aux1, aux2 )
//                    A, B row-major with K contiguous ("K-major"); B is an nn.Linear weight
//                    ([out,in]) for a forward layer, or a transposed copy for a backward layer.
//   gemm_tn_kernel : P[s][M,N] = sum_{k in split s} X0[k,M].Y0[k,N] (+ X1[k,M].Y1[k,N])
//                    both operands "MN-major" (reduction over the slow dimension: the rows of
//                    two activation arrays) - the weight-gradient contraction.
//
// Warp roles (192 threads): warp0 = TMA producer, warp1 = TMEM owner + MMA issuer,
// warps2..5 = epilogue (TMEM lane quarter = warp_id % 4).  Two CTAs are co-resident per SM
// (<= 113 KB smem, 256 TMEM columns each) so one CTA's epilogue overlaps the other's main loop.
#pragma once
#include "ptx_sm100.cuh"

namespace ardae {

enum EpiMode : int {
  EPI_LINEAR = 0,    // out = pre
  EPI_RELU = 1,      // out = max(pre, 0)
  EPI_SOFTPLUS = 2,  // out = softplus(pre)          (torch semantics, threshold 20)
  EPI_MUL_SIG = 3,   // out = pre * sig(aux1)        aux1 = softplus OUTPUT u, sig = 1 - exp(-u)
  EPI_MUL_STEP = 4,  // out = pre * (aux1 > 0)       relu backward, aux1 = relu output
  EPI_TANGENT = 5,   // out = pre * sig(aux1) ; out2 = aux2 * pre * (1 - sig(aux1))
  EPI_ADJOINT = 6,   // out = pre * sig(aux1) + aux2
  EPI_NUM_MODES = 7
};

constexpr int kBlockM = 128;
constexpr int kBlockK = 32;  // tf32 elements = one 128-byte swizzle row
constexpr int kUmmaK = 8;    // 32 bytes / sizeof(tf32)
constexpr int kTileBytes = kBlockM * 32 * 4;  // 128 rows x 32 floats staging / operand tile
constexpr int kGemmThreads = 192;
constexpr int kNumEpiStagingTiles = 6;

struct alignas(64) GemmNTParams {
  CUtensorMap tmA;     // dims {K, M}, box {32, 128}, SWIZZLE_128B
  CUtensorMap tmB;     // dims {K, N}, box {32, BLOCK_N}
  CUtensorMap tmAux1;  // dims {N, M}, box {32, 128}
  CUtensorMap tmAux2;
  CUtensorMap tmOut;
  CUtensorMap tmOut2;
  int M, N, K;
  float alpha;
  const float* bias;        // [N] or null
  const float* group_bias;  // [ceil(M/group), ldg] or null : row m uses row m/group
  int group, ldg;
  const float* row_scale;   // [M] or null   pre += row_scale[m] * col_vec[n]
  const float* col_vec;     // [N]
  float* colsum;            // [N] or null   colsum[n]  += colsum_scale * sum_m out[m,n]
  float* colsum_w;          // or null       colsum_w[n*colsum_w_stride] += sum_m out[m,n]*row_w[m]
  const float* row_w;       // [M]
  float* colsum2;           // [N] or null   colsum2[n] += sum_m out2[m,n]   (EPI_TANGENT only)
  float colsum_scale;
  int colsum_w_stride;
  int round_out;            // round stored outputs to tf32 (rna)
  int a_k_wrap;             // if > 0 the A operand's k coordinate wraps: col = (kb*32) % a_k_wrap
};

struct alignas(64) GemmTNParams {
  CUtensorMap tmX0;  // dims {Mx, K}, box {32, 32}
  CUtensorMap tmY0;  // dims {Ny, K}, box {32, 32}
  CUtensorMap tmX1;
  CUtensorMap tmY1;
  int M, N, K;       // output M x N, reduction length K (rows)
  int npairs;        // 1 or 2
  int kb_per_split;  // k-blocks (of 32 rows) handled by one split
  float* partial;    // [nsplit][Mpad][Npad], Mpad = gridDim.y*128, Npad = gridDim.z*BLOCK_N
};

__device__ __forceinline__ float softplus_f(float x) {
  // max(x,0) + log1p(exp(-|x|)); equals torch's thresholded softplus to fp32 precision.
  float e = __expf(-fabsf(x));
  float l = (e < 1e-4f) ? (e - 0.5f * e * e) : __logf(1.0f + e);
  return fmaxf(x, 0.0f) + l;
}
// sigmoid(a) given u = softplus(a):  1 - exp(-u)
__device__ __forceinline__ void sig_from_softplus(float u, float& s, float& one_minus_s) {
  float e = __expf(-u);
  one_minus_s = e;
  s = (u < 0.01f) ? u * (1.0f - u * (0.5f - u * (1.0f / 6.0f))) : 1.0f - e;
}

// After this, lane l holds sum over lanes of v[l] (recursive halving, 31 shuffles).
__device__ __forceinline__ float warp_transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      float send = up ? v[i] : v[i + off];
      float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

template <int BLOCK_N>
struct GemmNTConfig {
  static constexpr int kStageA = kBlockM * kBlockK * 4;
  static constexpr int kStageB = BLOCK_N * kBlockK * 4;
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kNumStages = (BLOCK_N >= 256) ? 2 : (BLOCK_N >= 128 ? 3 : 4);
  static constexpr int kPipeBytes = kStage * kNumStages;
  static constexpr int kEpiBytes = kNumEpiStagingTiles * kTileBytes;
  static constexpr int kDataBytes = kPipeBytes > kEpiBytes ? kPipeBytes : kEpiBytes;
  static constexpr int kSmemBytes = kDataBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int kTmemCols = BLOCK_N < 32 ? 32 : BLOCK_N;
};

// SPLIT (modes LINEAR / RELU / SOFTPLUS only): the fp32 result is stored as a tf32 pair
//   out = hi = rna(res), out2 = lo = rna(res - hi)
// so that a following "3xTF32" GEMM (A' = [hi | lo | hi] via a_k_wrap, B' = [W_hi | W_hi | W_lo])
// reproduces an fp32-accurate product on the tf32 tensor pipe.
template <int BLOCK_N, int MODE, bool SPLIT>
__global__ void __launch_bounds__(kGemmThreads, 2)
gemm_nt_kernel(const __grid_constant__ GemmNTParams p) {
  using Cfg = GemmNTConfig<BLOCK_N>;
  constexpr int NSTAGE = Cfg::kNumStages;
  constexpr bool kHasAux1 = MODE >= EPI_MUL_SIG;
  constexpr bool kHasAux2 = MODE >= EPI_TANGENT;
  constexpr bool kHasOut2 = (MODE == EPI_TANGENT) || SPLIT;
  static_assert(!SPLIT || MODE <= EPI_SOFTPLUS, "SPLIT only for plain activations");
  static_assert(BLOCK_N % 32 == 0 && BLOCK_N >= 32 && BLOCK_N <= 256, "BLOCK_N");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kDataBytes);
  uint64_t* empty_bar = full_bar + NSTAGE;
  uint64_t* tmem_full_bar = empty_bar + NSTAGE;
  uint64_t* aux_bar = tmem_full_bar + 1;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kBlockM;
  const int n0 = blockIdx.y * BLOCK_N;
  const int num_kb = (p.K + kBlockK - 1) / kBlockK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.tmA);
    ptx::prefetch_tmap(&p.tmB);
    ptx::prefetch_tmap(&p.tmOut);
    if (kHasAux1) ptx::prefetch_tmap(&p.tmAux1);
    if (kHasAux2) ptx::prefetch_tmap(&p.tmAux2);
    if (kHasOut2) ptx::prefetch_tmap(&p.tmOut2);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < NSTAGE; ++s) {
        ptx::mbar_init(&full_bar[s], 1);
        ptx::mbar_init(&empty_bar[s], 1);
      }
      ptx::mbar_init(tmem_full_bar, 1);
      ptx::mbar_init(&aux_bar[0], 1);
      ptx::mbar_init(&aux_bar[1], 1);
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
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % NSTAGE;
        const uint32_t ph = (kb / NSTAGE) & 1;
        ptx::mbar_wait(&empty_bar[s], ph ^ 1);
        ptx::mbar_expect_tx(&full_bar[s], Cfg::kStage);
        uint8_t* sa = smem + s * Cfg::kStage;
        int ka = kb * kBlockK;
        if (p.a_k_wrap > 0) ka %= p.a_k_wrap;
        ptx::tma_load_2d(sa, &p.tmA, &full_bar[s], ka, m0);
        ptx::tma_load_2d(sa + Cfg::kStageA, &p.tmB, &full_bar[s], kb * kBlockK, n0);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_tf32(kBlockM, BLOCK_N, 0, 0);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % NSTAGE;
        const uint32_t ph = (kb / NSTAGE) & 1;
        ptx::mbar_wait(&full_bar[s], ph);
        ptx::tc_fence_after();
        const uint32_t a_addr = ptx::smem_u32(smem + s * Cfg::kStage);
        const uint32_t b_addr = a_addr + Cfg::kStageA;
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
          const uint64_t adesc = ptx::make_smem_desc_sw128(a_addr + k * kUmmaK * 4, 0, 1024);
          const uint64_t bdesc = ptx::make_smem_desc_sw128(b_addr + k * kUmmaK * 4, 0, 1024);
          ptx::umma_tf32(tmem_base, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs retire
      }
      ptx::umma_commit(tmem_full_bar);  // accumulator complete (and every stage drained)
    }
  } else {
    // ------------------------------------------------------------ epilogue warps
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, +32) are accessible to this warp
    const int r = quarter * 32 + lane;  // row inside the tile
    const int m = m0 + r;
    const bool row_ok = m < p.M;
    const bool leader = (warp == 2 && lane == 0);
    constexpr int NCHUNK = BLOCK_N / 32;
    uint8_t* aux1_buf[2] = {smem + 0 * kTileBytes, smem + 1 * kTileBytes};
    uint8_t* aux2_buf[2] = {smem + 2 * kTileBytes, smem + 3 * kTileBytes};
    uint8_t* out_buf = smem + 4 * kTileBytes;
    uint8_t* out2_buf = smem + 5 * kTileBytes;
    constexpr uint32_t kAuxBytes = kTileBytes * ((kHasAux1 ? 1 : 0) + (kHasAux2 ? 1 : 0));

    ptx::mbar_wait(tmem_full_bar, 0);
    ptx::tc_fence_after();

    if (kHasAux1 && leader) {
      ptx::mbar_expect_tx(&aux_bar[0], kAuxBytes);
      ptx::tma_load_2d(aux1_buf[0], &p.tmAux1, &aux_bar[0], n0, m0);
      if (kHasAux2) ptx::tma_load_2d(aux2_buf[0], &p.tmAux2, &aux_bar[0], n0, m0);
    }
    const float rs = (p.row_scale != nullptr && row_ok) ? p.row_scale[m] : 0.0f;
    const float rw = (p.colsum_w != nullptr && row_ok) ? p.row_w[m] : 0.0f;
    const float* gb_row =
        (p.group_bias != nullptr) ? p.group_bias + static_cast<size_t>((row_ok ? m : 0) / p.group) * p.ldg
                                  : nullptr;
    const int swz = r & 7;
    const uint32_t row_off = static_cast<uint32_t>(r) * 128u;

#pragma unroll 1
    for (int c = 0; c < NCHUNK; ++c) {
      const int nc = n0 + c * 32;
      if (nc >= p.N) break;  // uniform across the CTA
      if (leader) ptx::tma_store_wait_read<0>();
      ptx::named_bar_sync(1, 128);
      if (kHasAux1 && leader && (c + 1 < NCHUNK) && (nc + 32 < p.N)) {
        const int b = (c + 1) & 1;
        ptx::mbar_expect_tx(&aux_bar[b], kAuxBytes);
        ptx::tma_load_2d(aux1_buf[b], &p.tmAux1, &aux_bar[b], nc + 32, m0);
        if (kHasAux2) ptx::tma_load_2d(aux2_buf[b], &p.tmAux2, &aux_bar[b], nc + 32, m0);
      }
      uint32_t accu[32];
      ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + c * 32, accu);
      if (kHasAux1) ptx::mbar_wait(&aux_bar[c & 1], (c >> 1) & 1);
      ptx::tmem_ld_wait();

      float v[32];  // becomes `out` in place
      const uint8_t* a1 = aux1_buf[c & 1] + row_off;
      const uint8_t* a2 = aux2_buf[c & 1] + row_off;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t soff = static_cast<uint32_t>((q ^ swz) << 4);
        float4 x1 = make_float4(0.f, 0.f, 0.f, 0.f), x2 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kHasAux1) x1 = *reinterpret_cast<const float4*>(a1 + soff);
        if (kHasAux2) x2 = *reinterpret_cast<const float4*>(a2 + soff);
        const float aux1v[4] = {x1.x, x1.y, x1.z, x1.w};
        const float aux2v[4] = {x2.x, x2.y, x2.z, x2.w};
        float o[4], o2[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = nc + q * 4 + j;
          const int colc = col < p.N ? col : p.N - 1;
          float pre = p.alpha * __uint_as_float(accu[q * 4 + j]);
          if (p.bias != nullptr) pre += __ldg(p.bias + colc);
          if (gb_row != nullptr) pre += __ldg(gb_row + colc);
          if (p.row_scale != nullptr) pre += rs * __ldg(p.col_vec + colc);
          float res, res2 = 0.0f;
          if (MODE == EPI_LINEAR) {
            res = pre;
          } else if (MODE == EPI_RELU) {
            res = fmaxf(pre, 0.0f);
          } else if (MODE == EPI_SOFTPLUS) {
            res = softplus_f(pre);
          } else if (MODE == EPI_MUL_STEP) {
            res = aux1v[j] > 0.0f ? pre : 0.0f;
          } else {
            float s, oms;
            sig_from_softplus(aux1v[j], s, oms);
            if (MODE == EPI_MUL_SIG) {
              res = pre * s;
            } else if (MODE == EPI_TANGENT) {
              res = pre * s;
              res2 = aux2v[j] * pre * oms;
            } else {  // EPI_ADJOINT
              res = pre * s + aux2v[j];
            }
          }
          if (SPLIT) {
            const float hi = ptx::round_tf32(res);
            res2 = ptx::round_tf32(res - hi);
            res = hi;
          } else if (p.round_out) {
            res = ptx::round_tf32(res);
            res2 = ptx::round_tf32(res2);
          }
          o[j] = res;
          o2[j] = res2;
          v[q * 4 + j] = (row_ok && col < p.N) ? res : 0.0f;
        }
        *reinterpret_cast<float4*>(out_buf + row_off + soff) = make_float4(o[0], o[1], o[2], o[3]);
        if (kHasOut2)
          *reinterpret_cast<float4*>(out2_buf + row_off + soff) =
              make_float4(o2[0], o2[1], o2[2], o2[3]);
      }
      ptx::fence_proxy_async_smem();
      ptx::named_bar_sync(2, 128);
      if (leader) {
        ptx::tma_store_2d(&p.tmOut, out_buf, nc, m0);
        if (kHasOut2) ptx::tma_store_2d(&p.tmOut2, out2_buf, nc, m0);
        ptx::tma_store_commit();
      }
      if (p.colsum_w != nullptr) {
        float w[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) w[i] = v[i] * rw;
        const float t = warp_transpose_reduce32(w, lane);
        if (nc + lane < p.N)
          atomicAdd(p.colsum_w + static_cast<size_t>(nc + lane) * p.colsum_w_stride, t);
      }
      if (p.colsum != nullptr) {
        const float t = warp_transpose_reduce32(v, lane);
        if (nc + lane < p.N) atomicAdd(p.colsum + nc + lane, p.colsum_scale * t);
      }
      if (MODE == EPI_TANGENT && p.colsum2 != nullptr) {
        // re-read this thread's own out2 row from the staging tile (still intact until the
        // next chunk's barrier) instead of keeping 32 more registers live
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 t4 =
              *reinterpret_cast<const float4*>(out2_buf + row_off + ((q ^ swz) << 4));
          const bool ok = row_ok;
          v[q * 4 + 0] = (ok && nc + q * 4 + 0 < p.N) ? t4.x : 0.0f;
          v[q * 4 + 1] = (ok && nc + q * 4 + 1 < p.N) ? t4.y : 0.0f;
          v[q * 4 + 2] = (ok && nc + q * 4 + 2 < p.N) ? t4.z : 0.0f;
          v[q * 4 + 3] = (ok && nc + q * 4 + 3 < p.N) ? t4.w : 0.0f;
        }
        const float t = warp_transpose_reduce32(v, lane);
        if (nc + lane < p.N) atomicAdd(p.colsum2 + nc + lane, t);
      }
    }
    if (leader) ptx::tma_store_wait_all<0>();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------
// Weight-gradient contraction: both operands MN-major (TMA boxes of 32 rows x 32 floats).
template <int BLOCK_N>
struct GemmTNConfig {
  static constexpr int kBoxBytes = 32 * 32 * 4;  // 4 KB: 32 k-rows x 32 floats (128 B rows)
  static constexpr int kStageA = (kBlockM / 32) * kBoxBytes;
  static constexpr int kStageB = (BLOCK_N / 32) * kBoxBytes;
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kNumStages = (BLOCK_N >= 256) ? 2 : (BLOCK_N >= 128 ? 3 : 4);
  static constexpr int kDataBytes = kStage * kNumStages;
  static constexpr int kSmemBytes = kDataBytes + 1024 + 256;
  static constexpr int kTmemCols = BLOCK_N < 32 ? 32 : BLOCK_N;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(kGemmThreads, 2)
gemm_tn_kernel(const __grid_constant__ GemmTNParams p) {
  using Cfg = GemmTNConfig<BLOCK_N>;
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
  const int total_kb = (p.K + kBlockK - 1) / kBlockK;
  const int kb_begin = split * p.kb_per_split;
  int kb_end = kb_begin + p.kb_per_split;
  if (kb_end > total_kb) kb_end = total_kb;
  const int nkb = kb_end > kb_begin ? kb_end - kb_begin : 0;
  const int iters = nkb * p.npairs;  // pair-major: all k-blocks of pair 0, then pair 1

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.tmX0);
    ptx::prefetch_tmap(&p.tmY0);
    if (p.npairs > 1) {
      ptx::prefetch_tmap(&p.tmX1);
      ptx::prefetch_tmap(&p.tmY1);
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
    if (lane == 0) {
      for (int it = 0; it < iters; ++it) {
        const int s = it % NSTAGE;
        const uint32_t ph = (it / NSTAGE) & 1;
        const int pair = it / nkb;
        const int k0 = (kb_begin + (it - pair * nkb)) * kBlockK;
        const CUtensorMap* tx = pair == 0 ? &p.tmX0 : &p.tmX1;
        const CUtensorMap* ty = pair == 0 ? &p.tmY0 : &p.tmY1;
        ptx::mbar_wait(&empty_bar[s], ph ^ 1);
        ptx::mbar_expect_tx(&full_bar[s], Cfg::kStage);
        uint8_t* sa = smem + s * Cfg::kStage;
        uint8_t* sb = sa + Cfg::kStageA;
#pragma unroll
        for (int b = 0; b < kBlockM / 32; ++b)
          ptx::tma_load_2d(sa + b * Cfg::kBoxBytes, tx, &full_bar[s], m0 + b * 32, k0);
#pragma unroll
        for (int b = 0; b < BLOCK_N / 32; ++b)
          ptx::tma_load_2d(sb + b * Cfg::kBoxBytes, ty, &full_bar[s], n0 + b * 32, k0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_tf32(kBlockM, BLOCK_N, 1, 1);
      for (int it = 0; it < iters; ++it) {
        const int s = it % NSTAGE;
        const uint32_t ph = (it / NSTAGE) & 1;
        ptx::mbar_wait(&full_bar[s], ph);
        ptx::tc_fence_after();
        const uint32_t a_addr = ptx::smem_u32(smem + s * Cfg::kStage);
        const uint32_t b_addr = a_addr + Cfg::kStageA;
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
          // MN-major tf32 must use the 128B swizzle with 32-byte atoms (layout type 1; TMA mode
          // SWIZZLE_128B_ATOM_32B): atom = 32 floats (MN) x 4 k-rows.  LBO = stride between
          // MN atoms (one TMA box of 32 k-rows), SBO = stride between 4-row k groups.
          const uint64_t adesc = ptx::make_smem_desc(a_addr + k * 1024, Cfg::kBoxBytes, 512, 1);
          const uint64_t bdesc = ptx::make_smem_desc(b_addr + k * 1024, Cfg::kBoxBytes, 512, 1);
          ptx::umma_tf32(tmem_base, adesc, bdesc, idesc, (it | k) != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&empty_bar[s]);
      }
      ptx::umma_commit(tmem_full_bar);
    }
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int Mpad = gridDim.y * kBlockM;
    const int Npad = gridDim.z * BLOCK_N;
    float* dst = p.partial + (static_cast<size_t>(split) * Mpad + (m0 + r)) * Npad + n0;
    if (iters > 0) {
      ptx::mbar_wait(tmem_full_bar, 0);
      ptx::tc_fence_after();
    }
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
      for (int q = 0; q < 8; ++q) {
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

// out[m*ld + n] = beta*out + scale * sum_s partial[s][m][n]   (m < M, n < N)
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int nsplit, int Mpad,
                                     int Npad, float* __restrict__ out, int M, int N, int ld,
                                     float scale, float beta) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * N) return;
  const int m = idx / N, n = idx - m * N;
  const size_t plane = static_cast<size_t>(Mpad) * Npad;
  const float* src = partial + static_cast<size_t>(m) * Npad + n;
  float acc = 0.0f;
  for (int s = 0; s < nsplit; ++s) acc += src[s * plane];
  float* o = out + static_cast<size_t>(m) * ld + n;
  *o = (beta != 0.0f ? beta * (*o) : 0.0f) + scale * acc;
}

}  // namespace ardae
