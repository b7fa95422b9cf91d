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
  int debug_flags;          // profiling experiments only: 1 = skip TMA stores, 2 = skip epilogue math
  int vec_ok;               // bias / group_bias / col_vec are 16-byte aligned: float4 broadcast loads
  int prefetch_tiles;       // > 0: prefetch the A / aux tiles of row tile (this + prefetch_tiles) into L2
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
  float* red_out;    // non-null: every split adds its tile straight into out[M, N] (pitch red_ld) with
  int red_ld;        //           red.global.add (no partial buffer, no reduce launch; summation order not fixed)
  int red_vec;       // rows of red_out are 16-byte aligned: red.global.add.v4.f32
  int npad;          // padded N of the partial buffer (n tiles x BLOCK_N)
  long long* dbg;    // profiling only: per CTA [4] clock64 stamps (entry, accumulator complete, epilogue done, -)
  int pdl;           // launched with programmatic stream serialization (see gemm_tn_body)
};

// Batched launch: grid.z selects one of up to kMaxTNBatch same-shape contractions (the 2L-1 [H,H] weight gradients of
// a CDAE update).  One launch = one ramp-up and one tail for all of them; CTAs of different problems co-reside, so
// one CTA's prologue / red.add epilogue overlaps another's main loop and HBM keeps streaming.
constexpr int kMaxTNBatch = 10;
struct alignas(64) GemmTNBatchParams {
  GemmTNParams prob[kMaxTNBatch];
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

// ---- lean epilogue arithmetic: the chain epilogues are instruction-issue bound (measured with clock64 stamps:
// ~2700 cycles per 32-column chunk with __expf/__logf range fix-ups, cvt.rna Inf checks and generic-space LD/ST),
// so: raw MUFU ex2/lg2, integer round-to-nearest to tf32, explicit shared-space vector accesses.
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_ftz(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// round-to-nearest (ties away) to tf32 on finite inputs: 2 integer ops instead of cvt.rna's Inf/NaN handling
__device__ __forceinline__ float round_tf32_fast(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
// sigmoid(a) from u = softplus(a) >= 0:  s = 1 - exp(-u)  (u(1 - u/2) below 2^-7: relative error < 1e-5), 1 - s = exp(-u)
__device__ __forceinline__ void sig_fast(float u, float& s, float& one_minus_s) {
  const float e = ex2_ftz(u * -1.4426950408889634f);
  one_minus_s = e;
  s = (u < 0.0078125f) ? fmaf(-0.5f * u, u, u) : 1.0f - e;
}
// max(x,0) + log1p(exp(-|x|)) == torch softplus (threshold 20) to fp32 rounding
__device__ __forceinline__ float softplus_fast(float x) {
  const float e = ex2_ftz(fabsf(x) * -1.4426950408889634f);
  const float l = (e < 1e-4f) ? fmaf(-0.5f * e, e, e) : lg2_ftz(1.0f + e) * 0.6931471805599453f;
  return fmaxf(x, 0.0f) + l;
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
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


// add[j] += scale * src[j], j < 32 (all lanes read the same addresses: broadcast loads)
__device__ __forceinline__ void add_cols32(const float* __restrict__ src, float scale, bool vec,
                                           int remaining, float (&add)[32]) {
  if (vec) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(src) + q);
      add[q * 4 + 0] = fmaf(scale, t.x, add[q * 4 + 0]);
      add[q * 4 + 1] = fmaf(scale, t.y, add[q * 4 + 1]);
      add[q * 4 + 2] = fmaf(scale, t.z, add[q * 4 + 2]);
      add[q * 4 + 3] = fmaf(scale, t.w, add[q * 4 + 3]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < remaining) add[j] = fmaf(scale, __ldg(src + j), add[j]);
  }
}

// One 32-column chunk of the fused epilogue for the thread that owns one accumulator row.
//   accu: the 32 fp32 accumulator columns; a1/a2: this row of the aux staging tiles (swizzled 16-byte
//   chunks); o1/o2: this row of the out staging tiles.  v[] returns the (unmasked) `out` values.
// Kept lean on purpose: the epilogue is instruction-issue bound (round-1 ncu: 1500 warp-instructions
// per chunk with per-element predicated bias loads; now ~1 FFMA + the activation per element).
template <int MODE, bool SPLIT, bool ROUND>
__device__ __forceinline__ void epilogue_chunk_body(const GemmNTParams& p, const uint32_t (&accu)[32],
                                                    const float (&add)[32], const uint8_t* a1,
                                                    const uint8_t* a2, uint8_t* o1, uint8_t* o2, int swz,
                                                    float (&v)[32]) {
  constexpr bool kHasAux1 = MODE >= EPI_MUL_SIG;
  constexpr bool kHasAux2 = MODE >= EPI_TANGENT;
  constexpr bool kHasOut2 = (MODE == EPI_TANGENT) || SPLIT;
  const float alpha = p.alpha;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const uint32_t soff = static_cast<uint32_t>((q ^ swz) << 4);
    float4 x1 = make_float4(0.f, 0.f, 0.f, 0.f), x2 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kHasAux1) x1 = *reinterpret_cast<const float4*>(a1 + soff);
    if (kHasAux2) x2 = *reinterpret_cast<const float4*>(a2 + soff);
    const float aux1v[4] = {x1.x, x1.y, x1.z, x1.w};
    const float aux2v[4] = {x2.x, x2.y, x2.z, x2.w};
    float o[4], ob[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float pre = fmaf(alpha, __uint_as_float(accu[q * 4 + j]), add[q * 4 + j]);
      float res, res2 = 0.0f;
      if (MODE == EPI_LINEAR) {
        res = pre;
      } else if (MODE == EPI_RELU) {
        res = fmaxf(pre, 0.0f);
      } else if (MODE == EPI_SOFTPLUS) {
        res = softplus_fast(pre);
      } else if (MODE == EPI_MUL_STEP) {
        res = aux1v[j] > 0.0f ? pre : 0.0f;
      } else {
        float sg, oms;
        sig_fast(aux1v[j], sg, oms);
        if (MODE == EPI_MUL_SIG) {
          res = pre * sg;
        } else if (MODE == EPI_TANGENT) {
          res = pre * sg;
          res2 = aux2v[j] * pre * oms;
        } else {  // EPI_ADJOINT
          res = fmaf(pre, sg, aux2v[j]);
        }
      }
      if (SPLIT) {
        const float hi = round_tf32_fast(res);
        res2 = round_tf32_fast(res - hi);
        res = hi;
      } else if (ROUND) {
        res = round_tf32_fast(res);
        if (kHasOut2) res2 = round_tf32_fast(res2);
      }
      o[j] = res;
      ob[j] = res2;
      v[q * 4 + j] = res;
    }
    *reinterpret_cast<float4*>(o1 + soff) = make_float4(o[0], o[1], o[2], o[3]);
    if (kHasOut2) *reinterpret_cast<float4*>(o2 + soff) = make_float4(ob[0], ob[1], ob[2], ob[3]);
  }
}

template <int MODE, bool SPLIT>
__device__ __forceinline__ void epilogue_chunk(const GemmNTParams& p, const uint32_t (&accu)[32],
                                               const uint8_t* a1, const uint8_t* a2, uint8_t* o1,
                                               uint8_t* o2, int swz, int nc, float rs, const float* gb_row,
                                               float (&v)[32]) {
  float add[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) add[j] = 0.0f;
  const bool vec = p.vec_ok && (nc + 32 <= p.N);
  const int remaining = p.N - nc;
  if (p.bias != nullptr) add_cols32(p.bias + nc, 1.0f, vec, remaining, add);
  if (gb_row != nullptr) add_cols32(gb_row + nc, 1.0f, vec, remaining, add);
  if (p.row_scale != nullptr) add_cols32(p.col_vec + nc, rs, vec, remaining, add);
  if (SPLIT || p.round_out)
    epilogue_chunk_body<MODE, SPLIT, true>(p, accu, add, a1, a2, o1, o2, swz, v);
  else
    epilogue_chunk_body<MODE, SPLIT, false>(p, accu, add, a1, a2, o1, o2, swz, v);
}

// Column sums of one chunk (see GemmNTParams::colsum*).  o2 = this row of the out2 staging tile.
template <int MODE>
__device__ __forceinline__ void epilogue_colsums(const GemmNTParams& p, float (&v)[32], const uint8_t* o2,
                                                 int swz, int nc, bool row_ok, float rw, int lane) {
  if (p.colsum == nullptr && p.colsum_w == nullptr && (MODE != EPI_TANGENT || p.colsum2 == nullptr)) return;
  const bool full = nc + 32 <= p.N;
  if (!row_ok || !full) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = (row_ok && nc + i < p.N) ? v[i] : 0.0f;
  }
  if (p.colsum_w != nullptr) {
    float w[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) w[i] = v[i] * rw;
    const float t = warp_transpose_reduce32(w, lane);
    if (nc + lane < p.N) atomicAdd(p.colsum_w + static_cast<size_t>(nc + lane) * p.colsum_w_stride, t);
  }
  if (p.colsum != nullptr) {
    const float t = warp_transpose_reduce32(v, lane);
    if (nc + lane < p.N) atomicAdd(p.colsum + nc + lane, p.colsum_scale * t);
  }
  if (MODE == EPI_TANGENT && p.colsum2 != nullptr) {
    // re-read this thread's own out2 row from the staging tile instead of keeping 32 more registers live
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 t4 = *reinterpret_cast<const float4*>(o2 + ((q ^ swz) << 4));
      v[q * 4 + 0] = (row_ok && nc + q * 4 + 0 < p.N) ? t4.x : 0.0f;
      v[q * 4 + 1] = (row_ok && nc + q * 4 + 1 < p.N) ? t4.y : 0.0f;
      v[q * 4 + 2] = (row_ok && nc + q * 4 + 2 < p.N) ? t4.z : 0.0f;
      v[q * 4 + 3] = (row_ok && nc + q * 4 + 3 < p.N) ? t4.w : 0.0f;
    }
    const float t = warp_transpose_reduce32(v, lane);
    if (nc + lane < p.N) atomicAdd(p.colsum2 + nc + lane, t);
  }
}

// DEEP: latency-optimised variant for grids that cannot fill the GPU (the B-row chains: context
// branch, encode(std=0), score, model forward/backward).  Narrow tiles (more CTAs, 1-2 epilogue chunks)
// and 8 TMA stages so the whole K loop of an L2-resident problem is in flight at once; one CTA per SM.
// CG2: CTA-pair (cta_group::2) variant: two CTAs on the SMs of one TPC compute a 256 x BLOCK_N tile;
// each stages its own 128 rows of A and HALF of the weight k-block, so only half of the weight bytes
// enter each SM (the NT kernels are bound by the ~68 GB/s per-SM TMA/L2 port, round-1 measurement).
template <int BLOCK_N, bool DEEP = false, bool CG2 = false>
struct GemmNTConfig {
  static constexpr int kStageA = kBlockM * kBlockK * 4;
  static constexpr int kStageB = (CG2 ? BLOCK_N / 2 : BLOCK_N) * kBlockK * 4;
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kNumStages = DEEP ? 8 : (CG2 ? 3 : ((BLOCK_N >= 256) ? 2 : (BLOCK_N >= 128 ? 3 : 4)));
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
// MC: launched as clusters of 2 CTAs (adjacent row tiles).  Each CTA fetches half of every weight
// (B) k-block and TMA-multicasts it to both, halving the L2->SM weight traffic that bounds the
// 3xTF32 forward sweep (768 KB of weights per 128-row tile otherwise).
template <int BLOCK_N, int MODE, bool SPLIT, bool MC, bool DEEP = false, bool CG2 = false>
__global__ void __launch_bounds__(kGemmThreads, 2)
gemm_nt_kernel(const __grid_constant__ GemmNTParams p) {
  using Cfg = GemmNTConfig<BLOCK_N, DEEP, CG2>;
  static_assert(!(MC && CG2), "MC and CG2 are exclusive");
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
  uint64_t* aux_bar = tmem_full_bar + 1;  // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_bar + 4);

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
        ptx::mbar_init(&empty_bar[s], MC ? 2 : 1);  // MC: both CTAs' MMAs release a stage
      }
      ptx::mbar_init(tmem_full_bar, 1);
      for (int a = 0; a < 4; ++a) ptx::mbar_init(&aux_bar[a], 1);
      ptx::fence_mbar_init();
    }
    __syncwarp();
    if (CG2) {
      ptx::tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols);
      ptx::tmem_relinquish_2sm();
    } else {
      ptx::tmem_alloc(tmem_slot, Cfg::kTmemCols);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (MC || CG2) ptx::cluster_sync_all();  // the peer's barriers exist before anything remote touches them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t pair_rank = (MC || CG2) ? ptx::cluster_ctarank() : 0;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (ptx::elect_one()) {  // one elected lane: tcgen05 / TMA issue stays on the uniform datapath
      const uint32_t rank = pair_rank;
      if (p.prefetch_tiles > 0) {
        // the CTA that will run on this SM slot a full wave later streams its operands from HBM: pull them
        // into L2 now so its TMA loads see L2-hit latency instead of a loaded-HBM round trip
        const int mp = m0 + p.prefetch_tiles * kBlockM;
        if (mp < p.M) {
          const int nkb_a = p.a_k_wrap > 0 ? p.a_k_wrap / kBlockK : num_kb;
          for (int kb = 0; kb < nkb_a; ++kb) ptx::tma_prefetch_l2_2d(&p.tmA, kb * kBlockK, mp);
          if (kHasAux1 || kHasAux2) {
            for (int c = 0; c * 32 < BLOCK_N && n0 + c * 32 < p.N; ++c) {
              if (kHasAux1) ptx::tma_prefetch_l2_2d(&p.tmAux1, n0 + c * 32, mp);
              if (kHasAux2) ptx::tma_prefetch_l2_2d(&p.tmAux2, n0 + c * 32, mp);
            }
          }
        }
      }
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % NSTAGE;
        const uint32_t ph = (kb / NSTAGE) & 1;
        ptx::mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * Cfg::kStage;
        int ka = kb * kBlockK;
        if (p.a_k_wrap > 0) ka %= p.a_k_wrap;
        if (CG2) {
          // both CTAs' tiles complete on the LEADER's full barrier, armed by the leader for both
          if (rank == 0) ptx::mbar_expect_tx(&full_bar[s], 2 * Cfg::kStage);
          ptx::tma_load_2d_2sm(sa, &p.tmA, &full_bar[s], ka, m0);
          ptx::tma_load_2d_2sm(sa + Cfg::kStageA, &p.tmB, &full_bar[s], kb * kBlockK,
                               n0 + static_cast<int>(rank) * (BLOCK_N / 2));
          continue;
        }
        ptx::mbar_expect_tx(&full_bar[s], Cfg::kStage);
        ptx::tma_load_2d(sa, &p.tmA, &full_bar[s], ka, m0);
        if (MC) {
          // this CTA's half of the weight k-block (box = BLOCK_N/2 rows), delivered to both CTAs
          constexpr int kHalfRows = BLOCK_N / 2;
          ptx::tma_load_2d_mc(sa + Cfg::kStageA + rank * (kHalfRows * kBlockK * 4), &p.tmB, &full_bar[s],
                              kb * kBlockK, n0 + static_cast<int>(rank) * kHalfRows, 0x3);
        } else {
          ptx::tma_load_2d(sa + Cfg::kStageA, &p.tmB, &full_bar[s], kb * kBlockK, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if ((!CG2 || pair_rank == 0) && ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_tf32(CG2 ? 2 * kBlockM : kBlockM, BLOCK_N, 0, 0);
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
          if (CG2)
            ptx::umma_tf32_2sm(tmem_base, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          else
            ptx::umma_tf32(tmem_base, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        if (CG2)
          ptx::umma_commit_2sm(&empty_bar[s], 0x3);  // frees the stage in both CTAs of the pair
        else if (MC)
          ptx::umma_commit_mc(&empty_bar[s], 0x3);  // frees this stage in BOTH CTAs' bookkeeping
        else
          ptx::umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs retire
      }
      if (CG2)
        ptx::umma_commit_2sm(tmem_full_bar, 0x3);  // both CTAs' epilogues may read their TMEM halves
      else
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
    // The six 16 KB staging tiles alias the (drained) main-loop stages:
    //   aux ring: NAUX slots of (aux1 [+ aux2]); out ring: NOUT tiles (+ NOUT for out2)
    constexpr int kAuxSlot = (kHasAux1 ? kTileBytes : 0) + (kHasAux2 ? kTileBytes : 0);
    constexpr int NAUX = !kHasAux1 ? 0 : (kHasAux2 ? 2 : 4);
    constexpr int NOUT = (MODE == EPI_TANGENT) ? 1 : (kHasAux1 ? 2 : (kHasOut2 ? 3 : 6));
    static_assert(NAUX * kAuxSlot + NOUT * kTileBytes * (kHasOut2 ? 2 : 1) <= kNumEpiStagingTiles * kTileBytes, "staging");
    uint8_t* aux_base = smem;
    uint8_t* out_base = smem + NAUX * kAuxSlot;
    uint8_t* out2_base = out_base + NOUT * kTileBytes;
    int nchunks = (p.N - n0 + 31) / 32;
    if (nchunks > NCHUNK) nchunks = NCHUNK;

    ptx::mbar_wait(tmem_full_bar, 0);
    ptx::tc_fence_after();

    if (kHasAux1 && leader) {
      for (int c = 0; c < NAUX && c < nchunks; ++c) {
        ptx::mbar_expect_tx(&aux_bar[c], kAuxSlot);
        ptx::tma_load_2d(aux_base + c * kAuxSlot, &p.tmAux1, &aux_bar[c], n0 + c * 32, m0);
        if (kHasAux2)
          ptx::tma_load_2d(aux_base + c * kAuxSlot + kTileBytes, &p.tmAux2, &aux_bar[c], n0 + c * 32, m0);
      }
    }
    const float rs = (p.row_scale != nullptr && row_ok) ? p.row_scale[m] : 0.0f;
    const float rw = (p.colsum_w != nullptr && row_ok) ? p.row_w[m] : 0.0f;
    const float* gb_row =
        (p.group_bias != nullptr) ? p.group_bias + static_cast<size_t>((row_ok ? m : 0) / p.group) * p.ldg
                                  : nullptr;
    const int swz = r & 7;
    const uint32_t row_off = static_cast<uint32_t>(r) * 128u;

#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
      const int nc = n0 + c * 32;
      uint32_t accu[32];
      ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + c * 32, accu);
      const int a = NAUX > 0 ? c % (NAUX > 0 ? NAUX : 1) : 0;
      if (kHasAux1) ptx::mbar_wait(&aux_bar[a], (c / (NAUX > 0 ? NAUX : 1)) & 1);
      ptx::tmem_ld_wait();
      const uint8_t* slot = aux_base + a * kAuxSlot;
      uint8_t* ob = out_base + (c % NOUT) * kTileBytes;
      uint8_t* ob2 = out2_base + (c % NOUT) * kTileBytes;
      if (NOUT == 1) {
        // single staging tile: the previous store must have drained it before anyone writes
        if (leader) ptx::tma_store_wait_read<0>();
        ptx::named_bar_sync(1, 128);
      }
      float v[32];  // `out` values of this thread's row (masked), for the column sums
      if (!(p.debug_flags & 2))
        epilogue_chunk<MODE, SPLIT>(p, accu, slot + row_off, slot + kTileBytes + row_off, ob + row_off,
                                    ob2 + row_off, swz, nc, rs, gb_row, v);
      ptx::fence_proxy_async_smem();
      // ring of NOUT >= 2 tiles: the tile chunk c+1 will write was last read by store(c+1-NOUT);
      // the leader waits for it BEFORE the barrier, so passing the barrier publishes "free"
      if (NOUT >= 2 && leader) ptx::tma_store_wait_read<(NOUT >= 2 ? NOUT - 2 : 0)>();
      ptx::named_bar_sync(2, 128);
      if (leader) {
        if (!(p.debug_flags & 1)) {
          ptx::tma_store_2d(&p.tmOut, ob, nc, m0);
          if (kHasOut2) ptx::tma_store_2d(&p.tmOut2, ob2, nc, m0);
        }
        ptx::tma_store_commit();
        if (kHasAux1 && c + NAUX < nchunks) {  // every thread is past its reads of slot a
          ptx::mbar_expect_tx(&aux_bar[a], kAuxSlot);
          ptx::tma_load_2d(aux_base + a * kAuxSlot, &p.tmAux1, &aux_bar[a], nc + NAUX * 32, m0);
          if (kHasAux2)
            ptx::tma_load_2d(aux_base + a * kAuxSlot + kTileBytes, &p.tmAux2, &aux_bar[a], nc + NAUX * 32, m0);
        }
      }
      epilogue_colsums<MODE>(p, v, ob2 + row_off, swz, nc, row_ok, rw, lane);
    }
    if (leader) ptx::tma_store_wait_all<0>();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (MC || CG2) ptx::cluster_sync_all();  // no CTA exits while its peer may still signal its barriers
  if (warp == 1) {
    ptx::tc_fence_after();
    if (CG2)
      ptx::tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
    else
      ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}


// ------------------------------------------------------------------------------------------
// Persistent variant for the hot shape (N <= 256, many row tiles): one CTA per SM loops over row
// tiles; the accumulator is double-buffered in TMEM (2 x 256 columns) so the epilogue of tile i
// overlaps the TMA/MMA main loop of tile i+1; aux tiles arrive through their own TMA ring fed by a
// dedicated producer warp; out tiles are double-buffered and leave through TMA stores.
// Warps: 0 = A/B TMA producer, 1 = MMA issuer + TMEM owner, 2 = aux TMA producer, 3 = idle,
//        4..7 = epilogue (TMEM lane quarter = warp & 3).
template <int MODE, bool SPLIT>
struct GemmNTPersistConfig {
  static constexpr bool kHasAux1 = MODE >= EPI_MUL_SIG;
  static constexpr bool kHasAux2 = MODE >= EPI_TANGENT;
  static constexpr bool kHasOut2 = (MODE == EPI_TANGENT) || SPLIT;
  static constexpr int kBlockN = 256;
  static constexpr int kStageA = kBlockM * kBlockK * 4;
  static constexpr int kStageB = kBlockN * kBlockK * 4;
  static constexpr int kStage = kStageA + kStageB;                       // 48 KB
  static constexpr int kNumStages = 2;
  static constexpr int kAuxSlotBytes = (kHasAux1 ? kTileBytes : 0) + (kHasAux2 ? kTileBytes : 0);
  static constexpr int kNumAux = !kHasAux1 ? 0 : (kHasAux2 ? 2 : 4);
  // ring of out staging tiles: as deep as the 224 KB budget allows, so several TMA stores stay in flight
  static constexpr int kNumOut = (MODE == EPI_TANGENT) ? 2 : (kHasAux1 ? 4 : (kHasOut2 ? 4 : 8));
  static constexpr int kOutBytes = kNumOut * kTileBytes * (kHasOut2 ? 2 : 1);
  static constexpr int kOffAux = kNumStages * kStage;
  static constexpr int kOffOut = kOffAux + kNumAux * kAuxSlotBytes;
  static constexpr int kDataBytes = kOffOut + kOutBytes;
  static constexpr int kSmemBytes = kDataBytes + 1024 + 512;
  static constexpr int kThreads = 256;
};

template <int MODE, bool SPLIT>
__global__ void __launch_bounds__(256, 1)
gemm_nt_persist_kernel(const __grid_constant__ GemmNTParams p) {
  using Cfg = GemmNTPersistConfig<MODE, SPLIT>;
  constexpr int NSTAGE = Cfg::kNumStages;
  constexpr int NAUX = Cfg::kNumAux;
  constexpr int NAUXD = NAUX > 0 ? NAUX : 1;  // divisor in code that only runs when NAUX > 0
  constexpr bool kHasAux1 = Cfg::kHasAux1, kHasAux2 = Cfg::kHasAux2, kHasOut2 = Cfg::kHasOut2;
  constexpr int NCHUNK = 8;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kDataBytes);
  uint64_t* empty_bar = full_bar + NSTAGE;
  uint64_t* tmem_full = empty_bar + NSTAGE;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;      // [2]
  uint64_t* aux_full = tmem_empty + 2;       // [NAUX]
  uint64_t* aux_empty = aux_full + (NAUX > 0 ? NAUX : 1);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_empty + (NAUX > 0 ? NAUX : 1));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = (p.K + kBlockK - 1) / kBlockK;
  const int num_tiles = (p.M + kBlockM - 1) / kBlockM;

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
      for (int b = 0; b < 2; ++b) {
        ptx::mbar_init(&tmem_full[b], 1);
        ptx::mbar_init(&tmem_empty[b], 4);  // one arrive per epilogue warp
      }
      for (int a = 0; a < NAUX; ++a) {
        ptx::mbar_init(&aux_full[a], 1);
        ptx::mbar_init(&aux_empty[a], 4);
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
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ A/B producer
    if (ptx::elect_one()) {  // one elected lane: tcgen05 / TMA issue stays on the uniform datapath
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = tile * kBlockM;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % NSTAGE;
          const uint32_t ph = (it / NSTAGE) & 1;
          ptx::mbar_wait(&empty_bar[s], ph ^ 1);
          ptx::mbar_expect_tx(&full_bar[s], Cfg::kStage);
          uint8_t* sa = smem + s * Cfg::kStage;
          int ka = kb * kBlockK;
          if (p.a_k_wrap > 0) ka %= p.a_k_wrap;
          ptx::tma_load_2d(sa, &p.tmA, &full_bar[s], ka, m0);
          ptx::tma_load_2d(sa + Cfg::kStageA, &p.tmB, &full_bar[s], kb * kBlockK, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (ptx::elect_one()) {  // one elected lane: tcgen05 / TMA issue stays on the uniform datapath
      constexpr uint32_t idesc = ptx::make_idesc_tf32(kBlockM, 256, 0, 0);
      int it = 0, t = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
        const int buf = t & 1;
        const uint32_t tph = (t >> 1) & 1;
        ptx::mbar_wait(&tmem_empty[buf], tph ^ 1);  // epilogue has drained this accumulator
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * 256;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % NSTAGE;
          const uint32_t ph = (it / NSTAGE) & 1;
          ptx::mbar_wait(&full_bar[s], ph);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem + s * Cfg::kStage);
          const uint32_t b_addr = a_addr + Cfg::kStageA;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            const uint64_t adesc = ptx::make_smem_desc_sw128(a_addr + k * kUmmaK * 4, 0, 1024);
            const uint64_t bdesc = ptx::make_smem_desc_sw128(b_addr + k * kUmmaK * 4, 0, 1024);
            ptx::umma_tf32(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          ptx::umma_commit(&empty_bar[s]);
        }
        ptx::umma_commit(&tmem_full[buf]);
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------ aux producer
    if (kHasAux1 && lane == 0) {
      int cc = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = tile * kBlockM;
        for (int c = 0; c < NCHUNK; ++c) {
          if (c * 32 >= p.N) break;
          const int a = cc % NAUXD;
          const uint32_t ph = (cc / NAUXD) & 1;
          ptx::mbar_wait(&aux_empty[a], ph ^ 1);
          ptx::mbar_expect_tx(&aux_full[a], Cfg::kAuxSlotBytes);
          uint8_t* slot = smem + Cfg::kOffAux + a * Cfg::kAuxSlotBytes;
          ptx::tma_load_2d(slot, &p.tmAux1, &aux_full[a], c * 32, m0);
          if (kHasAux2) ptx::tma_load_2d(slot + kTileBytes, &p.tmAux2, &aux_full[a], c * 32, m0);
          ++cc;
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue warps
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const bool leader = (warp == 4 && lane == 0);
    const int swz = r & 7;
    const uint32_t row_off = static_cast<uint32_t>(r) * 128u;
    uint8_t* out_base = smem + Cfg::kOffOut;
    int cc = 0, t = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
      const int m0 = tile * kBlockM;
      const int m = m0 + r;
      const bool row_ok = m < p.M;
      const int buf = t & 1;
      const float rs = (p.row_scale != nullptr && row_ok) ? p.row_scale[m] : 0.0f;
      const float rw = (p.colsum_w != nullptr && row_ok) ? p.row_w[m] : 0.0f;
      const float* gb_row = (p.group_bias != nullptr)
                                ? p.group_bias + static_cast<size_t>((row_ok ? m : 0) / p.group) * p.ldg
                                : nullptr;
      ptx::mbar_wait(&tmem_full[buf], (t >> 1) & 1);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < NCHUNK; ++c) {
        const int nc = c * 32;
        if (nc >= p.N) break;
        uint32_t accu[32];
        ptx::tmem_ld_32x32(tmem_base + buf * 256 + (static_cast<uint32_t>(quarter * 32) << 16) + nc, accu);
        const int a = kHasAux1 ? cc % NAUXD : 0;
        if (kHasAux1) ptx::mbar_wait(&aux_full[a], (cc / NAUXD) & 1);
        ptx::tmem_ld_wait();
        const uint8_t* slot = smem + Cfg::kOffAux + a * Cfg::kAuxSlotBytes;
        uint8_t* ob = out_base + (cc % Cfg::kNumOut) * kTileBytes;
        uint8_t* ob2 = out_base + (Cfg::kNumOut + cc % Cfg::kNumOut) * kTileBytes;
        float v[32];
        if (!(p.debug_flags & 2))
          epilogue_chunk<MODE, SPLIT>(p, accu, slot + row_off, slot + kTileBytes + row_off, ob + row_off,
                                      ob2 + row_off, swz, nc, rs, gb_row, v);
        if (kHasAux1) {
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&aux_empty[a]);  // this warp is done with the aux slot
        }
        if (c == NCHUNK - 1 || nc + 32 >= p.N) {
          // last TMEM read of this tile: hand the accumulator back to the MMA warp
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&tmem_empty[buf]);
        }
        ptx::fence_proxy_async_smem();
        // the store of chunk cc-1 must have finished READING its staging buffer before chunk cc+1
        // overwrites it; waiting here (before the barrier) makes that visible to all 128 threads
        if (leader) ptx::tma_store_wait_read<Cfg::kNumOut - 2>();
        ptx::named_bar_sync(1, 128);
        if (leader) {
          if (!(p.debug_flags & 1)) {
            ptx::tma_store_2d(&p.tmOut, ob, nc, m0);
            if (kHasOut2) ptx::tma_store_2d(&p.tmOut2, ob2, nc, m0);
          }
          ptx::tma_store_commit();
        }
        epilogue_colsums<MODE>(p, v, ob2 + row_off, swz, nc, row_ok, rw, lane);
        ++cc;
      }
    }
    if (leader) ptx::tma_store_wait_all<0>();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------
// Weight-gradient contraction: both operands MN-major (TMA boxes of 32 rows x 32 floats).
// MT = 2 ("full M", BLOCK_N = 256 only): one CTA owns a 256 x 256 output -- two accumulators fill the 512 TMEM
// columns, one CTA per SM -- so every X and Y tile enters an SM exactly once (with MT = 1 the two M-tile CTAs both
// fetch Y: 1.5x the algorithmic bytes through the L2 -> SM path of an HBM-bound kernel).
template <int BLOCK_N, int MT = 1>
struct GemmTNConfig {
  static constexpr int kBoxBytes = 32 * 32 * 4;  // 4 KB: 32 k-rows x 32 floats (128 B rows)
  static constexpr int kStageA = MT * (kBlockM / 32) * kBoxBytes;
  static constexpr int kStageB = (BLOCK_N / 32) * kBoxBytes;
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kNumStages = MT == 2 ? 3 : ((BLOCK_N >= 256) ? 2 : (BLOCK_N >= 128 ? 3 : 4));
  static constexpr int kDataBytes = kStage * kNumStages;
  static constexpr int kSmemBytes = kDataBytes + 1024 + 256;
  static constexpr int kTmemCols = MT * (BLOCK_N < 32 ? 32 : BLOCK_N);
  static_assert(MT == 1 || BLOCK_N == 256, "full-M variant: 256 x 256 tiles");
};

// Body shared by the single-problem kernel (n-tile = blockIdx.z) and the batched kernel (problem = blockIdx.z).
// `p` must live in kernel-parameter space (TMA descriptors are taken by address).
template <int BLOCK_N, int MT>
__device__ __forceinline__ void gemm_tn_body(const GemmTNParams& p, const int ntile) {
  using Cfg = GemmTNConfig<BLOCK_N, MT>;
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
  const int m0 = blockIdx.y * (MT * kBlockM);
  const int n0 = ntile * BLOCK_N;
  const int total_kb = (p.K + kBlockK - 1) / kBlockK;
  const int kb_begin = split * p.kb_per_split;
  int kb_end = kb_begin + p.kb_per_split;
  if (kb_end > total_kb) kb_end = total_kb;
  const int nkb = kb_end > kb_begin ? kb_end - kb_begin : 0;
  const int iters = nkb * p.npairs;  // pair-major: all k-blocks of pair 0, then pair 1
  long long* dbg = p.dbg ? p.dbg + (static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 4 : nullptr;
  if (dbg && threadIdx.x == 0) dbg[0] = clock64();
  // Programmatic dependent launch: consecutive weight-gradient contractions are independent of each other (read-only
  // inputs, disjoint outputs), so the next one may start filling SM slots as this one's CTAs retire instead of waiting
  // for the whole grid to drain.  No-ops unless the launch carries the programmatic-serialization attribute.
  if (p.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

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
    if (ptx::elect_one()) {  // one elected lane: tcgen05 / TMA issue stays on the uniform datapath
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
        for (int b = 0; b < MT * kBlockM / 32; ++b)
          ptx::tma_load_2d(sa + b * Cfg::kBoxBytes, tx, &full_bar[s], m0 + b * 32, k0);
#pragma unroll
        for (int b = 0; b < BLOCK_N / 32; ++b)
          ptx::tma_load_2d(sb + b * Cfg::kBoxBytes, ty, &full_bar[s], n0 + b * 32, k0);
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {  // one elected lane: tcgen05 / TMA issue stays on the uniform datapath
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
          if (MT == 2) {  // output rows [128, 256): X boxes 4..7, second accumulator
            const uint64_t adesc2 =
                ptx::make_smem_desc(a_addr + (kBlockM / 32) * Cfg::kBoxBytes + k * 1024, Cfg::kBoxBytes, 512, 1);
            ptx::umma_tf32(tmem_base + BLOCK_N, adesc2, bdesc, idesc, (it | k) != 0 ? 1u : 0u);
          }
        }
        ptx::umma_commit(&empty_bar[s]);
      }
      ptx::umma_commit(tmem_full_bar);
    }
  } else {
    const int quarter = warp & 3;
    const int Mpad = gridDim.y * (MT * kBlockM);
    const int Npad = p.npad;
    if (iters > 0) {
      ptx::mbar_wait(tmem_full_bar, 0);
      ptx::tc_fence_after();
    }
    if (dbg && warp == 2 && lane == 0) dbg[1] = clock64();
    // stream order for whatever follows: this grid does not finish (its reductions do not land) before its predecessor has
    if (p.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
#pragma unroll 1
    for (int mh = 0; mh < MT; ++mh) {
    const int r = mh * kBlockM + quarter * 32 + lane;
    const uint32_t tacc = tmem_base + mh * BLOCK_N;
    float* dst = p.partial + (static_cast<size_t>(split) * Mpad + (m0 + r)) * Npad + n0;
    if (p.red_out != nullptr) {
      // split-K without a second pass: accumulate this split's tile into the gradient with L2 reductions
      if (iters > 0) {
        const bool row_ok = m0 + r < p.M;  // tcgen05.ld is warp-collective: only the reductions are predicated
        float* orow = p.red_out + static_cast<size_t>(row_ok ? m0 + r : 0) * p.red_ld + n0;
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          if (n0 + c * 32 >= p.N) break;
          uint32_t accu[32];
          ptx::tmem_ld_32x32(tacc + (static_cast<uint32_t>(quarter * 32) << 16) + c * 32, accu);
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
#pragma unroll 1
    for (int c = 0; c < BLOCK_N / 32; ++c) {
      uint32_t accu[32];
      if (iters > 0) {
        ptx::tmem_ld_32x32(tacc + (static_cast<uint32_t>(quarter * 32) << 16) + c * 32, accu);
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
    }  // mh
    if (dbg && warp == 2 && lane == 0) dbg[2] = clock64();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BLOCK_N, int MT = 1>
__global__ void __launch_bounds__(kGemmThreads, MT == 2 ? 1 : 2)
gemm_tn_kernel(const __grid_constant__ GemmTNParams p) {
  gemm_tn_body<BLOCK_N, MT>(p, blockIdx.z);
}

template <int BLOCK_N>
__global__ void __launch_bounds__(kGemmThreads, 2)
gemm_tn_batch_kernel(const __grid_constant__ GemmTNBatchParams pb) {
  gemm_tn_body<BLOCK_N, 1>(pb.prob[blockIdx.z], 0);
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
