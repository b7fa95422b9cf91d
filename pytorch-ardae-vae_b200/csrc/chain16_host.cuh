// Host-side preparation (TMA tensor maps) and launch of the 16-bit-spill layer-chain kernel (chain16_sm100.cuh).
#pragma once
#include <cstdio>
#include <vector>

#include "chain16_sm100.cuh"
#include "chain_host.cuh"
#include "chain_s3h_sm100.cuh"
#include "chain16w_sm100.cuh"

namespace ardae {

// 2-D bf16 tensor map.  inner = contiguous dimension (elements); pitch in elements.
inline int encode_tmap_2d_bf16(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer,
                               uint64_t pitch_elems, uint32_t box_inner, uint32_t box_outer,
                               CUtensorMapSwizzle swizzle, CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16) {
  PFN_encodeTiled fn = get_encode_fn();
  if (fn == nullptr) return fail(-10, "cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail(-11, "TMA base pointer not 16-byte aligned");
  if ((pitch_elems * 2) % 16 != 0) return fail(-12, "TMA row pitch not a multiple of 16 bytes");
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, dtype, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled (bf16) failed (%d): inner=%llu outer=%llu pitch=%llu box=%ux%u",
             static_cast<int>(r), (unsigned long long)inner, (unsigned long long)outer,
             (unsigned long long)pitch_elems, box_inner, box_outer);
    return fail(-13, buf);
  }
  return 0;
}

// A bf16 [rows, cols] row-major array (pitch in elements).
struct Mat16 {
  uint16_t* p = nullptr;
  int rows = 0, cols = 0, ld = 0;
  Mat16() {}
  Mat16(uint16_t* p_, int r, int c, int l) : p(p_), rows(r), cols(c), ld(l) {}
};

struct Chain16LayerDesc {
  const float* W = nullptr; int ldw = 0;  // fp32 B operand [nout, kin] K-major (SOFTPLUS3: [H, 3*kin] = [Whi | Whi | Wlo])
  int kin = 0;                            // 0 = H
  int nout = 0;                           // 0 = H; < H: narrow linear last layer writing fp32 rows to out32
  int w_rows = 0;                         // rows of W that exist in memory (0 = nout); missing rows read as zero
  const uint16_t* aux1 = nullptr; int ld1 = 0;
  const uint16_t* aux2 = nullptr; int ld2 = 0;
  uint16_t* out = nullptr; int ldo = 0;
  uint16_t* out2 = nullptr; int ldo2 = 0;
  float* out32 = nullptr; int ld_out32 = 0;
  const float* bias = nullptr;
  const float* group_bias = nullptr; int group = 1, ldg = 0;
  const float* col_vec = nullptr;
  float* colsum = nullptr; float colsum_scale = 1.0f;
  float* colsum2 = nullptr;
  float* colsum_w = nullptr; int colsum_w_stride = 1;
  // SOFTPLUS3 on the fp16 pipe (chain_s3h_sm100.cuh): fp16 [H, 2*kin16] = [Whi | Wlo], scaled by 2^4
  const uint16_t* W16 = nullptr; int ldw16 = 0; int kin16 = 0;
};

struct Chain16Desc {
  int mode = CHAIN_MUL_SIG;
  int M = 0, H = 0;
  // initial activation: either bf16 [M, H] (A0_16) or fp32 [M, kin0] (A0_32; SOFTPLUS3: hi part, with A0lo)
  const uint16_t* A0_16 = nullptr; int lda0_16 = 0;
  const float* A0_32 = nullptr; int lda0_32 = 0;
  const float* A0lo = nullptr; int lda0lo = 0;
  const float* row_scale = nullptr;
  // chain_s3h only: the last layer also writes delta_L = -w_o * sig(v_L) (bf16), the start of the score sweep
  const float* wo = nullptr;
  uint16_t* delta16 = nullptr; int ld_delta16 = 0;
  std::vector<Chain16LayerDesc> layers;
};

// The fp16-pipe primal sweep serves H = 128 / 256 (64-wide k-blocks split evenly over the two N-halves).
// ARDAE_S3_FP16=0 keeps the tf32 variant (A/B measurements).
inline bool s3h_supported(int H) {
  static int env = -1;
  if (env < 0) {
    const char* e = std::getenv("ARDAE_S3_FP16");
    env = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return env != 0 && (H == 128 || H == 256);
}

struct PreparedChainS3h {
  ChainS3hParams params;
  dim3 grid;
};

// SOFTPLUS3 chain on the fp16 pipe: A0_32 / A0lo = the tf32 (hi, lo) pair of the initial activation [M, kin0]
inline int prepare_chain_s3h(const Chain16Desc& d, PreparedChainS3h* out) {
  const int nl = static_cast<int>(d.layers.size());
  if (d.M <= 0 || d.mode != CHAIN_SOFTPLUS3 || !s3h_supported(d.H) || nl < 1 || nl > kChainMaxLayers)
    return fail(-2, "chain_s3h: unsupported shape");
  if (!d.A0_32 || !d.A0lo || d.lda0_32 != d.lda0lo) return fail(-2, "chain_s3h: needs the fp32 (hi, lo) initial activation");
  PreparedChainS3h pr;
  std::memset(&pr.params, 0, sizeof(pr.params));
  ChainS3hParams& p = pr.params;
  const int H = d.H;
  const int kin0 = d.layers[0].kin > 0 ? d.layers[0].kin : H;
  if (kin0 % 32 != 0 || kin0 > H) return fail(-2, "chain_s3h: first-layer input width must be a multiple of 32, <= H");
  p.a0_hi = d.A0_32; p.a0_lo = d.A0lo; p.a0_ld = d.lda0_32; p.kin0 = kin0;
  p.row_scale = d.row_scale; p.M = d.M; p.H = H; p.nlayers = nl;
  uintptr_t align_or = reinterpret_cast<uintptr_t>(d.A0_32) | reinterpret_cast<uintptr_t>(d.A0lo) |
                       (static_cast<uintptr_t>(d.lda0_32) * 4);
  int rc;
  for (int l = 0; l < nl; ++l) {
    const Chain16LayerDesc& s = d.layers[l];
    ChainS3hLayerParams& q = p.layer[l];
    const int kin = s.kin > 0 ? s.kin : H;
    const int kin16 = (kin + 63) / 64 * 64;
    if (l > 0 && kin != H) return fail(-2, "chain_s3h: only the first layer may be narrower than H on input");
    if (!s.W16 || s.kin16 != kin16 || !s.out) return fail(-2, "chain_s3h: missing fp16 weight operand / output");
    if ((rc = encode_tmap_2d_bf16(&q.tmW, s.W16, 2 * kin16, H, s.ldw16, kS3hBlockK, H / 2, CU_TENSOR_MAP_SWIZZLE_128B,
                                  CU_TENSOR_MAP_DATA_TYPE_FLOAT16)))
      return rc;
    q.kin16 = kin16;
    q.bias = s.bias; q.group_bias = s.group_bias; q.col_vec = s.col_vec;
    q.group = s.group > 0 ? s.group : 1; q.ldg = s.ldg;
    q.out16 = s.out; q.ld_out16 = s.ldo;
    if (s.col_vec && !d.row_scale) return fail(-2, "chain_s3h: col_vec needs row_scale");
    align_or |= reinterpret_cast<uintptr_t>(s.bias) | reinterpret_cast<uintptr_t>(s.group_bias) |
                reinterpret_cast<uintptr_t>(s.col_vec) | (static_cast<uintptr_t>(s.ldg) * 4) |
                reinterpret_cast<uintptr_t>(s.out) | (static_cast<uintptr_t>(s.ldo) * 2);
  }
  if (d.delta16 != nullptr) {
    if (!d.wo) return fail(-2, "chain_s3h: delta output needs w_o");
    p.wo = d.wo; p.delta16 = d.delta16; p.ld_delta16 = d.ld_delta16;
    align_or |= reinterpret_cast<uintptr_t>(d.wo) | reinterpret_cast<uintptr_t>(d.delta16) |
                (static_cast<uintptr_t>(d.ld_delta16) * 2);
  }
  if ((align_or & 15) != 0) return fail(-2, "chain_s3h: operands must be 16-byte aligned");
  p.vec_ok = 1;
  pr.grid = dim3((d.M + kBlockM - 1) / kBlockM, 1, 1);
  ARDAE_CUDA_OK(cudaFuncSetAttribute(reinterpret_cast<const void*>(&chain_s3h_kernel),
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, ChainS3hConfig::kSmemBytes));
  *out = pr;
  return 0;
}

inline int launch_prepared_chain_s3h(const PreparedChainS3h& pr, cudaStream_t stream) {
  void* args[1] = {const_cast<ChainS3hParams*>(&pr.params)};
  ARDAE_CUDA_OK(cudaLaunchKernel(reinterpret_cast<const void*>(&chain_s3h_kernel), pr.grid,
                                 dim3(ChainS3hConfig::kThreads), args, ChainS3hConfig::kSmemBytes, stream));
  return 0;
}

struct PreparedChain16 {
  Chain16Params params;
  const void* fn = nullptr;
  dim3 grid;
  int smem = 0, threads = 0;
  int cluster = 1;
};

// CTA pairs (cta_group::2) for the tf32 sweeps: each SM ingests half of the weight stream.  Opt-in (ARDAE_CHAIN16_PAIR=1):
// measured on B200 it is correct but SLOWER (score / tangent / adjoint sweeps 0.48 / 0.63 / 0.51 ms vs 0.36 / 0.50 / 0.42):
// the lock-step of the pair costs more than the halved weight stream saves (same finding as round 1 for the fp32 kernel).
inline bool chain16_pair_default() {
  static int env = -1;
  if (env < 0) {
    const char* e = std::getenv("ARDAE_CHAIN16_PAIR");
    env = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return env != 0;
}

// Weight multicast across CTA pairs (chain16w_kernel<MODE, false, true>): opt-in, ARDAE_CHAIN16_MC=1.  Measured on B200:
// halving the L2 -> SM weight traffic this way changes nothing (0.365 / 0.501 / 0.410 ms vs 0.367 / 0.496 / 0.416), i.e.
// the sweeps are not bound by the weight stream.
inline bool chain16_mc_default() {
  static int env = -1;
  if (env < 0) {
    const char* e = std::getenv("ARDAE_CHAIN16_MC");
    env = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return env != 0;
}

inline int prepare_chain16(const Chain16Desc& d, PreparedChain16* out) {
  const int nl = static_cast<int>(d.layers.size());
  if (d.M <= 0 || !chain_supported(d.H, nl)) return fail(-2, "chain16: unsupported shape");
  if (d.mode < 0 || d.mode >= CHAIN_NUM_MODES) return fail(-2, "chain16: bad mode");
  const bool s3 = d.mode == CHAIN_SOFTPLUS3;
  const bool aux2 = d.mode == CHAIN_TANGENT || d.mode == CHAIN_ADJOINT, out2 = d.mode == CHAIN_TANGENT;
  const int H = d.H;
  PreparedChain16 pr;
  std::memset(&pr.params, 0, sizeof(pr.params));
  Chain16Params& p = pr.params;
  int rc;
  static int warps_env = -1;
  if (warps_env < 0) {
    const char* e = std::getenv("ARDAE_CHAIN16_WARPS");
    warps_env = e ? std::atoi(e) : 16;
  }
  const bool w16 = warps_env != 8 && !s3;  // 16 epilogue warps (chain16w_sm100.cuh) unless ARDAE_CHAIN16_WARPS=8
  // pairs need 8-row aligned weight halves (H/4 and nout/2 multiples of 8) and at least two row tiles
  bool cg2 = w16 && chain16_pair_default() && (H / 4) % 8 == 0 && d.M > kBlockM;
  bool mc = w16 && !cg2 && chain16_mc_default() && (H / 4) % 8 == 0 && d.M > kBlockM;
  for (const Chain16LayerDesc& s : d.layers)
    if (s.nout > 0 && s.nout != H && (s.nout / 2) % 8 != 0) cg2 = mc = false;
  const int kin0 = d.layers[0].kin > 0 ? d.layers[0].kin : H;
  if (kin0 % 32 != 0 || kin0 > H) return fail(-2, "chain16: first-layer input width must be a multiple of 32, <= H");
  uintptr_t align_or = 0;
  if (s3) {
    if (!d.A0_32 || !d.A0lo) return fail(-2, "chain16: SOFTPLUS3 needs the fp32 (hi, lo) initial activation");
    if ((rc = encode_tmap_2d(&p.tmA0, d.A0_32, kin0, d.M, d.lda0_32, 32, kBlockM))) return rc;
    p.a0_f32 = d.A0lo; p.a0_ld = d.lda0lo; p.a0_mode = 0;
    align_or |= reinterpret_cast<uintptr_t>(d.A0lo) | (static_cast<uintptr_t>(d.lda0lo) * 4);
  } else if (d.A0_16) {
    if (kin0 != H) return fail(-2, "chain16: a bf16 initial activation must be H wide");
    if ((rc = encode_tmap_2d_bf16(&p.tmA0, d.A0_16, H, d.M, d.lda0_16, 32, kBlockM, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    p.a0_mode = 1;
  } else {
    if (!d.A0_32) return fail(-2, "chain16: missing initial activation");
    p.a0_f32 = d.A0_32; p.a0_ld = d.lda0_32; p.a0_mode = 0;
    if ((reinterpret_cast<uintptr_t>(d.A0_32) & 15) != 0 || d.lda0_32 % 4 != 0)
      return fail(-2, "chain16: fp32 initial activation must be 16-byte aligned");
  }
  p.row_scale = d.row_scale;
  p.M = d.M; p.H = H; p.nlayers = nl;
  for (int l = 0; l < nl; ++l) {
    const Chain16LayerDesc& s = d.layers[l];
    Chain16LayerParams& q = p.layer[l];
    const int kin = s.kin > 0 ? s.kin : H, nout = s.nout > 0 ? s.nout : H;
    if (l > 0 && kin != H) return fail(-2, "chain16: only the first layer may be narrower than H on input");
    if (nout != H && (l != nl - 1 || s3 || nout % 32 != 0 || nout > H / 2 || d.mode != CHAIN_MUL_SIG || !s.out32))
      return fail(-2, "chain16: only the last MUL_SIG layer may be narrow (multiple of 32, <= H/2, fp32 out32)");
    const bool narrow = nout != H;
    if (!s.W || (!narrow && !s3 && (!s.aux1 || !s.out)) || (!narrow && aux2 && !s.aux2) || (out2 && !s.out2) ||
        (s3 && !s.out))
      return fail(-2, "chain16: missing operand pointer");
    if ((rc = encode_tmap_2d(&q.tmW, s.W, s3 ? 3 * kin : kin, s.w_rows > 0 ? s.w_rows : nout, s.ldw, kBlockK,
                             (narrow ? nout : H / 2) / ((cg2 || mc) ? 2 : 1))))
      return rc;
    if (!narrow && !s3) {
      if ((rc = encode_tmap_2d_bf16(&q.tmAux1, s.aux1, H, d.M, s.ld1, 32, kBlockM, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
      if (aux2 && (rc = encode_tmap_2d_bf16(&q.tmAux2, s.aux2, H, d.M, s.ld2, 32, kBlockM, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
      if ((rc = encode_tmap_2d_bf16(&q.tmOut, s.out, H, d.M, s.ldo, 32, kBlockM, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
      if (out2 && (rc = encode_tmap_2d_bf16(&q.tmOut2, s.out2, H, d.M, s.ldo2, 32, kBlockM, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    }
    if (s3) {
      q.out16 = s.out; q.ld_out16 = s.ldo;
      if ((reinterpret_cast<uintptr_t>(s.out) & 15) != 0 || (s.ldo * 2) % 16 != 0)
        return fail(-2, "chain16: SOFTPLUS3 spill rows must be 16-byte aligned");
    }
    q.out32 = s.out32; q.ld_out32 = s.ld_out32;
    if (narrow && ((reinterpret_cast<uintptr_t>(s.out32) & 15) != 0 || s.ld_out32 % 4 != 0 || s.ld_out32 < nout))
      return fail(-2, "chain16: narrow-layer output rows must be 16-byte aligned and nout wide");
    q.kin = kin; q.nout = nout;
    q.bias = s.bias; q.group_bias = s.group_bias; q.col_vec = s.col_vec;
    q.group = s.group > 0 ? s.group : 1; q.ldg = s.ldg;
    q.colsum = s.colsum; q.colsum_scale = s.colsum_scale; q.colsum2 = s.colsum2;
    q.colsum_w = s.colsum_w; q.colsum_w_stride = s.colsum_w_stride;
    if ((s.col_vec || s.colsum_w) && !d.row_scale) return fail(-2, "chain16: col_vec / colsum_w need row_scale");
    align_or |= reinterpret_cast<uintptr_t>(s.bias) | reinterpret_cast<uintptr_t>(s.group_bias) |
                reinterpret_cast<uintptr_t>(s.col_vec) | (static_cast<uintptr_t>(s.ldg) * 4);
  }
  p.vec_ok = (align_or & 15) == 0 ? 1 : 0;
  if (s3 && !p.vec_ok) return fail(-2, "chain16: SOFTPLUS3 operands must be 16-byte aligned");
#define ARDAE_CHAIN16_CASE(MODE_)                                                                               \
  case MODE_:                                                                                                   \
    pr.fn = reinterpret_cast<const void*>(&chain16_kernel<MODE_>);                                              \
    pr.smem = Chain16Config<MODE_>::kSmemBytes; pr.threads = Chain16Config<MODE_>::kThreads;                   \
    break;
#define ARDAE_CHAIN16W_CASE(MODE_)                                                                              \
  case MODE_:                                                                                                   \
    pr.fn = cg2 ? reinterpret_cast<const void*>(&chain16w_kernel<MODE_, true>)                                  \
                : (mc ? reinterpret_cast<const void*>(&chain16w_kernel<MODE_, false, true>)                     \
                      : (w16 ? reinterpret_cast<const void*>(&chain16w_kernel<MODE_, false>)                    \
                             : reinterpret_cast<const void*>(&chain16_kernel<MODE_>)));                         \
    pr.smem = Chain16Config<MODE_>::kSmemBytes;                                                                 \
    pr.threads = w16 ? Chain16wConfig::kThreads : Chain16Config<MODE_>::kThreads;                               \
    break;
  switch (d.mode) {
    ARDAE_CHAIN16W_CASE(CHAIN_MUL_SIG)
    ARDAE_CHAIN16W_CASE(CHAIN_TANGENT)
    ARDAE_CHAIN16W_CASE(CHAIN_ADJOINT)
    default:
    ARDAE_CHAIN16_CASE(CHAIN_SOFTPLUS3)
  }
#undef ARDAE_CHAIN16_CASE
#undef ARDAE_CHAIN16W_CASE
  pr.grid = dim3((d.M + kBlockM - 1) / kBlockM, 1, 1);
  if ((cg2 || mc) && std::getenv("ARDAE_VERBOSE") != nullptr)
    std::fprintf(stderr, "chain16: mode %d M=%d runs as CTA pairs (%s)\n", d.mode, d.M, cg2 ? "cta_group::2" : "weight multicast");
  if (cg2 || mc) {
    pr.cluster = 2;
    pr.grid.x = (pr.grid.x + 1) / 2 * 2;  // an odd tail CTA works on an out-of-range tile (TMA clips)
  }
  ARDAE_CUDA_OK(cudaFuncSetAttribute(pr.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, pr.smem));
  *out = pr;
  return 0;
}

inline int launch_prepared_chain16(const PreparedChain16& pr, cudaStream_t stream) {
  void* args[1] = {const_cast<Chain16Params*>(&pr.params)};
  if (pr.cluster > 1) {
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = pr.grid; cfg.blockDim = dim3(pr.threads); cfg.dynamicSmemBytes = pr.smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = pr.cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    ARDAE_CUDA_OK(cudaLaunchKernelExC(&cfg, pr.fn, args));
    return 0;
  }
  ARDAE_CUDA_OK(cudaLaunchKernel(pr.fn, pr.grid, dim3(pr.threads), args, pr.smem, stream));
  return 0;
}

}  // namespace ardae
