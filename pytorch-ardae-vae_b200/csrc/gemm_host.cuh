// Host-side preparation (TMA tensor maps, tile selection) and launch of the GEMM kernels.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "gemm_sm100.cuh"

namespace ardae {

// ------------------------------------------------------------------ error plumbing
inline std::string& last_error_string() {
  static thread_local std::string s;
  return s;
}
inline int fail(int code, const std::string& msg) {
  last_error_string() = msg;
  return code;
}
#define ARDAE_CUDA_OK(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      return ::ardae::fail(static_cast<int>(_e), std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

// ------------------------------------------------------------------ driver entry point
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 2-D fp32 tensor map, 128-byte swizzle.  inner = contiguous dimension (elements).
inline int encode_tmap_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer,
                          uint64_t pitch_elems, uint32_t box_inner, uint32_t box_outer,
                          CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  PFN_encodeTiled fn = get_encode_fn();
  if (fn == nullptr) return fail(-10, "cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0)
    return fail(-11, "TMA base pointer not 16-byte aligned");
  if ((pitch_elems * 4) % 16 != 0) return fail(-12, "TMA row pitch not a multiple of 16 bytes");
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {pitch_elems * 4};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf),
             "cuTensorMapEncodeTiled failed (%d): inner=%llu outer=%llu pitch=%llu box=%ux%u",
             static_cast<int>(r), (unsigned long long)inner, (unsigned long long)outer,
             (unsigned long long)pitch_elems, box_inner, box_outer);
    return fail(-13, buf);
  }
  return 0;
}

// ------------------------------------------------------------------ NT GEMM
struct GemmNTDesc {
  const float* A = nullptr; int lda = 0;   // [M, K]
  const float* B = nullptr; int ldb = 0;   // [N, K]
  float* out = nullptr; int ldo = 0;       // [M, N]
  float* out2 = nullptr; int ldo2 = 0;
  const float* aux1 = nullptr; int ld1 = 0;
  const float* aux2 = nullptr; int ld2 = 0;
  int M = 0, N = 0, K = 0;
  int mode = EPI_LINEAR;
  float alpha = 1.0f;
  const float* bias = nullptr;
  const float* group_bias = nullptr; int group = 1, ldg = 0;
  const float* row_scale = nullptr; const float* col_vec = nullptr;
  float* colsum = nullptr; float colsum_scale = 1.0f;
  float* colsum_w = nullptr; const float* row_w = nullptr; int colsum_w_stride = 1;
  float* colsum2 = nullptr;
  int round_out = 1;
  int split_out = 0;   // modes LINEAR/RELU/SOFTPLUS: out = tf32 hi part, out2 = tf32 lo part
  int a_k_wrap = 0;
  int force_block_n = 0;
  int persist = 0;     // 0 auto, 1 force the persistent kernel, -1 force the tile-per-CTA kernel
  int deep = 0;        // 0 auto (small grids), 1 force the deep-pipeline narrow-tile variant, -1 off
  int prefetch = 0;    // 0 auto (one resident wave ahead for big grids), -1 off, > 0 explicit distance in row tiles
  int pair = 0;        // 0 auto, 1 force the CTA-pair (cta_group::2) kernel for 256-wide tiles, -1 off
  int multicast = 0;   // 0 auto (2-CTA weight multicast for 256-wide tiles with >= 2 row tiles), 1 on, -1 off
  int debug_flags = 0;
};

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

template <bool SPLIT>
inline const void* nt_persist_kernel_for_mode(int mode, int* smem) {
#define ARDAE_PK(M)                                                                   \
  case M:                                                                             \
    *smem = GemmNTPersistConfig<M, SPLIT>::kSmemBytes;                                \
    return reinterpret_cast<const void*>(&gemm_nt_persist_kernel<M, SPLIT>);
  switch (mode) {
    ARDAE_PK(EPI_LINEAR)
    ARDAE_PK(EPI_RELU)
    ARDAE_PK(EPI_SOFTPLUS)
    default: break;
  }
  if (!SPLIT) {
    switch (mode) {
      case EPI_MUL_SIG: *smem = GemmNTPersistConfig<EPI_MUL_SIG, false>::kSmemBytes; return reinterpret_cast<const void*>(&gemm_nt_persist_kernel<EPI_MUL_SIG, false>);
      case EPI_MUL_STEP: *smem = GemmNTPersistConfig<EPI_MUL_STEP, false>::kSmemBytes; return reinterpret_cast<const void*>(&gemm_nt_persist_kernel<EPI_MUL_STEP, false>);
      case EPI_TANGENT: *smem = GemmNTPersistConfig<EPI_TANGENT, false>::kSmemBytes; return reinterpret_cast<const void*>(&gemm_nt_persist_kernel<EPI_TANGENT, false>);
      case EPI_ADJOINT: *smem = GemmNTPersistConfig<EPI_ADJOINT, false>::kSmemBytes; return reinterpret_cast<const void*>(&gemm_nt_persist_kernel<EPI_ADJOINT, false>);
      default: break;
    }
  }
#undef ARDAE_PK
  return nullptr;
}

struct PreparedNT {
  GemmNTParams params;
  const void* fn = nullptr;
  dim3 grid;
  int smem = 0;
  int threads = kGemmThreads;
  int cluster = 1;
};

template <int BLOCK_N, bool MC, bool DEEP = false, bool CG2 = false>
inline const void* nt_kernel_for_mode(int mode, int split) {
  if (split) {
    switch (mode) {
      case EPI_LINEAR: return reinterpret_cast<const void*>(&gemm_nt_kernel<BLOCK_N, EPI_LINEAR, true, MC, DEEP, CG2>);
      case EPI_RELU: return reinterpret_cast<const void*>(&gemm_nt_kernel<BLOCK_N, EPI_RELU, true, MC, DEEP, CG2>);
      case EPI_SOFTPLUS: return reinterpret_cast<const void*>(&gemm_nt_kernel<BLOCK_N, EPI_SOFTPLUS, true, MC, DEEP, CG2>);
      default: return nullptr;
    }
  }
  switch (mode) {
    case EPI_LINEAR: return reinterpret_cast<const void*>(&gemm_nt_kernel<BLOCK_N, EPI_LINEAR, false, MC, DEEP, CG2>);
    case EPI_RELU: return reinterpret_cast<const void*>(&gemm_nt_kernel<BLOCK_N, EPI_RELU, false, MC, DEEP, CG2>);
    case EPI_SOFTPLUS: return reinterpret_cast<const void*>(&gemm_nt_kernel<BLOCK_N, EPI_SOFTPLUS, false, MC, DEEP, CG2>);
    case EPI_MUL_SIG: return reinterpret_cast<const void*>(&gemm_nt_kernel<BLOCK_N, EPI_MUL_SIG, false, MC, DEEP, CG2>);
    case EPI_MUL_STEP: return reinterpret_cast<const void*>(&gemm_nt_kernel<BLOCK_N, EPI_MUL_STEP, false, MC, DEEP, CG2>);
    case EPI_TANGENT: return reinterpret_cast<const void*>(&gemm_nt_kernel<BLOCK_N, EPI_TANGENT, false, MC, DEEP, CG2>);
    case EPI_ADJOINT: return reinterpret_cast<const void*>(&gemm_nt_kernel<BLOCK_N, EPI_ADJOINT, false, MC, DEEP, CG2>);
    default: return nullptr;
  }
}

inline int pick_block_n(int N) {
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  if (N <= 128) return 128;
  const int pad256 = ((N + 255) / 256) * 256;
  const int pad128 = ((N + 127) / 128) * 128;
  return pad128 < pad256 ? 128 : 256;
}

inline int prepare_gemm_nt(const GemmNTDesc& d, PreparedNT* out) {
  if (d.M <= 0 || d.N <= 0 || d.K <= 0) return fail(-2, "gemm_nt: empty problem");
  if (d.mode < 0 || d.mode >= EPI_NUM_MODES) return fail(-2, "gemm_nt: bad mode");
  const bool has_aux1 = d.mode >= EPI_MUL_SIG, has_aux2 = d.mode >= EPI_TANGENT;
  const bool has_out2 = d.mode == EPI_TANGENT || d.split_out;
  if (d.split_out && d.mode > EPI_SOFTPLUS) return fail(-2, "gemm_nt: split_out needs a plain activation mode");
  if (!d.A || !d.B || !d.out || (has_aux1 && !d.aux1) || (has_aux2 && !d.aux2) ||
      (has_out2 && !d.out2))
    return fail(-2, "gemm_nt: missing operand pointer");
  int bn = d.force_block_n ? d.force_block_n : pick_block_n(d.N);
  // small grids (B-row chains) are latency bound: narrow tiles + deep TMA pipeline
  const int tiles_m_pre = (d.M + kBlockM - 1) / kBlockM;
  bool deep = false;
  if (!d.force_block_n && d.persist <= 0 && d.deep >= 0 &&
      (d.deep > 0 || tiles_m_pre * ((d.N + bn - 1) / bn) * 4 <= num_sms())) {
    bn = d.N <= 32 ? 32 : 64;
    deep = true;
  }
  PreparedNT pr;
  std::memset(&pr.params, 0, sizeof(pr.params));
  GemmNTParams& p = pr.params;
  int rc;
  const uint64_t a_inner = d.a_k_wrap > 0 ? d.a_k_wrap : d.K;
  if (d.a_k_wrap > 0 && d.a_k_wrap % kBlockK != 0) return fail(-2, "gemm_nt: a_k_wrap must be a multiple of 32");
  if ((rc = encode_tmap_2d(&p.tmA, d.A, a_inner, d.M, d.lda, kBlockK, kBlockM))) return rc;
  const int tiles_m0 = (d.M + kBlockM - 1) / kBlockM;
  // measured (gemm_selftest): weight multicast does not pay at 2 CTAs/SM (the main loop is bound by
  // bytes-in-flight x TMA latency, not by L2 bandwidth) -> opt-in only
  const bool mc = d.persist <= 0 && bn == 256 && d.multicast > 0;
  const bool cg2 = !mc && d.persist <= 0 && bn == 256 &&
                   (d.pair > 0 || (d.pair == 0 && tiles_m0 >= 2 * num_sms() && d.K >= 512));  // pays for the 3xTF32 sweeps (measured)
  if ((rc = encode_tmap_2d(&p.tmB, d.B, d.K, d.N, d.ldb, kBlockK, (mc || cg2) ? bn / 2 : bn))) return rc;
  if ((rc = encode_tmap_2d(&p.tmOut, d.out, d.N, d.M, d.ldo, 32, kBlockM))) return rc;
  if (has_out2 && (rc = encode_tmap_2d(&p.tmOut2, d.out2, d.N, d.M, d.ldo2, 32, kBlockM))) return rc;
  if (has_aux1 && (rc = encode_tmap_2d(&p.tmAux1, d.aux1, d.N, d.M, d.ld1, 32, kBlockM))) return rc;
  if (has_aux2 && (rc = encode_tmap_2d(&p.tmAux2, d.aux2, d.N, d.M, d.ld2, 32, kBlockM))) return rc;
  p.M = d.M; p.N = d.N; p.K = d.K; p.alpha = d.alpha;
  p.bias = d.bias; p.group_bias = d.group_bias; p.group = d.group > 0 ? d.group : 1; p.ldg = d.ldg;
  p.row_scale = d.row_scale; p.col_vec = d.col_vec;
  p.colsum = d.colsum; p.colsum_w = d.colsum_w; p.row_w = d.row_w;
  p.colsum2 = d.colsum2; p.colsum_scale = d.colsum_scale; p.colsum_w_stride = d.colsum_w_stride;
  p.round_out = d.round_out; p.a_k_wrap = d.a_k_wrap; p.debug_flags = d.debug_flags;
  p.vec_ok = (((reinterpret_cast<uintptr_t>(d.bias) | reinterpret_cast<uintptr_t>(d.group_bias) |
                reinterpret_cast<uintptr_t>(d.col_vec)) & 15) == 0 && d.ldg % 4 == 0) ? 1 : 0;
  if (p.row_scale && !p.col_vec) return fail(-2, "gemm_nt: row_scale without col_vec");
  if (p.colsum_w && !p.row_w) return fail(-2, "gemm_nt: colsum_w without row_w");
  switch (bn) {
    case 32:
      if (deep) { pr.fn = nt_kernel_for_mode<32, false, true>(d.mode, d.split_out); pr.smem = GemmNTConfig<32, true>::kSmemBytes; }
      else { pr.fn = nt_kernel_for_mode<32, false>(d.mode, d.split_out); pr.smem = GemmNTConfig<32>::kSmemBytes; }
      break;
    case 64:
      if (deep) { pr.fn = nt_kernel_for_mode<64, false, true>(d.mode, d.split_out); pr.smem = GemmNTConfig<64, true>::kSmemBytes; }
      else { pr.fn = nt_kernel_for_mode<64, false>(d.mode, d.split_out); pr.smem = GemmNTConfig<64>::kSmemBytes; }
      break;
    case 128: pr.fn = nt_kernel_for_mode<128, false>(d.mode, d.split_out); pr.smem = GemmNTConfig<128>::kSmemBytes; break;
    case 256:
      pr.fn = cg2 ? nt_kernel_for_mode<256, false, false, true>(d.mode, d.split_out)
                  : (mc ? nt_kernel_for_mode<256, true>(d.mode, d.split_out) : nt_kernel_for_mode<256, false>(d.mode, d.split_out));
      pr.smem = cg2 ? GemmNTConfig<256, false, true>::kSmemBytes : GemmNTConfig<256>::kSmemBytes;
      break;
    default: return fail(-2, "gemm_nt: bad BLOCK_N");
  }
  pr.grid = dim3((d.M + kBlockM - 1) / kBlockM, (d.N + bn - 1) / bn, 1);
  if (mc || cg2) {
    pr.cluster = 2;
    pr.grid.x = (pr.grid.x + 1) / 2 * 2;  // an odd tail CTA works on an out-of-range tile (TMA clips)
  }
  const int tiles_m = (d.M + kBlockM - 1) / kBlockM;
  // measured on B200 (tests/native/gemm_selftest): with the lean epilogue the tile-per-CTA kernel at
  // 2 CTAs/SM is on par with the persistent one, so the persistent kernel is opt-in
  const bool want_persist = d.persist > 0;
  if (want_persist) {
    if (bn != 256 || d.N > 256) return fail(-2, "gemm_nt: persistent kernel needs N <= 256");
    int smem = 0;
    const void* fn = d.split_out ? nt_persist_kernel_for_mode<true>(d.mode, &smem)
                                 : nt_persist_kernel_for_mode<false>(d.mode, &smem);
    if (fn == nullptr) return fail(-2, "gemm_nt: no persistent kernel for this mode");
    pr.fn = fn; pr.smem = smem; pr.threads = 256;
    pr.grid = dim3(tiles_m < num_sms() ? tiles_m : num_sms(), 1, 1);
  }
  pr.params.prefetch_tiles = 0;
  if (!want_persist) {
    if (d.prefetch > 0) pr.params.prefetch_tiles = d.prefetch;
    // auto = off: measured on B200, L2 prefetch of the next wave's tiles slows every sweep by 15-25 %
    // (extra L2/HBM contention; the kernels are bandwidth-, not latency-limited)
  }
  ARDAE_CUDA_OK(cudaFuncSetAttribute(pr.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, pr.smem));
  *out = pr;
  return 0;
}

inline int launch_prepared_nt(const PreparedNT& pr, cudaStream_t stream) {
  void* args[1] = {const_cast<GemmNTParams*>(&pr.params)};
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

// ------------------------------------------------------------------ TN GEMM (weight gradient)
struct GemmTNDesc {
  const float* X0 = nullptr; int ldx0 = 0;  // [K, M]
  const float* Y0 = nullptr; int ldy0 = 0;  // [K, N]
  const float* X1 = nullptr; int ldx1 = 0;  // optional second pair
  const float* Y1 = nullptr; int ldy1 = 0;
  int M = 0, N = 0, K = 0;
  float* out = nullptr; int ldo = 0;  // [M, N] destination
  float scale = 1.0f, beta = 0.0f;    // out = beta*out + scale*(X0^T Y0 + X1^T Y1)
  float* workspace = nullptr;         // split-K partials
  size_t workspace_bytes = 0;
  int target_ctas = 296;
  int atomic = 0;                     // 0 auto (ARDAE_TN_ATOMIC, default on), 1 red.global.add epilogue, -1 two-pass reduce
  int full_m = 0;                     // 1: one CTA per SM owns a whole 256 x 256 output (opt-in, also ARDAE_TN_FULL_M=1)
};

struct PreparedTN {
  GemmTNParams params;
  const void* fn = nullptr;
  dim3 grid;
  int smem = 0;
  // reduce
  int nsplit = 0, Mpad = 0, Npad = 0;
  float* out = nullptr; int M = 0, N = 0, ldo = 0;
  float scale = 1.0f, beta = 0.0f;
  bool atomic = false;
};

// Split-K combine: L2 reductions straight into the gradient (default) or partial tiles + a reduce launch
// (ARDAE_TN_ATOMIC=0 / ARDAE_DETERMINISTIC=1: fixed summation order, bit-reproducible gradients).
inline bool tn_atomic_default() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("ARDAE_TN_ATOMIC");
    const char* d = std::getenv("ARDAE_DETERMINISTIC");
    v = 1;
    if (e != nullptr && e[0] == '0') v = 0;
    if (d != nullptr && d[0] == '1') v = 0;
  }
  return v != 0;
}

inline bool tn_full_m(int M, int N, int request) {
  static int env = -2;
  if (env == -2) {
    const char* e = std::getenv("ARDAE_TN_FULL_M");
    env = e ? std::atoi(e) : 0;
  }
  const int req = request != 0 ? request : env;
  // opt-in: measured on B200 it is correct but no faster (98 -> 104 us per [256,256] contraction): the kernel is not
  // bound by the L2 -> SM path
  return req > 0 && M > kBlockM && M <= 2 * kBlockM && pick_block_n(N) == 256 && N <= 256;
}

inline size_t tn_workspace_bytes(int M, int N, int K, int target_ctas = 296) {
  {
    const char* e = std::getenv("ARDAE_TN_CTAS");
    if (e && std::atoi(e) > target_ctas) target_ctas = std::atoi(e);
  }
  const int bn = pick_block_n(N);
  if (tn_full_m(M, N, 0)) {  // one CTA per SM
    const int total_kb = (K + kBlockK - 1) / kBlockK;
    int nsplit = num_sms();
    if (nsplit > total_kb) nsplit = total_kb;
    const size_t full = static_cast<size_t>(nsplit) * 2 * kBlockM * bn * sizeof(float);
    const int mt2 = 2, nt2 = 1;
    int ns2 = target_ctas / (mt2 * nt2);
    if (ns2 > total_kb) ns2 = total_kb;
    const size_t split = static_cast<size_t>(ns2) * mt2 * kBlockM * nt2 * bn * sizeof(float);
    return full > split ? full : split;
  }
  const int mt = (M + kBlockM - 1) / kBlockM, nt = (N + bn - 1) / bn;
  const int total_kb = (K + kBlockK - 1) / kBlockK;
  int nsplit = target_ctas / (mt * nt);
  if (nsplit < 1) nsplit = 1;
  if (nsplit > total_kb) nsplit = total_kb;
  return static_cast<size_t>(nsplit) * mt * kBlockM * nt * bn * sizeof(float);
}

inline int tn_target_ctas(int requested) {
  static int env = -1;
  if (env < 0) {
    const char* e = std::getenv("ARDAE_TN_CTAS");
    env = e ? std::atoi(e) : 0;
  }
  return env > 0 ? env : requested;
}

inline int prepare_gemm_tn(const GemmTNDesc& d_in, PreparedTN* out) {
  GemmTNDesc d = d_in;
  d.target_ctas = tn_target_ctas(d.target_ctas);
  if (d.M <= 0 || d.N <= 0 || d.K <= 0) return fail(-2, "gemm_tn: empty problem");
  if (!d.X0 || !d.Y0 || !d.out || !d.workspace) return fail(-2, "gemm_tn: missing pointer");
  const int bn = pick_block_n(d.N);
  const bool full_m = tn_full_m(d.M, d.N, d.full_m);
  const int mrows = full_m ? 2 * kBlockM : kBlockM;  // output rows per CTA
  const int mt = (d.M + mrows - 1) / mrows, nt = (d.N + bn - 1) / bn;
  const int total_kb = (d.K + kBlockK - 1) / kBlockK;
  int nsplit = (full_m ? num_sms() : d.target_ctas) / (mt * nt);
  if (nsplit < 1) nsplit = 1;
  if (nsplit > total_kb) nsplit = total_kb;
  const int kb_per_split = (total_kb + nsplit - 1) / nsplit;
  nsplit = (total_kb + kb_per_split - 1) / kb_per_split;
  const size_t need = static_cast<size_t>(nsplit) * mt * mrows * nt * bn * sizeof(float);
  if (need > d.workspace_bytes) return fail(-3, "gemm_tn: workspace too small");
  PreparedTN pr;
  std::memset(&pr.params, 0, sizeof(pr.params));
  GemmTNParams& p = pr.params;
  int rc;
  if ((rc = encode_tmap_2d(&p.tmX0, d.X0, d.M, d.K, d.ldx0, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
  if ((rc = encode_tmap_2d(&p.tmY0, d.Y0, d.N, d.K, d.ldy0, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
  p.npairs = 1;
  if (d.X1 != nullptr) {
    if (!d.Y1) return fail(-2, "gemm_tn: X1 without Y1");
    if ((rc = encode_tmap_2d(&p.tmX1, d.X1, d.M, d.K, d.ldx1, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
    if ((rc = encode_tmap_2d(&p.tmY1, d.Y1, d.N, d.K, d.ldy1, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
    p.npairs = 2;
  }
  p.M = d.M; p.N = d.N; p.K = d.K; p.kb_per_split = kb_per_split; p.partial = d.workspace;
  // vector reductions need 16-byte aligned rows (the [H, 2H+1] first neglogprob layer keeps the two-pass path)
  const bool red_vec = (reinterpret_cast<uintptr_t>(d.out) & 15) == 0 && d.ldo % 4 == 0 && d.N % 4 == 0;
  pr.atomic = (d.atomic > 0 || (d.atomic == 0 && tn_atomic_default() && red_vec)) && d.scale == 1.0f && d.beta == 1.0f;
  if (pr.atomic) {
    p.red_out = d.out; p.red_ld = d.ldo; p.red_vec = red_vec ? 1 : 0;
  }
  switch (bn) {
    case 32: pr.fn = reinterpret_cast<const void*>(&gemm_tn_kernel<32>); pr.smem = GemmTNConfig<32>::kSmemBytes; break;
    case 64: pr.fn = reinterpret_cast<const void*>(&gemm_tn_kernel<64>); pr.smem = GemmTNConfig<64>::kSmemBytes; break;
    case 128: pr.fn = reinterpret_cast<const void*>(&gemm_tn_kernel<128>); pr.smem = GemmTNConfig<128>::kSmemBytes; break;
    case 256:
      if (full_m) { pr.fn = reinterpret_cast<const void*>(&gemm_tn_kernel<256, 2>); pr.smem = GemmTNConfig<256, 2>::kSmemBytes; }
      else { pr.fn = reinterpret_cast<const void*>(&gemm_tn_kernel<256>); pr.smem = GemmTNConfig<256>::kSmemBytes; }
      break;
    default: return fail(-2, "gemm_tn: bad BLOCK_N");
  }
  pr.grid = dim3(nsplit, mt, nt);
  pr.nsplit = nsplit; pr.Mpad = mt * mrows; pr.Npad = nt * bn;
  p.npad = pr.Npad;
  pr.out = d.out; pr.M = d.M; pr.N = d.N; pr.ldo = d.ldo; pr.scale = d.scale; pr.beta = d.beta;
  ARDAE_CUDA_OK(cudaFuncSetAttribute(pr.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, pr.smem));
  *out = pr;
  return 0;
}

// ---- batched weight gradients: same-shape, red.add-combined contractions in ONE launch (grid.z = problem)
struct PreparedTNBatch {
  GemmTNBatchParams params;
  const void* fn = nullptr;
  dim3 grid;
  int smem = 0;
  int count = 0;
};

// true if `d` can join a batch (vector-reduction epilogue, 256-wide single n tile, no full-M variant)
inline bool tn_batchable(const GemmTNDesc& d) {
  static int env = -2;
  if (env == -2) {
    // opt-in: measured on B200 the batched launch is SLOWER (116 us vs 102 us per contraction: co-resident CTAs of
    // different problems interleave their HBM streams), so the default stays one launch per contraction
    const char* e = std::getenv("ARDAE_TN_BATCH");
    env = e ? std::atoi(e) : 0;
  }
  const bool red_vec = (reinterpret_cast<uintptr_t>(d.out) & 15) == 0 && d.ldo % 4 == 0 && d.N % 4 == 0;
  return env > 0 && tn_atomic_default() && red_vec && d.atomic >= 0 && d.scale == 1.0f && d.beta == 1.0f &&
         pick_block_n(d.N) == 256 && d.N <= 256 && !tn_full_m(d.M, d.N, d.full_m);
}

inline int prepare_gemm_tn_batch(const std::vector<GemmTNDesc>& ds, PreparedTNBatch* out) {
  const int n = static_cast<int>(ds.size());
  if (n < 1 || n > kMaxTNBatch) return fail(-2, "gemm_tn_batch: 1..10 problems");
  PreparedTNBatch pb;
  std::memset(&pb.params, 0, sizeof(pb.params));
  PreparedTN first;
  for (int i = 0; i < n; ++i) {
    if (!tn_batchable(ds[i])) return fail(-2, "gemm_tn_batch: problem is not batchable");
    PreparedTN pr;
    int rc = prepare_gemm_tn(ds[i], &pr);
    if (rc) return rc;
    if (!pr.atomic) return fail(-2, "gemm_tn_batch: needs the red.add epilogue");
    if (i == 0) first = pr;
    else if (pr.grid.x != first.grid.x || pr.grid.y != first.grid.y || pr.grid.z != 1 || pr.fn != first.fn)
      return fail(-2, "gemm_tn_batch: problems must have the same shape");
    pb.params.prob[i] = pr.params;
  }
  if (first.grid.z != 1) return fail(-2, "gemm_tn_batch: single n tile only");
  pb.fn = reinterpret_cast<const void*>(&gemm_tn_batch_kernel<256>);
  pb.smem = GemmTNConfig<256>::kSmemBytes;
  pb.grid = dim3(first.grid.x, first.grid.y, n);
  pb.count = n;
  ARDAE_CUDA_OK(cudaFuncSetAttribute(pb.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, pb.smem));
  *out = pb;
  return 0;
}

inline int launch_prepared_tn_batch(const PreparedTNBatch& pb, cudaStream_t stream) {
  void* args[1] = {const_cast<GemmTNBatchParams*>(&pb.params)};
  ARDAE_CUDA_OK(cudaLaunchKernel(pb.fn, pb.grid, dim3(kGemmThreads), args, pb.smem, stream));
  return 0;
}

inline int launch_prepared_tn(const PreparedTN& pr, cudaStream_t stream) {
  void* args[1] = {const_cast<GemmTNParams*>(&pr.params)};
  static int pdl_env = -1;
  if (pdl_env < 0) {
    const char* e = std::getenv("ARDAE_TN_PDL");
    pdl_env = e ? std::atoi(e) : 0;
  }
  if (pr.atomic && pdl_env > 0) {
    // overlap this contraction's ramp-up with the tail of the previous launch in the stream (PDL)
    PreparedTN q = pr;
    q.params.pdl = 1;
    void* qargs[1] = {&q.params};
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = pr.grid; cfg.blockDim = dim3(kGemmThreads); cfg.dynamicSmemBytes = pr.smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    ARDAE_CUDA_OK(cudaLaunchKernelExC(&cfg, pr.fn, qargs));
    return 0;
  }
  ARDAE_CUDA_OK(cudaLaunchKernel(pr.fn, pr.grid, dim3(kGemmThreads), args, pr.smem, stream));
  if (pr.atomic) return 0;
  const int total = pr.M * pr.N;
  splitk_reduce_kernel<<<(total + 255) / 256, 256, 0, stream>>>(
      pr.params.partial, pr.nsplit, pr.Mpad, pr.Npad, pr.out, pr.M, pr.N, pr.ldo, pr.scale, pr.beta);
  ARDAE_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace ardae
