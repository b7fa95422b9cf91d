// Host-side preparation (TMA tensor maps) and launch of the fused layer-chain kernel (chain_sm100.cuh).
#pragma once
#include <cstdlib>
#include <vector>

#include "chain_sm100.cuh"
#include "gemm_host.cuh"

namespace ardae {

struct ChainLayerDesc {
  const float* W = nullptr; int ldw = 0;      // B operand [H, H] K-major (SOFTPLUS3: [H, 3H] = [Whi | Whi | Wlo])
  const float* aux1 = nullptr; int ld1 = 0;   // [M, H]
  const float* aux2 = nullptr; int ld2 = 0;
  float* out = nullptr; int ldo = 0;          // [M, H]
  float* out2 = nullptr; int ldo2 = 0;        // TANGENT
  float* out_lo = nullptr; int ld_out_lo = 0; // SOFTPLUS3: optional lo spill
  const float* bias = nullptr;                // SOFTPLUS3
  const float* group_bias = nullptr; int group = 1, ldg = 0;
  const float* col_vec = nullptr;             // with ChainDesc::row_scale
  float* colsum = nullptr; float colsum_scale = 1.0f;
  float* colsum2 = nullptr;
  float* colsum_w = nullptr; int colsum_w_stride = 1;
};

struct ChainDesc {
  int mode = CHAIN_MUL_SIG;
  int M = 0, H = 0;
  const float* A0 = nullptr; int lda0 = 0;        // initial activation [M, H] (SOFTPLUS3: hi part)
  const float* A0lo = nullptr; int lda0lo = 0;    // SOFTPLUS3: lo part
  const float* row_scale = nullptr;               // [M]
  int pair = 0;   // CTA-pair (cta_group::2) kernel: 1 on, 0 / -1 off (ARDAE_CHAIN_PAIR=1 switches the default)
  int multicast = 0;  // clusters of 2 / 4 / 8 CTAs sharing one multicast weight stream (1 = 8), -1 off, 0 = default (ARDAE_CHAIN_MC)
  std::vector<ChainLayerDesc> layers;
};

struct PreparedChain {
  ChainParams params;
  const void* fn = nullptr;
  dim3 grid;
  int smem = 0, threads = 0;
  int cluster = 1;
};

// The chain kernel serves H -> H layers with H a multiple of 64 up to 256 (two N-halves of <= 128 columns each).
inline bool chain_supported(int H, int nlayers) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = std::getenv("ARDAE_CHAIN");
    enabled = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return enabled && H >= 64 && H <= 256 && H % 64 == 0 && nlayers >= 1 && nlayers <= kChainMaxLayers;
}

inline int prepare_chain(const ChainDesc& d, PreparedChain* out) {
  const int nl = static_cast<int>(d.layers.size());
  if (d.M <= 0 || !chain_supported(d.H, nl)) return fail(-2, "chain: unsupported shape");
  if (d.mode < 0 || d.mode >= CHAIN_NUM_MODES) return fail(-2, "chain: bad mode");
  const bool s3 = d.mode == CHAIN_SOFTPLUS3;
  const bool aux2 = d.mode == CHAIN_TANGENT || d.mode == CHAIN_ADJOINT, out2 = d.mode == CHAIN_TANGENT;
  if (!d.A0 || (s3 && !d.A0lo)) return fail(-2, "chain: missing initial activation");
  PreparedChain pr;
  std::memset(&pr.params, 0, sizeof(pr.params));
  ChainParams& p = pr.params;
  int rc;
  bool cg2_box;
  {
    const char* e = std::getenv("ARDAE_CHAIN_PAIR");
    const int pe = e ? std::atoi(e) : 0;
    const int pq = d.pair != 0 ? d.pair : pe;
    cg2_box = pq > 0;
  }
  static int mc_env = -2;
  if (mc_env == -2) {
    const char* e = std::getenv("ARDAE_CHAIN_MC");
    mc_env = e ? std::atoi(e) : 0;
  }
  // weight multicast over clusters of 8: needs 8-row-aligned slices (H >= 128) and enough tiles to fill clusters
  int mcs = d.multicast > 0 ? (d.multicast == 1 ? 8 : d.multicast) : (d.multicast == 0 && mc_env > 0 ? (mc_env == 1 ? 8 : mc_env) : 0);
  if (mcs != 2 && mcs != 4 && mcs != 8) mcs = 0;
  if (cg2_box || (d.H / 2) % (8 * (mcs ? mcs : 1)) != 0 || (d.M + kBlockM - 1) / kBlockM < mcs) mcs = 0;
  const bool mc8 = mcs > 0;
  if ((rc = encode_tmap_2d(&p.tmA0, d.A0, d.H, d.M, d.lda0, 32, kBlockM))) return rc;
  p.a0_lo = d.A0lo; p.a0_lo_ld = d.lda0lo; p.row_scale = d.row_scale;
  p.M = d.M; p.H = d.H; p.nlayers = nl;
  uintptr_t align_or = s3 ? (reinterpret_cast<uintptr_t>(d.A0lo) | static_cast<uintptr_t>(d.lda0lo) * 4) : 0;
  for (int l = 0; l < nl; ++l) {
    const ChainLayerDesc& s = d.layers[l];
    ChainLayerParams& q = p.layer[l];
    if (!s.W || !s.out || (!s3 && !s.aux1) || (aux2 && !s.aux2) || (out2 && !s.out2))
      return fail(-2, "chain: missing operand pointer");
    if ((rc = encode_tmap_2d(&q.tmW, s.W, s3 ? 3 * d.H : d.H, d.H, s.ldw, kBlockK, mc8 ? d.H / (2 * mcs) : (cg2_box ? d.H / 4 : d.H / 2)))) return rc;
    if (!s3 && (rc = encode_tmap_2d(&q.tmAux1, s.aux1, d.H, d.M, s.ld1, 32, kBlockM))) return rc;
    if (aux2 && (rc = encode_tmap_2d(&q.tmAux2, s.aux2, d.H, d.M, s.ld2, 32, kBlockM))) return rc;
    if ((rc = encode_tmap_2d(&q.tmOut, s.out, d.H, d.M, s.ldo, 32, kBlockM))) return rc;
    if (out2 && (rc = encode_tmap_2d(&q.tmOut2, s.out2, d.H, d.M, s.ldo2, 32, kBlockM))) return rc;
    q.bias = s.bias; q.group_bias = s.group_bias; q.col_vec = s.col_vec;
    q.group = s.group > 0 ? s.group : 1; q.ldg = s.ldg;
    q.out_lo = s.out_lo; q.ld_out_lo = s.ld_out_lo;
    align_or |= reinterpret_cast<uintptr_t>(s.out_lo) | (static_cast<uintptr_t>(s.ld_out_lo) * 4);
    q.colsum = s.colsum; q.colsum_scale = s.colsum_scale; q.colsum2 = s.colsum2;
    q.colsum_w = s.colsum_w; q.colsum_w_stride = s.colsum_w_stride;
    if ((s.col_vec || s.colsum_w) && !d.row_scale) return fail(-2, "chain: col_vec / colsum_w need row_scale");
    align_or |= reinterpret_cast<uintptr_t>(s.bias) | reinterpret_cast<uintptr_t>(s.group_bias) |
                reinterpret_cast<uintptr_t>(s.col_vec) | (static_cast<uintptr_t>(s.ldg) * 4);
  }
  p.vec_ok = (align_or & 15) == 0 ? 1 : 0;
  if (s3 && !p.vec_ok) return fail(-2, "chain: SOFTPLUS3 operands must be 16-byte aligned");
  // CTA pairs halve the weight bytes each SM ingests; opt-in (ChainDesc::pair = 1 or ARDAE_CHAIN_PAIR=1)
  static int pair_env = -2;
  if (pair_env == -2) {
    const char* e = std::getenv("ARDAE_CHAIN_PAIR");
    pair_env = e ? std::atoi(e) : 0;
  }
  const int pair_req = d.pair != 0 ? d.pair : pair_env;
  const bool cg2 = pair_req > 0;  // measured on B200: the pair's lock-step costs more than the halved weight stream saves
#define ARDAE_CHAIN_CASE(MODE_)                                                                      \
  case MODE_:                                                                                        \
    if (mcs == 8) { pr.fn = reinterpret_cast<const void*>(&chain_kernel<MODE_, false, 8>); pr.smem = ChainConfig<MODE_, false, 8>::kSmemBytes; pr.threads = ChainConfig<MODE_, false, 8>::kThreads; } \
    else if (mcs == 4) { pr.fn = reinterpret_cast<const void*>(&chain_kernel<MODE_, false, 4>); pr.smem = ChainConfig<MODE_, false, 4>::kSmemBytes; pr.threads = ChainConfig<MODE_, false, 4>::kThreads; } \
    else if (mcs == 2) { pr.fn = reinterpret_cast<const void*>(&chain_kernel<MODE_, false, 2>); pr.smem = ChainConfig<MODE_, false, 2>::kSmemBytes; pr.threads = ChainConfig<MODE_, false, 2>::kThreads; } \
    else if (cg2) { pr.fn = reinterpret_cast<const void*>(&chain_kernel<MODE_, true>); pr.smem = ChainConfig<MODE_, true>::kSmemBytes; pr.threads = ChainConfig<MODE_, true>::kThreads; } \
    else { pr.fn = reinterpret_cast<const void*>(&chain_kernel<MODE_, false>); pr.smem = ChainConfig<MODE_, false>::kSmemBytes; pr.threads = ChainConfig<MODE_, false>::kThreads; } \
    break;
  switch (d.mode) {
    ARDAE_CHAIN_CASE(CHAIN_MUL_SIG)
    ARDAE_CHAIN_CASE(CHAIN_TANGENT)
    ARDAE_CHAIN_CASE(CHAIN_ADJOINT)
    default:
    ARDAE_CHAIN_CASE(CHAIN_SOFTPLUS3)
  }
#undef ARDAE_CHAIN_CASE
  pr.grid = dim3((d.M + kBlockM - 1) / kBlockM, 1, 1);
  if (cg2) {
    pr.cluster = 2;
    pr.grid.x = (pr.grid.x + 1) / 2 * 2;  // an odd tail CTA works on an out-of-range tile (TMA clips)
  }
  if (mc8) {
    pr.cluster = mcs;
    pr.grid.x = (pr.grid.x + mcs - 1) / mcs * mcs;  // tail CTAs work on out-of-range tiles but still serve their weight slices
  }
  ARDAE_CUDA_OK(cudaFuncSetAttribute(pr.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, pr.smem));
  *out = pr;
  return 0;
}

inline int launch_prepared_chain(const PreparedChain& pr, cudaStream_t stream) {
  void* args[1] = {const_cast<ChainParams*>(&pr.params)};
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
