// Fused layer-chain kernel: a run of H -> H layers of one AR-DAE sweep executed by ONE launch, each CTA
// carrying its 128-row tile through every layer with the activation operand resident ON CHIP.
//
//   per layer l:   acc[128, H] = A_l[128, H] . W_l[H, H]^T          (tcgen05.mma kind::tf32, fp32 accumulate in TMEM)
//                  (out_l, A_{l+1}) = epilogue_l(acc ; aux1_l, aux2_l)   A_{l+1} stays in TMEM / shared memory
//
// Why: the layer-per-launch path re-reads every activation tile from HBM as the next layer's A operand and
// pushes A + W + aux through the ~80 GB/s-per-SM L2 port (round-1 profile: NT kernels at 27-44 % DRAM, 9-21 %
// tensor pipe).  Here A never leaves the SM; per layer a CTA only streams the (L2-resident) weight, the aux
// tiles the epilogue needs (softplus outputs of the primal sweep, deltas) and writes the spill the later
// sweeps / weight-gradient contractions read.
//
// Modes (reference math: SURVEY.md 8a-3, models/graddae/mlp.py:400-444 double back-prop):
//   CHAIN_MUL_SIG   sweep 2 (score backward)  out = acc * sig(aux1)                              A' = out
//   CHAIN_TANGENT   sweep 3 (tangent forward) out = acc * sig(aux1), out2 = aux2*acc*(1-sig)     A' = out
//   CHAIN_ADJOINT   sweep 4 (adjoint backward) out = acc * sig(aux1) + aux2                      A' = out
//   CHAIN_SOFTPLUS3 sweep 1 (primal forward, "3xTF32"): out = softplus(acc + bias terms) split into tf32
//                   hi/lo; hi lives in shared memory (MMA operand AND TMA-store source of the spill), lo in TMEM;
//                   acc = hi.Whi + lo.Whi + hi.Wlo  (fp32-accurate products on the tf32 pipe)
// sig(aux1) = 1 - exp(-u) with u the stored softplus OUTPUT (as in gemm_sm100.cuh).
//
// On-chip budget (H = 256): TMEM 512 columns = accumulator [0,256) + A operand [256,512);
// shared memory 224 KB = 3 x 32 KB weight k-block ring + 64 KB aux ring + 64 KB out staging
// (SOFTPLUS3: weight ring + 128 KB hi tiles).  One CTA per SM; 12 warps:
//   warp 0 weight TMA producer | warp 1 MMA issuer + TMEM owner | warp 2 aux / A0 TMA producer | warp 3 idle
//   warps 4-11 epilogue: two groups of four warps (TMEM lane quarter = warp & 3); group g owns the 32-column
//   chunks c with c % 2 == g, has its own staging tiles, named barrier and TMA-store issuing thread.
// Pipelining inside a CTA: every layer's MMA is issued as two N-halves (accumulator columns [0,H/2) then [H/2,H)).
// The epilogue of half 0 runs underneath the MMA of half 1 (it may overwrite A chunk c only once the half-1 MMAs
// of k-block c have retired: `kfree[c]`), and the next layer's MMA starts on k-blocks [0,H/64) as soon as the
// half-0 epilogue has produced them (`a_ready[0]`) while the half-1 epilogue is still running.
#pragma once
#include "gemm_sm100.cuh"

namespace ardae {

constexpr int kChainMaxLayers = 10;
enum ChainMode : int { CHAIN_MUL_SIG = 0, CHAIN_TANGENT = 1, CHAIN_ADJOINT = 2, CHAIN_SOFTPLUS3 = 3, CHAIN_NUM_MODES = 4 };

struct alignas(64) ChainLayerParams {
  CUtensorMap tmW;     // B operand [H rows (out), K] K-major, box {32, H}; SOFTPLUS3: [H, 3H] = [Whi | Whi | Wlo]
  CUtensorMap tmAux1;  // [M, H], box {32, 128}
  CUtensorMap tmAux2;
  CUtensorMap tmOut;
  CUtensorMap tmOut2;
  const float* bias;        // SOFTPLUS3: [H] or null
  const float* group_bias;  // SOFTPLUS3: row m adds group_bias[(m / group) * ldg + n]
  const float* col_vec;     // SOFTPLUS3: adds row_scale[m] * col_vec[n]
  float* out_lo;            // SOFTPLUS3: optional spill of the lo part (row pitch ld_out_lo) for a following chain
  float* colsum;            // colsum[n]  += colsum_scale * sum_m out[m, n]
  float* colsum2;           // TANGENT: colsum2[n] += sum_m out2[m, n]
  float* colsum_w;          // colsum_w[n * stride] += sum_m out[m, n] * row_scale[m]
  float colsum_scale;
  int colsum_w_stride;
  int group, ldg;
  int ld_out_lo;
};

struct alignas(64) ChainParams {
  CUtensorMap tmA0;         // initial activation [M, H] (SOFTPLUS3: its hi part), box {32, 128}
  const float* a0_lo;       // SOFTPLUS3: lo part of the initial activation, row pitch a0_lo_ld
  const float* row_scale;   // [M] (sigma) or null
  int a0_lo_ld;
  int M, H, nlayers;
  int vec_ok;
  long long* debug_times;   // profiling only: [grid][kChainMaxLayers][8] clock64 stamps (null in production)
  ChainLayerParams layer[kChainMaxLayers];
};

template <int MODE, bool CG2 = false, int MC = 1>
struct ChainConfig {
  static constexpr bool kS3 = MODE == CHAIN_SOFTPLUS3;
  static constexpr bool kAux2 = MODE == CHAIN_TANGENT || MODE == CHAIN_ADJOINT;
  static constexpr bool kOut2 = MODE == CHAIN_TANGENT;
  static constexpr int kGroups = 2;
  // one k-block of one N-half ([<=128, 32] weight rows); a CTA pair (CG2) stages half of it in each CTA
  static constexpr int kWStage = (CG2 ? 64 : 128) * kBlockK * 4;
  // score-type sweep of the residual CDAE (one aux tile per slot): 8 weight stages + 6 slots, as in chain16_sm100.cuh
  // (the weight ring is bound by stages in flight x L2 -> SM latency); the other modes keep the 6 + 8-tile split
#ifndef ARDAE_CHAIN_WSTAGES_1
#define ARDAE_CHAIN_WSTAGES_1 8
#endif
  static constexpr bool kDeepW = !kS3 && !kAux2 && !CG2 && MC == 1;
  static constexpr int kNumWStages = CG2 ? 12 : (kDeepW ? ARDAE_CHAIN_WSTAGES_1 : 6);
  // One ring serves aux loads AND out stores: a slot receives the aux tile(s) of a chunk by TMA, the epilogue
  // overwrites them IN PLACE with the out tile(s), the TMA store leaves from the same bytes, and the slot is
  // recycled once that store has read it.  128 KB of HBM traffic in flight per SM instead of 64.
  static constexpr int kAuxSlot = kS3 ? 0 : (kAux2 ? 2 : 1) * kTileBytes;
  static constexpr int kAuxTiles = kDeepW ? 14 - kNumWStages : 8;
  static constexpr int kNumAux = kS3 ? 0 : (kAux2 ? 4 : kAuxTiles);
  static constexpr int kOffAux = kNumWStages * kWStage;        // SOFTPLUS3: the 8 hi tiles start here
  static constexpr int kDataBytes = kOffAux + kAuxTiles * kTileBytes;
  static constexpr int kSmemBytes = kDataBytes + 1024 + 512;
  static_assert(kSmemBytes <= 232448 && kNumAux <= 8, "shared memory budget");
  static constexpr int kThreads = 128 + kGroups * 128;
};

// MC > 1: launched as clusters of MC CTAs (adjacent row tiles, independent MMAs).  Every weight stage is fetched ONCE
// per cluster: CTA r loads rows [r*HH/MC, (r+1)*HH/MC) of the k-block and TMA-multicasts them into the same stage of
// all MC CTAs; a stage is recycled when all MC MMA threads have retired it (multicast tcgen05.commit).  The chains
// re-stream 256 KB (3xTF32: 512 KB) of weights per tile and layer from L2 -- as much L2 -> SM traffic as their HBM
// data -- so the tf32 sweeps are L2-bandwidth limited; multicast divides that stream by MC.
// CG2: launched as clusters of two CTAs (adjacent row tiles).  The leader issues tcgen05.mma.cta_group::2 for the
// 256-row pair; every weight k-block is staged half in each CTA, so each SM ingests HALF of the weight stream
// (the 3xTF32 sweep moves 512 KB of weights per tile and layer and is bound by that stream) and the same 96 KB
// ring holds twice as many k-blocks.
template <int MODE, bool CG2 = false, int MC = 1>
__global__ void __launch_bounds__(ChainConfig<MODE, CG2, MC>::kThreads, 1)
chain_kernel(const __grid_constant__ ChainParams p) {
  using Cfg = ChainConfig<MODE, CG2, MC>;
  static_assert(!(CG2 && MC > 1), "CTA pairs and weight multicast are exclusive");
  constexpr uint16_t kMcMask = static_cast<uint16_t>((1u << MC) - 1u);
  constexpr bool S3 = Cfg::kS3, HAS_AUX2 = Cfg::kAux2, HAS_OUT2 = Cfg::kOut2;
  constexpr int G = Cfg::kGroups, NW = Cfg::kNumWStages;
  constexpr int NAUX = Cfg::kNumAux > 0 ? Cfg::kNumAux : 1;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* w_full = reinterpret_cast<uint64_t*>(smem + Cfg::kDataBytes);
  uint64_t* w_empty = w_full + NW;
  uint64_t* aux_full = w_empty + NW;   // [8]
  uint64_t* aux_empty = aux_full + 8;  // [8]
  uint64_t* acc_full = aux_empty + 8;  // [2] accumulator half h of the current layer is complete
  uint64_t* a_ready = acc_full + 2;    // [2] A chunks of half h are written and accumulator half h is drained
  uint64_t* kfree = a_ready + 2;       // [4] the current layer no longer reads A chunk c (c < NB/2)
  uint64_t* a0_full = kfree + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a0_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kBlockM;
  const int H = p.H;
  const int NB = H >> 5;   // 32-column chunks == k-blocks (even: H is a multiple of 64)
  const int NB0 = NB >> 1;  // chunks per N-half
  const int HH = H >> 1;    // columns per N-half
  const int nl = p.nlayers;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.tmA0);
    ptx::prefetch_tmap(&p.layer[0].tmW);
    ptx::prefetch_tmap(&p.layer[0].tmOut);
    if (!S3) ptx::prefetch_tmap(&p.layer[0].tmAux1);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < NW; ++s) {
        ptx::mbar_init(&w_full[s], 1);
        ptx::mbar_init(&w_empty[s], MC);  // MC > 1: every CTA of the cluster retires the (shared) stage
      }
      static_assert(NW <= 12, "barrier area");
      for (int a = 0; a < 8; ++a) {
        ptx::mbar_init(&aux_full[a], 1);
        ptx::mbar_init(&aux_empty[a], 1);  // the store-issuing thread of the group that consumed the slot
      }
      for (int h = 0; h < 2; ++h) {
        ptx::mbar_init(&acc_full[h], 1);
        ptx::mbar_init(&a_ready[h], (CG2 ? 2 : 1) * 4 * G);  // every epilogue warp (of both CTAs of a pair)
      }
      for (int c = 0; c < 4; ++c) ptx::mbar_init(&kfree[c], 1);
      ptx::mbar_init(a0_full, 1);
      ptx::fence_mbar_init();
    }
    __syncwarp();
    if (CG2) {
      ptx::tmem_alloc_2sm(tmem_slot, 512);
      ptx::tmem_relinquish_2sm();
    } else {
      ptx::tmem_alloc(tmem_slot, 512);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CG2 || MC > 1) ptx::cluster_sync_all();  // the peers' barriers exist before anything remote touches them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t rank = (CG2 || MC > 1) ? ptx::cluster_ctarank() : 0;
  const uint32_t acc_t = tmem_base;       // accumulator columns [0, H)
  const uint32_t a_t = tmem_base + 256;   // A operand (SOFTPLUS3: its lo part) columns [256, 256 + H)
  uint8_t* hi_tiles = smem + Cfg::kOffAux;  // SOFTPLUS3 only

  if (warp == 0) {
    // ------------------------------------------------------------ weight producer
    if (ptx::elect_one()) {  // one elected lane: tcgen05 / TMA issue stays on the uniform datapath
      const uint32_t wbytes = static_cast<uint32_t>(HH) * kBlockK * 4;  // per pair when CG2 (half lands in each CTA)
      const int wrows = CG2 ? HH / 2 : HH;
      int it = 0;
      for (int l = 0; l < nl; ++l) {
        const CUtensorMap* tw = &p.layer[l].tmW;
        const int nst = S3 ? 2 * NB : NB;
        for (int h = 0; h < 2; ++h) {
          for (int j = 0; j < nst; ++j, ++it) {
            const int s = it % NW;
            const uint32_t ph = (it / NW) & 1;
            ptx::mbar_wait(&w_empty[s], ph ^ 1);
            // SOFTPLUS3: k-block kb of Whi (columns [0,H)) then of Wlo (columns [2H,3H)); both serve hi, Whi also lo
            const int kc = S3 ? (((j & 1) ? 2 * H : 0) + (j >> 1) * kBlockK) : j * kBlockK;
            if (CG2) {
              // both CTAs' halves complete on the LEADER's full barrier, armed by the leader for both
              if (rank == 0) ptx::mbar_expect_tx(&w_full[s], wbytes);
              ptx::tma_load_2d_2sm(smem + s * Cfg::kWStage, tw, &w_full[s], kc, h * HH + static_cast<int>(rank) * wrows);
            } else if (MC > 1) {
              // this CTA's 1/MC slice of the k-block, delivered to the same stage of every CTA of the cluster; the
              // local barrier expects the whole stage (MC slices from MC producers)
              const int srows = HH / MC;
              ptx::mbar_expect_tx(&w_full[s], wbytes);
              ptx::tma_load_2d_mc(smem + s * Cfg::kWStage + rank * (srows * kBlockK * 4), tw, &w_full[s], kc,
                                  h * HH + static_cast<int>(rank) * srows, kMcMask);
            } else {
              ptx::mbar_expect_tx(&w_full[s], wbytes);
              ptx::tma_load_2d(smem + s * Cfg::kWStage, tw, &w_full[s], kc, h * HH);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if ((!CG2 || rank == 0) && ptx::elect_one()) {  // CG2: the leader issues for the pair; otherwise every CTA
      const uint32_t idesc = ptx::make_idesc_tf32(CG2 ? 2 * kBlockM : kBlockM, HH, 0, 0);
      auto mma_ss = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t acc) {
        if (CG2) ptx::umma_tf32_2sm(d, a, b, idesc, acc); else ptx::umma_tf32(d, a, b, idesc, acc);
      };
      auto mma_ts = [&](uint32_t d, uint32_t a, uint64_t b, uint32_t acc) {
        if (CG2) ptx::umma_tf32_ts_2sm(d, a, b, idesc, acc); else ptx::umma_tf32_ts(d, a, b, idesc, acc);
      };
      auto commit = [&](uint64_t* bar) {  // CG2: the same barrier in both CTAs of the pair
        if (CG2) ptx::umma_commit_2sm(bar, 0x3); else ptx::umma_commit(bar);
      };
      auto commit_w = [&](uint64_t* bar) {  // weight stage retired: MC > 1 tells every CTA of the cluster
        if (MC > 1) ptx::umma_commit_mc(bar, kMcMask); else commit(bar);
      };
      int it = 0;
      for (int l = 0; l < nl; ++l) {
        long long* dbg = p.debug_times ? p.debug_times + (static_cast<size_t>(blockIdx.x) * kChainMaxLayers + l) * 8 : nullptr;
        if (dbg) dbg[0] = clock64();
        long long wait_w = 0, wait_a1 = 0;
        if (S3 && l == 0) ptx::mbar_wait(a0_full, 0);  // hi tiles of the initial activation have landed (TMA)
        for (int h = 0; h < 2; ++h) {
          const uint32_t d_t = acc_t + h * HH;
          for (int kb = 0; kb < NB; ++kb) {
            if (h == 0 && kb == 0) {
              if (CG2) ptx::mbar_wait_cluster(&a_ready[0], l & 1); else
              ptx::mbar_wait(&a_ready[0], l & 1);  // A chunks [0, NB/2) written, accumulator half 0 drained
              ptx::tc_fence_after();
              if (dbg) dbg[1] = clock64();
            }
            if (h == 0 && kb == NB0) {
              long long t0 = dbg ? clock64() : 0;
              if (CG2) ptx::mbar_wait_cluster(&a_ready[1], l & 1); else
              ptx::mbar_wait(&a_ready[1], l & 1);  // A chunks [NB/2, NB) written, accumulator half 1 drained
              ptx::tc_fence_after();
              if (dbg) wait_a1 += clock64() - t0;
            }
            {
              const int s = it % NW;
              const uint32_t ph = (it / NW) & 1;
              long long t0 = dbg ? clock64() : 0;
              ptx::mbar_wait(&w_full[s], ph);
              if (dbg) wait_w += clock64() - t0;
              ptx::tc_fence_after();
              const uint32_t b_addr = ptx::smem_u32(smem + s * Cfg::kWStage);
              if (S3) {
                const uint32_t h_addr = ptx::smem_u32(hi_tiles + kb * kTileBytes);
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                  const uint64_t adesc = ptx::make_smem_desc_sw128(h_addr + k * kUmmaK * 4, 0, 1024);
                  const uint64_t bdesc = ptx::make_smem_desc_sw128(b_addr + k * kUmmaK * 4, 0, 1024);
                  mma_ss(d_t, adesc, bdesc, (kb | k) != 0 ? 1u : 0u);
                }
              }
#pragma unroll
              for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                const uint64_t bdesc = ptx::make_smem_desc_sw128(b_addr + k * kUmmaK * 4, 0, 1024);
                mma_ts(d_t, a_t + kb * kBlockK + k * kUmmaK, bdesc, (S3 || (kb | k) != 0) ? 1u : 0u);
              }
              commit_w(&w_empty[s]);
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
                mma_ss(d_t, adesc, bdesc, 1u);
              }
              commit_w(&w_empty[s]);
              ++it;
            }
            // second half: this layer is done with A chunk kb -> the half-0 epilogue may overwrite it
            if (h == 1 && kb < NB0) commit(&kfree[kb]);
          }
          commit(&acc_full[h]);
        }
        if (dbg) {
          dbg[2] = clock64();
          long long* dbg2 = dbg + static_cast<size_t>(gridDim.x) * kChainMaxLayers * 8;
          dbg2[0] = wait_w; dbg2[1] = wait_a1;
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------ aux / initial-activation producer
    if (ptx::elect_one()) {  // one elected lane: tcgen05 / TMA issue stays on the uniform datapath
      if (S3) {
        if (CG2) {  // both CTAs' hi tiles complete on the leader's barrier (its MMA thread is the only consumer)
          if (rank == 0) ptx::mbar_expect_tx(a0_full, 2u * static_cast<uint32_t>(NB) * kTileBytes);
          for (int c = 0; c < NB; ++c) ptx::tma_load_2d_2sm(hi_tiles + c * kTileBytes, &p.tmA0, a0_full, c * 32, m0);
        } else {
          ptx::mbar_expect_tx(a0_full, static_cast<uint32_t>(NB) * kTileBytes);
          for (int c = 0; c < NB; ++c) ptx::tma_load_2d(hi_tiles + c * kTileBytes, &p.tmA0, a0_full, c * 32, m0);
        }
      } else {
        int it = 0;
        for (int l = -1; l < nl; ++l) {
          for (int c = 0; c < NB; ++c, ++it) {
            const int a = it % NAUX;
            const uint32_t ph = (it / NAUX) & 1;
            ptx::mbar_wait(&aux_empty[a], ph ^ 1);
            uint8_t* slot = smem + Cfg::kOffAux + a * Cfg::kAuxSlot;
            if (l < 0) {
              ptx::mbar_expect_tx(&aux_full[a], kTileBytes);
              ptx::tma_load_2d(slot, &p.tmA0, &aux_full[a], c * 32, m0);
            } else {
              ptx::mbar_expect_tx(&aux_full[a], Cfg::kAuxSlot);
              ptx::tma_load_2d(slot, &p.layer[l].tmAux1, &aux_full[a], c * 32, m0);
              if (HAS_AUX2) ptx::tma_load_2d(slot + kTileBytes, &p.layer[l].tmAux2, &aux_full[a], c * 32, m0);
            }
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
    const int swz = r & 7;
    const uint32_t row_off = static_cast<uint32_t>(r) * 128u;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t bar_a = 1 + 2 * g, bar_b = 2 + 2 * g;
    const float rs = (p.row_scale != nullptr && row_ok) ? p.row_scale[m] : 0.0f;

    // ---- pseudo-layer -1: bring the initial activation into TMEM
    for (int c = g; c < NB; c += G) {
      uint32_t v[32];
      if (S3) {
        const float* src = p.a0_lo + static_cast<size_t>(row_ok ? m : 0) * p.a0_lo_ld + c * 32;
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
        const uint32_t src = ptx::smem_u32(smem + Cfg::kOffAux + a * Cfg::kAuxSlot) + row_off;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 t = lds128(src + ((q ^ swz) << 4));
          v[q * 4 + 0] = __float_as_uint(t.x); v[q * 4 + 1] = __float_as_uint(t.y);
          v[q * 4 + 2] = __float_as_uint(t.z); v[q * 4 + 3] = __float_as_uint(t.w);
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
      if (CG2) {
        ptx::mbar_arrive_cluster(&a_ready[0], 0);
        ptx::mbar_arrive_cluster(&a_ready[1], 0);
      } else {
        ptx::mbar_arrive(&a_ready[0]);
        ptx::mbar_arrive(&a_ready[1]);
      }
    }

    int prev_slot = -1;  // slot whose TMA store may still be reading it (released one chunk later)
#pragma unroll 1
    for (int l = 0; l < nl; ++l) {
      const ChainLayerParams& L = p.layer[l];
      const bool last = (l == nl - 1);
      const float* gb_row = (S3 && L.group_bias != nullptr)
                                ? L.group_bias + static_cast<size_t>((row_ok ? m : 0) / L.group) * L.ldg
                                : nullptr;
      long long* dbg = (p.debug_times && leader && g == 0)
                           ? p.debug_times + (static_cast<size_t>(blockIdx.x) * kChainMaxLayers + l) * 8 : nullptr;
      if (dbg) { dbg[3] = clock64(); dbg[6] = 0; dbg[7] = 0; }
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
      {
        long long tw0 = 0;
        if (dbg) tw0 = clock64();
        ptx::mbar_wait(&acc_full[h], l & 1);
        ptx::tc_fence_after();
        if (dbg) { dbg[7] += clock64() - tw0; if (h == 0) dbg[4] = clock64(); }
      }
      if (S3 && h == 0) {
        // the hi tiles are rewritten in place: the TMA stores that spilled the previous layer must have read them
        if (leader) ptx::tma_store_wait_read<0>();
        ptx::named_bar_sync(bar_a, 128);
      }
#pragma unroll 1
      for (int c = h * NB0 + g; c < (h + 1) * NB0; c += G) {
        const int nc = c * 32;
        if (h == 0) {  // the half-1 MMAs of this layer still read A chunk c until kfree[c] fires
          ptx::mbar_wait(&kfree[c], l & 1);
          ptx::tc_fence_after();
        }
        uint32_t accu[32];
        ptx::tmem_ld_32x32(acc_t + lane_addr + nc, accu);
        const int it = NB * (l + 1) + c;
        const int a = it % NAUX;
        long long tw0 = 0;
        if (dbg) tw0 = clock64();
        if (!S3) ptx::mbar_wait(&aux_full[a], (it / NAUX) & 1);
        if (dbg) dbg[6] += clock64() - tw0;  // time spent waiting for aux tiles
        ptx::tmem_ld_wait();
        // aux tile(s) in, out tile(s) written over them in place (same thread, same bytes)
        uint8_t* slot = S3 ? hi_tiles + c * kTileBytes : smem + Cfg::kOffAux + a * Cfg::kAuxSlot;
        const uint32_t s1 = ptx::smem_u32(slot) + row_off;  // aux1 / out (and hi tile) row of this thread
        const uint32_t s2 = s1 + kTileBytes;                // aux2 / out2
        float v[32];  // A operand of the next layer (tf32): `out`, or the lo part for SOFTPLUS3
        if (S3) {
          float add[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) add[j] = 0.0f;
          const bool vec = p.vec_ok != 0;
          if (L.bias != nullptr) add_cols32(L.bias + nc, 1.0f, vec, 32, add);
          if (gb_row != nullptr) add_cols32(gb_row + nc, 1.0f, vec, 32, add);
          if (L.col_vec != nullptr) add_cols32(L.col_vec + nc, rs, vec, 32, add);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float res = softplus_fast(__uint_as_float(accu[q * 4 + j]) + add[q * 4 + j]);
              const float hi = round_tf32_fast(res);
              o[j] = hi;
              v[q * 4 + j] = round_tf32_fast(res - hi);
            }
            sts128(s1 + ((q ^ swz) << 4), o[0], o[1], o[2], o[3]);
          }
          if (L.out_lo != nullptr && row_ok) {  // one layer per step at most: plain row stores
            float4* dst = reinterpret_cast<float4*>(L.out_lo + static_cast<size_t>(m) * L.ld_out_lo + nc);
#pragma unroll
            for (int q = 0; q < 8; ++q) dst[q] = make_float4(v[q * 4 + 0], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
          }
        } else {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            float4 x1[4], x2[4];
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
              const uint32_t soff = static_cast<uint32_t>(((half * 4 + qq) ^ swz) << 4);
              x1[qq] = lds128(s1 + soff);
              if (HAS_AUX2) x2[qq] = lds128(s2 + soff);
            }
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
              const int q = half * 4 + qq;
              const uint32_t soff = static_cast<uint32_t>((q ^ swz) << 4);
              const float u[4] = {x1[qq].x, x1[qq].y, x1[qq].z, x1[qq].w};
              float o[4], ob[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float w2 = HAS_AUX2 ? (j == 0 ? x2[qq].x : (j == 1 ? x2[qq].y : (j == 2 ? x2[qq].z : x2[qq].w))) : 0.0f;
                const float pre = __uint_as_float(accu[q * 4 + j]);
                float sg, oms;
                sig_fast(u[j], sg, oms);
                float res, res2 = 0.0f;
                if (MODE == CHAIN_MUL_SIG) {
                  res = pre * sg;
                } else if (MODE == CHAIN_TANGENT) {
                  res = pre * sg;
                  res2 = round_tf32_fast(w2 * pre * oms);
                } else {
                  res = fmaf(pre, sg, w2);
                }
                res = round_tf32_fast(res);
                o[j] = res;
                ob[j] = res2;
                v[q * 4 + j] = res;
              }
              sts128(s1 + soff, o[0], o[1], o[2], o[3]);
              if (HAS_OUT2) sts128(s2 + soff, ob[0], ob[1], ob[2], ob[3]);
            }
          }
        }
        if (!last) {
          uint32_t vu[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) vu[j] = __float_as_uint(v[j]);
          ptx::tmem_st_32x32(a_t + lane_addr + nc, vu);
        }
        ptx::fence_proxy_async_smem();
        ptx::named_bar_sync(bar_b, 128);
        if (leader) {
          ptx::tma_store_2d(&L.tmOut, slot, nc, m0);
          if (HAS_OUT2) ptx::tma_store_2d(&L.tmOut2, slot + kTileBytes, nc, m0);
          ptx::tma_store_commit();
          if (!S3) {
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
            for (int q = 0; q < 8; ++q) {
              const float4 t4 = lds128(s2 + ((q ^ swz) << 4));
              v[q * 4 + 0] = row_ok ? t4.x : 0.0f; v[q * 4 + 1] = row_ok ? t4.y : 0.0f;
              v[q * 4 + 2] = row_ok ? t4.z : 0.0f; v[q * 4 + 3] = row_ok ? t4.w : 0.0f;
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
      if (lane == 0) {
        if (CG2) ptx::mbar_arrive_cluster(&a_ready[h], 0); else ptx::mbar_arrive(&a_ready[h]);
      }
      if (dbg && h == 1) dbg[5] = clock64();
      if (HAS_OUT2 && L.colsum2 != nullptr) ptx::named_bar_sync(bar_a, 128);  // colsum2 re-reads of the slot are done
      if (!S3 && leader && prev_slot >= 0) {  // do not sit on a slot while waiting for the next accumulator half
        ptx::tma_store_wait_read<0>();
        ptx::mbar_arrive(&aux_empty[prev_slot]);
        prev_slot = -1;
      }
      }  // h
    }
    if (leader) ptx::tma_store_wait_all<0>();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (CG2 || MC > 1) ptx::cluster_sync_all();  // no CTA exits while a peer may still signal its barriers / write its memory
  if (warp == 1) {
    ptx::tc_fence_after();
    if (CG2) ptx::tmem_dealloc_2sm(tmem_base, 512); else ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace ardae
