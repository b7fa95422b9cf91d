// 16-epilogue-warp variant of the bf16-spill chain kernel (chain16_sm100.cuh) for the score / tangent / adjoint sweeps.
// The 8-warp kernel is paced by epilogue instruction issue (two warps per scheduler; measured period per layer and CTA:
// 5.3 / 7.4 / 5.8 us against an MMA floor of 2.2 us and an HBM share of 3.0 / 6.0 / 4.5 us).  Here four groups of four
// warps own the chunks c % 4 == g and work through them in 16-column halves, which keeps the kernel at <= 96 registers
// per thread for 640 threads.  Producer / MMA / aux-producer warps, barriers, the in-place aux -> out slot ring and the
// TMA stores are those of chain16_kernel; ARDAE_CHAIN16_WARPS=8 selects the 8-warp kernel (A/B measurements).
#pragma once
#include "chain_s3h_sm100.cuh"

namespace ardae {

struct Chain16wConfig {
  static constexpr int kGroups = 4;
  static constexpr int kThreads = 128 + kGroups * 128;
};

// registers -> TMEM: thread i of the warp writes lane base+i, 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
// After this, lanes l and l ^ 16 hold the sum over all 32 lanes of v[l & 15] (15 + 1 shuffles).
__device__ __forceinline__ float warp_transpose_reduce16(float (&v)[16], int lane) {
#pragma unroll
  for (int off = 8; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16);
}

// CG2: launched as clusters of two CTAs (adjacent row tiles); the leader issues tcgen05.mma.cta_group::2 for the 256-row
// pair and every weight k-block is staged HALF in each CTA, so each SM ingests half of the weight stream -- the tf32
// sweeps move 256 KB of weights per tile and layer through the L2 -> SM path, more than their aux / spill bytes, and
// are bound by it.
// MC: clusters of two CTAs on adjacent row tiles that stay independent (own MMAs, own accumulators) but SHARE the weight
// stream: each CTA fetches half of every weight k-block and TMA-multicasts it into both shared memories, so the
// L2 -> SM weight traffic per SM halves without the pair's MMA lock-step (a stage is recycled when both CTAs' MMAs
// have consumed it: the multicast commit arrives on both empty barriers).
template <int MODE, bool CG2, bool MC = false>
__global__ void __launch_bounds__(Chain16wConfig::kThreads, 1)
chain16w_kernel(const __grid_constant__ Chain16Params p) {
  static_assert(!(CG2 && MC), "CG2 and MC are exclusive");
  using Cfg = Chain16Config<MODE>;
  static_assert(MODE != CHAIN_SOFTPLUS3, "the primal sweep has its own kernels");
  constexpr bool S3 = false, HAS_AUX2 = Cfg::kAux2, HAS_OUT2 = Cfg::kOut2;
  constexpr int G = Chain16wConfig::kGroups;
  constexpr int NW = CG2 ? 2 * Cfg::kNumWStages : Cfg::kNumWStages;  // CG2: twice as many half-size stages
  constexpr int kWStage = CG2 ? Cfg::kWStage / 2 : Cfg::kWStage;
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
        ptx::mbar_init(&w_empty[s], MC ? 2 : 1);  // MC: both CTAs' MMAs release a stage
      }
      for (int a = 0; a < 16; ++a) {
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
  if (CG2 || MC) ptx::cluster_sync_all();  // the peer's barriers exist before anything remote touches them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t rank = (CG2 || MC) ? ptx::cluster_ctarank() : 0;
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
        const int nrows = narrow ? L.nout : HH;  // B rows of one N-half (per pair when CG2: half lands in each CTA)
        const uint32_t wbytes = static_cast<uint32_t>(nrows) * kBlockK * 4;
        const int wrows = (CG2 || MC) ? nrows / 2 : nrows;
        for (int h = 0; h < (narrow ? 1 : 2); ++h) {
          for (int j = 0; j < nst; ++j, ++it) {
            const int s = it % NW;
            const uint32_t ph = (it / NW) & 1;
            ptx::mbar_wait(&w_empty[s], ph ^ 1);
            const int kc = j * kBlockK;
            if (CG2) {
              // both CTAs' halves complete on the LEADER's full barrier, armed by the leader for both
              if (rank == 0) ptx::mbar_expect_tx(&w_full[s], wbytes);
              ptx::tma_load_2d_2sm(smem + s * kWStage, tw, &w_full[s], kc, h * HH + static_cast<int>(rank) * wrows);
            } else if (MC) {
              // the whole k-block lands here (this CTA's half + the peer's half); this CTA issues its half for both
              ptx::mbar_expect_tx(&w_full[s], wbytes);
              ptx::tma_load_2d_mc(smem + s * kWStage + rank * (static_cast<uint32_t>(wrows) * kBlockK * 4), tw, &w_full[s],
                                  kc, h * HH + static_cast<int>(rank) * wrows, 0x3);
            } else {
              ptx::mbar_expect_tx(&w_full[s], wbytes);
              ptx::tma_load_2d(smem + s * kWStage, tw, &w_full[s], kc, h * HH);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (CG2: the leader issues for the pair)
    if ((!CG2 || rank == 0) && ptx::elect_one()) {
      auto mma_ts = [&](uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
        if (CG2) ptx::umma_tf32_ts_2sm(d, a, b, idesc, acc); else ptx::umma_tf32_ts(d, a, b, idesc, acc);
      };
      auto commit = [&](uint64_t* bar) {  // CG2: the same barrier in both CTAs of the pair
        if (CG2) ptx::umma_commit_2sm(bar, 0x3); else ptx::umma_commit(bar);
      };
      auto commit_w = [&](uint64_t* bar) {  // weight stage consumed (MC: tell both CTAs' producers)
        if (MC) ptx::umma_commit_mc(bar, 0x3); else commit(bar);
      };
      auto wait_a = [&](uint64_t* bar, uint32_t ph) {
        if (CG2) ptx::mbar_wait_cluster(bar, ph); else ptx::mbar_wait(bar, ph);
        ptx::tc_fence_after();
      };
      int it = 0;
      for (int l = 0; l < nl; ++l) {
        const Chain16LayerParams& L = p.layer[l];
        const int NBl = L.kin >> 5;
        const bool narrow = L.nout < H;
        const uint32_t idesc = ptx::make_idesc_tf32(CG2 ? 2 * kBlockM : kBlockM, narrow ? L.nout : HH, 0, 0);
        for (int h = 0; h < (narrow ? 1 : 2); ++h) {
          const uint32_t d_t = acc_t + h * HH;
          for (int kb = 0; kb < NBl; ++kb) {
            if (h == 0 && kb == 0) wait_a(&a_ready[0], l & 1);    // A chunks [0, NB/2) written, accumulator half 0 drained
            if (h == 0 && kb == NB0) wait_a(&a_ready[1], l & 1);  // A chunks [NB/2, NB) written, accumulator half 1 drained
            const int s = it % NW;
            const uint32_t ph = (it / NW) & 1;
            ptx::mbar_wait(&w_full[s], ph);
            ptx::tc_fence_after();
            const uint32_t b_addr = ptx::smem_u32(smem + s * kWStage);
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
              const uint64_t bdesc = ptx::make_smem_desc_sw128(b_addr + k * kUmmaK * 4, 0, 1024);
              mma_ts(d_t, a_t + kb * kBlockK + k * kUmmaK, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            commit_w(&w_empty[s]);
            ++it;
            // second half: this layer is done with A chunk kb -> the half-0 epilogue may overwrite it
            if (h == 1 && kb < NB0) commit(&kfree[kb]);
          }
          if (h == 0 && NBl <= NB0 && !narrow) wait_a(&a_ready[1], l & 1);  // short first layer: consume this phase too
          if (h == 1)
            for (int c = NBl; c < NB0; ++c) commit(&kfree[c]);  // chunks this layer never read
          commit(&acc_full[h]);
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
    // ------------------------------------------------------------ epilogue warps: four groups of four warps, group g
    // owns the 32-column chunks c % 4 == g and works through them in two 16-column halves (<= 96 registers / thread)
    const int g = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int m = m0 + r;
    const bool row_ok = m < p.M;
    const bool leader = (quarter == 0 && lane == 0);
    const int sw16 = (r >> 1) & 3;               // bf16 tiles (64-byte rows, SWIZZLE_64B)
    const uint32_t row_off16 = static_cast<uint32_t>(r) * 64u;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t bar_a = 1 + 2 * g, bar_b = 2 + 2 * g;
    const float rs = (p.row_scale != nullptr && row_ok) ? p.row_scale[m] : 0.0f;

    // ---- pseudo-layer -1: bring the initial activation into TMEM
    for (int c = g; c < NBin0; c += G) {
      if (!a0_ring) {
        const float* src = p.a0_f32 + static_cast<size_t>(row_ok ? m : 0) * p.a0_ld + c * 32;
#pragma unroll 1
        for (int sub = 0; sub < 2; ++sub) {
          uint32_t v[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row_ok) t = __ldg(reinterpret_cast<const float4*>(src + sub * 16) + q);
            v[q * 4 + 0] = __float_as_uint(t.x); v[q * 4 + 1] = __float_as_uint(t.y);
            v[q * 4 + 2] = __float_as_uint(t.z); v[q * 4 + 3] = __float_as_uint(t.w);
          }
          tmem_st_32x16(a_t + lane_addr + c * 32 + sub * 16, v);
        }
      } else {
        const int it = c;
        const int a = it % NAUX;
        ptx::mbar_wait(&aux_full[a], (it / NAUX) & 1);
        const uint32_t src = ptx::smem_u32(smem + Cfg::kOffAux + a * Cfg::kAuxSlot) + row_off16;
#pragma unroll 1
        for (int sub = 0; sub < 2; ++sub) {
          uint32_t v[16];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const uint4 t = lds128u(src + (((sub * 2 + q) ^ sw16) << 4));
            const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              v[q * 8 + 2 * i] = w[i] << 16;
              v[q * 8 + 2 * i + 1] = w[i] & 0xFFFF0000u;
            }
          }
          tmem_st_32x16(a_t + lane_addr + c * 32 + sub * 16, v);
        }
        ptx::named_bar_sync(bar_a, 128);  // all four warps have read the slot
        if (leader) ptx::mbar_arrive(&aux_empty[a]);
      }
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
      const Chain16LayerParams& L = p.layer[l];
      const bool last = (l == nl - 1);
      if (L.nout < H) {
        // ---- narrow linear last layer (score sweep: g = delta a_1 . A_1): fp32 rows straight to global
        ptx::mbar_wait(&acc_full[0], l & 1);
        ptx::tc_fence_after();
        for (int c = g; c * 32 < L.nout; c += G) {
#pragma unroll 1
          for (int sub = 0; sub < 2; ++sub) {
            uint32_t accu[16];
            tmem_ld_32x16(acc_t + lane_addr + c * 32 + sub * 16, accu);
            ptx::tmem_ld_wait();
            if (row_ok) {
              float4* dst = reinterpret_cast<float4*>(L.out32 + static_cast<size_t>(m) * L.ld_out32 + c * 32 + sub * 16);
#pragma unroll
              for (int q = 0; q < 4; ++q)
                dst[q] = make_float4(__uint_as_float(accu[q * 4 + 0]), __uint_as_float(accu[q * 4 + 1]),
                                     __uint_as_float(accu[q * 4 + 2]), __uint_as_float(accu[q * 4 + 3]));
            }
          }
        }
        break;
      }
      const bool want_cs = L.colsum != nullptr || L.colsum_w != nullptr;
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        long long* dbg = (p.dbg != nullptr && blockIdx.x == 0 && quarter == 1 && lane == 0)
                             ? p.dbg + ((static_cast<size_t>(l) * 2 + h) * G + g) * 8 : nullptr;
        if (dbg) dbg[0] = clock64();
        ptx::mbar_wait(&acc_full[h], l & 1);
        ptx::tc_fence_after();
        if (dbg) dbg[1] = clock64();
#pragma unroll 1
        for (int c = h * NB0 + g; c < (h + 1) * NB0; c += G) {
          const int nc = c * 32;
          if (h == 0) {  // the half-1 MMAs of this layer still read A chunk c until kfree[c] fires
            ptx::mbar_wait(&kfree[c], l & 1);
            ptx::tc_fence_after();
          }
          if (dbg) dbg[2] = clock64();
          const int it = ring_base + NB * l + c;
          const int a = it % NAUX;
          ptx::mbar_wait(&aux_full[a], (it / NAUX) & 1);
          if (dbg) dbg[3] = clock64();
          uint8_t* slot = smem + Cfg::kOffAux + a * Cfg::kAuxSlot;  // aux tile(s) in, out tile(s) over them in place
          const uint32_t s1 = ptx::smem_u32(slot) + row_off16;      // aux1 / out row of this thread
          const uint32_t s2 = s1 + kTile16Bytes;                    // aux2 / out2
#pragma unroll 1
          for (int sub = 0; sub < 2; ++sub) {
            uint32_t accu[16];
            tmem_ld_32x16(acc_t + lane_addr + nc + sub * 16, accu);
            uint4 x1[2], x2[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const uint32_t soff = static_cast<uint32_t>(((sub * 2 + q) ^ sw16) << 4);
              x1[q] = lds128u(s1 + soff);
              if (HAS_AUX2) x2[q] = lds128u(s2 + soff);
            }
            ptx::tmem_ld_wait();
            float v[16];  // A operand of the next layer (tf32)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const uint32_t soff = static_cast<uint32_t>(((sub * 2 + q) ^ sw16) << 4);
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
            if (!last) {
              uint32_t vu[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) vu[j] = __float_as_uint(v[j]);
              tmem_st_32x16(a_t + lane_addr + nc + sub * 16, vu);
            }
            // ---- fused column sums (bias gradients, d w_sigma, d w_o) of these 16 columns
            if (want_cs) {
              if (!row_ok) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = 0.0f;
              }
              if (L.colsum_w != nullptr) {
                float w[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) w[i] = v[i] * rs;
                const float t = warp_transpose_reduce16(w, lane);
                if (lane < 16) atomicAdd(L.colsum_w + static_cast<size_t>(nc + sub * 16 + lane) * L.colsum_w_stride, t);
              }
              if (L.colsum != nullptr) {
                const float t = warp_transpose_reduce16(v, lane);
                if (lane < 16) atomicAdd(L.colsum + nc + sub * 16 + lane, L.colsum_scale * t);
              }
            }
          }
          if (dbg) dbg[4] = clock64();
          ptx::fence_proxy_async_smem();
          ptx::named_bar_sync(bar_b, 128);
          if (dbg) dbg[5] = clock64();
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
          if (HAS_OUT2 && L.colsum2 != nullptr) {
            // column sums of out2 from the staging tile (still intact: the next write to it happens after this thread
            // passes the next barrier)
#pragma unroll 1
            for (int sub = 0; sub < 2; ++sub) {
              float v2[16];
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                const uint4 t4 = lds128u(s2 + (((sub * 2 + q) ^ sw16) << 4));
                const uint32_t w[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  v2[q * 8 + 2 * i] = row_ok ? bf16_lo(w[i]) : 0.0f;
                  v2[q * 8 + 2 * i + 1] = row_ok ? bf16_hi(w[i]) : 0.0f;
                }
              }
              const float t = warp_transpose_reduce16(v2, lane);
              if (lane < 16) atomicAdd(L.colsum2 + nc + sub * 16 + lane, t);
            }
          }
        }
        // A chunks of this half written (TMEM stores complete, shared-memory writes fenced), accumulator half drained
        if (dbg) dbg[6] = clock64();
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG2) ptx::mbar_arrive_cluster(&a_ready[h], 0); else ptx::mbar_arrive(&a_ready[h]);
        }
        if (dbg) dbg[7] = clock64();
        if (HAS_OUT2 && L.colsum2 != nullptr) ptx::named_bar_sync(bar_a, 128);  // colsum2 re-reads of the slot are done
        if (leader && prev_slot >= 0) {  // do not sit on a slot while waiting for the next accumulator half
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
  if (CG2 || MC) ptx::cluster_sync_all();  // no CTA exits while its peer may still signal its barriers / read its memory
  if (warp == 1) {
    ptx::tc_fence_after();
    if (CG2) ptx::tmem_dealloc_2sm(tmem_base, 512); else ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace ardae
