// Small HBM-bound kernels around the GEMM chain: weight derivation (tf32-rounded copies and
// transposes), noise / perturbation prologues, loss epilogues, reductions, optimizers.
// All are grid-stride, coalesced over the contiguous dimension, vectorised where rows allow it.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "gemm_sm100.cuh"  // softplus_f
#include "ptx_sm100.cuh"

namespace ardae {

constexpr float kLog2Pi = 1.8378770664093453f;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Block-wide sum; result valid in thread 0.  blockDim.x <= 1024.
__device__ __forceinline__ float block_sum(float v) {
  __shared__ float red[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    v = (lane < (blockDim.x + 31) / 32) ? red[lane] : 0.0f;
    v = warp_sum(v);
  }
  return v;
}

// bfloat16 <-> fp32 (round to nearest even; finite inputs)
__device__ __forceinline__ uint16_t f32_to_bf16_rn(float x) {
  const uint32_t u = __float_as_uint(x);
  return static_cast<uint16_t>((u + 0x7FFFu + ((u >> 16) & 1u)) >> 16);
}
__device__ __forceinline__ float bf16_to_f32(uint16_t h) { return __uint_as_float(static_cast<uint32_t>(h) << 16); }

// ---------------------------------------------------------------- Philox4x32-10 + Box-Muller
// Replay counter: a CUDA-graph-captured step re-launches the same kernel arguments on every replay, so anything that
// must change per step (Philox seeds, Adam's bias-correction step) is offset by a device-resident counter that a tiny
// kernel bumps at the end of each replay.  `ctr == nullptr` (eager calls) means offset 0.
typedef unsigned long long replay_ctr_t;
inline const replay_ctr_t*& replay_counter() {
  static const replay_ctr_t* p = nullptr;
  return p;
}
// The host derives the seed of draw k of iteration t as base + (t * kReplaySeedStride + k) * GOLDEN (k < stride);
// a replay advances t by the counter, so replayed and eager iterations draw from disjoint seed sets (a stride of 1
// made draw k+1 of replay r collide with draw k of replay r+1).
constexpr unsigned long long kReplaySeedStride = 64ull;
__device__ __forceinline__ uint64_t replay_seed(uint64_t seed, const replay_ctr_t* ctr) {
  return ctr != nullptr ? seed + static_cast<uint64_t>(*ctr) * (0x9E3779B97F4A7C15ull * kReplaySeedStride) : seed;
}
__global__ void bump_counter_kernel(replay_ctr_t* ctr) { *ctr += 1ull; }

struct Philox {
  static __device__ __forceinline__ uint4 round4(uint4 c, uint2 k) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    return make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
  }
  static __device__ __forceinline__ uint4 gen(uint64_t seed, uint64_t ctr, uint32_t stream) {
    uint4 c = make_uint4(static_cast<uint32_t>(ctr), static_cast<uint32_t>(ctr >> 32), stream, 0u);
    uint2 k = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      c = round4(c, k);
      k.x += 0x9E3779B9u;
      k.y += 0xBB67AE85u;
    }
    return c;
  }
};
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  const float u1 = (static_cast<float>(a) + 1.0f) * 2.3283064365386963e-10f;  // (0,1]
  const float u2 = static_cast<float>(b) * 2.3283064365386963e-10f;
  const float r = sqrtf(-2.0f * __logf(u1));
  float s, c;
  __sincosf(6.283185307179586f * u2, &s, &c);
  return make_float2(r * c, r * s);
}
// out[i] = N(0,1), i < n   (4 normals per Philox call)
__global__ void randn_kernel(float* __restrict__ out, size_t n, uint64_t seed, uint32_t stream,
                             const replay_ctr_t* __restrict__ ctr) {
  seed = replay_seed(seed, ctr);
  const size_t nq = (n + 3) / 4;
  for (size_t q = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; q < nq;
       q += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint4 r = Philox::gen(seed, q, stream);
    const float2 a = box_muller(r.x, r.y), b = box_muller(r.z, r.w);
    const float v[4] = {a.x, a.y, b.x, b.y};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (q * 4 + j < n) out[q * 4 + j] = v[j];
  }
}

// ---------------------------------------------------------------- weight derivation
// For every parameter matrix W[rows, cols] (nn.Linear layout) produce, tf32-rounded:
//   dst [rows, ld]   straight copy (ld = padded pitch, pad columns zero)
//   dstT[cols, ldT]  transpose     (the B operand of the backward-data GEMMs)
//   dst3[rows, ld3] "3xTF32" B operand [W_hi | W_hi | W_lo], each segment kp columns wide
//                   (kp = cols rounded up to the 32-column k-block; pad columns stay zero)
struct DeriveItem {
  const float* src;
  float* dst;
  float* dstT;
  float* dst3;
  int rows, cols, src_ld, ld, ldT, kp, ld3;
  int first_block;  // prefix sum over 32x32 tiles
  int tiles_x;      // ceil(ld/32)
  uint16_t* dst16;  // fp16 [rows, 2*k16] = [W_hi | W_lo] of w * 2^4 (chain_s3h_sm100.cuh), pad columns zero
  int k16;
};
__global__ void derive_weights_kernel(const DeriveItem* __restrict__ items, int nitems) {
  __shared__ float tile[32][33];
  int it = 0;
  while (it + 1 < nitems && static_cast<int>(blockIdx.x) >= items[it + 1].first_block) ++it;
  const DeriveItem d = items[it];
  const int t = blockIdx.x - d.first_block;
  const int tx = t % d.tiles_x, ty = t / d.tiles_x;
  const int c0 = tx * 32, r0 = ty * 32;
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int r = r0 + ly + j, c = c0 + lx;
    float v = 0.0f;
    if (r < d.rows && c < d.cols) {
      const float w = d.src[static_cast<size_t>(r) * d.src_ld + c];
      v = ptx::round_tf32(w);
      if (d.dst3 != nullptr) {
        float* o = d.dst3 + static_cast<size_t>(r) * d.ld3 + c;
        o[0] = v;
        o[d.kp] = v;
        o[2 * d.kp] = ptx::round_tf32(w - v);
      }
      if (d.dst16 != nullptr) {
        const float ws = fminf(fmaxf(w * 16.0f, -65000.0f), 65000.0f);
        const __half hi = __float2half_rn(ws);
        const __half lo = __float2half_rn(ws - __half2float(hi));
        uint16_t* o = d.dst16 + static_cast<size_t>(r) * (2 * d.k16) + c;
        o[0] = __half_as_ushort(hi);
        o[d.k16] = __half_as_ushort(lo);
      }
    }
    tile[ly + j][lx] = v;
    if (d.dst != nullptr && r < d.rows && c < d.ld) d.dst[static_cast<size_t>(r) * d.ld + c] = v;
  }
  __syncthreads();
  if (d.dstT != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      const int c = c0 + ly + j, r = r0 + lx;  // transposed: row index c of dstT, col r
      // an item owns round_up(rows, 4) columns of its dstT rows (several items may share one buffer)
      if (c < d.cols && r < ((d.rows + 3) & ~3)) d.dstT[static_cast<size_t>(c) * d.ldT + r] = (r < d.rows) ? tile[lx][ly + j] : 0.0f;
    }
  }
}

// ---------------------------------------------------------------- generic 2-D copy / affine
// dst[r, c] = round?( a * src[r, c] + b ),  c < cols;  zero for cols <= c < dst_cols
__global__ void copy2d_kernel(const float* __restrict__ src, int src_ld, float* __restrict__ dst,
                              int dst_ld, int rows, int cols, int dst_cols, float a, float b,
                              int round) {
  const size_t total = static_cast<size_t>(rows) * dst_cols;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / dst_cols), c = static_cast<int>(i - static_cast<size_t>(r) * dst_cols);
    float v = 0.0f;
    if (c < cols) {
      v = a * src[static_cast<size_t>(r) * src_ld + c] + b;
      if (round) v = ptx::round_tf32(v);
    }
    dst[static_cast<size_t>(r) * dst_ld + c] = v;
  }
}

// dst[r, c] = hi, dst[r, kp + c] = lo of (a * src[r, c] + b), c < cols  (tf32 pair, see gemm SPLIT)
__global__ void split2d_kernel(const float* __restrict__ src, int src_ld, float* __restrict__ dst,
                               int dst_ld, int rows, int cols, int kp, float a, float b) {
  const size_t total = static_cast<size_t>(rows) * cols;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / cols), c = static_cast<int>(i - static_cast<size_t>(r) * cols);
    const float v = a * src[static_cast<size_t>(r) * src_ld + c] + b;
    const float hi = ptx::round_tf32(v);
    dst[static_cast<size_t>(r) * dst_ld + c] = hi;
    dst[static_cast<size_t>(r) * dst_ld + kp + c] = ptx::round_tf32(v - hi);
  }
}

// out[b, c] = sum_{k<S} in[(b*S+k), c]
__global__ void group_sum_kernel(const float* __restrict__ in, int in_ld, float* __restrict__ out,
                                 int out_ld, int B, int S, int cols, int round) {
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    float acc = 0.0f;
    const float* p = in + static_cast<size_t>(b) * S * in_ld + c;
    for (int k = 0; k < S; ++k) acc += p[static_cast<size_t>(k) * in_ld];
    out[static_cast<size_t>(b) * out_ld + c] = round ? ptx::round_tf32(acc) : acc;
  }
}

// out[i] = (u_i < p[i]) ? 1 : 0, u ~ U[0,1) from Philox: dynamic binarisation of grey-level images on the device
// (the reference applies torch.bernoulli as a DataLoader transform: datasets/mnist.py:39-40)
__global__ void bernoulli_kernel(const float* __restrict__ p, float* __restrict__ out, size_t n, uint64_t seed,
                                 const replay_ctr_t* __restrict__ ctr) {
  seed = replay_seed(seed, ctr);
  const size_t nq = (n + 3) / 4;
  for (size_t q = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; q < nq;
       q += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint4 r = Philox::gen(seed, q, 17u);
    const uint32_t u[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (q * 4 + j < n) out[q * 4 + j] = (static_cast<float>(u[j] >> 8) * 5.9604644775390625e-08f < p[q * 4 + j]) ? 1.0f : 0.0f;
  }
}

// dst pair (hi | lo at +kp) = act(src) : activation of an fp32 [rows, cols] block, stored as a tf32 pair
// (mean-code branch of the encoder: the noise half of fc layer 0 vanishes for eps = 0)
__global__ void act_split_kernel(const float* __restrict__ src, int ld, float* __restrict__ dst, int ldd, int rows,
                                 int cols, int kp, int act) {
  const size_t total = static_cast<size_t>(rows) * cols;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / cols), c = static_cast<int>(i - static_cast<size_t>(r) * cols);
    const float a = src[static_cast<size_t>(r) * ld + c];
    const float v = act == 1 ? softplus_f(a) : fmaxf(a, 0.0f);
    const float hi = ptx::round_tf32(v);
    dst[static_cast<size_t>(r) * ldd + c] = hi;
    dst[static_cast<size_t>(r) * ldd + kp + c] = ptx::round_tf32(v - hi);
  }
}

// ---------------------------------------------------------------- CDAE prologue / epilogues
// x~[n, j] = x[n, j] + sigma[n] * eps[n, j]  (models/graddae/mlp.py:21-23), stored as the tf32
// pair xt[n, j] = hi, xt[n, kp + j] = lo.  If gen_eps, eps is first drawn here (Philox) and
// stored; eps == nullptr means "no noise" (glogprob).
__global__ void cdae_perturb_kernel(const float* __restrict__ x, const float* __restrict__ sigma,
                                    float* __restrict__ eps, float* __restrict__ xt, int N, int d,
                                    int ldx, int kp, int gen_eps, uint64_t seed,
                                    const replay_ctr_t* __restrict__ ctr, uint16_t* __restrict__ x16h = nullptr,
                                    uint16_t* __restrict__ x16l = nullptr, int ld16 = 0) {
  seed = replay_seed(seed, ctr);
  const size_t total = static_cast<size_t>(N) * d;
  for (size_t e = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(e / d), j = static_cast<int>(e - static_cast<size_t>(n) * d);
    float v = x[e];
    if (eps != nullptr) {
      float ev;
      if (gen_eps) {
        const uint4 r = Philox::gen(seed, e >> 2, 7u);
        const float2 a = box_muller(r.x, r.y), b = box_muller(r.z, r.w);
        const float q[4] = {a.x, a.y, b.x, b.y};
        ev = q[e & 3];
        eps[e] = ev;
      } else {
        ev = eps[e];
      }
      v += sigma[n] * ev;
    }
    const float hi = ptx::round_tf32(v);
    xt[static_cast<size_t>(n) * ldx + j] = hi;
    xt[static_cast<size_t>(n) * ldx + kp + j] = ptx::round_tf32(v - hi);
    if (x16h != nullptr) {  // bf16 (hi, lo) pair of x~: the Y operand of the first-layer weight-gradient contraction
      const uint16_t h16 = f32_to_bf16_rn(v);
      x16h[static_cast<size_t>(n) * ld16 + j] = h16;
      x16l[static_cast<size_t>(n) * ld16 + j] = f32_to_bf16_rn(v - bf16_to_f32(h16));
    }
  }
}

// bf16 variant of cdae_init_delta_kernel (16-bit spill plans): 8 elements per thread; H % 8 == 0
__global__ void cdae_init_delta16_kernel(const uint16_t* __restrict__ vL, int ld, const float* __restrict__ wo,
                                         uint16_t* __restrict__ dp, int ld_dp, int N, int H) {
  const int H8 = H >> 3;
  const size_t total = static_cast<size_t>(N) * H8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i / H8), j = static_cast<int>(i - static_cast<size_t>(n) * H8) * 8;
    const uint4 u4 = *reinterpret_cast<const uint4*>(vL + static_cast<size_t>(n) * ld + j);
    const uint32_t uw[4] = {u4.x, u4.y, u4.z, u4.w};
    uint32_t ow[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float o[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float u = __uint_as_float(e ? (uw[k] & 0xFFFF0000u) : (uw[k] << 16));
        const float ex = __expf(-u);
        const float s = (u < 0.01f) ? u * (1.0f - u * (0.5f - u * (1.0f / 6.0f))) : 1.0f - ex;
        o[e] = -__ldg(wo + j + 2 * k + e) * s;
      }
      ow[k] = static_cast<uint32_t>(f32_to_bf16_rn(o[0])) | (static_cast<uint32_t>(f32_to_bf16_rn(o[1])) << 16);
    }
    *reinterpret_cast<uint4*>(dp + static_cast<size_t>(n) * ld_dp + j) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
}

// out[b, c] = sum_{k<S} in[(b*S+k), c]  with a bf16 input array
__global__ void group_sum16_kernel(const uint16_t* __restrict__ in, int in_ld, float* __restrict__ out, int out_ld,
                                   int B, int S, int cols, int round) {
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    float acc = 0.0f;
    const uint16_t* p = in + static_cast<size_t>(b) * S * in_ld + c;
    for (int k = 0; k < S; ++k) acc += bf16_to_f32(p[static_cast<size_t>(k) * in_ld]);
    out[static_cast<size_t>(b) * out_ld + c] = round ? ptx::round_tf32(acc) : acc;
  }
}

// dpL[n, j] = -w_o[j] * sig(vL[n, j]),  sig from the stored softplus output (float4 over j; H % 4 == 0)
__global__ void cdae_init_delta_kernel(const float* __restrict__ vL, int ld,
                                       const float* __restrict__ wo, float* __restrict__ dp,
                                       int ld_dp, int N, int H) {
  const int H4 = H >> 2;
  const size_t total = static_cast<size_t>(N) * H4;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i / H4), j = static_cast<int>(i - static_cast<size_t>(n) * H4) * 4;
    const float4 u4 = *reinterpret_cast<const float4*>(vL + static_cast<size_t>(n) * ld + j);
    const float4 w4 = __ldg(reinterpret_cast<const float4*>(wo + j));
    const float u[4] = {u4.x, u4.y, u4.z, u4.w}, w[4] = {w4.x, w4.y, w4.z, w4.w};
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float e = __expf(-u[k]);
      const float s = (u[k] < 0.01f) ? u[k] * (1.0f - u[k] * (0.5f - u[k] * (1.0f / 6.0f))) : 1.0f - e;
      o[k] = ptx::round_tf32(-w[k] * s);
    }
    *reinterpret_cast<float4*>(dp + static_cast<size_t>(n) * ld_dp + j) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// resid = sigma*g + eps ; loss += inv_count * sum resid^2 ; r = 2*inv_count*sigma*resid
// (models/graddae/mlp.py:395-398,441 and SURVEY 8a-3).  g, r: [N, ld]; eps: [N, d].
__global__ void cdae_loss_kernel(const float* __restrict__ g, int ld, const float* __restrict__ sigma,
                                 const float* __restrict__ eps, float* __restrict__ r, int N, int d,
                                 float inv_count, float* __restrict__ loss_out,
                                 float* __restrict__ score_out, uint16_t* __restrict__ r16h = nullptr,
                                 uint16_t* __restrict__ r16l = nullptr, int ld16 = 0) {
  float acc = 0.0f;
  const size_t total = static_cast<size_t>(N) * ld;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i / ld), j = static_cast<int>(i - static_cast<size_t>(n) * ld);
    float rv = 0.0f;
    if (j < d) {
      const float s = sigma[n];
      const float gv = g[i];
      const float res = s * gv + eps[static_cast<size_t>(n) * d + j];
      acc += res * res;
      rv = ptx::round_tf32(2.0f * inv_count * s * res);
      if (score_out) score_out[static_cast<size_t>(n) * d + j] = gv;
      if (r16h != nullptr) {
        const uint16_t h16 = f32_to_bf16_rn(rv);
        r16h[static_cast<size_t>(n) * ld16 + j] = h16;
        r16l[static_cast<size_t>(n) * ld16 + j] = f32_to_bf16_rn(rv - bf16_to_f32(h16));
      }
    }
    if (r) r[i] = rv;
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(loss_out, acc * inv_count);
}

// dst[n*d + j] = src[n*ld + j]
__global__ void unpad_kernel(const float* __restrict__ src, int ld, float* __restrict__ dst, int N,
                             int d, float scale) {
  const size_t total = static_cast<size_t>(N) * d;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i / d), j = static_cast<int>(i - static_cast<size_t>(n) * d);
    dst[i] = scale * src[static_cast<size_t>(n) * ld + j];
  }
}

// zbuf[r, j] = hi[r, j] + lo[r, j]; optionally also to a contiguous user buffer
__global__ void pair_sum_kernel(const float* __restrict__ hi, const float* __restrict__ lo, int ld,
                                float* __restrict__ zbuf, int ldz, float* __restrict__ out, int R, int w) {
  const size_t total = static_cast<size_t>(R) * w;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / w), j = static_cast<int>(i - static_cast<size_t>(r) * w);
    const float v = hi[static_cast<size_t>(r) * ld + j] + lo[static_cast<size_t>(r) * ld + j];
    zbuf[static_cast<size_t>(r) * ldz + j] = v;
    if (out != nullptr) out[i] = v;
  }
}

// colsum[c] += scale * sum_r X[r, c]   (one block per 32 columns x a slab of rows)
__global__ void colsum_kernel(const float* __restrict__ X, int ld, int rows, int cols,
                              float* __restrict__ out, float scale) {
  __shared__ float red[8][33];
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lx;
  float acc = 0.0f;
  if (c < cols)
    for (int r = blockIdx.y * 8 + ly; r < rows; r += gridDim.y * 8) acc += X[static_cast<size_t>(r) * ld + c];
  red[ly][lx] = acc;
  __syncthreads();
  if (ly == 0) {
#pragma unroll
    for (int j = 1; j < 8; ++j) acc += red[j][lx];
    if (c < cols) atomicAdd(out + c, scale * acc);
  }
}

// ---------------------------------------------------------------- ELBO terms (model update)
// Bernoulli decoder (models/ivae/mnist.py:240-249, utils/vae.py:21-30, utils/energy.py:69-77):
//   recon_r = sum_px softplus(l) - x*l ; prior_r = 0.5*sum_d (z^2 + log 2pi)
//   sums[0] += (recon_r + beta*prior_r)/R_total ; sums[1] += recon_r/R_total ; sums[2] += prior_r/R_total
//   dlogit[r, px] = gscale * (sigmoid(l) - x)      (gscale = 1/R_total; written tf32-rounded)
// One block per row r; x row index = r / nz.
__global__ void bern_elbo_kernel(const float* __restrict__ logit, int ldl, const float* __restrict__ x,
                                 int D, const float* __restrict__ z, int ldz, int zd, int nz, float beta,
                                 float inv_rows, float* __restrict__ sums, float* __restrict__ dlogit,
                                 int ldd, const float* __restrict__ beta_dev = nullptr) {
  if (beta_dev != nullptr) beta = *beta_dev;
  const int r = blockIdx.x;
  const float* l = logit + static_cast<size_t>(r) * ldl;
  const float* xr = x + static_cast<size_t>(r / nz) * D;
  float rec = 0.0f;
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    const float lv = l[j], xv = xr[j];
    const float e = __expf(-fabsf(lv));
    const float sp = fmaxf(lv, 0.0f) + ((e < 1e-4f) ? (e - 0.5f * e * e) : __logf(1.0f + e));
    rec += sp - xv * lv;
    if (dlogit != nullptr) {
      const float sg = (lv >= 0.0f) ? 1.0f / (1.0f + e) : e / (1.0f + e);
      dlogit[static_cast<size_t>(r) * ldd + j] = ptx::round_tf32(inv_rows * (sg - xv));
    }
  }
  float pri = 0.0f;
  for (int j = threadIdx.x; j < zd; j += blockDim.x) {
    const float zv = z[static_cast<size_t>(r) * ldz + j];
    pri += 0.5f * (zv * zv + kLog2Pi);
  }
  rec = block_sum(rec);
  pri = block_sum(pri);
  if (threadIdx.x == 0) {
    atomicAdd(sums + 0, (rec + beta * pri) * inv_rows);
    atomicAdd(sums + 1, rec * inv_rows);
    atomicAdd(sums + 2, pri * inv_rows);
  }
}
// Gaussian decoder (models/ivae/toy.py:794-803, utils/vae.py:36-52):
//   recon_r = 0.5*sum_D [lv + (x-mu)^2/exp(lv) + log 2pi]
//   heads [R, 2D] = [mu | logvar] ; dheads = gscale * [ -(x-mu)/e^lv | 0.5*(1 - (x-mu)^2/e^lv) ]
__global__ void gauss_elbo_kernel(const float* __restrict__ heads, int ldh, int Dp,
                                  const float* __restrict__ x,
                                  int D, const float* __restrict__ z, int ldz, int zd, int nz, float beta,
                                  float inv_rows, float* __restrict__ sums, float* __restrict__ dheads,
                                  int ldd, const float* __restrict__ beta_dev = nullptr) {
  if (beta_dev != nullptr) beta = *beta_dev;
  const int r = blockIdx.x;
  const float* hrow = heads + static_cast<size_t>(r) * ldh;
  const float* xr = x + static_cast<size_t>(r / nz) * D;
  float rec = 0.0f;
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    const float mu = hrow[j], lv = hrow[Dp + j], df = xr[j] - mu;
    const float iv = __expf(-lv);
    rec += 0.5f * (lv + df * df * iv + kLog2Pi);
    if (dheads != nullptr) {
      dheads[static_cast<size_t>(r) * ldd + j] = ptx::round_tf32(-inv_rows * df * iv);
      dheads[static_cast<size_t>(r) * ldd + Dp + j] = ptx::round_tf32(inv_rows * 0.5f * (1.0f - df * df * iv));
    }
  }
  float pri = 0.0f;
  for (int j = threadIdx.x; j < zd; j += blockDim.x) {
    const float zv = z[static_cast<size_t>(r) * ldz + j];
    pri += 0.5f * (zv * zv + kLog2Pi);
  }
  rec = block_sum(rec);
  pri = block_sum(pri);
  if (threadIdx.x == 0) {
    atomicAdd(sums + 0, (rec + beta * pri) * inv_rows);
    atomicAdd(sums + 1, rec * inv_rows);
    atomicAdd(sums + 2, pri * inv_rows);
  }
}
// dz_total[r, j] = loss_scale * (dz_dec[r, j] + beta*inv_rows*z[r, j]) + gz_scale*gz[r, j]  (tf32)
__global__ void dz_total_kernel(const float* __restrict__ dz_dec, int ld_dec, const float* __restrict__ z,
                                int ldz, const float* __restrict__ gz, float gz_scale, float loss_scale,
                                float beta_inv_rows,
                                float* __restrict__ out, int ld_out, int R, int zd,
                                const float* __restrict__ beta_dev = nullptr, float inv_rows = 0.0f) {
  // beta_dev: beta lives in a device scalar (annealing under CUDA-graph replay); gz_scale then excludes beta
  if (beta_dev != nullptr) {
    const float b = *beta_dev;
    beta_inv_rows = b * inv_rows;
    gz_scale *= b;
  }
  const size_t total = static_cast<size_t>(R) * zd;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / zd), j = static_cast<int>(i - static_cast<size_t>(r) * zd);
    float v = 0.0f;
    if (loss_scale != 0.0f)
      v = loss_scale * (dz_dec[static_cast<size_t>(r) * ld_dec + j] + beta_inv_rows * z[static_cast<size_t>(r) * ldz + j]);
    if (gz != nullptr) v += gz_scale * gz[i];
    out[static_cast<size_t>(r) * ld_out + j] = ptx::round_tf32(v);
  }
}

// ---------------------------------------------------------------- Gaussian reparametrisation stages of the hierarchical
// (aux*) encoders: models/ivae/auxmnist.py:33-40,88-101, models/vae/auxmnist.py:24-29
// out[r, j] = mu[g, j] + exp(0.5 * lv[g, j]) * eps[r, j],  g = r / group  (eps == nullptr: std = 0 -> out = mu).
// heads [rows/group, ldh] = [mu | lv at +lvoff].  Written as a tf32 pair (hi | lo at +kp), optionally also plain.
__global__ void aux_reparam_kernel(const float* __restrict__ heads, int ldh, int lvoff, const float* __restrict__ eps,
                                   int ld_eps, int R, int w, int group, float* __restrict__ pair, int ldp, int kp,
                                   float* __restrict__ plain, int ld_plain, float* __restrict__ user_out) {
  const size_t total = static_cast<size_t>(R) * w;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / w), j = static_cast<int>(i - static_cast<size_t>(r) * w);
    const float* hg = heads + static_cast<size_t>(r / group) * ldh;
    float v = hg[j];
    if (eps != nullptr) v += __expf(0.5f * hg[lvoff + j]) * eps[static_cast<size_t>(r) * ld_eps + j];
    const float hi = ptx::round_tf32(v);
    pair[static_cast<size_t>(r) * ldp + j] = hi;
    pair[static_cast<size_t>(r) * ldp + kp + j] = ptx::round_tf32(v - hi);
    if (plain != nullptr) plain[static_cast<size_t>(r) * ld_plain + j] = v;
    if (user_out != nullptr) user_out[i] = v;
  }
}
// dheads[r, j] = dz[r, j] ; dheads[r, lvoff + j] = dz[r, j] * 0.5 * exp(0.5 lv[r, j]) * eps[r, j]   (tf32-rounded)
__global__ void aux_reparam_bwd_kernel(const float* __restrict__ dz, int ldz, const float* __restrict__ heads, int ldh,
                                       int lvoff, const float* __restrict__ eps, int ld_eps, int R, int w,
                                       float* __restrict__ dheads, int ldd) {
  const size_t total = static_cast<size_t>(R) * w;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / w), j = static_cast<int>(i - static_cast<size_t>(r) * w);
    const float g = dz[static_cast<size_t>(r) * ldz + j];
    const float e = eps != nullptr ? eps[static_cast<size_t>(r) * ld_eps + j] : 0.0f;
    dheads[static_cast<size_t>(r) * ldd + j] = ptx::round_tf32(g);
    dheads[static_cast<size_t>(r) * ldd + lvoff + j] =
        ptx::round_tf32(g * 0.5f * __expf(0.5f * heads[static_cast<size_t>(r) * ldh + lvoff + j]) * e);
  }
}
// Same through a per-data-row head: dheads[b, j] = sum_k dz[b*S+k, j] ; dheads[b, lvoff+j] = sum_k dz * 0.5 exp(0.5 lv[b,j]) eps
__global__ void aux_reparam_bwd_group_kernel(const float* __restrict__ dz, int ldz, const float* __restrict__ heads,
                                             int ldh, int lvoff, const float* __restrict__ eps, int ld_eps, int B, int S,
                                             int w, float* __restrict__ dheads, int ldd) {
  const int b = blockIdx.x;
  for (int j = threadIdx.x; j < w; j += blockDim.x) {
    float a0 = 0.0f, a1 = 0.0f;
    for (int k = 0; k < S; ++k) {
      const size_t r = static_cast<size_t>(b) * S + k;
      const float g = dz[r * ldz + j];
      a0 += g;
      if (eps != nullptr) a1 += g * eps[r * ld_eps + j];
    }
    dheads[static_cast<size_t>(b) * ldd + j] = ptx::round_tf32(a0);
    dheads[static_cast<size_t>(b) * ldd + lvoff + j] =
        ptx::round_tf32(a1 * 0.5f * __expf(0.5f * heads[static_cast<size_t>(b) * ldh + lvoff + j]));
  }
}

// ---------------------------------------------------------------- sigma schedule (ivae_ardae.py:753-767)
// One block per data row b.  lsm = S*(z - zbar); s_b = delta * mean_d std_k(lsm) (unbiased over nz);
// outputs: x_out[(b*nz + k)*nstd + t, :] = lsm[b,k,:] ; sigma_out[same] = s_b * xi[same] ;
// std_out[b] = s_b.  xi == nullptr -> drawn with Philox.
__global__ void sigma_schedule_kernel(const float* __restrict__ z, const float* __restrict__ zbar,
                                      int nz, int d, int nstd, float S, float delta,
                                      const float* __restrict__ xi, uint64_t seed,
                                      float* __restrict__ x_out, float* __restrict__ sigma_out,
                                      float* __restrict__ std_out, const replay_ctr_t* __restrict__ ctr) {
  seed = replay_seed(seed, ctr);
  extern __shared__ float sm[];  // [blockDim.x]
  const int b = blockIdx.x;
  const float* zb = z + static_cast<size_t>(b) * nz * d;
  const float* zm = zbar + static_cast<size_t>(b) * d;
  // per-dimension unbiased std over the nz samples (two passes, like torch.std).  Element i = k*d + j of the row's
  // [nz, d] block is read by thread i % blockDim: coalesced, every thread busy.  When blockDim % d == 0 a thread
  // always sees the same dimension j, so per-thread partial sums reduce per dimension through shared memory.
  float acc = 0.0f;
  const int total = nz * d;
  if (blockDim.x % d == 0 && d <= static_cast<int>(blockDim.x)) {
    float* red = sm + 4;                 // [blockDim.x]
    float* stat = sm + 4 + blockDim.x;   // [d]
    const int j = threadIdx.x % d;
    const float m0 = zm[j];
    float s1 = 0.0f;
    for (int i = threadIdx.x; i < total; i += blockDim.x) s1 += S * (zb[i] - m0);
    red[threadIdx.x] = s1;
    __syncthreads();
    if (threadIdx.x < d) {
      float t = 0.0f;
      for (int q = threadIdx.x; q < static_cast<int>(blockDim.x); q += d) t += red[q];
      stat[threadIdx.x] = t / nz;
    }
    __syncthreads();
    const float mean = stat[j];
    float s2 = 0.0f;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const float v = S * (zb[i] - m0) - mean;
      s2 += v * v;
    }
    __syncthreads();
    red[threadIdx.x] = s2;
    __syncthreads();
    if (threadIdx.x < d) {
      float t = 0.0f;
      for (int q = threadIdx.x; q < static_cast<int>(blockDim.x); q += d) t += red[q];
      acc = sqrtf(t / (nz - 1));
    }
  } else {
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    const float m0 = zm[j];
    float mean = 0.0f;
    for (int k = 0; k < nz; ++k) mean += S * (zb[static_cast<size_t>(k) * d + j] - m0);
    mean /= nz;
    float var = 0.0f;
    for (int k = 0; k < nz; ++k) {
      const float v = S * (zb[static_cast<size_t>(k) * d + j] - m0) - mean;
      var += v * v;
    }
    acc += sqrtf(var / (nz - 1));
  }
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) sm[0] = delta * acc / d;
  __syncthreads();
  const float sb = sm[0];
  if (threadIdx.x == 0 && std_out) std_out[b] = sb;
  const int rows = nz * nstd;
  for (int i = threadIdx.x; i < rows * d; i += blockDim.x) {
    const int row = i / d, j = i - row * d;
    const int k = row / nstd;
    x_out[(static_cast<size_t>(b) * rows + row) * d + j] = S * (zb[static_cast<size_t>(k) * d + j] - zm[j]);
  }
  for (int row = threadIdx.x; row < rows; row += blockDim.x) {
    const size_t e = static_cast<size_t>(b) * rows + row;
    float xv;
    if (xi != nullptr) {
      xv = xi[e];
    } else {
      const uint4 r = Philox::gen(seed, e >> 2, 11u);
      const float2 p = box_muller(r.x, r.y), q = box_muller(r.z, r.w);
      const float v4[4] = {p.x, p.y, q.x, q.y};
      xv = v4[e & 3];
    }
    sigma_out[e] = sb * xv;
  }
}
// x_out[r, :] = S * (z[r, :] - zbar[r / nz, :])
__global__ void scaled_diff_kernel(const float* __restrict__ z, const float* __restrict__ zbar, int R,
                                   int nz, int d, float S, float* __restrict__ out) {
  const size_t total = static_cast<size_t>(R) * d;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / d), j = static_cast<int>(i - static_cast<size_t>(r) * d);
    out[i] = S * (z[i] - zbar[static_cast<size_t>(r / nz) * d + j]);
  }
}

// ---------------------------------------------------------------- IWS evaluator (ivae/mnist.py:378-437)
// One block per image i.  z[i] = [S, d] encoder samples.  Computes the moment-matched Gaussian proposal
// (mean, unbiased covariance: utils/stat.py:127-158), its Cholesky factor L (what MultivariateNormal
// builds), newz_k = mu + L eta_k (rsample), and lw0_k = log N(newz_k; 0, I) - log N(newz_k; mu, Sigma).
// newz is written as the tf32 pair the decoder's first 3xTF32 GEMM consumes.  d <= 64.
// eta == nullptr -> Philox draw.  Dynamic smem: (d*d + 2*d + 64*d) floats.
__global__ void iws_moments_kernel(const float* __restrict__ z, int ldz, int S, int d,
                                   const float* __restrict__ eta, uint64_t seed,
                                   float* __restrict__ newz_pair, int ldp, int kp,
                                   float* __restrict__ lw0, int* __restrict__ status, float diag_eps = 0.0f) {
  extern __shared__ float sm[];
  float* cov = sm;               // [d*d]
  float* mu = cov + d * d;       // [d]
  float* red = mu + d;           // [d] scratch
  float* tile = red + d;         // [64*d] sample tile
  const int img = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const float* zi = z + static_cast<size_t>(img) * S * ldz;
  // ---- mean
  for (int j = tid; j < d; j += nt) {
    float a = 0.0f;
    for (int k = 0; k < S; ++k) a += zi[static_cast<size_t>(k) * ldz + j];
    mu[j] = a / S;
  }
  for (int e = tid; e < d * d; e += nt) cov[e] = 0.0f;
  __syncthreads();
  // ---- covariance (lower + upper, thread e owns entry (e / d, e % d)), samples staged through smem
  const int per = (d * d + nt - 1) / nt;
  for (int k0 = 0; k0 < S; k0 += 64) {
    const int kn = min(64, S - k0);
    for (int t = tid; t < kn * d; t += nt) {
      const int kk = t / d, j = t - kk * d;
      tile[t] = zi[static_cast<size_t>(k0 + kk) * ldz + j] - mu[j];
    }
    __syncthreads();
    for (int q = 0; q < per; ++q) {
      const int e = tid + q * nt;
      if (e < d * d) {
        const int a = e / d, b = e - a * d;
        float acc = 0.0f;
        for (int kk = 0; kk < kn; ++kk) acc += tile[kk * d + a] * tile[kk * d + b];
        cov[e] += acc;
      }
    }
    __syncthreads();
  }
  // (diag_eps: the hierarchical models add 1e-5 I to the covariance, models/ivae/auxmnist.py:350)
  for (int e = tid; e < d * d; e += nt) cov[e] = cov[e] * (1.0f / (S - 1)) + ((e / d == e % d) ? diag_eps : 0.0f);
  __syncthreads();
  // ---- Cholesky (in place, lower), warp 0; column j at a time
  if (tid < 32) {
    for (int j = 0; j < d; ++j) {
      float diag = cov[j * d + j];
      for (int k = 0; k < j; ++k) diag -= cov[j * d + k] * cov[j * d + k];
      if (!(diag > 0.0f)) {
        if (tid == 0 && status) atomicExch(status, 1 + img);
        diag = 1e-30f;
      }
      const float ljj = sqrtf(diag);
      __syncwarp();
      for (int i = j + 1 + tid; i < d; i += 32) {
        float v = cov[i * d + j];
        for (int k = 0; k < j; ++k) v -= cov[i * d + k] * cov[j * d + k];
        cov[i * d + j] = v / ljj;
      }
      if (tid == 0) cov[j * d + j] = ljj;
      __syncwarp();
    }
  }
  __syncthreads();
  float logdet = 0.0f;
  for (int j = 0; j < d; ++j) logdet += __logf(cov[j * d + j]);
  // ---- samples: newz = mu + L eta ; lw0 = logprior - logq
  for (int k = tid; k < S; k += nt) {
    const size_t row = static_cast<size_t>(img) * S + k;
    float ev[64];
    float e2 = 0.0f;
    for (int j = 0; j < d; ++j) {
      float v;
      if (eta != nullptr) {
        v = eta[row * d + j];
      } else {
        const size_t e = row * d + j;
        const uint4 r = Philox::gen(seed, e >> 2, 13u);
        const float2 p0 = box_muller(r.x, r.y), p1 = box_muller(r.z, r.w);
        const float q4[4] = {p0.x, p0.y, p1.x, p1.y};
        v = q4[e & 3];
      }
      ev[j] = v;
      e2 += v * v;
    }
    float z2 = 0.0f;
    for (int i = 0; i < d; ++i) {
      float v = mu[i];
      for (int j = 0; j <= i; ++j) v += cov[i * d + j] * ev[j];
      z2 += v * v;
      const float hi = ptx::round_tf32(v);
      newz_pair[row * ldp + i] = hi;
      newz_pair[row * ldp + kp + i] = ptx::round_tf32(v - hi);
    }
    // logprior = -0.5*(z2 + d*log2pi) ; logq = -0.5*e2 - logdet - 0.5*d*log2pi
    lw0[row] = -0.5f * z2 + 0.5f * e2 + logdet;
  }
}
// w_r = lw0_r + loglik_r ; loglik from the decoder heads.  One block per row.
//   kind 1 (Bernoulli): -sum_px softplus(l) - x*l        kind 0 (Gaussian): -0.5*sum [(x-mu)^2/e^lv + lv + log2pi]
__global__ void iws_loglik_kernel(const float* __restrict__ heads, int ldh, int Dp, const float* __restrict__ x,
                                  int D, int S, int kind, const float* __restrict__ lw0, float* __restrict__ w) {
  const int r = blockIdx.x;
  const float* h = heads + static_cast<size_t>(r) * ldh;
  const float* xr = x + static_cast<size_t>(r / S) * D;
  float acc = 0.0f;
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    if (kind == 1) {
      const float lv = h[j];
      const float e = __expf(-fabsf(lv));
      const float sp = fmaxf(lv, 0.0f) + ((e < 1e-4f) ? (e - 0.5f * e * e) : __logf(1.0f + e));
      acc -= sp - xr[j] * lv;
    } else {
      const float m = h[j], lv = h[Dp + j], df = xr[j] - m;
      acc -= 0.5f * (df * df * __expf(-lv) + lv + kLog2Pi);
    }
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) w[r] = acc + lw0[r];
}
// out[i] = log(mean_k exp(w[i,k] - max) + 1e-10) + max   (ivae/mnist.py:431-434: NOT a plain logsumexp)
__global__ void iws_logmeanexp_kernel(const float* __restrict__ w, int S, float* __restrict__ out,
                                      float* __restrict__ total) {
  __shared__ float sh_max;
  const int i = blockIdx.x;
  const float* wi = w + static_cast<size_t>(i) * S;
  float m = -INFINITY;
  for (int k = threadIdx.x; k < S; k += blockDim.x) m = fmaxf(m, wi[k]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float redm[32];
  if ((threadIdx.x & 31) == 0) redm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x + 31) / 32 ? redm[threadIdx.x] : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) sh_max = m;
  }
  __syncthreads();
  m = sh_max;
  float s = 0.0f;
  for (int k = threadIdx.x; k < S; k += blockDim.x) s += __expf(wi[k] - m);
  s = block_sum(s);
  if (threadIdx.x == 0) {
    const float v = logf(s / S + 1e-10f) + m;
    out[i] = v;
    if (total) atomicAdd(total, v);
  }
}

// ---------------------------------------------------------------- optimizers (one flat pass)
// Reference Adam (utils/optim.py:59-106, PyTorch-1.2 epsilon placement):
//   m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g*g ; p -= (lr/bc1) * m / ((sqrt(v)+eps)/sqrt(bc2))
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, size_t n, float lr, int step0,
                            float b1, float b2, float eps, float gscale, const replay_ctr_t* __restrict__ ctr) {
  // bias corrections of step t = step0 + replay counter, in double like the host formula (utils/optim.py:96-98)
  __shared__ float bc[2];
  if (threadIdx.x == 0) {
    const double t = static_cast<double>(step0) + (ctr != nullptr ? static_cast<double>(*ctr) : 0.0);
    const double bc1 = 1.0 - pow(static_cast<double>(b1), t), bc2 = 1.0 - pow(static_cast<double>(b2), t);
    bc[0] = static_cast<float>(static_cast<double>(lr) / bc1);
    bc[1] = static_cast<float>(1.0 / sqrt(bc2));
  }
  __syncthreads();
  const float lr_bc1 = bc[0], inv_sqrt_bc2 = bc[1];
  const size_t n4 = n / 4;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float4 pp = p4[i], gg = g4[i], mm = m4[i], vv = v4[i];
    float* pa = &pp.x; float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = gscale * ga[j];
      ma[j] = b1 * ma[j] + (1.0f - b1) * gr;
      va[j] = b2 * va[j] + (1.0f - b2) * gr * gr;
      const float denom = (sqrtf(va[j]) + eps) * inv_sqrt_bc2;
      pa[j] -= lr_bc1 * ma[j] / denom;
    }
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
  }
}
// torch.optim.RMSprop (centered=False): sq = a*sq + (1-a)*g*g ; avg = sqrt(sq)+eps ;
//   momentum > 0: buf = mu*buf + g/avg ; p -= lr*buf      else p -= lr*g/avg
__global__ void rmsprop_kernel(float* __restrict__ p, const float* __restrict__ g,
                               float* __restrict__ sq, float* __restrict__ buf, size_t n, float lr,
                               float alpha, float eps, float mu, float gscale) {
  const size_t n4 = n / 4;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* s4 = reinterpret_cast<float4*>(sq);
  float4* b4 = reinterpret_cast<float4*>(buf);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float4 pp = p4[i], gg = g4[i], ss = s4[i], bb = b4[i];
    float* pa = &pp.x; float* ga = &gg.x; float* sa = &ss.x; float* ba = &bb.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = gscale * ga[j];
      sa[j] = alpha * sa[j] + (1.0f - alpha) * gr * gr;
      const float avg = sqrtf(sa[j]) + eps;
      if (mu > 0.0f) {
        ba[j] = mu * ba[j] + gr / avg;
        pa[j] -= lr * ba[j];
      } else {
        pa[j] -= lr * gr / avg;
      }
    }
    p4[i] = pp; s4[i] = ss; b4[i] = bb;
  }
}

inline int grid_for(size_t total, int block = 256, int max_blocks = 148 * 8) {
  size_t b = (total + block - 1) / block;
  if (b < 1) b = 1;
  if (b > static_cast<size_t>(max_blocks)) b = max_blocks;
  return static_cast<int>(b);
}

}  // namespace ardae
