// 5x5 stride-2 pad-2 convolution family of the conv implicit-posterior VAE
// (reference: models/ivae/conv.py:70-72,84-96 encoder; models/vae/conv.py:112-131 decoder).
// These layers run on the B data rows only (<2 % of the step's FLOPs) - direct fp32 CUDA-core kernels in
// PyTorch's NCHW layout, exact fp32 arithmetic (no tf32 splitting needed); the N-row work of the model
// (fc4/fc5, CDAE) stays on the tcgen05 GEMMs.
//
// One geometry: out[oh] <-> in[ih] with oh*2 - 2 + kh = ih (conv) ; a ConvTranspose2d is the adjoint:
//   deconv forward  = conv backward-data,  deconv backward-data = conv forward (no bias),
//   deconv backward-weight = conv backward-weight with the roles of the two activations swapped
// (weight memory layout [C_big_out_of_conv, C_in_of_conv, 5, 5] is shared by both views).
#pragma once
#include "kernels.cuh"

namespace ardae {

enum ConvAct : int { CONV_ACT_NONE = 0, CONV_ACT_RELU = 1, CONV_ACT_SOFTPLUS = 2 };

__device__ __forceinline__ float conv_act(float x, int act) {
  if (act == CONV_ACT_RELU) return fmaxf(x, 0.0f);
  if (act == CONV_ACT_SOFTPLUS) return (x > 20.0f) ? x : log1pf(expf(x));
  return x;
}
// d act / d pre from the stored OUTPUT u
__device__ __forceinline__ float conv_dact_from_out(float u, int act) {
  if (act == CONV_ACT_RELU) return u > 0.0f ? 1.0f : 0.0f;
  if (act == CONV_ACT_SOFTPLUS) return -expm1f(-u);
  return 1.0f;
}

// out[b,co,oh,ow] = act( bias[co] + sum_{ci,kh,kw} (a*in[b,ci,2oh-2+kh,2ow-2+kw]+s) * W[co,ci,kh,kw] )
// `in` logical size Hi x Hi stored with row pitch in_pitch (>= Hi) and plane size in_pitch*in_rows;
// positions outside [0,Hi) are zero padding (the affine a*x+s applies to real pixels only).
__global__ void conv5s2_fwd_kernel(const float* __restrict__ in, int Ci, int Hi, int in_pitch, int in_rows,
                                   const float* __restrict__ W, const float* __restrict__ bias,
                                   float* __restrict__ out, int Co, int Ho, int out_pitch, int out_rows,
                                   int B, float a, float s, int act, const float* __restrict__ u_post) {
  const size_t total = static_cast<size_t>(B) * Co * Ho * Ho;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ow = static_cast<int>(i % Ho);
    const int oh = static_cast<int>((i / Ho) % Ho);
    const int co = static_cast<int>((i / (static_cast<size_t>(Ho) * Ho)) % Co);
    const int b = static_cast<int>(i / (static_cast<size_t>(Ho) * Ho * Co));
    float acc = bias ? bias[co] : 0.0f;
    const float* wb = W + static_cast<size_t>(co) * Ci * 25;
    const float* ib = in + static_cast<size_t>(b) * Ci * in_pitch * in_rows;
    for (int ci = 0; ci < Ci; ++ci) {
      const float* ip = ib + static_cast<size_t>(ci) * in_pitch * in_rows;
      const float* wp = wb + ci * 25;
#pragma unroll
      for (int kh = 0; kh < 5; ++kh) {
        const int ih = 2 * oh - 2 + kh;
        if (ih < 0 || ih >= Hi) continue;
#pragma unroll
        for (int kw = 0; kw < 5; ++kw) {
          const int iw = 2 * ow - 2 + kw;
          if (iw < 0 || iw >= Hi) continue;
          acc = fmaf(fmaf(a, ip[ih * in_pitch + iw], s), wp[kh * 5 + kw], acc);
        }
      }
    }
    const size_t o = (static_cast<size_t>(b) * Co + co) * out_pitch * out_rows + oh * out_pitch + ow;
    // u_post != null: backward-data use - multiply by d act / d pre taken from the stored output u_post
    out[o] = u_post ? acc * conv_dact_from_out(u_post[o], act) : conv_act(acc, act);
  }
}

// Adjoint: din[b,ci,ih,iw] = post( bias[ci] + sum_{co,kh,kw : 2oh-2+kh = ih} dout[b,co,oh,ow] * W[co,ci,kh,kw] )
//   as conv backward-data: post = multiply by dact(u_in[b,ci,ih,iw]) (u_in = stored output of the layer below)
//   as deconv forward    : post = act(. + bias)
// Only the Hi x Hi region is produced (din pitch/rows may be larger: zero-padded buffers).
__global__ void conv5s2_bwd_data_kernel(const float* __restrict__ dout, int Co, int Ho, int out_pitch, int out_rows,
                                        const float* __restrict__ W, const float* __restrict__ bias,
                                        float* __restrict__ din, int Ci, int Hi, int in_pitch, int in_rows,
                                        int B, int act, const float* __restrict__ u_in, int mode_dact) {
  const size_t total = static_cast<size_t>(B) * Ci * Hi * Hi;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int iw = static_cast<int>(i % Hi);
    const int ih = static_cast<int>((i / Hi) % Hi);
    const int ci = static_cast<int>((i / (static_cast<size_t>(Hi) * Hi)) % Ci);
    const int b = static_cast<int>(i / (static_cast<size_t>(Hi) * Hi * Ci));
    float acc = bias ? bias[ci] : 0.0f;
    const float* ob = dout + static_cast<size_t>(b) * Co * out_pitch * out_rows;
    for (int co = 0; co < Co; ++co) {
      const float* op = ob + static_cast<size_t>(co) * out_pitch * out_rows;
      const float* wp = W + (static_cast<size_t>(co) * Ci + ci) * 25;
#pragma unroll
      for (int kh = 0; kh < 5; ++kh) {
        const int t = ih + 2 - kh;
        if (t < 0 || (t & 1)) continue;
        const int oh = t >> 1;
        if (oh >= Ho) continue;
#pragma unroll
        for (int kw = 0; kw < 5; ++kw) {
          const int v = iw + 2 - kw;
          if (v < 0 || (v & 1)) continue;
          const int ow = v >> 1;
          if (ow >= Ho) continue;
          acc = fmaf(op[oh * out_pitch + ow], wp[kh * 5 + kw], acc);
        }
      }
    }
    const size_t o = (static_cast<size_t>(b) * Ci + ci) * in_pitch * in_rows + ih * in_pitch + iw;
    if (mode_dact) acc *= conv_dact_from_out(u_in[o], act);
    else acc = conv_act(acc, act);
    din[o] = acc;
  }
}

// dW[co,ci,kh,kw] += sum_{b,oh,ow} dout[b,co,oh,ow] * (a*in[b,ci,2oh-2+kh,2ow-2+kw]+s)
// One block per (co, ci); 25 taps accumulated per thread, block-reduced.  dbias[co] += sum dout (ci == 0).
__global__ void conv5s2_bwd_weight_kernel(const float* __restrict__ in, int Ci, int Hi, int in_pitch, int in_rows,
                                          const float* __restrict__ dout, int Co, int Ho, int out_pitch,
                                          int out_rows, float* __restrict__ dW, float* __restrict__ dbias, int B,
                                          float a, float s) {
  const int co = blockIdx.x / Ci, ci = blockIdx.x % Ci;
  float acc[25];
#pragma unroll
  for (int t = 0; t < 25; ++t) acc[t] = 0.0f;
  float bsum = 0.0f;
  const int per = Ho * Ho;
  for (int i = threadIdx.x; i < B * per; i += blockDim.x) {
    const int b = i / per, r = i - b * per, oh = r / Ho, ow = r - oh * Ho;
    const float g = dout[(static_cast<size_t>(b) * Co + co) * out_pitch * out_rows + oh * out_pitch + ow];
    bsum += g;
    const float* ip = in + (static_cast<size_t>(b) * Ci + ci) * in_pitch * in_rows;
#pragma unroll
    for (int kh = 0; kh < 5; ++kh) {
      const int ih = 2 * oh - 2 + kh;
      if (ih < 0 || ih >= Hi) continue;
#pragma unroll
      for (int kw = 0; kw < 5; ++kw) {
        const int iw = 2 * ow - 2 + kw;
        if (iw < 0 || iw >= Hi) continue;
        acc[kh * 5 + kw] = fmaf(g, fmaf(a, ip[ih * in_pitch + iw], s), acc[kh * 5 + kw]);
      }
    }
  }
  __shared__ float red[8][26];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t < 25; ++t) {
    const float v = warp_sum(acc[t]);
    if (lane == 0) red[w][t] = v;
  }
  bsum = warp_sum(bsum);
  if (lane == 0) red[w][25] = bsum;
  __syncthreads();
  if (threadIdx.x < 26) {
    float v = 0.0f;
    for (int k = 0; k < static_cast<int>(blockDim.x >> 5); ++k) v += red[k][threadIdx.x];
    if (threadIdx.x < 25) dW[(static_cast<size_t>(co) * Ci + ci) * 25 + threadIdx.x] += v;
    else if (ci == 0 && dbias != nullptr) dbias[co] += v;
  }
}

// out[c] += sum_{b,h,w} x[b,c,h,w] over the H x H region of planes with the given pitch / rows
__global__ void chan_sum_kernel(const float* __restrict__ x, int C, int H, int pitch, int rows, int B,
                                float* __restrict__ out) {
  const int c = blockIdx.x;
  float acc = 0.0f;
  const int per = H * H;
  for (int i = threadIdx.x; i < B * per; i += blockDim.x) {
    const int b = i / per, r = i - b * per, h = r / H, w = r - h * H;
    acc += x[(static_cast<size_t>(b) * C + c) * pitch * rows + h * pitch + w];
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) out[c] += acc;
}

// dst[r, c] = src[r, c] * dact(u[r, c])   (flat; used between the decoder's fc output and the deconv stack)
__global__ void mul_dact_kernel(const float* __restrict__ src, const float* __restrict__ u, float* __restrict__ dst,
                                size_t n, int act, int round) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float v = src[i] * conv_dact_from_out(u[i], act);
    dst[i] = round ? ptx::round_tf32(v) : v;
  }
}

}  // namespace ardae
