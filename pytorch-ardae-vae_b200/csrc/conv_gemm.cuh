// The 5x5 stride-2 pad-2 convolution family of the conv implicit-posterior VAE as GEMMs on the tcgen05 kernels
// (reference: models/ivae/conv.py:70-72,84-96 encoder; models/vae/conv.py:112-131 decoder).
//
// Activations are pixel-major matrices [rows * H*H, C] (one GEMM row per pixel, channels contiguous), so that
//   conv   forward:  out = act(im2col(in) . W^T + b)           W as [Co, Ci*25]  (PyTorch's own memory layout)
//   conv   backward: dW = dout^T . im2col(in) ; din = col2im(dout . W) * act'
//   deconv forward:  out = act(col2im(in . W) + b)              W as [Ci_deconv, Co_deconv*25] (ConvTranspose2d layout)
//   deconv backward: dW = in^T . im2col(dout) ; din = (im2col(dout) . W^T) * act'
// with one geometry for both directions: column k = c*25 + kh*5 + kw of pixel (oh, ow) of the SMALL grid pairs with
// pixel (2oh-2+kh, 2ow-2+kw) of the LARGE grid.  The forward products are 3xTF32 (fp32-accurate) through the pair
// layout [hi | lo]; the backward ones tf32, as everywhere else in the model plans.  The kernels below only move data:
// im2col / col2im gathers, the NCHW <-> pixel-major permutes at the two fully-connected boundaries
// (flatten order of ivae/conv.py:96 and vae/conv.py:123 is c*P + p) and the transposed weight derivation.
#pragma once
#include "conv_kernels.cuh"

namespace ardae {

// value(b, c, y, x) = p[b*sb + c*sc + y*sy + x*sx] for 0 <= y, x < H; zero outside
struct ImgView {
  float* p;
  long long sb;
  int sc, sy, sx;
  int C, H;
};
inline ImgView img_pix(float* p, int ld, int C, int H, int pitch) {  // pixel-major [rows*pitch*pitch, C], logical H <= pitch
  ImgView v;
  v.p = p; v.sb = static_cast<long long>(pitch) * pitch * ld; v.sc = 1; v.sy = pitch * ld; v.sx = ld; v.C = C; v.H = H;
  return v;
}
inline ImgView img_nchw(float* p, int C, int H) {
  ImgView v;
  v.p = p; v.sb = static_cast<long long>(C) * H * H; v.sc = H * H; v.sy = H; v.sx = 1; v.C = C; v.H = H;
  return v;
}

// col[(b*Ho + oh)*Ho + ow, c*25 + kh*5 + kw] = a * in(b, c, 2oh-2+kh, 2ow-2+kw) + s   (0 outside the image)
// kp > 0: written as a tf32 pair (hi at k, lo at kp + k); kp == 0: one tf32-rounded value.
// A block walks rows; KT threads cover the K columns of one row (256 / KT rows per pass), so the row decode is
// per pass and the column decode (constant divisors) is hoisted out of the row loop.
__global__ void im2col5s2_kernel(ImgView in, int Ho, float a, float s, float* __restrict__ col, int ldc, int kp, int B,
                                 int KT) {
  const int K = in.C * 25;
  const int rpb = blockDim.x / KT;                       // rows per pass
  const int kt = threadIdx.x % KT, rl = threadIdx.x / KT;
  const size_t rows = static_cast<size_t>(B) * Ho * Ho;
  for (size_t row = static_cast<size_t>(blockIdx.x) * rpb + rl; row < rows; row += static_cast<size_t>(gridDim.x) * rpb) {
    const int ow = static_cast<int>(row % Ho), oh = static_cast<int>((row / Ho) % Ho);
    const size_t b = row / (static_cast<size_t>(Ho) * Ho);
    const float* ib = in.p + b * in.sb;
    float* cr = col + row * ldc;
    for (int k = kt; k < K; k += KT) {
      const int c = k / 25, t = k - c * 25, kh = t / 5, kw = t - kh * 5;
      const int y = 2 * oh - 2 + kh, x = 2 * ow - 2 + kw;
      float v = 0.0f;
      if (y >= 0 && y < in.H && x >= 0 && x < in.H) v = fmaf(a, ib[c * in.sc + y * in.sy + x * in.sx], s);
      const float hi = ptx::round_tf32(v);
      cr[k] = hi;
      if (kp > 0) cr[kp + k] = ptx::round_tf32(v - hi);
    }
  }
}
inline void launch_im2col5s2(const ImgView& in, int Ho, float a, float s, float* col, int ldc, int kp, int B, cudaStream_t st) {
  const int K = in.C * 25;
  int KT = 32;
  while (KT < K && KT < 256) KT *= 2;
  const size_t rows = static_cast<size_t>(B) * Ho * Ho;
  const size_t passes = (rows + (256 / KT) - 1) / (256 / KT);
  const int grid = static_cast<int>(std::min<size_t>(passes, 148 * 16));
  im2col5s2_kernel<<<grid, 256, 0, st>>>(in, Ho, a, s, col, ldc, kp, B, KT);
}

enum Col2ImPost : int { C2I_ACT_PAIR = 0, C2I_PLAIN = 1, C2I_MUL_DACT = 2 };

// out(b, c, y, x) = post( sum_{kh,kw : (y+2-kh), (x+2-kw) even, in range} cols[(b, (y+2-kh)/2, (x+2-kw)/2), c*25 + kh*5 + kw] )
//   C2I_ACT_PAIR : act(v + bias[c]) written as a tf32 pair (lo at + kp * out.sc)           (deconv forward)
//   C2I_PLAIN    : v + bias[c], fp32                                                       (logits)
//   C2I_MUL_DACT : v * act'(u(b, c, y, x)), tf32-rounded; u = stored activated output      (conv backward-data)
__global__ void col2im5s2_kernel(const float* __restrict__ cols, int ldc, int Ho, ImgView out, const float* __restrict__ bias,
                                 int act, int post, int kp, const float* __restrict__ u, int B) {
  const int C = out.C, H = out.H;
  const size_t total = static_cast<size_t>(B) * H * H * C;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    // channel fastest for pixel-major outputs, x fastest for NCHW ones (coalesced writes either way)
    int c, y, x;
    size_t b;
    if (out.sc == 1) {
      c = static_cast<int>(i % C);
      x = static_cast<int>((i / C) % H);
      y = static_cast<int>((i / (static_cast<size_t>(C) * H)) % H);
      b = i / (static_cast<size_t>(C) * H * H);
    } else {
      x = static_cast<int>(i % H);
      y = static_cast<int>((i / H) % H);
      c = static_cast<int>((i / (static_cast<size_t>(H) * H)) % C);
      b = i / (static_cast<size_t>(C) * H * H);
    }
    float v = 0.0f;
    const float* cb = cols + b * Ho * Ho * ldc + c * 25;
#pragma unroll
    for (int kh = 0; kh < 5; ++kh) {
      const int ty = y + 2 - kh;
      if (ty < 0 || (ty & 1) || (ty >> 1) >= Ho) continue;
#pragma unroll
      for (int kw = 0; kw < 5; ++kw) {
        const int tx = x + 2 - kw;
        if (tx < 0 || (tx & 1) || (tx >> 1) >= Ho) continue;
        v += cb[static_cast<size_t>((ty >> 1) * Ho + (tx >> 1)) * ldc + kh * 5 + kw];
      }
    }
    const size_t o = b * out.sb + static_cast<size_t>(c) * out.sc + y * out.sy + x * out.sx;
    if (post == C2I_MUL_DACT) {
      out.p[o] = ptx::round_tf32(v * conv_dact_from_out(u[o], act));
    } else {
      if (bias != nullptr) v += bias[c];
      if (post == C2I_PLAIN) {
        out.p[o] = v;
      } else {
        v = conv_act(v, act);
        const float hi = ptx::round_tf32(v);
        out.p[o] = hi;
        out.p[o + static_cast<size_t>(kp) * out.sc] = ptx::round_tf32(v - hi);
      }
    }
  }
}

// NCHW-flattened rows [R, C*P] (element c*P + p)  <->  pixel-major [R*P, C].  src_kp / dst_kp > 0: tf32 pairs
// (hi | lo at +kp); dst_kp == 0: one value, tf32-rounded if `round`.
template <bool TO_PIX>
__global__ void chw_pix_permute_kernel(const float* __restrict__ src, int src_ld, int src_kp, float* __restrict__ dst,
                                       int dst_ld, int dst_kp, int R, int C, int P, int round) {
  const size_t total = static_cast<size_t>(R) * C * P;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    // iterate in destination order
    size_t r;
    int c, p;
    if (TO_PIX) {
      c = static_cast<int>(i % C); p = static_cast<int>((i / C) % P); r = i / (static_cast<size_t>(C) * P);
    } else {
      p = static_cast<int>(i % P); c = static_cast<int>((i / P) % C); r = i / (static_cast<size_t>(C) * P);
    }
    const size_t chw = r * (TO_PIX ? src_ld : dst_ld) + static_cast<size_t>(c) * P + p;
    const size_t pix = (r * P + p) * (TO_PIX ? dst_ld : src_ld) + c;
    const size_t si = TO_PIX ? chw : pix, di = TO_PIX ? pix : chw;
    float v = src[si];
    if (src_kp > 0) v += src[si + src_kp];
    if (dst_kp > 0) {
      const float hi = ptx::round_tf32(v);
      dst[di] = hi;
      dst[di + dst_kp] = ptx::round_tf32(v - hi);
    } else {
      dst[di] = round ? ptx::round_tf32(v) : v;
    }
  }
}

// Deconv weights W [rows = Ci_deconv, cols = Co_deconv*25]:  t3 [cols, 3*kp] = [W^T_hi | W^T_hi | W^T_lo] (the 3xTF32 B
// operand of in . W) and wr [rows, cols] = tf32-rounded copy (B operand of im2col(dout) . W^T).
__global__ void derive_t3_kernel(const float* __restrict__ W, int rows, int cols, float* __restrict__ t3, int kp,
                                 float* __restrict__ wr, int ldr) {
  const size_t total = static_cast<size_t>(rows) * cols;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / cols), c = static_cast<int>(i - static_cast<size_t>(r) * cols);
    const float w = W[i];
    const float hi = ptx::round_tf32(w);
    float* o = t3 + static_cast<size_t>(c) * 3 * kp + r;
    o[0] = hi;
    o[kp] = hi;
    o[2 * kp] = ptx::round_tf32(w - hi);
    if (wr != nullptr) wr[static_cast<size_t>(r) * ldr + c] = hi;
  }
}

// out[c] += sum over rows and the P positions of x[r, c*P + p]   (NCHW-flattened rows; grid (C, splits))
__global__ void chan_sum_nchw_kernel(const float* __restrict__ x, int ld, int R, int C, int P, float* __restrict__ out) {
  const int c = blockIdx.x;
  float acc = 0.0f;
  const size_t total = static_cast<size_t>(R) * P;
  for (size_t i = blockIdx.y * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.y) * blockDim.x) {
    const size_t r = i / P;
    const int p = static_cast<int>(i - r * P);
    acc += x[r * ld + static_cast<size_t>(c) * P + p];
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(out + c, acc);
}

// read at plan-build time (not cached: tests build both variants in one process)
inline bool conv_gemm_enabled() {
  const char* e = std::getenv("ARDAE_CONV_GEMM");
  return !(e && std::atoi(e) == 0);
}

}  // namespace ardae
