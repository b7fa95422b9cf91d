// Stand-alone GPU self-test + micro-benchmark for the 16-bit-spill chain kernel (chain16_sm100.cuh) and the bf16
// weight-gradient contraction (gemm_tn16.cuh).  No Python, no torch.
//   chain16_selftest            correctness: all modes, ragged M, H = 256 / 128 / 64, d-wide first layer (kin < H),
//                               narrow fp32 last layer of the score sweep, 1 / 2 / 4-pair contractions
//   chain16_selftest bench      + timing at the config-2 shape (M = 131072, H = 256, 2L = 10 layers, d = 32)
// The CPU reference runs the whole chain in double precision with the operand roundings the kernel applies
// (tf32 chain operands, bf16 spill); tolerances are bf16 / tf32 rounding bounds, stated next to each check.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <random>
#include <vector>

#include <cuda_fp16.h>

#include "chain16_host.cuh"
#include "gemm_tn16.cuh"

using namespace ardae;

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e = (x);                                                             \
    if (e != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(2);                                                                       \
    }                                                                                \
  } while (0)

static std::mt19937 rng(9876);
static float tf32_rna(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u = (u + 0x1000u) & 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}
static uint16_t bf16_rn(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  return static_cast<uint16_t>((u + 0x7FFFu + ((u >> 16) & 1u)) >> 16);
}
static float bf16_f(uint16_t h) {
  uint32_t u = static_cast<uint32_t>(h) << 16;
  float x;
  memcpy(&x, &u, 4);
  return x;
}
static std::vector<float> randn(size_t n, float scale, bool tf32 = false, bool pos = false) {
  std::normal_distribution<float> d(0.f, 1.f);
  std::vector<float> v(n);
  for (auto& x : v) {
    x = scale * d(rng);
    if (pos) x = std::fabs(x);
    if (tf32) x = tf32_rna(x);
  }
  return v;
}
static std::vector<uint16_t> to16(const std::vector<float>& v) {
  std::vector<uint16_t> o(v.size());
  for (size_t i = 0; i < v.size(); ++i) o[i] = bf16_rn(v[i]);
  return o;
}
template <class T>
static T* dev(const std::vector<T>& h) {
  T* d;
  CK(cudaMalloc(&d, h.size() * sizeof(T) + 16));
  CK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}
template <class T>
static T* dev_fill(size_t n, int byte) {
  T* d;
  CK(cudaMalloc(&d, n * sizeof(T) + 16));
  CK(cudaMemset(d, byte, n * sizeof(T)));
  return d;
}
template <class T>
static std::vector<T> host(const T* d, size_t n) {
  std::vector<T> h(n);
  CK(cudaMemcpy(h.data(), d, n * sizeof(T), cudaMemcpyDeviceToHost));
  return h;
}
static double softplus_d(double x) { return x > 20 ? x : std::log1p(std::exp(x)); }

static int g_fail = 0;
static float time_it(const std::function<void()>& f, int reps);
struct Cmp {
  double max_err = 0, max_ref = 0;
  long bad = 0, n = 0;
  void add(double got, double ref, double tol) {
    const double e = std::fabs(got - ref);
    if (e > max_err) max_err = e;
    if (std::fabs(ref) > max_ref) max_ref = std::fabs(ref);
    if (!(e <= tol)) ++bad;
    ++n;
  }
  // The reference cascades through the CPU-side chain, so a handful of elements whose tf32 / bf16 rounding falls on the
  // other side of a tie exceed the per-element bound by a fraction of an ulp: tolerate 1e-3 of the elements as long as
  // the largest error stays below 1 % of the largest value (indexing / swizzle / synchronisation bugs give O(1) errors).
  void report(const char* what, int mode, int l) {
    const bool fail = !(max_err == max_err) || bad > n / 1000 + 2 || max_err > 1e-2 * max_ref + 1e-6;
    printf("  mode %d layer %d %-8s max_err=%.3e max_ref=%.3e bad=%ld/%ld %s\n", mode, l, what, max_err, max_ref, bad, n,
           fail ? "FAIL" : "ok");
    if (fail) ++g_fail;
  }
};

// kin0: input width of the first layer (forward sweeps); dn > 0: append a narrow linear last layer with dn real
// weight rows (score sweep)
static void test_chain16(int mode, int M, int H, int nl, int kin0, int dn) {
  printf("chain16 mode %d M=%d H=%d layers=%d kin0=%d narrow=%d\n", mode, M, H, nl, kin0, dn);
  const bool s3 = mode == CHAIN_SOFTPLUS3, aux2 = mode == CHAIN_TANGENT || mode == CHAIN_ADJOINT, out2 = mode == CHAIN_TANGENT;
  const bool a0_16 = mode == CHAIN_MUL_SIG || mode == CHAIN_ADJOINT;
  const int group = 48, ng = (M + group - 1) / group;
  std::vector<float> sigma = randn(M, 1.0f);
  float* dsig = dev(sigma);
  // ---- initial activation
  std::vector<float> A0f, A0lo;      // fp32 paths
  std::vector<uint16_t> A016;        // bf16 path
  std::vector<double> Ain((size_t)M * H, 0.0), Ainlo((size_t)M * H, 0.0);  // CPU chain operand (width = current kin)
  float *dA0f = nullptr, *dA0lo = nullptr;
  uint16_t* dA016 = nullptr;
  int kin = a0_16 ? H : kin0;
  if (a0_16) {
    A016 = to16(randn((size_t)M * H, 1.0f));
    dA016 = dev(A016);
    for (size_t i = 0; i < A016.size(); ++i) Ain[i] = bf16_f(A016[i]);
  } else if (s3) {
    std::vector<float> full = randn((size_t)M * kin0, 30.0f);
    A0f.resize(full.size()); A0lo.resize(full.size());
    for (size_t i = 0; i < full.size(); ++i) { A0f[i] = tf32_rna(full[i]); A0lo[i] = tf32_rna(full[i] - A0f[i]); }
    dA0f = dev(A0f); dA0lo = dev(A0lo);
    for (int m = 0; m < M; ++m)
      for (int k = 0; k < kin0; ++k) { Ain[(size_t)m * H + k] = A0f[(size_t)m * kin0 + k]; Ainlo[(size_t)m * H + k] = A0lo[(size_t)m * kin0 + k]; }
  } else {
    A0f = randn((size_t)M * kin0, 1.0f, true);
    dA0f = dev(A0f);
    for (int m = 0; m < M; ++m)
      for (int k = 0; k < kin0; ++k) Ain[(size_t)m * H + k] = A0f[(size_t)m * kin0 + k];
  }
  struct LayerBufs {
    std::vector<float> W, Wfull, bias, gb, colv;
    std::vector<uint16_t> aux1, aux2;
    float *dW = nullptr, *dbias = nullptr, *dgb = nullptr, *dcolv = nullptr, *dcs = nullptr, *dcs2 = nullptr, *dcsw = nullptr, *dout32 = nullptr;
    uint16_t *daux1 = nullptr, *daux2 = nullptr, *dout = nullptr, *dout2 = nullptr;
    int kin = 0, nout = 0, ldw = 0;
  };
  const int ntot = nl + (dn > 0 ? 1 : 0);
  std::vector<LayerBufs> Ls(ntot);
  Chain16Desc d;
  d.mode = mode; d.M = M; d.H = H; d.row_scale = dsig;
  if (a0_16) { d.A0_16 = dA016; d.lda0_16 = H; }
  else { d.A0_32 = dA0f; d.lda0_32 = kin0; d.A0lo = dA0lo; d.lda0lo = kin0; }
  for (int l = 0; l < ntot; ++l) {
    LayerBufs& b = Ls[l];
    const bool narrow = l == nl;
    b.kin = (l == 0) ? kin : H;
    b.nout = narrow ? 32 : H;
    const int wrows = narrow ? dn : H;
    b.ldw = s3 ? 3 * b.kin : b.kin;
    b.Wfull = randn((size_t)wrows * b.kin, s3 ? 0.05f : 0.08f, !s3);
    b.W.assign((size_t)wrows * b.ldw, 0.f);
    for (int o = 0; o < wrows; ++o)
      for (int i = 0; i < b.kin; ++i) {
        const float w = b.Wfull[(size_t)o * b.kin + i];
        if (s3) {
          const float hi = tf32_rna(w), lo = tf32_rna(w - hi);
          b.W[(size_t)o * b.ldw + i] = hi; b.W[(size_t)o * b.ldw + b.kin + i] = hi; b.W[(size_t)o * b.ldw + 2 * b.kin + i] = lo;
        } else {
          b.W[(size_t)o * b.ldw + i] = w;
        }
      }
    b.dW = dev(b.W);
    Chain16LayerDesc q;
    q.W = b.dW; q.ldw = b.ldw; q.kin = b.kin;
    if (narrow) {
      b.dout32 = dev_fill<float>((size_t)M * 32, 0xFF);
      q.nout = 32; q.w_rows = dn; q.out32 = b.dout32; q.ld_out32 = 32;
    } else {
      b.aux1 = to16(randn((size_t)M * H, 1.5f, false, true));
      b.aux2 = to16(randn((size_t)M * H, 1.0f));
      b.bias = randn(H, 1.0f); b.gb = randn((size_t)ng * H, 1.0f); b.colv = randn(H, 1.0f);
      b.daux1 = dev(b.aux1); b.daux2 = dev(b.aux2);
      b.dbias = dev(b.bias); b.dgb = dev(b.gb); b.dcolv = dev(b.colv);
      b.dout = dev_fill<uint16_t>((size_t)M * H, 0xFF); b.dout2 = dev_fill<uint16_t>((size_t)M * H, 0xFF);
      b.dcs = dev_fill<float>(H, 0); b.dcs2 = dev_fill<float>(H, 0); b.dcsw = dev_fill<float>((size_t)H * 3, 0);
      q.aux1 = b.daux1; q.ld1 = H; q.aux2 = b.daux2; q.ld2 = H;
      q.out = (mode == CHAIN_ADJOINT) ? b.daux2 : b.dout; q.ldo = H;   // adjoint: in place over aux2 (the t buffer)
      q.out2 = b.dout2; q.ldo2 = H;
      if (s3) {
        if (l % 2 == 0) q.bias = b.dbias;
        else { q.group_bias = b.dgb; q.group = group; q.ldg = H; q.col_vec = b.dcolv; }
      } else {
        q.colsum = b.dcs; q.colsum_scale = -1.0f;
        if (out2) q.colsum2 = b.dcs2;
        if (l == nl - 1) { q.colsum_w = b.dcsw; q.colsum_w_stride = 3; }
      }
    }
    d.layers.push_back(q);
  }
  PreparedChain16 pr;
  int rc = prepare_chain16(d, &pr);
  if (rc) { printf("  prepare failed %d: %s\n", rc, last_error_string().c_str()); ++g_fail; return; }
  rc = launch_prepared_chain16(pr, 0);
  if (rc) { printf("  launch failed %d: %s\n", rc, last_error_string().c_str()); ++g_fail; return; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  kernel error %s\n", cudaGetErrorString(e)); exit(3); }

  for (int l = 0; l < ntot; ++l) {
    LayerBufs& b = Ls[l];
    const bool narrow = l == nl;
    const int nout = narrow ? 32 : H, wrows = narrow ? dn : H;
    std::vector<uint16_t> out, o2;
    std::vector<float> o32;
    if (narrow) o32 = host(b.dout32, (size_t)M * 32);
    else { out = host(mode == CHAIN_ADJOINT ? b.daux2 : b.dout, (size_t)M * H); o2 = host(b.dout2, (size_t)M * H); }
    Cmp c1, c2, c3, c4, c5;
    std::vector<double> rcs(H, 0.0), rcs2(H, 0.0), rcsw(H, 0.0), mcs(H, 0.0);
    std::vector<double> next((size_t)M * H, 0.0), nextlo((size_t)M * H, 0.0);
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < nout; ++n) {
        double acc = 0, mag = 0;
        if (n < wrows)
          for (int k = 0; k < b.kin; ++k) {
            const double a = Ain[(size_t)m * H + k];
            if (s3) {
              const double whi = b.W[(size_t)n * b.ldw + k], wlo = b.W[(size_t)n * b.ldw + 2 * b.kin + k];
              acc += a * whi + Ainlo[(size_t)m * H + k] * whi + a * wlo;
              mag += std::fabs(a * whi);
            } else {
              const double w = b.Wfull[(size_t)n * b.kin + k];
              acc += a * w;
              mag += std::fabs(a * w);
            }
          }
        if (narrow) {
          c1.add(o32[(size_t)m * 32 + n], acc, 1e-4 * mag + 1e-6);  // fp32 accumulation of tf32 products; operands cascade from the CPU chain
          continue;
        }
        const size_t idx = (size_t)m * H + n;
        const double u = bf16_f(b.aux1[idx]), sg = 1.0 - std::exp(-u), x2 = bf16_f(b.aux2[idx]);
        double ref, ref2 = 0;
        if (mode == CHAIN_MUL_SIG) ref = acc * sg;
        else if (mode == CHAIN_TANGENT) { ref = acc * sg; ref2 = x2 * acc * (1.0 - sg); }
        else if (mode == CHAIN_ADJOINT) ref = acc * sg + x2;
        else {
          double pre = acc;
          if (l % 2 == 0) pre += b.bias[n];
          else pre += b.gb[(size_t)(m / group) * H + n] + (double)sigma[m] * b.colv[n];
          ref = softplus_d(pre);
        }
        // stored value: one bf16 rounding (half an ulp = 2^-8 relative at the bottom of a binade) on top of the
        // accumulation error and of the operand differences that cascade from the CPU-side chain
        const double tol = 1e-4 * mag + 4.0e-3 * std::fabs(ref) + 1e-6;
        c1.add(bf16_f(out[idx]), ref, tol);
        if (out2) c2.add(bf16_f(o2[idx]), ref2, 1e-4 * mag * std::fabs(x2) + 4.0e-3 * std::fabs(ref2) + 1e-6);
        // the chain continues from the fp32 result rounded to tf32 (SOFTPLUS3: the hi / lo pair)
        const float rf = (float)ref;
        if (s3) { const float hi = tf32_rna(rf); next[idx] = hi; nextlo[idx] = tf32_rna(rf - hi); }
        else next[idx] = tf32_rna(rf);
        rcs[n] += -1.0 * next[idx]; rcs2[n] += bf16_f(o2[idx]); rcsw[n] += next[idx] * sigma[m];
        mcs[n] += std::fabs(ref) * (1.0 + std::fabs(sigma[m])) + (out2 ? std::fabs(ref2) : 0.0) + mag * 1e-2;
      }
    c1.report(narrow ? "out32" : "out", mode, l);
    if (out2) c2.report("out2", mode, l);
    if (!s3 && !narrow) {
      auto cs = host(b.dcs, H), cs2 = host(b.dcs2, H), csw = host(b.dcsw, (size_t)H * 3);
      for (int n = 0; n < H; ++n) {
        c3.add(cs[n], rcs[n], 1e-3 * mcs[n] + 1e-5);
        if (out2) c4.add(cs2[n], rcs2[n], 1e-5 * mcs[n] + 1e-5);
        if (l == nl - 1) c5.add(csw[(size_t)n * 3], rcsw[n], 1e-3 * mcs[n] + 1e-5);
      }
      c3.report("colsum", mode, l);
      if (out2) c4.report("colsum2", mode, l);
      if (l == nl - 1) c5.report("colsum_w", mode, l);
    }
    Ain = next; Ainlo = nextlo;
  }
  for (auto& b : Ls) {
    cudaFree(b.dW); cudaFree(b.daux1); cudaFree(b.daux2); cudaFree(b.dout); cudaFree(b.dout2); cudaFree(b.dbias);
    cudaFree(b.dgb); cudaFree(b.dcolv); cudaFree(b.dcs); cudaFree(b.dcs2); cudaFree(b.dcsw); cudaFree(b.dout32);
  }
  cudaFree(dA0f); cudaFree(dA0lo); cudaFree(dA016); cudaFree(dsig);
}

// ---- primal sweep on the fp16 pipe (chain_s3h_sm100.cuh): checked against the EXACT double-precision chain (the
// three-product scheme is fp32-accurate: relative error ~2^-21 of sum |a w|), bf16 rounding of the stored value on top
static uint16_t f16_bits(float x) {
  __half h = __float2half_rn(x);
  uint16_t u;
  memcpy(&u, &h, 2);
  return u;
}
static float f16_val(uint16_t u) {
  __half h;
  memcpy(&h, &u, 2);
  return __half2float(h);
}
static void run_s3h(int M, int H, int nl, int kin0, bool bench) {
  printf("%s s3h M=%d H=%d layers=%d kin0=%d\n", bench ? "bench" : "test", M, H, nl, kin0);
  const int group = 48, ng = (M + group - 1) / group;
  std::vector<float> sigma = randn(M, 1.0f);
  float* dsig = dev(sigma);
  std::vector<float> full = randn((size_t)M * kin0, 3000.0f), a0h(full.size()), a0l(full.size());
  for (size_t i = 0; i < full.size(); ++i) { a0h[i] = tf32_rna(full[i]); a0l[i] = tf32_rna(full[i] - a0h[i]); }
  float *da0h = dev(a0h), *da0l = dev(a0l);
  std::vector<double> Ain((size_t)M * H, 0.0);
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < kin0; ++k) Ain[(size_t)m * H + k] = (double)a0h[(size_t)m * kin0 + k] + a0l[(size_t)m * kin0 + k];
  struct LB { std::vector<float> W, bias, gb, colv; std::vector<uint16_t> W16; uint16_t *dW16, *dout; float *dbias, *dgb, *dcolv; int kin, k16; };
  std::vector<LB> Ls(nl);
  Chain16Desc d;
  d.mode = CHAIN_SOFTPLUS3; d.M = M; d.H = H; d.row_scale = dsig;
  d.A0_32 = da0h; d.lda0_32 = kin0; d.A0lo = da0l; d.lda0lo = kin0;
  for (int l = 0; l < nl; ++l) {
    LB& b = Ls[l];
    b.kin = l == 0 ? kin0 : H;
    b.k16 = (b.kin + 63) / 64 * 64;
    b.W = randn((size_t)H * b.kin, l == 0 ? 0.2f : 0.05f);
    b.W16.assign((size_t)H * 2 * b.k16, 0);
    for (int o = 0; o < H; ++o)
      for (int i = 0; i < b.kin; ++i) {
        const float ws = b.W[(size_t)o * b.kin + i] * 16.0f;
        const uint16_t hi = f16_bits(ws);
        b.W16[(size_t)o * 2 * b.k16 + i] = hi;
        b.W16[(size_t)o * 2 * b.k16 + b.k16 + i] = f16_bits(ws - f16_val(hi));
      }
    b.bias = randn(H, 1.0f); b.gb = randn((size_t)ng * H, 1.0f); b.colv = randn(H, 1.0f);
    b.dW16 = dev(b.W16); b.dbias = dev(b.bias); b.dgb = dev(b.gb); b.dcolv = dev(b.colv);
    b.dout = dev_fill<uint16_t>((size_t)M * H, 0xFF);
    Chain16LayerDesc q;
    q.kin = b.kin; q.W16 = b.dW16; q.ldw16 = 2 * b.k16; q.kin16 = b.k16; q.out = b.dout; q.ldo = H;
    if (l % 2 == 0) q.bias = b.dbias;
    else { q.group_bias = b.dgb; q.group = group; q.ldg = H; q.col_vec = b.dcolv; }
    d.layers.push_back(q);
  }
  PreparedChainS3h pr;
  int rc = prepare_chain_s3h(d, &pr);
  if (rc) { printf("  prepare failed %d: %s\n", rc, last_error_string().c_str()); ++g_fail; return; }
  rc = launch_prepared_chain_s3h(pr, 0);
  cudaError_t e = cudaDeviceSynchronize();
  if (rc || e != cudaSuccess) { printf("  launch failed %d %s\n", rc, cudaGetErrorString(e)); exit(3); }
  if (bench) {
    const float ms = time_it([&]() { launch_prepared_chain_s3h(pr, 0); }, 5);
    const double flops = 2.0 * M * ((double)H * H * (nl - 1) + (double)H * kin0) * 3.0;
    printf("bench s3h M=%d H=%d layers=%d: %.3f ms (%.1f us/layer) tensor %.0f TFLOP/s executed (fp16 pipe), %.0f algorithmic\n", M, H,
           nl, ms, ms * 1e3 / nl, flops / ms * 1e-9, flops / 3 / ms * 1e-9);
  } else {
    for (int l = 0; l < nl; ++l) {
      LB& b = Ls[l];
      auto out = host(b.dout, (size_t)M * H);
      Cmp c1;
      std::vector<double> next((size_t)M * H);
      for (int m = 0; m < M; ++m)
        for (int n = 0; n < H; ++n) {
          double acc = 0, mag = 0;
          for (int k = 0; k < b.kin; ++k) {
            const double t = Ain[(size_t)m * H + k] * b.W[(size_t)n * b.kin + k];
            acc += t; mag += std::fabs(t);
          }
          double pre = acc;
          if (l % 2 == 0) pre += b.bias[n];
          else pre += b.gb[(size_t)(m / group) * H + n] + (double)sigma[m] * b.colv[n];
          const double ref = softplus_d(pre);
          next[(size_t)m * H + n] = ref;
          // fp32-accurate products (3 x 2^-22 per term, random signs), one bf16 rounding of the stored value
          c1.add(bf16_f(out[(size_t)m * H + n]), ref, 3e-6 * mag + 4.0e-3 * std::fabs(ref) + 1e-5);
        }
      c1.report("s3h out", 3, l);
      Ain = next;
    }
  }
  for (auto& b : Ls) { cudaFree(b.dW16); cudaFree(b.dbias); cudaFree(b.dgb); cudaFree(b.dcolv); cudaFree(b.dout); }
  cudaFree(da0h); cudaFree(da0l); cudaFree(dsig);
}

static void test_tn16(int M, int N, int Ny, int K, int npairs, int ldo, int atomic) {
  printf("tn16 M=%d N=%d Ny=%d K=%d pairs=%d ldo=%d atomic=%d\n", M, N, Ny, K, npairs, ldo, atomic);
  std::vector<std::vector<uint16_t>> X(npairs), Y(npairs);
  std::vector<uint16_t*> dX(npairs), dY(npairs);
  GemmTN16Desc d;
  for (int q = 0; q < npairs; ++q) {
    X[q] = to16(randn((size_t)K * M, 1.0f));
    std::vector<float> y = randn((size_t)K * Ny, 1.0f);
    for (int k = 0; k < K; ++k)
      for (int n = N; n < Ny; ++n) y[(size_t)k * Ny + n] = 0.f;  // pad columns are zero in the plan
    Y[q] = to16(y);
    dX[q] = dev(X[q]); dY[q] = dev(Y[q]);
    d.X[q] = dX[q]; d.ldx[q] = M; d.Y[q] = dY[q]; d.ldy[q] = Ny;
  }
  d.npairs = npairs; d.M = M; d.N = N; d.Ny = Ny; d.K = K;
  std::vector<float> out0 = randn((size_t)M * ldo, 1.0f);
  float* dout = dev(out0);
  d.out = dout; d.ldo = ldo; d.atomic = atomic;
  d.workspace_bytes = tn16_workspace_bytes(M, N, K);
  d.workspace = dev_fill<float>(d.workspace_bytes / 4, 0);
  PreparedTN16 pr;
  int rc = prepare_gemm_tn16(d, &pr);
  if (rc) { printf("  prepare failed %d: %s\n", rc, last_error_string().c_str()); ++g_fail; return; }
  rc = launch_prepared_tn16(pr, 0);
  cudaError_t e = cudaDeviceSynchronize();
  if (rc || e != cudaSuccess) { printf("  launch failed %d %s\n", rc, cudaGetErrorString(e)); exit(3); }
  auto out = host(dout, (size_t)M * ldo);
  Cmp c;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = 0, mag = 0;
      for (int q = 0; q < npairs; ++q)
        for (int k = 0; k < K; ++k) {
          const double t = (double)bf16_f(X[q][(size_t)k * M + m]) * bf16_f(Y[q][(size_t)k * Ny + n]);
          acc += t; mag += std::fabs(t);
        }
      c.add(out[(size_t)m * ldo + n], acc + out0[(size_t)m * ldo + n], 2e-6 * mag + 1e-5);  // exact products, fp32 accumulation
    }
  c.report("tn16", npairs, atomic);
  for (int q = 0; q < npairs; ++q) { cudaFree(dX[q]); cudaFree(dY[q]); }
  cudaFree(dout); cudaFree(d.workspace);
}

static float time_it(const std::function<void()>& f, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 2; ++i) f();
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps;
}

static void bench_chain16(int mode, int M, int H, int nl, int kin0, bool narrow) {
  const bool s3 = mode == CHAIN_SOFTPLUS3, aux2 = mode == CHAIN_TANGENT || mode == CHAIN_ADJOINT, out2 = mode == CHAIN_TANGENT;
  const bool a0_16 = mode == CHAIN_MUL_SIG || mode == CHAIN_ADJOINT;
  const size_t n = (size_t)M * H;
  Chain16Desc d;
  d.mode = mode; d.M = M; d.H = H;
  float* dsig = dev_fill<float>(M, 0);
  d.row_scale = dsig;
  std::vector<void*> bufs;
  if (a0_16) { d.A0_16 = dev_fill<uint16_t>(n, 0x3c); d.lda0_16 = H; }
  else { d.A0_32 = dev_fill<float>((size_t)M * kin0, 0); d.lda0_32 = kin0; d.A0lo = dev_fill<float>((size_t)M * kin0, 0); d.lda0lo = kin0; }
  float* dbias = dev_fill<float>(H, 0);
  const int ntot = nl + (narrow ? 1 : 0);
  for (int l = 0; l < ntot; ++l) {
    const int kin = (l == 0 && !a0_16) ? kin0 : H;
    const int ldw = s3 ? 3 * kin : kin;
    std::vector<float> W = randn((size_t)H * ldw, 0.05f, true);  // a distinct weight per layer (L2 traffic as in the step)
    float* dW = dev(W);
    bufs.push_back(dW);
    Chain16LayerDesc q;
    q.W = dW; q.ldw = ldw; q.kin = kin;
    if (l == nl) {
      q.nout = 32; q.w_rows = 32; q.out32 = dev_fill<float>((size_t)M * 32, 0); q.ld_out32 = 32;
      bufs.push_back(q.out32);
    } else {
      uint16_t *a1 = dev_fill<uint16_t>(n, 0x3c), *a2 = aux2 ? dev_fill<uint16_t>(n, 0) : nullptr;
      uint16_t *o = (mode == CHAIN_ADJOINT) ? a2 : dev_fill<uint16_t>(n, 0), *o2 = out2 ? dev_fill<uint16_t>(n, 0) : nullptr;
      bufs.push_back(a1); bufs.push_back(a2); if (mode != CHAIN_ADJOINT) bufs.push_back(o); bufs.push_back(o2);
      q.aux1 = a1; q.ld1 = H; q.aux2 = a2; q.ld2 = H; q.out = o; q.ldo = H; q.out2 = o2; q.ldo2 = H;
      if (s3) q.bias = dbias;
    }
    d.layers.push_back(q);
  }
  PreparedChain16 pr;
  int rc = prepare_chain16(d, &pr);
  if (rc) { printf("bench prepare failed %d: %s\n", rc, last_error_string().c_str()); ++g_fail; return; }
  if (getenv("CHAIN16_TIMES")) {
    const int nlay = (int)d.layers.size();
    long long* dt;
    const size_t nn = (size_t)nlay * 2 * 4 * 8;
    CK(cudaMalloc(&dt, nn * sizeof(long long)));
    CK(cudaMemset(dt, 0, nn * sizeof(long long)));
    pr.params.dbg = dt;
    launch_prepared_chain16(pr, 0);
    CK(cudaDeviceSynchronize());
    std::vector<long long> hdt(nn);
    CK(cudaMemcpy(hdt.data(), dt, nn * sizeof(long long), cudaMemcpyDeviceToHost));
    const long long base = hdt[0];
    printf("  times mode %d (cycles, CTA 0, warp quarter 1): layer half group | start  acc_full  kfree  aux_full  computed  bar_b  pre_arrive  arrived\n", mode);
    for (int l = 0; l < nlay && l < 4; ++l)
      for (int h = 0; h < 2; ++h)
        for (int g = 0; g < 4; ++g) {
          const long long* t = &hdt[((size_t)(l * 2 + h) * 4 + g) * 8];
          printf("    %d %d %d | %7lld %7lld %7lld %7lld %7lld %7lld %7lld %7lld\n", l, h, g, t[0] - base, t[1] - base, t[2] - base,
                 t[3] - base, t[4] - base, t[5] - base, t[6] - base, t[7] - base);
        }
    pr.params.dbg = nullptr;
    cudaFree(dt);
  }
  const float ms = time_it([&]() { launch_prepared_chain16(pr, 0); }, 5);
  const double arrays = s3 ? 1.0 : (1.0 + (aux2 ? 1 : 0) + 1.0 + (out2 ? 1 : 0));  // per layer: aux reads + out writes
  const double bytes = (arrays * nl + (a0_16 ? 1.0 : 0.0)) * n * 2.0;
  const double flops = 2.0 * M * (double)H * H * (nl - (a0_16 ? 0 : 1)) * (s3 ? 3.0 : 1.0);
  printf("bench16 mode %d M=%d H=%d layers=%d%s: %.3f ms  (%.1f us/layer)  HBM %.0f GB/s  tensor %.0f TFLOP/s (executed)\n",
         mode, M, H, ntot, narrow ? " (last narrow)" : "", ms, ms * 1e3 / ntot, bytes / ms * 1e-6, flops / ms * 1e-9);
  for (void* b : bufs) if (b) cudaFree(b);
  cudaFree(dsig); cudaFree(dbias);
  cudaFree(const_cast<uint16_t*>(d.A0_16)); cudaFree(const_cast<float*>(d.A0_32)); cudaFree(const_cast<float*>(d.A0lo));
}

// several same-shape contractions in one launch (gemm_tn16_multi_kernel): each output checked like test_tn16
static void run_tn16_multi(int M, int N, int K, int nprob, bool bench) {
  printf("%s tn16_multi M=%d N=%d K=%d problems=%d\n", bench ? "bench" : "test", M, N, K, nprob);
  std::vector<std::vector<uint16_t>> X(2 * nprob), Y(2 * nprob);
  std::vector<uint16_t*> dX(2 * nprob), dY(2 * nprob);
  std::vector<std::vector<float>> out0(nprob);
  std::vector<float*> dout(nprob);
  std::vector<GemmTN16Desc> ds(nprob);
  float* ws = dev_fill<float>(tn16_workspace_bytes(M, N, K) / 4, 0);
  for (int q = 0; q < nprob; ++q) {
    GemmTN16Desc& d = ds[q];
    d.npairs = (!bench && q % 3 == 2) ? 1 : 2;
    for (int k = 0; k < d.npairs; ++k) {
      const int i = 2 * q + k;
      if (bench) { dX[i] = dev_fill<uint16_t>((size_t)K * M, 0x3c); dY[i] = dev_fill<uint16_t>((size_t)K * N, 0x3c); }
      else {
        X[i] = to16(randn((size_t)K * M, 1.0f)); Y[i] = to16(randn((size_t)K * N, 1.0f));
        dX[i] = dev(X[i]); dY[i] = dev(Y[i]);
      }
      d.X[k] = dX[i]; d.ldx[k] = M; d.Y[k] = dY[i]; d.ldy[k] = N;
    }
    d.M = M; d.N = N; d.Ny = N; d.K = K;
    out0[q] = randn((size_t)M * N, 1.0f);
    dout[q] = dev(out0[q]);
    d.out = dout[q]; d.ldo = N; d.workspace = ws; d.workspace_bytes = tn16_workspace_bytes(M, N, K);
  }
  PreparedTN16Multi pr;
  int rc = prepare_gemm_tn16_multi(ds, &pr);
  if (rc) { printf("  prepare failed %d: %s\n", rc, last_error_string().c_str()); ++g_fail; return; }
  rc = launch_prepared_tn16_multi(pr, 0);
  cudaError_t e = cudaDeviceSynchronize();
  if (rc || e != cudaSuccess) { printf("  launch failed %d %s\n", rc, cudaGetErrorString(e)); exit(3); }
  if (bench) {
    const float ms = time_it([&]() { launch_prepared_tn16_multi(pr, 0); }, 5);
    const double bytes = (double)nprob * 2 * K * (M + N) * 2.0;
    printf("bench tn16_multi M=%d N=%d K=%d problems=%d grid=%dx%dx%d: %.1f us (%.1f us per problem)  HBM %.0f GB/s\n", M, N, K,
           nprob, pr.grid.x, pr.grid.y, pr.grid.z, ms * 1e3, ms * 1e3 / nprob, bytes / ms * 1e-6);
  } else {
    for (int q = 0; q < nprob; ++q) {
      auto out = host(dout[q], (size_t)M * N);
      Cmp c;
      for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
          double acc = 0, mag = 0;
          for (int k2 = 0; k2 < ds[q].npairs; ++k2)
            for (int k = 0; k < K; ++k) {
              const double t = (double)bf16_f(X[2 * q + k2][(size_t)k * M + m]) * bf16_f(Y[2 * q + k2][(size_t)k * N + n]);
              acc += t; mag += std::fabs(t);
            }
          c.add(out[(size_t)m * N + n], acc + out0[q][(size_t)m * N + n], 2e-6 * mag + 1e-5);
        }
      c.report("tn16multi", q, ds[q].npairs);
    }
  }
  for (int i = 0; i < 2 * nprob; ++i) { if (dX[i]) cudaFree(dX[i]); if (dY[i]) cudaFree(dY[i]); }
  for (int q = 0; q < nprob; ++q) cudaFree(dout[q]);
  cudaFree(ws);
}

static void bench_tn16(int M, int N, int K, int npairs) {
  GemmTN16Desc d;
  for (int q = 0; q < npairs; ++q) {
    d.X[q] = dev_fill<uint16_t>((size_t)K * M, 0x3c); d.ldx[q] = M;
    d.Y[q] = dev_fill<uint16_t>((size_t)K * N, 0x3c); d.ldy[q] = N;
  }
  d.npairs = npairs; d.M = M; d.N = N; d.Ny = N; d.K = K;
  d.out = dev_fill<float>((size_t)M * N, 0); d.ldo = N;
  d.workspace_bytes = tn16_workspace_bytes(M, N, K);
  d.workspace = dev_fill<float>(d.workspace_bytes / 4, 0);
  PreparedTN16 pr;
  int rc = prepare_gemm_tn16(d, &pr);
  if (rc) { printf("bench prepare failed %d: %s\n", rc, last_error_string().c_str()); ++g_fail; return; }
  const float ms = time_it([&]() { launch_prepared_tn16(pr, 0); }, 10);
  const double bytes = (double)npairs * K * (M + N) * 2.0;
  printf("bench tn16 M=%d N=%d K=%d pairs=%d grid=%dx%dx%d: %.1f us  HBM %.0f GB/s  tensor %.0f TFLOP/s\n", M, N, K, npairs,
         pr.grid.x, pr.grid.y, pr.grid.z, ms * 1e3, bytes / ms * 1e-6, 2.0 * npairs * K * (double)M * N / ms * 1e-9);
  for (int q = 0; q < npairs; ++q) { cudaFree(const_cast<uint16_t*>(d.X[q])); cudaFree(const_cast<uint16_t*>(d.Y[q])); }
  cudaFree(d.out); cudaFree(d.workspace);
}

int main(int argc, char** argv) {
  const bool bench = argc > 1 && !strcmp(argv[1], "bench");
  const char* only = argc > 2 ? argv[2] : "";
  if (!*only || !strcmp(only, "tn")) {
    test_tn16(256, 256, 256, 1000, 1, 256, 1);
    test_tn16(256, 256, 256, 4096, 2, 256, 1);
    test_tn16(256, 256, 256, 777, 2, 513, -1);   // odd pitch: partial tiles + reduce launch
    test_tn16(256, 32, 64, 2048, 4, 32, 1);      // d-wide contraction, four pairs, zero-padded Y
    test_tn16(128, 2, 64, 300, 4, 2, 0);         // toy d = 2 (scalar reductions)
    test_tn16(64, 64, 64, 129, 1, 64, 1);
    run_tn16_multi(256, 256, 2000, 4, false);
    run_tn16_multi(128, 128, 700, 3, false);
  }
  if (!*only || !strcmp(only, "chain")) {
    for (int mode = 0; mode < CHAIN_NUM_MODES; ++mode) {
      const bool fwd = mode == CHAIN_TANGENT || mode == CHAIN_SOFTPLUS3;
      test_chain16(mode, 300, 256, 3, fwd ? 32 : 256, mode == CHAIN_MUL_SIG ? 20 : 0);
      test_chain16(mode, 128, 64, 2, fwd ? 32 : 64, mode == CHAIN_MUL_SIG ? 32 : 0);
      test_chain16(mode, 1000, 128, 4, fwd ? 64 : 128, mode == CHAIN_MUL_SIG ? 2 : 0);
      test_chain16(mode, 520, 256, 3, 256, 0);
    }
  }
  if (!*only || !strcmp(only, "s3h")) {
    run_s3h(300, 256, 3, 32, false);
    run_s3h(1000, 128, 4, 64, false);
    run_s3h(520, 256, 3, 256, false);
    run_s3h(200, 256, 4, 96, false);
  }
  if (bench) {
    run_s3h(131072, 256, 10, 32, true);
    for (int mode = 0; mode < CHAIN_NUM_MODES; ++mode) {
      const bool fwd = mode == CHAIN_TANGENT || mode == CHAIN_SOFTPLUS3;
      bench_chain16(mode, 131072, 256, fwd ? 10 : 9, fwd ? 32 : 256, mode == CHAIN_MUL_SIG);
    }
    bench_tn16(256, 256, 131072, 2);
    run_tn16_multi(256, 256, 131072, 8, true);
    bench_tn16(256, 64, 131072, 4);
  }
  printf(g_fail ? "CHAIN16 SELFTEST FAILED (%d)\n" : "CHAIN16 SELFTEST OK\n", g_fail);
  return g_fail ? 1 : 0;
}
