// Stand-alone GPU self-test for the tcgen05 GEMM kernels (no Python, no torch).
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo \
//              -I pytorch-ardae-vae_b200/csrc tests/native/gemm_selftest.cu -o gemm_selftest
// Inputs are generated exactly representable in tf32, so the CPU reference (double accumulate)
// must match to fp32 accumulation error whatever rounding the tensor core applies.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "gemm_host.cuh"

using namespace ardae;

static float tf32_trunc(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}
static std::mt19937 rng(1234);
static void fill(std::vector<float>& v, float scale = 1.0f) {
  std::normal_distribution<float> d(0.f, 1.f);
  for (auto& x : v) x = tf32_trunc(scale * d(rng));
}
static void fill_pos(std::vector<float>& v) {
  std::normal_distribution<float> d(0.f, 1.f);
  for (auto& x : v) x = tf32_trunc(std::fabs(d(rng)) + 0.0f);
}
#define CK(x)                                                                     \
  do {                                                                            \
    cudaError_t e = (x);                                                          \
    if (e != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(2);                                                                    \
    }                                                                             \
  } while (0)

template <class T>
static T* dev(const std::vector<T>& h) {
  T* d;
  CK(cudaMalloc(&d, h.size() * sizeof(T)));
  CK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}
static std::vector<float> host(const float* d, size_t n) {
  std::vector<float> h(n);
  CK(cudaMemcpy(h.data(), d, n * sizeof(float), cudaMemcpyDeviceToHost));
  return h;
}

static double softplus_d(double x) { return x > 20 ? x : std::log1p(std::exp(x)); }

struct Cmp {
  double max_abs = 0, max_ref = 0;
  int bad = 0;
  void add(double got, double ref, double tol_abs) {
    double e = std::fabs(got - ref);
    if (e > max_abs) max_abs = e;
    if (std::fabs(ref) > max_ref) max_ref = std::fabs(ref);
    if (!(e <= tol_abs)) ++bad;
  }
};

static int g_fail = 0;
static void report(const char* name, const Cmp& c) {
  printf("%-44s max_abs_err=%.3e max_ref=%.3e bad=%d  %s\n", name, c.max_abs, c.max_ref, c.bad,
         c.bad ? "FAIL" : "ok");
  if (c.bad) ++g_fail;
}

// mode-generic NT test
static int g_persist = 0;
static int g_split = 0;
static int g_mc = -1;
static int g_deep = -1;
static int g_pair = -1;
static void test_nt(const char* name, int M, int N, int K, int mode, int force_bn, bool use_bias,
                    bool use_group, bool use_rank1, bool use_colsum) {
  const int lda = (K + 3) / 4 * 4, ldb = lda, ldo = (N + 3) / 4 * 4;
  std::vector<float> A((size_t)M * lda), B((size_t)N * ldb), bias(N), colv(N), rows(M), roww(M);
  std::vector<float> aux1((size_t)M * ldo), aux2((size_t)M * ldo);
  const int group = 48;
  const int ng = (M + group - 1) / group;
  std::vector<float> gb((size_t)ng * ldo);
  fill(A, 0.5f); fill(B, 0.25f); fill(bias); fill(colv); fill(rows); fill(roww); fill(gb);
  if (mode == EPI_MUL_STEP) fill(aux1); else fill_pos(aux1);
  fill(aux2);
  float *dA = dev(A), *dB = dev(B), *dbias = dev(bias), *dcolv = dev(colv), *drows = dev(rows),
        *droww = dev(roww), *daux1 = dev(aux1), *daux2 = dev(aux2), *dgb = dev(gb);
  float *dout, *dout2, *dcs, *dcsw;
  CK(cudaMalloc(&dout, (size_t)M * ldo * 4)); CK(cudaMemset(dout, 0xFF, (size_t)M * ldo * 4));
  CK(cudaMalloc(&dout2, (size_t)M * ldo * 4)); CK(cudaMemset(dout2, 0xFF, (size_t)M * ldo * 4));
  CK(cudaMalloc(&dcs, N * 4)); CK(cudaMemset(dcs, 0, N * 4));
  CK(cudaMalloc(&dcsw, N * 4)); CK(cudaMemset(dcsw, 0, N * 4));
  GemmNTDesc d;
  d.A = dA; d.lda = lda; d.B = dB; d.ldb = ldb; d.out = dout; d.ldo = ldo; d.out2 = dout2; d.ldo2 = ldo;
  d.aux1 = daux1; d.ld1 = ldo; d.aux2 = daux2; d.ld2 = ldo;
  d.M = M; d.N = N; d.K = K; d.mode = mode; d.alpha = 0.75f; d.round_out = 0; d.force_block_n = force_bn;
  d.persist = g_persist; d.split_out = g_split; d.multicast = g_mc; d.deep = g_deep; d.pair = g_pair;
  if (use_bias) d.bias = dbias;
  if (use_group) { d.group_bias = dgb; d.group = group; d.ldg = ldo; }
  if (use_rank1) { d.row_scale = drows; d.col_vec = dcolv; }
  if (use_colsum) { d.colsum = dcs; d.colsum_w = dcsw; d.row_w = droww; }
  PreparedNT pr;
  int rc = prepare_gemm_nt(d, &pr);
  if (rc) { printf("%s: prepare failed %d: %s\n", name, rc, last_error_string().c_str()); ++g_fail; return; }
  rc = launch_prepared_nt(pr, 0);
  if (rc) { printf("%s: launch failed %d: %s\n", name, rc, last_error_string().c_str()); ++g_fail; return; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: kernel error %s\n", name, cudaGetErrorString(e)); exit(3); }
  auto out = host(dout, (size_t)M * ldo), out2 = host(dout2, (size_t)M * ldo);
  auto cs = host(dcs, N), csw = host(dcsw, N);
  Cmp c1, c2, c3, c4;
  std::vector<double> rcs(N, 0.0), rcsw(N, 0.0);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = 0, mag = 0;
      for (int k = 0; k < K; ++k) {
        double t = (double)A[(size_t)m * lda + k] * B[(size_t)n * ldb + k];
        acc += t; mag += std::fabs(t);
      }
      double pre = 0.75 * acc;
      if (use_bias) pre += bias[n];
      if (use_group) pre += gb[(size_t)(m / group) * ldo + n];
      if (use_rank1) pre += (double)rows[m] * colv[n];
      double a1 = aux1[(size_t)m * ldo + n], a2 = aux2[(size_t)m * ldo + n];
      double s = 1.0 - std::exp(-a1);
      double r1 = 0, r2 = 0;
      switch (mode) {
        case EPI_LINEAR: r1 = pre; break;
        case EPI_RELU: r1 = pre > 0 ? pre : 0; break;
        case EPI_SOFTPLUS: r1 = softplus_d(pre); break;
        case EPI_MUL_SIG: r1 = pre * s; break;
        case EPI_MUL_STEP: r1 = a1 > 0 ? pre : 0; break;
        case EPI_TANGENT: r1 = pre * s; r2 = a2 * pre * (1 - s); break;
        case EPI_ADJOINT: r1 = pre * s + a2; break;
      }
      double tol = 2e-5 * (mag + std::fabs(pre) + 1.0) * (1.0 + std::fabs(a2));
      if (g_split) {  // out = tf32 hi, out2 = lo: their sum reproduces the fp32 result
        c1.add((double)out[(size_t)m * ldo + n] + out2[(size_t)m * ldo + n], r1, tol);
      } else {
        c1.add(out[(size_t)m * ldo + n], r1, tol);
      }
      if (mode == EPI_TANGENT) c2.add(out2[(size_t)m * ldo + n], r2, tol);
      rcs[n] += g_split ? (double)out[(size_t)m * ldo + n] : r1; rcsw[n] += (g_split ? (double)out[(size_t)m * ldo + n] : r1) * roww[m];
    }
  char buf[128];
  snprintf(buf, sizeof(buf), "%s out", name); report(buf, c1);
  if (mode == EPI_TANGENT) { snprintf(buf, sizeof(buf), "%s out2", name); report(buf, c2); }
  if (use_colsum) {
    for (int n = 0; n < N; ++n) { c3.add(cs[n], rcs[n], 1e-3 * (1 + std::fabs(rcs[n])) + 2e-4 * M); c4.add(csw[n], rcsw[n], 1e-3 * (1 + std::fabs(rcsw[n])) + 2e-4 * M); }
    snprintf(buf, sizeof(buf), "%s colsum", name); report(buf, c3);
    snprintf(buf, sizeof(buf), "%s colsum_w", name); report(buf, c4);
  }
  if (c1.bad) {
    printf("   first row got/ref:");
    for (int n = 0; n < 8 && n < N; ++n) {
      double acc = 0; for (int k = 0; k < K; ++k) acc += (double)A[k] * B[(size_t)n * ldb + k];
      printf(" %.4f/%.4f", out[n], 0.75 * acc);
    }
    printf("\n");
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dbias); cudaFree(dcolv); cudaFree(drows); cudaFree(droww);
  cudaFree(daux1); cudaFree(daux2); cudaFree(dgb); cudaFree(dout); cudaFree(dout2); cudaFree(dcs); cudaFree(dcsw);
}

static void test_tn(const char* name, int M, int N, int K, bool two) {
  const int ldx = (M + 3) / 4 * 4, ldy = (N + 3) / 4 * 4, ldo = N + 4;
  std::vector<float> X0((size_t)K * ldx), Y0((size_t)K * ldy), X1((size_t)K * ldx), Y1((size_t)K * ldy);
  std::vector<float> O((size_t)M * ldo);
  fill(X0, 0.5f); fill(Y0, 0.5f); fill(X1, 0.5f); fill(Y1, 0.5f); fill(O);
  float *dX0 = dev(X0), *dY0 = dev(Y0), *dX1 = dev(X1), *dY1 = dev(Y1), *dO = dev(O);
  size_t wsb = tn_workspace_bytes(M, N, K);
  float* ws; CK(cudaMalloc(&ws, wsb)); CK(cudaMemset(ws, 0xFF, wsb));
  GemmTNDesc d;
  d.X0 = dX0; d.ldx0 = ldx; d.Y0 = dY0; d.ldy0 = ldy;
  if (two) { d.X1 = dX1; d.ldx1 = ldx; d.Y1 = dY1; d.ldy1 = ldy; }
  d.M = M; d.N = N; d.K = K; d.out = dO; d.ldo = ldo; d.scale = 0.5f; d.beta = 1.0f;
  d.workspace = ws; d.workspace_bytes = wsb;
  PreparedTN pr;
  int rc = prepare_gemm_tn(d, &pr);
  if (rc) { printf("%s: prepare failed %d: %s\n", name, rc, last_error_string().c_str()); ++g_fail; return; }
  rc = launch_prepared_tn(pr, 0);
  if (rc) { printf("%s: launch failed %d: %s\n", name, rc, last_error_string().c_str()); ++g_fail; return; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: kernel error %s\n", name, cudaGetErrorString(e)); exit(3); }
  auto out = host(dO, (size_t)M * ldo);
  Cmp c;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = 0, mag = 0;
      for (int k = 0; k < K; ++k) {
        double t = (double)X0[(size_t)k * ldx + m] * Y0[(size_t)k * ldy + n];
        if (two) t += (double)X1[(size_t)k * ldx + m] * Y1[(size_t)k * ldy + n];
        acc += t; mag += std::fabs(t);
      }
      c.add(out[(size_t)m * ldo + n], O[(size_t)m * ldo + n] + 0.5 * acc, 2e-5 * (mag + 1));
    }
  report(name, c);
  if (c.bad) {
    printf("   first row got/ref:");
    for (int n = 0; n < 8 && n < N; ++n) {
      double acc = 0; for (int k = 0; k < K; ++k) acc += (double)X0[(size_t)k * ldx] * Y0[(size_t)k * ldy + n] + (two ? (double)X1[(size_t)k * ldx] * Y1[(size_t)k * ldy + n] : 0.0);
      printf(" %.4f/%.4f", out[n], O[n] + 0.5 * acc);
    }
    printf("\n");
  }
  cudaFree(dX0); cudaFree(dY0); cudaFree(dX1); cudaFree(dY1); cudaFree(dO); cudaFree(ws);
}

static int g_dbg = 0;
static int g_prefetch = -1;
static void bench_nt(int M, int N, int K, int mode, int persist = 0, int split = 0, int mc = -1, int deep = -1, int pair = -1) {
  float *A, *B, *O, *O2, *X1, *X2;
  CK(cudaMalloc(&A, (size_t)M * K * 4)); CK(cudaMalloc(&B, (size_t)N * K * 4));
  CK(cudaMalloc(&O, (size_t)M * N * 4)); CK(cudaMalloc(&O2, (size_t)M * N * 4));
  CK(cudaMalloc(&X1, (size_t)M * N * 4)); CK(cudaMalloc(&X2, (size_t)M * N * 4));
  CK(cudaMemset(A, 0, (size_t)M * K * 4)); CK(cudaMemset(B, 0, (size_t)N * K * 4));
  CK(cudaMemset(X1, 0, (size_t)M * N * 4)); CK(cudaMemset(X2, 0, (size_t)M * N * 4));
  GemmNTDesc d;
  d.A = A; d.lda = K; d.B = B; d.ldb = K; d.out = O; d.ldo = N; d.out2 = O2; d.ldo2 = N;
  d.aux1 = X1; d.ld1 = N; d.aux2 = X2; d.ld2 = N; d.M = M; d.N = N; d.K = K; d.mode = mode;
  d.persist = persist; d.split_out = split; d.debug_flags = g_dbg; d.multicast = mc; d.deep = deep; d.pair = pair; d.prefetch = g_prefetch;
  PreparedNT pr;
  if (prepare_gemm_nt(d, &pr)) { printf("bench prepare failed: %s\n", last_error_string().c_str()); return; }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) launch_prepared_nt(pr, 0);
  cudaEventRecord(e0);
  const int iters = 10;
  for (int i = 0; i < iters; ++i) launch_prepared_nt(pr, 0);
  cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
  int narr = 2 + (mode >= EPI_MUL_SIG) + (mode >= EPI_TANGENT) + (mode == EPI_TANGENT);
  double bytes = (double)M * K * 4 + (double)(narr - 1) * M * N * 4;
  printf("bench NT%s%s M=%d N=%d K=%d mode=%d: %.1f us  %.1f TFLOP/s  %.0f GB/s (%d arrays)\n", persist > 0 ? "(persist)" : "", mc > 0 ? "(mc)" : (pair > 0 ? "(cg2)" : ""), M, N, K, mode,
         ms * 1e3, 2.0 * M * N * K / ms * 1e-9, bytes / ms * 1e-6, narr);
  cudaFree(A); cudaFree(B); cudaFree(O); cudaFree(O2); cudaFree(X1); cudaFree(X2);
}
static void bench_tn(int M, int N, int K, bool two) {
  float *X, *Y, *X1, *Y1, *O, *ws;
  CK(cudaMalloc(&X, (size_t)K * M * 4)); CK(cudaMalloc(&Y, (size_t)K * N * 4));
  CK(cudaMalloc(&X1, (size_t)K * M * 4)); CK(cudaMalloc(&Y1, (size_t)K * N * 4));
  CK(cudaMalloc(&O, (size_t)M * N * 4));
  CK(cudaMemset(X, 0, (size_t)K * M * 4)); CK(cudaMemset(Y, 0, (size_t)K * N * 4));
  CK(cudaMemset(X1, 0, (size_t)K * M * 4)); CK(cudaMemset(Y1, 0, (size_t)K * N * 4));
  size_t wsb = tn_workspace_bytes(M, N, K); CK(cudaMalloc(&ws, wsb));
  GemmTNDesc d;
  d.X0 = X; d.ldx0 = M; d.Y0 = Y; d.ldy0 = N; if (two) { d.X1 = X1; d.ldx1 = M; d.Y1 = Y1; d.ldy1 = N; }
  d.M = M; d.N = N; d.K = K; d.out = O; d.ldo = N; d.workspace = ws; d.workspace_bytes = wsb;
  if (getenv("TN_ATOMIC")) { d.scale = 1.0f; d.beta = 1.0f; }  // the red.global.add path of the training plan
  PreparedTN pr;
  if (prepare_gemm_tn(d, &pr)) { printf("bench prepare failed: %s\n", last_error_string().c_str()); return; }
  if (getenv("TN_TIMES")) {
    const size_t nc = (size_t)pr.grid.x * pr.grid.y;
    long long* dt; CK(cudaMalloc(&dt, nc * 4 * sizeof(long long))); CK(cudaMemset(dt, 0, nc * 4 * sizeof(long long)));
    pr.params.dbg = dt;
    launch_prepared_tn(pr, 0); CK(cudaDeviceSynchronize());
    launch_prepared_tn(pr, 0); CK(cudaDeviceSynchronize());
    std::vector<long long> h(nc * 4);
    CK(cudaMemcpy(h.data(), dt, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    long long t0 = h[0], tend = 0; double main_sum = 0, epi_sum = 0; long long main_max = 0, epi_max = 0, start_max = 0;
    for (size_t c = 0; c < nc; ++c) { if (h[c * 4] < t0) t0 = h[c * 4]; }
    for (size_t c = 0; c < nc; ++c) {
      const long long s0 = h[c * 4] - t0, mn = h[c * 4 + 1] - h[c * 4], ep = h[c * 4 + 2] - h[c * 4 + 1];
      main_sum += mn; epi_sum += ep;
      if (mn > main_max) main_max = mn; if (ep > epi_max) epi_max = ep; if (s0 > start_max) start_max = s0;
      if (h[c * 4 + 2] - t0 > tend) tend = h[c * 4 + 2] - t0;
    }
    printf("  TN phases (cycles, %zu CTAs; clocks of different SMs are only roughly aligned): latest start %lld, main loop avg %.0f max %lld, "
           "epilogue avg %.0f max %lld, last end %lld\n", nc, start_max, main_sum / nc, main_max, epi_sum / nc, epi_max, tend);
    pr.params.dbg = nullptr; cudaFree(dt);
  }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) launch_prepared_tn(pr, 0);
  cudaEventRecord(e0);
  const int iters = 10;
  for (int i = 0; i < iters; ++i) launch_prepared_tn(pr, 0);
  cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
  double fl = 2.0 * M * N * K * (two ? 2 : 1), bytes = (double)K * (M + N) * 4 * (two ? 2 : 1);
  printf("bench TN M=%d N=%d K=%d pairs=%d: %.1f us  %.1f TFLOP/s  %.0f GB/s  (grid %d,%d,%d)\n", M, N, K,
         two ? 2 : 1, ms * 1e3, fl / ms * 1e-9, bytes / ms * 1e-6, pr.grid.x, pr.grid.y, pr.grid.z);
  cudaFree(X); cudaFree(Y); cudaFree(X1); cudaFree(Y1); cudaFree(O); cudaFree(ws);
}

int main(int argc, char** argv) {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  const bool quick = argc > 1 && !strcmp(argv[1], "quick");
  if (argc > 1 && !strcmp(argv[1], "pf")) {
    for (int pf : {-1, 148, 296, 444}) {
      g_prefetch = pf;
      printf("prefetch distance %d tiles\n", pf);
      bench_nt(131072, 256, 768, EPI_SOFTPLUS, -1, 1, -1, -1, 1);
      bench_nt(131072, 256, 256, EPI_LINEAR, -1);
      bench_nt(131072, 256, 256, EPI_MUL_SIG, -1);
      bench_nt(131072, 256, 256, EPI_TANGENT, -1);
      bench_nt(131072, 256, 256, EPI_ADJOINT, -1);
    }
    return 0;
  }
  if (argc > 1 && !strcmp(argv[1], "tn")) {  // weight-gradient contraction only (TN_ATOMIC=1 TN_TIMES=1 for the phase stamps)
    bench_tn(256, 256, 131072, true);
    return 0;
  }
  if (argc > 5 && !strcmp(argv[1], "one")) {  // one <mode> <K> <split> <pair>: a single bench config (for ncu)
    bench_nt(131072, 256, atoi(argv[3]), atoi(argv[2]), -1, atoi(argv[4]), -1, -1, atoi(argv[5]));
    return 0;
  }
  if (argc > 1 && !strcmp(argv[1], "dbg")) {
    for (int dbg : {0, 1, 2, 3}) {
      g_dbg = dbg;
      printf("debug_flags=%d\n", dbg);
      bench_nt(131072, 256, 768, EPI_SOFTPLUS, -1, 1);
      bench_nt(131072, 256, 768, EPI_SOFTPLUS, 1, 1);
      bench_nt(131072, 256, 256, EPI_MUL_SIG, 1);
      bench_nt(131072, 256, 256, EPI_TANGENT, 1);
      bench_nt(131072, 256, 32, EPI_SOFTPLUS, -1);
      bench_nt(131072, 256, 256, EPI_LINEAR, -1);
      bench_nt(131072, 256, 256, EPI_MUL_SIG, -1);
      bench_nt(131072, 256, 256, EPI_TANGENT, -1);
    }
    return 0;
  }
  test_nt("NT linear 128x32x32 bn32", 128, 32, 32, EPI_LINEAR, 32, false, false, false, false);
  test_nt("NT linear 128x256x256", 128, 256, 256, EPI_LINEAR, 0, false, false, false, false);
  test_nt("NT linear 256x256x64 bn128", 256, 256, 64, EPI_LINEAR, 128, false, false, false, false);
  test_nt("NT linear 256x64x96 bn64", 256, 64, 96, EPI_LINEAR, 64, true, false, false, false);
  test_tn("TN 128x32 K=64", 128, 32, 64, false);
  test_tn("TN 256x256 K=1024", 256, 256, 1024, false);
  if (!quick) {
    test_nt("NT softplus ragged 200x300x100", 200, 300, 100, EPI_SOFTPLUS, 0, true, true, true, true);
    test_nt("NT relu ragged 130x784x300", 130, 784, 300, EPI_RELU, 0, true, false, false, true);
    test_nt("NT mul_sig 384x256x256", 384, 256, 256, EPI_MUL_SIG, 0, false, false, false, true);
    test_nt("NT mul_step 384x256x256", 384, 256, 256, EPI_MUL_STEP, 0, false, false, false, false);
    test_nt("NT tangent 384x256x256", 384, 256, 256, EPI_TANGENT, 0, false, false, false, true);
    test_nt("NT adjoint 384x256x256", 384, 256, 256, EPI_ADJOINT, 0, false, false, false, true);
    test_nt("NT adjoint ragged 100x40x8", 100, 40, 8, EPI_ADJOINT, 0, true, false, false, true);
    test_nt("NT linear K=4 (tiny) 256x256x4", 256, 256, 4, EPI_LINEAR, 0, true, false, true, false);
    g_persist = 1;
    test_nt("P NT linear 1000x256x256", 1000, 256, 256, EPI_LINEAR, 0, true, false, false, true);
    test_nt("P NT softplus grp+rank1 20000x256x64", 20000, 256, 64, EPI_SOFTPLUS, 0, true, true, true, true);
    test_nt("P NT mul_sig 20000x256x256", 20000, 256, 256, EPI_MUL_SIG, 0, false, false, false, true);
    test_nt("P NT mul_step 3000x200x96", 3000, 200, 96, EPI_MUL_STEP, 0, false, false, false, true);
    test_nt("P NT tangent 20000x256x256", 20000, 256, 256, EPI_TANGENT, 0, false, false, false, true);
    test_nt("P NT adjoint 20000x256x256", 20000, 256, 256, EPI_ADJOINT, 0, true, false, false, true);
    test_nt("P NT adjoint ragged 20001x200x40", 20001, 200, 40, EPI_ADJOINT, 0, false, false, false, true);
    g_split = 1;
    test_nt("P NT softplus split 20000x256x256", 20000, 256, 256, EPI_SOFTPLUS, 0, true, false, false, true);
    g_persist = -1;
    test_nt("NT softplus split 500x256x256", 500, 256, 256, EPI_SOFTPLUS, 0, true, false, false, true);
    g_split = 0; g_persist = 0;
    g_mc = 1;
    test_nt("MC NT linear 1000x256x256", 1000, 256, 256, EPI_LINEAR, 0, true, false, false, true);
    test_nt("MC NT tangent 20000x256x256", 20000, 256, 256, EPI_TANGENT, 0, false, false, false, true);
    test_nt("MC NT adjoint ragged 2945x200x40", 2945, 200, 40, EPI_ADJOINT, 0, true, false, false, true);
    g_split = 1;
    test_nt("MC NT softplus split 20000x256x768", 20000, 256, 768, EPI_SOFTPLUS, 0, true, false, false, true);
    g_split = 0; g_mc = -1;
    g_pair = 1;
    test_nt("CG2 NT linear 1000x256x256", 1000, 256, 256, EPI_LINEAR, 0, true, false, false, true);
    test_nt("CG2 NT tangent 20000x256x256", 20000, 256, 256, EPI_TANGENT, 0, false, false, false, true);
    test_nt("CG2 NT adjoint ragged 2945x200x40", 2945, 200, 40, EPI_ADJOINT, 0, true, false, false, true);
    g_split = 1;
    test_nt("CG2 NT softplus split 20000x256x768", 20000, 256, 768, EPI_SOFTPLUS, 0, true, false, false, true);
    g_split = 0; g_pair = -1;
    g_deep = 0;  // auto: small grids take the deep-pipeline narrow-tile variant
    test_nt("DEEP NT softplus 512x300x784", 512, 300, 784, EPI_SOFTPLUS, 0, true, true, true, true);
    test_nt("DEEP NT tangent 512x256x256", 512, 256, 256, EPI_TANGENT, 0, false, false, false, true);
    test_nt("DEEP NT adjoint 130x20x300", 130, 20, 300, EPI_ADJOINT, 0, true, false, false, true);
    test_nt("DEEP NT mul_step 512x784x300", 512, 784, 300, EPI_MUL_STEP, 0, false, false, false, true);
    g_split = 1;
    test_nt("DEEP NT softplus split 512x256x768", 512, 256, 768, EPI_SOFTPLUS, 0, true, false, false, true);
    g_split = 0; g_deep = -1;
    test_tn("TN two pairs 256x256 K=4096", 256, 256, 4096, true);
    test_tn("TN ragged 300x100 K=500", 300, 100, 500, true);
    test_tn("TN 256x36 K=2048", 256, 36, 2048, false);
    test_tn("TN 784x300 K=512", 784, 300, 512, false);
    for (int mode : {EPI_LINEAR, EPI_SOFTPLUS, EPI_MUL_SIG, EPI_TANGENT, EPI_ADJOINT}) {
      bench_nt(131072, 256, 256, mode, -1);
      bench_nt(131072, 256, 256, mode, 1);
    }
    bench_nt(131072, 256, 768, EPI_SOFTPLUS, -1, 1);
    bench_nt(131072, 256, 768, EPI_SOFTPLUS, 1, 1);
    bench_nt(131072, 256, 768, EPI_SOFTPLUS, -1, 1, 1);
    bench_nt(131072, 256, 768, EPI_SOFTPLUS, -1, 1, -1, -1, 1);
    for (int mode : {EPI_LINEAR, EPI_MUL_SIG, EPI_TANGENT, EPI_ADJOINT}) bench_nt(131072, 256, 256, mode, -1, 0, -1, -1, 1);
    for (int deep : {-1, 1}) {
      printf("small-M chain shapes, deep=%d\n", deep);
      bench_nt(512, 256, 768, EPI_SOFTPLUS, -1, 1, -1, deep);
      bench_nt(512, 256, 256, EPI_MUL_SIG, -1, 0, -1, deep);
      bench_nt(512, 300, 960, EPI_SOFTPLUS, -1, 1, -1, deep);
      bench_nt(512, 784, 960, EPI_LINEAR, -1, 0, -1, deep);
      bench_nt(512, 300, 784, EPI_MUL_SIG, -1, 0, -1, deep);
    }
    for (int mode : {EPI_LINEAR, EPI_MUL_SIG, EPI_TANGENT, EPI_ADJOINT}) bench_nt(131072, 256, 256, mode, -1, 0, 1);
    bench_nt(131072, 256, 32, EPI_SOFTPLUS);
    bench_nt(131072, 32, 256, EPI_LINEAR);
    bench_tn(256, 256, 131072, false);
    bench_tn(256, 256, 131072, true);
  }
  printf(g_fail ? "SELFTEST FAILED (%d)\n" : "SELFTEST PASSED\n", g_fail);
  return g_fail ? 1 : 0;
}
