// Plan infrastructure: row-major matrix views, a bump allocator over the caller's workspace,
// and a recorded list of stream-ordered launches (GEMMs with pre-encoded TMA maps + small kernels).
#pragma once
#include <functional>
#include <memory>
#include <vector>

#include "chain16_host.cuh"
#include "chain_host.cuh"
#include "dec_iws_sm100.cuh"
#include "enc_sample_sm100.cuh"
#include "gemm_host.cuh"
#include "gemm_tn16.cuh"

namespace ardae {

struct Mat {
  float* p = nullptr;
  int rows = 0, cols = 0, ld = 0;
  Mat() {}
  Mat(float* p_, int r, int c, int l) : p(p_), rows(r), cols(c), ld(l) {}
  Mat cols_from(int c0, int n) const { return Mat(p ? p + c0 : nullptr, rows, n, ld); }
  Mat rows_from(int r0, int n) const {
    return Mat(p ? p + static_cast<size_t>(r0) * ld : nullptr, n, cols, ld);
  }
};

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Bump allocator.  In dry mode (base == nullptr) it only measures.
struct Workspace {
  uint8_t* base = nullptr;
  size_t size = 0, off = 0;
  bool dry = true;
  float* floats(size_t n) {
    off = (off + 255) & ~static_cast<size_t>(255);
    float* p = dry ? reinterpret_cast<float*>(static_cast<uintptr_t>(256) + off)
                   : reinterpret_cast<float*>(base + off);
    off += n * sizeof(float);
    return p;
  }
  // bf16 [rows, cols], dense rows (cols % 8 == 0: 16-byte row pitch for TMA)
  Mat16 mat16(int rows, int cols) {
    return Mat16(reinterpret_cast<uint16_t*>(floats((static_cast<size_t>(rows) * cols + 1) / 2)), rows, cols, cols);
  }
  // [rows, cols] with a pitch that satisfies TMA (multiple of 4 floats)
  Mat mat(int rows, int cols) {
    const int ld = round_up(cols, 4);
    return Mat(floats(static_cast<size_t>(rows) * ld), rows, cols, ld);
  }
  bool fits() const { return dry || off <= size; }
};

struct Plan {
  bool dry = true;
  int error = 0;
  // ops carry a lane: 0 = the caller's stream, 1 = a plan-owned side stream used between fork()/join()
  // for the B-row chains (context branch) that are independent of the N-row sweeps.
  std::vector<std::function<int(cudaStream_t)>> ops;
  std::vector<int> lanes;   // per op: 0 main, 1 side, 2 = fork marker, 3 = join marker
  int cur_lane = 0;
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int num_gemm_nt = 0, num_gemm_tn = 0, num_small = 0, num_chain = 0;
  // optional per-op timing (bench.py roofline): CUDA events around every op of the LAST run, on the op's stream
  std::vector<const char*> tags;
  bool profile = false;
  std::vector<cudaEvent_t> ev0, ev1;

  void tag_last(const char* t) {
    if (!dry) tags.resize(ops.size(), "aux"), tags.back() = t;
  }
  // elapsed ms of every op of the last run (synchronises the device); returns the number of ops written
  int read_profile(int max_ops, const char** out_tags, float* out_ms) {
    if (ev0.size() != ops.size()) return 0;
    cudaDeviceSynchronize();
    int n = 0;
    for (size_t i = 0; i < ops.size() && n < max_ops; ++i) {
      if (lanes[i] >= 2) continue;
      float ms = 0.0f;
      if (cudaEventElapsedTime(&ms, ev0[i], ev1[i]) != cudaSuccess) { cudaGetLastError(); continue; }
      out_tags[n] = i < tags.size() ? tags[i] : "aux";
      out_ms[n] = ms;
      ++n;
    }
    return n;
  }

  void fork() {
    cur_lane = 1;
    if (dry) return;
    ops.push_back(nullptr);
    lanes.push_back(2);
  }
  void join() {
    cur_lane = 0;
    if (dry) return;
    ops.push_back(nullptr);
    lanes.push_back(3);
  }

  void fail_with(int rc) {
    if (error == 0) error = rc;
  }
  void add(std::function<int(cudaStream_t)> f) {
    ++num_small;
    if (!dry) {
      ops.push_back(std::move(f));
      lanes.push_back(cur_lane);
    }
  }
  void nt(const GemmNTDesc& d) {
    ++num_gemm_nt;
    if (dry) return;
    PreparedNT pr;
    int rc = prepare_gemm_nt(d, &pr);
    if (rc) return fail_with(rc);
    ops.push_back([pr](cudaStream_t s) { return launch_prepared_nt(pr, s); });
    lanes.push_back(cur_lane);
    tag_last("gemm_nt");
  }
  void chain(const ChainDesc& d) {
    ++num_chain;
    if (dry) return;
    PreparedChain pr;
    int rc = prepare_chain(d, &pr);
    if (rc) return fail_with(rc);
    auto sp = std::make_shared<PreparedChain>(pr);  // ~8 KB of tensor maps: keep one copy
    ops.push_back([sp](cudaStream_t s) { return launch_prepared_chain(*sp, s); });
    lanes.push_back(cur_lane);
    static const char* names[CHAIN_NUM_MODES] = {"chain_mul_sig", "chain_tangent", "chain_adjoint", "chain_softplus3"};
    tag_last(names[d.mode]);
  }
  void chain16(const Chain16Desc& d) {
    ++num_chain;
    if (dry) return;
    PreparedChain16 pr;
    int rc = prepare_chain16(d, &pr);
    if (rc) return fail_with(rc);
    auto sp = std::make_shared<PreparedChain16>(pr);
    ops.push_back([sp](cudaStream_t s) { return launch_prepared_chain16(*sp, s); });
    lanes.push_back(cur_lane);
    static const char* names[CHAIN_NUM_MODES] = {"chain_mul_sig", "chain_tangent", "chain_adjoint", "chain_softplus3"};
    tag_last(names[d.mode]);
  }
  void chain_s3h(const Chain16Desc& d) {
    ++num_chain;
    if (dry) return;
    PreparedChainS3h pr;
    int rc = prepare_chain_s3h(d, &pr);
    if (rc) return fail_with(rc);
    auto sp = std::make_shared<PreparedChainS3h>(pr);
    ops.push_back([sp](cudaStream_t s) { return launch_prepared_chain_s3h(*sp, s); });
    lanes.push_back(cur_lane);
    tag_last("chain_softplus3");
  }
  void tn16(const GemmTN16Desc& d) {
    ++num_gemm_tn;
    if (dry) return;
    PreparedTN16 pr;
    int rc = prepare_gemm_tn16(d, &pr);
    if (rc) return fail_with(rc);
    auto sp = std::make_shared<PreparedTN16>(pr);
    ops.push_back([sp](cudaStream_t s) { return launch_prepared_tn16(*sp, s); });
    lanes.push_back(cur_lane);
    tag_last("gemm_tn16");
  }
  // same-shape bf16 contractions as one launch where possible, else one launch each
  void tn16_multi(const std::vector<GemmTN16Desc>& ds) {
    if (dry || !tn16_multi_ok(ds)) {
      for (const GemmTN16Desc& d : ds) tn16(d);
      return;
    }
    ++num_gemm_tn;
    PreparedTN16Multi pr;
    int rc = prepare_gemm_tn16_multi(ds, &pr);
    if (rc) return fail_with(rc);
    auto sp = std::make_shared<PreparedTN16Multi>(pr);
    ops.push_back([sp](cudaStream_t s) { return launch_prepared_tn16_multi(*sp, s); });
    lanes.push_back(cur_lane);
    tag_last("gemm_tn16_multi");
  }
  void tn(const GemmTNDesc& d) {
    ++num_gemm_tn;
    if (dry) return;
    PreparedTN pr;
    int rc = prepare_gemm_tn(d, &pr);
    if (rc) return fail_with(rc);
    ops.push_back([pr](cudaStream_t s) { return launch_prepared_tn(pr, s); });
    lanes.push_back(cur_lane);
    tag_last("gemm_tn");
  }
  // same-shape weight-gradient contractions as one launch where possible, else one launch each
  void tn_batch(const std::vector<GemmTNDesc>& ds) {
    bool ok = ds.size() >= 2 && ds.size() <= static_cast<size_t>(kMaxTNBatch);
    for (const GemmTNDesc& d : ds) ok = ok && !dry && tn_batchable(d);
    if (ok && !dry) {
      for (size_t i = 1; i < ds.size(); ++i)
        ok = ok && ds[i].M == ds[0].M && ds[i].N == ds[0].N && ds[i].K == ds[0].K &&
             (ds[i].X1 != nullptr) == (ds[0].X1 != nullptr);
    }
    if (!ok) {
      for (const GemmTNDesc& d : ds) tn(d);
      return;
    }
    ++num_gemm_tn;  // one launch
    PreparedTNBatch pb;
    int rc = prepare_gemm_tn_batch(ds, &pb);
    if (rc) return fail_with(rc);
    auto sp = std::make_shared<PreparedTNBatch>(pb);
    ops.push_back([sp](cudaStream_t s) { return launch_prepared_tn_batch(*sp, s); });
    lanes.push_back(cur_lane);
    tag_last("gemm_tn_batch");
  }
  int run(cudaStream_t s) {
    if (profile && ev0.size() != ops.size()) {
      ev0.resize(ops.size()); ev1.resize(ops.size());
      for (size_t i = 0; i < ops.size(); ++i) {
        ARDAE_CUDA_OK(cudaEventCreate(&ev0[i]));
        ARDAE_CUDA_OK(cudaEventCreate(&ev1[i]));
      }
    }
    for (size_t i = 0; i < ops.size(); ++i) {
      const int lane = lanes[i];
      if (lane == 2) {  // fork: the side stream picks up after everything issued so far
        if (side == nullptr) {
          ARDAE_CUDA_OK(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
          ARDAE_CUDA_OK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
          ARDAE_CUDA_OK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
        }
        ARDAE_CUDA_OK(cudaEventRecord(ev_fork, s));
        ARDAE_CUDA_OK(cudaStreamWaitEvent(side, ev_fork, 0));
        continue;
      }
      if (lane == 3) {  // join: the main stream waits for the side lane
        ARDAE_CUDA_OK(cudaEventRecord(ev_join, side));
        ARDAE_CUDA_OK(cudaStreamWaitEvent(s, ev_join, 0));
        continue;
      }
      cudaStream_t st = lane == 1 ? side : s;
      if (profile) ARDAE_CUDA_OK(cudaEventRecord(ev0[i], st));
      int rc = ops[i](st);
      if (rc) return rc;
      if (profile) ARDAE_CUDA_OK(cudaEventRecord(ev1[i], st));
    }
    return 0;
  }
  ~Plan() {
    if (side) cudaStreamDestroy(side);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    for (cudaEvent_t e : ev0) cudaEventDestroy(e);
    for (cudaEvent_t e : ev1) cudaEventDestroy(e);
  }
  Plan() = default;
  Plan(const Plan&) = delete;
  Plan& operator=(const Plan&) = delete;
  void reset() {
    ops.clear(); lanes.clear(); tags.clear(); cur_lane = 0; error = 0;
    num_gemm_nt = num_gemm_tn = num_small = num_chain = 0;
  }
  int launches() const { return num_gemm_nt + 2 * num_gemm_tn + num_small + num_chain; }
};

// An activation stored as a tf32 pair: buf[rows, 2*kp] with hi in columns [0,w) and
// lo = rna(x - hi) in columns [kp, kp+w); kp = w rounded up to the 32-column k-block, pad
// columns are zero (the workspace is zeroed at plan creation and pads are never written).
// `split` storage (fused-chain plans): hi and lo are separate dense [rows, w] arrays -- a chain only streams the hi
// part as aux / spill, and a dense 1 KB row pitch keeps those streams on full DRAM pages (inside the interleaved
// pair buffer every other KB of a row is the unused lo half); lo exists only where a chain restarts from the pair.
struct Pair {
  Mat buf;
  int w = 0, kp = 0;
  bool split = false;
  Mat hi_m, lo_m;
  Mat hi() const { return split ? hi_m : buf.cols_from(0, w); }
  Mat lo() const { return split ? lo_m : buf.cols_from(kp, w); }
  Mat a3() const { return Mat(buf.p, buf.rows, 3 * kp, buf.ld); }  // [hi | lo | hi] via a_k_wrap (interleaved storage only)
};
inline Pair make_pair_split(Workspace& ws, int rows, int w, bool need_lo) {
  Pair p;
  p.w = w;
  p.kp = round_up(w, 32);
  p.split = true;
  p.hi_m = ws.mat(rows, w);
  if (need_lo) p.lo_m = ws.mat(rows, w);
  return p;
}
inline Pair make_pair(Workspace& ws, int rows, int w) {
  Pair p;
  p.w = w;
  p.kp = round_up(w, 32);
  p.buf = Mat(ws.floats(static_cast<size_t>(rows) * 2 * p.kp), rows, 2 * p.kp, 2 * p.kp);
  return p;
}
// A weight in the layouts the kernels consume.
struct W3 {
  Mat16 h16;   // fp16 [out, 2*k16] = [W_hi | W_lo] of w * 2^4 (fp16-pipe primal sweep), k16 = in rounded up to 64
  int k16 = 0;
  Mat b3;  // [out, 3*kp]  = [W_hi | W_hi | W_lo]  (3xTF32 B operand of the forward GEMMs)
  Mat T;   // [in, out]    tf32-rounded transpose   (B operand of the backward-data GEMMs)
  int in = 0, out = 0, kp = 0;
  Mat hi() const { return b3.cols_from(0, in); }  // [out, in] tf32-rounded (tangent sweep)
};

// Out[M,N] = epi(A . W^T + bias)   A:[M,K]  W:[N,K]
inline GemmNTDesc nt_desc(const Mat& A, const Mat& W, const Mat& out, int mode) {
  GemmNTDesc d;
  d.A = A.p; d.lda = A.ld; d.B = W.p; d.ldb = W.ld; d.out = out.p; d.ldo = out.ld;
  d.M = A.rows; d.K = A.cols; d.N = W.rows; d.mode = mode;
  return d;
}
inline void set_aux1(GemmNTDesc& d, const Mat& m) { d.aux1 = m.p; d.ld1 = m.ld; }
inline void set_aux2(GemmNTDesc& d, const Mat& m) { d.aux2 = m.p; d.ld2 = m.ld; }
inline void set_out2(GemmNTDesc& d, const Mat& m) { d.out2 = m.p; d.ldo2 = m.ld; }
// fp32-accurate forward layer on the tf32 pipe: out(pair) = act(A(pair) . W^T + ...)
inline GemmNTDesc nt3_desc(const Pair& A, const W3& W, const Pair& out, int mode) {
  GemmNTDesc d = nt_desc(A.a3(), W.b3, out.hi(), mode);
  d.a_k_wrap = 2 * A.kp;
  d.split_out = 1;
  set_out2(d, out.lo());
  return d;
}
// same, plain fp32 output (no split)
inline GemmNTDesc nt3_desc_plain(const Pair& A, const W3& W, const Mat& out, int mode) {
  GemmNTDesc d = nt_desc(A.a3(), W.b3, out, mode);
  d.a_k_wrap = 2 * A.kp;
  d.round_out = 0;
  return d;
}

}  // namespace ardae
