// extern "C" surface of libardae.so (declared in include/ardae.h).
#include "../../include/ardae.h"

#include <cmath>

#include "model.cuh"
#include "dp_fused.cuh"

using namespace ardae;

struct ardae_cdae_s {
  CdaePlan p;
};
struct ardae_model_s {
  ModelPlan p;
  const float* beta_dev = nullptr;
};

extern "C" {

ARDAE_API int ardae_version(void) { return ARDAE_VERSION; }
ARDAE_API const char* ardae_last_error(void) { return last_error_string().c_str(); }

ARDAE_API int ardae_check_device(int dev) {
  cudaDeviceProp prop;
  ARDAE_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) return fail(-20, std::string("device is not sm_100: ") + prop.name);
  if (get_encode_fn() == nullptr) return fail(-10, "cuTensorMapEncodeTiled entry point unavailable");
  return 0;
}

static void to_cfg(const ardae_cdae_config* c, CdaeConfig* o) {
  o->d = c->input_dim; o->c = c->context_dim; o->H = c->h_dim; o->L = c->num_hidden_layers;
  o->B = c->batch; o->S = c->samples; o->train = c->train; o->kind = c->kind;
}

ARDAE_API int ardae_cdae_workspace_bytes(const ardae_cdae_config* cfg, size_t* bytes) {
  if (!cfg || !bytes) return fail(-1, "null argument");
  CdaePlan p;
  to_cfg(cfg, &p.cfg);
  p.ws.dry = true;
  int rc = p.build(nullptr, nullptr);
  if (rc) return rc;
  *bytes = p.ws.off + 256;
  return 0;
}

ARDAE_API int ardae_cdae_create(const ardae_cdae_config* cfg, float* const* params, float* const* grads,
                      int num_tensors, void* workspace, size_t workspace_bytes, ardae_cdae_t* out) {
  if (!cfg || !params || !workspace || !out) return fail(-1, "null argument");
  if (cfg->train && !grads) return fail(-1, "train plan needs grads");
  std::unique_ptr<ardae_cdae_s> h(new ardae_cdae_s());
  to_cfg(cfg, &h->p.cfg);
  if (num_tensors != h->p.ntensors()) return fail(-2, "cdae: expected 6L+2 parameter tensors");
  h->p.ws.dry = true;
  int rc = h->p.build(nullptr, nullptr);  // measures (and fixes the split-K workspace need)
  if (rc) return rc;
  if (h->p.ws.off + 256 > workspace_bytes) return fail(-3, "cdae: workspace too small");
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return fail(-11, "workspace must be 256-byte aligned");
  for (int i = 0; i < num_tensors; ++i)
    if (reinterpret_cast<uintptr_t>(params[i]) & 15) return fail(-11, "parameter tensors must be 16-byte aligned");
  h->p.plan.reset();
  h->p.ws = Workspace();
  h->p.ws.dry = false;
  h->p.ws.base = static_cast<uint8_t*>(workspace);
  h->p.ws.size = workspace_bytes;
  // pad columns of the tf32-pair buffers are never written: they must start (and stay) zero
  ARDAE_CUDA_OK(cudaMemset(workspace, 0, workspace_bytes));
  rc = h->p.build(params, grads);
  if (rc) return rc;
  *out = h.release();
  return 0;
}

ARDAE_API void ardae_cdae_destroy(ardae_cdae_t h) { delete h; }
ARDAE_API int ardae_cdae_num_launches(ardae_cdae_t h) { return h ? h->p.plan.launches() + 2 : 0; }

ARDAE_API int ardae_cdae_set_profile(ardae_cdae_t h, int on) {
  if (!h) return fail(-1, "null argument");
  h->p.plan.profile = on != 0;
  return 0;
}
ARDAE_API int ardae_cdae_read_profile(ardae_cdae_t h, int max_ops, char* tags, float* ms, int* num_ops) {
  if (!h || !tags || !ms || !num_ops || max_ops <= 0) return fail(-1, "null argument");
  std::vector<const char*> t(max_ops);
  const int n = h->p.plan.read_profile(max_ops, t.data(), ms);
  for (int i = 0; i < n; ++i) {
    std::strncpy(tags + 16 * i, t[i], 15);
    tags[16 * i + 15] = 0;
  }
  *num_ops = n;
  return 0;
}

ARDAE_API int ardae_cdae_train(ardae_cdae_t h, const float* x, const float* ctx, const float* sigma,
                     float* eps, int gen_eps, uint64_t seed, float inv_count, float* loss_out,
                     float* score_out, void* stream) {
  if (!h || !x || !ctx || !sigma || !eps || !loss_out) return fail(-1, "null argument");
  if (!h->p.cfg.train) return fail(-2, "handle was created with train = 0");
  CdaeBindings& b = h->p.bind;
  b.x = x; b.ctx = ctx; b.sigma = sigma; b.eps = eps; b.gen_eps = gen_eps; b.seed = seed;
  b.inv_count = inv_count; b.loss_out = loss_out; b.score_out = score_out;
  return h->p.plan.run(static_cast<cudaStream_t>(stream));
}

ARDAE_API int ardae_cdae_score(ardae_cdae_t h, const float* x, const float* ctx, const float* sigma,
                     float* score_out, void* stream) {
  if (!h || !x || !ctx || !sigma || !score_out) return fail(-1, "null argument");
  if (h->p.cfg.train) return fail(-2, "handle was created with train = 1");
  CdaeBindings& b = h->p.bind;
  b = CdaeBindings();
  b.x = x; b.ctx = ctx; b.sigma = sigma; b.score_out = score_out;
  return h->p.plan.run(static_cast<cudaStream_t>(stream));
}

static void to_mcfg(const ardae_model_config* c, ModelConfig* o) {
  o->kind = c->kind; o->D = c->input_dim; o->n = c->noise_dim; o->h = c->h_dim; o->zd = c->z_dim;
  o->n_inp = c->n_inp; o->n_fc = c->n_fc; o->n_dec = c->n_dec; o->act = c->act; o->B = c->batch;
  o->nz = c->nz; o->mode = c->mode; o->img_h = c->img_h; o->img_c = c->img_c;
}

ARDAE_API int ardae_model_workspace_bytes(const ardae_model_config* cfg, size_t* bytes) {
  if (!cfg || !bytes) return fail(-1, "null argument");
  ModelPlan p;
  to_mcfg(cfg, &p.cfg);
  p.ws.dry = true;
  int rc = p.build(nullptr, nullptr);
  if (rc) return rc;
  *bytes = p.ws.off + 256;
  return 0;
}

ARDAE_API int ardae_model_create(const ardae_model_config* cfg, float* const* params, float* const* grads,
                                 int num_tensors, void* workspace, size_t workspace_bytes, ardae_model_t* out) {
  if (!cfg || !params || !workspace || !out) return fail(-1, "null argument");
  if (cfg->mode == 1 && !grads) return fail(-1, "forward+backward plan needs grads");
  std::unique_ptr<ardae_model_s> h(new ardae_model_s());
  to_mcfg(cfg, &h->p.cfg);
  if (num_tensors != h->p.ntensors()) return fail(-2, "model: unexpected number of parameter tensors");
  h->p.ws.dry = true;
  int rc = h->p.build(nullptr, nullptr);
  if (rc) return rc;
  if (h->p.ws.off + 256 > workspace_bytes) return fail(-3, "model: workspace too small");
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return fail(-11, "workspace must be 256-byte aligned");
  h->p.fwd.reset(); h->p.bwd_dec.reset(); h->p.bwd_enc.reset(); h->p.fwd_mean.reset();
  h->p.ws = Workspace();
  h->p.ws.dry = false;
  h->p.ws.base = static_cast<uint8_t*>(workspace);
  h->p.ws.size = workspace_bytes;
  ARDAE_CUDA_OK(cudaMemset(workspace, 0, workspace_bytes));
  rc = h->p.build(params, grads);
  if (rc) return rc;
  *out = h.release();
  return 0;
}

ARDAE_API void ardae_model_destroy(ardae_model_t h) { delete h; }
ARDAE_API int ardae_model_num_launches(ardae_model_t h, int which) {
  if (!h) return 0;
  if (which == 2) return h->p.fwd_mean.launches();
  return which == 0 ? h->p.fwd.launches() + 4 : h->p.bwd_dec.launches() + h->p.bwd_enc.launches();
}

ARDAE_API int ardae_model_encode(ardae_model_t h, const float* x, const float* noise, float* z_out, void* stream) {
  if (!h || !x || !z_out) return fail(-1, "null argument");
  if (h->p.cfg.mode != 0) return fail(-2, "handle was created with mode = 1 (use ardae_model_forward)");
  ModelBindings& b = h->p.bind;
  b = ModelBindings();
  b.x = x; b.noise = noise; b.z_out = z_out;
  return h->p.fwd.run(static_cast<cudaStream_t>(stream));
}

ARDAE_API int ardae_model_encode_with_mean(ardae_model_t h, const float* x, const float* noise, float* z_out,
                                           float* zbar_out, void* stream) {
  if (!h || !x || !z_out || !zbar_out) return fail(-1, "null argument");
  if (h->p.cfg.mode != 0) return fail(-2, "handle was created with mode = 1 (use ardae_model_forward)");
  ModelBindings& b = h->p.bind;
  b = ModelBindings();
  b.x = x; b.noise = noise; b.z_out = z_out; b.zbar_out = zbar_out;
  int rc = h->p.fwd.run(static_cast<cudaStream_t>(stream));
  if (rc) return rc;
  return h->p.fwd_mean.run(static_cast<cudaStream_t>(stream));
}

ARDAE_API int ardae_model_encode_hidden(ardae_model_t h, const float* x, const float* noise, float* z_out,
                                        float* zbar_out, float* hidden_out, void* stream) {
  if (!h || !x || !z_out || !zbar_out || !hidden_out) return fail(-1, "null argument");
  if (h->p.cfg.mode != 0) return fail(-2, "handle was created with mode = 1 (use ardae_model_forward)");
  if (h->p.cfg.kind != 3) return fail(-2, "encode_hidden: only the auxmnist kind has a hidden context");
  ModelBindings& b = h->p.bind;
  b = ModelBindings();
  b.x = x; b.noise = noise; b.z_out = z_out; b.zbar_out = zbar_out; b.hidden_out = hidden_out;
  int rc = h->p.fwd.run(static_cast<cudaStream_t>(stream));
  if (rc) return rc;
  return h->p.fwd_mean.run(static_cast<cudaStream_t>(stream));
}

ARDAE_API int ardae_model_forward(ardae_model_t h, const float* x, const float* noise, float beta, float inv_rows,
                                  float* z_out, float* sums, float* heads_out, void* stream) {
  if (!h || !x || !z_out || !sums) return fail(-1, "null argument");
  if (h->p.cfg.mode != 1) return fail(-2, "handle was created with mode = 0");
  ModelBindings& b = h->p.bind;
  b = ModelBindings();
  b.x = x; b.noise = noise; b.z_out = z_out; b.sums = sums; b.heads_out = heads_out; b.beta = beta;
  b.beta_dev = h->beta_dev;
  b.inv_rows = inv_rows;
  return h->p.fwd.run(static_cast<cudaStream_t>(stream));
}

ARDAE_API int ardae_model_decode(ardae_model_t h, const float* z, float* heads_out, void* stream) {
  if (!h || !z || !heads_out) return fail(-1, "null argument");
  if (h->p.cfg.mode != 3) return fail(-2, "handle was not created with mode = 3 (decode)");
  ModelBindings& b = h->p.bind;
  b = ModelBindings();
  b.z_in = z; b.heads_out = heads_out;
  return h->p.fwd.run(static_cast<cudaStream_t>(stream));
}

ARDAE_API int ardae_model_forward_inp(ardae_model_t h, const float* x, float* inp_out, void* stream) {
  if (!h || !x || !inp_out) return fail(-1, "null argument");
  if (h->p.cfg.mode != 4) return fail(-2, "handle was not created with mode = 4 (_forward_inp)");
  ModelBindings& b = h->p.bind;
  b = ModelBindings();
  b.x = x; b.inp_out = inp_out;
  return h->p.fwd.run(static_cast<cudaStream_t>(stream));
}

ARDAE_API int ardae_model_forward_all(ardae_model_t h, const float* inp, const float* noise, float* z_out, void* stream) {
  if (!h || !inp || !z_out) return fail(-1, "null argument");
  if (h->p.cfg.mode != 5) return fail(-2, "handle was not created with mode = 5 (_forward_all)");
  ModelBindings& b = h->p.bind;
  b = ModelBindings();
  b.inp_in = inp; b.noise = noise; b.z_out = z_out;
  return h->p.fwd.run(static_cast<cudaStream_t>(stream));
}

ARDAE_API int ardae_model_set_beta_device(ardae_model_t h, const float* beta_device) {
  if (!h) return fail(-1, "null argument");
  h->beta_dev = beta_device;
  h->p.bind.beta_dev = beta_device;
  return 0;
}

ARDAE_API int ardae_model_iws(ardae_model_t h, const float* x, const float* noise, const float* eta, uint64_t seed,
                              float* out, float* total, int* status, void* stream) {
  if (!h || !x || !noise || !out) return fail(-1, "null argument");
  if (h->p.cfg.mode != 2) return fail(-2, "handle was not created with mode = 2 (IWS)");
  ModelBindings& b = h->p.bind;
  b = ModelBindings();
  b.x = x; b.noise = noise; b.eta = eta; b.seed = seed; b.iws_out = out; b.iws_total = total; b.status = status;
  return h->p.fwd.run(static_cast<cudaStream_t>(stream));
}

// The two halves of ardae_model_backward as separate calls: the decoder half depends only on the ELBO forward, so the
// fused driver runs it on its side stream underneath the CDAE update; the encoder half needs the entropy gradient.
ARDAE_API int ardae_model_backward_decoder(ardae_model_t h, float loss_scale, void* stream) {
  if (!h) return fail(-1, "null argument");
  if (h->p.cfg.mode != 1) return fail(-2, "handle was created with mode = 0");
  if (loss_scale == 0.0f) return 0;
  h->p.bind.loss_scale = loss_scale;
  return h->p.bwd_dec.run(static_cast<cudaStream_t>(stream));
}

ARDAE_API int ardae_model_backward_encoder(ardae_model_t h, float loss_scale, const float* gz, float gz_scale,
                                           void* stream) {
  if (!h) return fail(-1, "null argument");
  if (h->p.cfg.mode != 1) return fail(-2, "handle was created with mode = 0");
  ModelBindings& b = h->p.bind;
  b.loss_scale = loss_scale; b.gz = gz; b.gz_scale = gz_scale;
  return h->p.bwd_enc.run(static_cast<cudaStream_t>(stream));
}

ARDAE_API int ardae_model_backward(ardae_model_t h, float loss_scale, const float* gz, float gz_scale, void* stream) {
  if (!h) return fail(-1, "null argument");
  if (h->p.cfg.mode != 1) return fail(-2, "handle was created with mode = 0");
  ModelBindings& b = h->p.bind;
  b.loss_scale = loss_scale; b.gz = gz; b.gz_scale = gz_scale;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (loss_scale != 0.0f) {
    int rc = h->p.bwd_dec.run(s);
    if (rc) return rc;
  }
  return h->p.bwd_enc.run(s);
}

ARDAE_API int ardae_sigma_schedule(const float* z, const float* zbar, int B, int nz, int d, int nstd, float S,
                                   float delta, const float* xi, uint64_t seed, float* x_out, float* sigma_out,
                                   float* std_out, void* stream) {
  if (!z || !zbar || !x_out || !sigma_out) return fail(-1, "null argument");
  if (B <= 0 || nz < 2 || d <= 0 || nstd <= 0) return fail(-2, "sigma_schedule: need B>0, nz>=2, d>0, nstd>0");
  sigma_schedule_kernel<<<B, 256, sizeof(float) * (4 + 256 + d), static_cast<cudaStream_t>(stream)>>>(
      z, zbar, nz, d, nstd, S, delta, xi, seed, x_out, sigma_out, std_out, replay_counter());
  ARDAE_CUDA_OK(cudaGetLastError());
  return 0;
}

ARDAE_API int ardae_scaled_diff(const float* z, const float* zbar, int R, int nz, int d, float S, float* out,
                                void* stream) {
  if (!z || !zbar || !out) return fail(-1, "null argument");
  scaled_diff_kernel<<<grid_for(static_cast<size_t>(R) * d), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      z, zbar, R, nz, d, S, out);
  ARDAE_CUDA_OK(cudaGetLastError());
  return 0;
}

ARDAE_API int ardae_adam_step(float* p, const float* g, float* exp_avg, float* exp_avg_sq, size_t n, float lr,
                              float beta1, float beta2, float eps, int step, float gscale, void* stream) {
  if (!p || !g || !exp_avg || !exp_avg_sq) return fail(-1, "null argument");
  if (n % 4 != 0 || ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                      reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15))
    return fail(-11, "optimizer arenas must be 16-byte aligned with n % 4 == 0");
  if (step < 1) return fail(-2, "adam: step must be >= 1");
  adam_kernel<<<grid_for(n / 4, 256, 148 * 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      p, g, exp_avg, exp_avg_sq, n, lr, step, beta1, beta2, eps, gscale, replay_counter());
  ARDAE_CUDA_OK(cudaGetLastError());
  return 0;
}

ARDAE_API int ardae_rmsprop_step(float* p, const float* g, float* square_avg, float* momentum_buf, size_t n,
                                 float lr, float alpha, float eps, float momentum, float gscale, void* stream) {
  if (!p || !g || !square_avg || !momentum_buf) return fail(-1, "null argument");
  if (n % 4 != 0 || ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                      reinterpret_cast<uintptr_t>(square_avg) | reinterpret_cast<uintptr_t>(momentum_buf)) & 15))
    return fail(-11, "optimizer arenas must be 16-byte aligned with n % 4 == 0");
  rmsprop_kernel<<<grid_for(n / 4, 256, 148 * 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      p, g, square_avg, momentum_buf, n, lr, alpha, eps, momentum, gscale);
  ARDAE_CUDA_OK(cudaGetLastError());
  return 0;
}

ARDAE_API int ardae_randn(float* out, size_t n, uint64_t seed, uint32_t stream_id, void* stream) {
  if (!out) return fail(-1, "null argument");
  if (n == 0) return 0;
  randn_kernel<<<grid_for((n + 3) / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(out, n, seed, stream_id,
                                                                                      replay_counter());
  ARDAE_CUDA_OK(cudaGetLastError());
  return 0;
}

ARDAE_API int ardae_bernoulli(const float* probs, float* out, size_t n, uint64_t seed, void* stream) {
  if (!probs || !out) return fail(-1, "null argument");
  if (n == 0) return 0;
  bernoulli_kernel<<<grid_for((n + 3) / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(probs, out, n, seed,
                                                                                          replay_counter());
  ARDAE_CUDA_OK(cudaGetLastError());
  return 0;
}

ARDAE_API int ardae_set_replay_counter(const unsigned long long* device_counter) {
  replay_counter() = device_counter;
  return 0;
}

ARDAE_API int ardae_bump_replay_counter(unsigned long long* device_counter, void* stream) {
  if (!device_counter) return fail(-1, "null argument");
  bump_counter_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(device_counter);
  ARDAE_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"

// ---------------------------------------------------------------- data parallel: peer memory + fused exchange / update
ARDAE_API int ardae_ipc_export(const void* ptr, unsigned char* handle64, size_t* offset) {
  if (!ptr || !handle64 || !offset) return fail(-1, "null argument");
  typedef CUresult (*PFN_range)(CUdeviceptr*, size_t*, CUdeviceptr);
  static PFN_range fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_range>(p);
  }
  if (fn == nullptr) return fail(-10, "cuMemGetAddressRange entry point unavailable");
  CUdeviceptr base = 0;
  size_t size = 0;
  if (fn(&base, &size, reinterpret_cast<CUdeviceptr>(ptr)) != CUDA_SUCCESS) return fail(-2, "ipc_export: not a device allocation");
  cudaIpcMemHandle_t h;
  ARDAE_CUDA_OK(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base)));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  std::memcpy(handle64, &h, 64);
  *offset = reinterpret_cast<uintptr_t>(ptr) - static_cast<uintptr_t>(base);
  return 0;
}

ARDAE_API int ardae_ipc_import(const unsigned char* handle64, size_t offset, void** out) {
  if (!handle64 || !out) return fail(-1, "null argument");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, 64);
  void* base = nullptr;
  ARDAE_CUDA_OK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
  *out = static_cast<unsigned char*>(base) + offset;
  return 0;
}

ARDAE_API int ardae_dp_xchg_bytes(size_t n, int world, size_t* bytes) {
  if (!bytes || world < 1 || world > kDpMaxWorld || n % 4 != 0) return fail(-2, "dp_xchg_bytes: bad arguments");
  *bytes = dp_xchg_bytes(n, world);
  return 0;
}

ARDAE_API int ardae_dp_fused_step(int kind, int rank, int world, float* p, const float* g, float* s1, float* s2, size_t n,
                                  void* const* peer_xchg, unsigned long long* epoch_barrier_status, float lr, float beta1,
                                  float beta2_or_alpha, float eps, float momentum, int step, float gscale, void* stream) {
  if (!p || !g || !s1 || !s2 || !peer_xchg || !epoch_barrier_status) return fail(-1, "null argument");
  if (world < 2 || world > kDpMaxWorld || rank < 0 || rank >= world) return fail(-2, "dp_fused_step: bad rank / world");
  if (kind != 0 && kind != 1) return fail(-2, "dp_fused_step: kind must be 0 (Adam) or 1 (RMSprop)");
  if (n % 4 != 0 || ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(s1) |
                      reinterpret_cast<uintptr_t>(s2)) & 15))
    return fail(-11, "optimizer arenas must be 16-byte aligned with n % 4 == 0");
  if (kind == 0 && step < 1) return fail(-2, "adam: step must be >= 1");
  DpFusedParams a;
  std::memset(&a, 0, sizeof(a));
  a.p = p; a.g = g; a.s1 = s1; a.s2 = s2; a.n = n; a.rank = rank; a.world = world; a.kind = kind;
  a.lr = lr; a.b1 = beta1; a.b2 = beta2_or_alpha; a.alpha = beta2_or_alpha; a.eps = eps; a.mu = momentum; a.gscale = gscale;
  a.step0 = step; a.ctr = replay_counter();
  a.epoch = epoch_barrier_status; a.barrier = epoch_barrier_status + 1;
  a.status = reinterpret_cast<int*>(epoch_barrier_status + 2);
  for (int r = 0; r < world; ++r) {
    if (!peer_xchg[r]) return fail(-1, "dp_fused_step: null exchange buffer");
    a.xchg[r] = static_cast<unsigned char*>(peer_xchg[r]);
  }
  dp_fused_step_kernel<<<kDpGrid, kDpBlock, 0, static_cast<cudaStream_t>(stream)>>>(a);
  ARDAE_CUDA_OK(cudaGetLastError());
  return 0;
}
