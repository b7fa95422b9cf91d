// extern "C" surface of libardae.so (declared in include/ardae.h).
#include "../../include/ardae.h"

#include "cdae.cuh"

using namespace ardae;

struct ardae_cdae_s {
  CdaePlan p;
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
  o->B = c->batch; o->S = c->samples; o->train = c->train;
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
  h->p.plan = Plan();
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

ARDAE_API int ardae_randn(float* out, size_t n, uint64_t seed, uint32_t stream_id, void* stream) {
  if (!out) return fail(-1, "null argument");
  if (n == 0) return 0;
  randn_kernel<<<grid_for((n + 3) / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(out, n, seed, stream_id);
  ARDAE_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
