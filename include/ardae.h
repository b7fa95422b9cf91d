/* libardae -- C ABI of the B200-native AR-DAE hot path.
 *
 * The reference (lim0606/pytorch-ardae-vae) has no FFI: its path sits behind Python nn.Module /
 * Optimizer objects.  This library is what a reference-side binding loads in place of
 * the PyTorch library calls those objects make; each entry point names the reference interface it
 * replaces (file:line, relative to the reference root).  INTEGRATION.md shows the ctypes stubs.
 *
 * Conventions
 *  - plain C: device pointers + sizes, no torch types.  All tensors fp32, row-major, contiguous
 *    unless a pitch is given.  The library never allocates or frees caller tensors; plans carve
 *    their scratch out of a caller-provided workspace (query the size first).
 *  - every call is stream-ordered and asynchronous on the cudaStream_t passed (as void*).
 *  - return value: 0 = ok, negative = invalid argument / unsupported shape, positive = CUDA error
 *    code; ardae_last_error() returns the message of the last failure on the calling thread.
 *  - handles are not thread-safe; calls sharing a handle must be serialized by the caller.
 *  - sm_100a only.  There is no CPU fallback: without a B200 every compute entry point fails.
 */
#ifndef ARDAE_H_
#define ARDAE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ARDAE_VERSION 100
#if defined(__GNUC__)
#define ARDAE_API __attribute__((visibility("default")))
#else
#define ARDAE_API
#endif

ARDAE_API int ardae_version(void);
ARDAE_API const char* ardae_last_error(void);
/* 0 if device `dev` is an sm_100 part and the TMA descriptor entry point resolves. */
ARDAE_API int ardae_check_device(int dev);

/* ------------------------------------------------------------------------------------------
 * Conditional AR-DAE, `net.MLPGradCARDAE` (models/graddae/mlp.py:341-483; constructed at
 * ivae_ardae.py:595-606 with enc_ctx = enc_input = True, softplus).
 * Parameter tensors are passed as 6L+2 device pointers in state_dict order:
 *   ctx_encode.layers.{0..L-2}.{weight,bias}, ctx_encode.fc.{weight,bias},
 *   inp_encode.(same), neglogprob.layers.{0..L-1}.{weight,bias}, neglogprob.fc.{weight,bias}
 * with nn.Linear [out,in] layout.  `grads` has the same order; gradients are ACCUMULATED into it
 * (autograd semantics); the entry of neglogprob.fc.bias is never written (the reference leaves
 * its .grad = None).  The pointers must stay valid and fixed for the life of the handle; their
 * CONTENTS may change between calls (weights are re-read every call).
 */
typedef struct ardae_cdae_s* ardae_cdae_t;

typedef struct {
  int input_dim;         /* d  */
  int context_dim;       /* c  */
  int h_dim;             /* H, multiple of 4 */
  int num_hidden_layers; /* L >= 2 */
  int batch;             /* B  distinct context rows */
  int samples;           /* S  rows per context row; N = B*S, row index b*S+k */
  int train;             /* 1: loss + gradients (forward + backward); 0: glogprob only */
  int kind;              /* 0: net.MLPGradCARDAE (--cdae mlp-grad, models/graddae/mlp.py:341-483): energy network, score by
                          *    back-propagation, double-backprop training gradient;
                          * 1: net.MLPResCARDAE (--cdae mlp-res, models/resdae/mlp.py:286-413): the last MLP is `dae`
                          *    ([d, H] output layer) and outputs the score itself; every tensor receives a gradient */
} ardae_cdae_config;

ARDAE_API int ardae_cdae_workspace_bytes(const ardae_cdae_config* cfg, size_t* bytes);
ARDAE_API int ardae_cdae_create(const ardae_cdae_config* cfg, float* const* params, float* const* grads,
                      int num_tensors, void* workspace, size_t workspace_bytes, ardae_cdae_t* out);
ARDAE_API void ardae_cdae_destroy(ardae_cdae_t h);
/* number of kernel launches one train / score call issues */
ARDAE_API int ardae_cdae_num_launches(ardae_cdae_t h);
/* Measurement hook (bench.py roofline, no reference counterpart): when on, every launch of the plan is bracketed by
 * CUDA events on the stream it is issued to; read_profile synchronises the device and returns, for the LAST call,
 * one 16-byte tag ("chain_tangent", "gemm_tn", ...) and the elapsed milliseconds per launch. */
ARDAE_API int ardae_cdae_set_profile(ardae_cdae_t h, int on);
ARDAE_API int ardae_cdae_read_profile(ardae_cdae_t h, int max_ops, char* tags, float* ms, int* num_ops);

/* Replaces ConditionalARDAE.forward + cdae_loss.backward() (graddae/mlp.py:400-444,
 * ivae_ardae.py:768-771).  x [N,d] = input flattened, ctx [B,c], sigma [N] (= std, signed),
 * eps [N,d]: the Gaussian noise of add_gaussian_noise (graddae/mlp.py:21-23) -- supplied by the
 * caller (gen_eps = 0) or drawn in-kernel with Philox from `seed` and written back (gen_eps = 1).
 * inv_count = 1/(N_total*d): the mse_loss mean (pass the GLOBAL row count under data parallelism
 * so that summing gradients over ranks reproduces the reference mean).
 * loss_out: device scalar (overwritten).  score_out: optional [N,d], receives g = grad log p. */
ARDAE_API int ardae_cdae_train(ardae_cdae_t h, const float* x, const float* ctx, const float* sigma,
                     float* eps, int gen_eps, uint64_t seed, float inv_count, float* loss_out,
                     float* score_out, void* stream);

/* Replaces ConditionalARDAE.glogprob (graddae/mlp.py:446-483; ivae_ardae.py:829).
 * Needs a handle created with train = 0.  score_out [N,d]. */
ARDAE_API int ardae_cdae_score(ardae_cdae_t h, const float* x, const float* ctx, const float* sigma,
                     float* score_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Implicit-posterior VAE with the noise-concat MLP encoder: `net.ToyIPVAE` (models/ivae/toy.py:
 * 30-109,154-194,694-858) and `net.MNISTIPVAE` (models/ivae/mnist.py:38-301), enc_type 'concat'.
 * Parameter tensors in state_dict order (weights [out,in]):
 *   encode.inp_encode (n_inp linears), encode.fc.layers (n_fc), encode.fc.fc,
 *   decode.main (n_dec linears), then decode.reparam.{mean_fn,logvar_fn} (toy) | logit_fn (mnist).
 * kind 0 = toy: Gaussian decoder, noise re-concatenated at every fc layer (layers.py:707-724);
 * kind 1 = mnist: x <- 2x-1, noise concatenated once, Bernoulli decoder.
 * kind 2 = conv: `net.ConvIPVAE` (models/ivae/conv.py:44-304 + models/vae/conv.py:79-136): tensors
 *   encode.conv1..3, encode.fc4, encode.fc5, decode.fc.layers.0, decode.fc.fc, decode.deconv1, decode.deconv2,
 *   decode.reparam.logit_fn (n_inp = 3, n_fc = 1, n_dec = 2, h_dim = 800 = fc4 width).
 * kind 3 = auxmnist: hierarchical `net.MNISTAuxIPVAE` (models/ivae/auxmnist.py:47-300 over models/vae/auxmnist.py:31-68,
 *   147-191): tensors encode.aux_encode.main (n_inp linears), encode.aux_encode.reparam.{mean_fn,logvar_fn},
 *   encode.encode.fc (n_fc linears, the first [h, D + noise_dim]), encode.encode.reparam.{mean_fn,logvar_fn},
 *   decode.main (n_dec linears), decode.reparam.logit_fn.  z0 = mu0(x) + exp(lv0(x)/2) eps0 is noise_dim wide,
 *   z = mu(x, z0) + exp(lv(x, z0)/2) eps; `noise` is [R, noise_dim + z_dim] = (eps0 | eps).  Modes 0-2.
 */
typedef struct ardae_model_s* ardae_model_t;

typedef struct {
  int kind;
  int input_dim, noise_dim, h_dim, z_dim;
  int n_inp, n_fc, n_dec; /* linears in inp_encode; hidden layers of encode.fc; linears in decode.main */
  int act;                /* 0 relu, 1 softplus */
  int batch, nz;          /* B data rows, nz noise samples per row: R = B*nz rows, row index b*nz+k */
  int mode;               /* 0: encode only; 1: forward + backward; 2: IWS log-likelihood;
                           * 3 / 4 / 5: decode(z) / encode._forward_inp(x) / encode._forward_all(inp, nos) (nz = 1) */
  int img_h, img_c;       /* kind 2 (ConvIPVAE): image height (= width) and channels; input_dim = img_c*img_h^2 */
} ardae_model_config;

ARDAE_API int ardae_model_workspace_bytes(const ardae_model_config* cfg, size_t* bytes);
ARDAE_API int ardae_model_create(const ardae_model_config* cfg, float* const* params, float* const* grads,
                                 int num_tensors, void* workspace, size_t workspace_bytes, ardae_model_t* out);
ARDAE_API void ardae_model_destroy(ardae_model_t h);
ARDAE_API int ardae_model_num_launches(ardae_model_t h, int which); /* 0 fwd, 1 bwd */

/* Replaces Encoder.forward / ImplicitPosteriorVAE.forward_hidden (toy.py:90-109,811-822):
 * z_out [R, z_dim] = f(x [B, D], noise [R, n]); noise == NULL means zeros (encode(std=0)). */
ARDAE_API int ardae_model_encode(ardae_model_t h, const float* x, const float* noise, float* z_out, void* stream);

/* z = encode(x, noise, nz) AND z-bar = encode(x, std=0) [B, z_dim] in one pass: the reference evaluates
 * model.encode(x, std=0) twice per update next to the sampling pass (ivae_ardae.py:735,748,749), each time recomputing
 * the input stack; here the mean code is two or three B-row launches on top of the sampling pass. */
ARDAE_API int ardae_model_encode_with_mean(ardae_model_t h, const float* x, const float* noise, float* z_out,
                                           float* zbar_out, void* stream);
/* kind 3 only: the same pass plus Encoder.forward_hidden(x, std=0) (models/ivae/auxmnist.py:123-131), the CDAE context
 * of cdae_ctx_type 'hidden1a' (ivae_ardae.py:739-741): hidden_out [B, 2*h_dim] = cat(h0, h) at eps0 = eps = 0. */
ARDAE_API int ardae_model_encode_hidden(ardae_model_t h, const float* x, const float* noise, float* z_out,
                                        float* zbar_out, float* hidden_out, void* stream);
/* Replaces ImplicitPosteriorVAE.forward (toy.py:824-858 / mnist.py:267-301, lmbd = 0): encoder,
 * decoder, per-row recon + beta*prior.  sums[3] (device) <- mean loss, recon, prior over
 * 1/inv_rows rows (pass the GLOBAL row count under data parallelism).  heads_out (optional):
 * head-major [nH][R][D] = logits (mnist) or mu, logvar (toy).  Keeps the tape for backward. */
ARDAE_API int ardae_model_forward(ardae_model_t h, const float* x, const float* noise, float beta, float inv_rows,
                                  float* z_out, float* sums, float* heads_out, void* stream);

/* Sub-module calls of the reference API (SURVEY 8b "must expose"), each on its own plan (nz = 1, batch = rows):
 *   mode 3  Decoder.forward(z) minus the sampler (ivae/toy.py:725-737, ivae/mnist.py:188-199, vae/conv.py:118-136):
 *           z [rows, z_dim] -> heads_out, head-major [n_heads][rows][input_dim] (toy: mu, logvar; mnist / conv: logit)
 *   mode 4  Encoder._forward_inp(x) (ivae/mnist.py:76-86; conv: the flattened conv3 features, ivae/conv.py:84-96):
 *           x [rows, input_dim] -> inp_out [rows, feat]
 *   mode 5  ConcatEncoder._forward_all(inp, nos) (ivae/mnist.py:161-165, ivae/toy.py:192-194, ivae/conv.py:107-115):
 *           inp [rows, feat], nos [rows, noise_dim] (NULL = zeros) -> z_out [rows, z_dim] */
ARDAE_API int ardae_model_decode(ardae_model_t h, const float* z, float* heads_out, void* stream);
ARDAE_API int ardae_model_forward_inp(ardae_model_t h, const float* x, float* inp_out, void* stream);
ARDAE_API int ardae_model_forward_all(ardae_model_t h, const float* inp, const float* noise, float* z_out, void* stream);

/* beta annealing (utils/msc.py:53-55 annealing_func, `--beta-annealing` of ivae_ardae.py:101) under CUDA-graph
 * replay: when `beta_device` is non-NULL every later forward / backward launch of this handle reads beta from that
 * device scalar at kernel time -- the float `beta` argument of ardae_model_forward is ignored and the `gz_scale` of
 * the backward calls must then EXCLUDE beta (the kernels multiply it in).  NULL restores the by-value behaviour. */
ARDAE_API int ardae_model_set_beta_device(ardae_model_t h, const float* beta_device);

/* Replaces model_loss.backward() and (S*(z - zbar)).backward(grad) (ivae_ardae.py:804,834) in one
 * pass: accumulates d(loss_scale*loss)/dtheta plus the pull-back of gz_scale*gz (an upstream
 * gradient on z, [R, z_dim], may be NULL) into `grads`.  loss_scale == 0 skips the decoder. */
ARDAE_API int ardae_model_backward(ardae_model_t h, float loss_scale, const float* gz, float gz_scale, void* stream);
/* The same backward in two calls (decoder + prior half first, then the encoder half with the upstream gradient on z):
 * model_loss.backward(retain_graph=True) at ivae_ardae.py:804 only needs the ELBO forward, so a driver can issue the
 * decoder half before the entropy-gradient estimate of :829 exists. */
ARDAE_API int ardae_model_backward_decoder(ardae_model_t h, float loss_scale, void* stream);
ARDAE_API int ardae_model_backward_encoder(ardae_model_t h, float loss_scale, const float* gz, float gz_scale,
                                           void* stream);

/* Replaces ImplicitPosteriorVAE.logprob = logprob_w_cov_gaussian_posterior (toy.py:878-939 /
 * mnist.py:378-437; called by evaluate_iws, ivae_ardae.py:644-673), batched over the images instead
 * of a Python loop.  Handle created with mode = 2, batch = images per call, nz = sample_size
 * (>= 2*z_dim, z_dim <= 64).  noise [B*S, n]: encoder noise; eta [B*S, z_dim]: the standard normal
 * behind MultivariateNormal.rsample (NULL = Philox from seed).  out [B] <- log(mean_k exp(w_k - max)
 * + 1e-10) + max per image; *total (device, optional) += sum_i out[i]; *status (device, optional)
 * <- 1+i if image i's sample covariance is not positive definite. */
ARDAE_API int ardae_model_iws(ardae_model_t h, const float* x, const float* noise, const float* eta, uint64_t seed,
                              float* out, float* total, int* status, void* stream);

/* ------------------------------------------------------------------------------------------
 * Step glue and optimizers */
/* sigma schedule, ivae_ardae.py:753-767: x_out [B*nz*nstd, d] = S*(z - zbar) (expanded over nstd),
 * std_out[b] = delta * mean_d std_k(S*(z-zbar)) (unbiased over nz), sigma_out = std_b * xi
 * (xi [B*nz*nstd] supplied, or NULL = Philox draw from seed).  z [B,nz,d], zbar [B,d]. */
ARDAE_API int ardae_sigma_schedule(const float* z, const float* zbar, int B, int nz, int d, int nstd, float S,
                                   float delta, const float* xi, uint64_t seed, float* x_out, float* sigma_out,
                                   float* std_out, void* stream);
/* out [R,d] = S*(z [R,d] - zbar [R/nz, d])  (ivae_ardae.py:827) */
ARDAE_API int ardae_scaled_diff(const float* z, const float* zbar, int R, int nz, int d, float S, float* out,
                                void* stream);
/* utils.Adam.step (utils/optim.py:49-108, PyTorch-1.2 epsilon placement) over a flat arena of n
 * floats (n % 4 == 0, 16-byte aligned); `step` is the 1-based step count; grads are pre-multiplied
 * by gscale (e.g. 1/world_size). */
ARDAE_API int ardae_adam_step(float* p, const float* g, float* exp_avg, float* exp_avg_sq, size_t n, float lr,
                              float beta1, float beta2, float eps, int step, float gscale, void* stream);
/* torch.optim.RMSprop.step (centered = False) as constructed at ivae_ardae.py:626. */
ARDAE_API int ardae_rmsprop_step(float* p, const float* g, float* square_avg, float* momentum_buf, size_t n,
                                 float lr, float alpha, float eps, float momentum, float gscale, void* stream);

/* ------------------------------------------------------------------------------------------
 * Utilities */
/* out[i] ~ N(0,1), Philox4x32-10 + Box-Muller (replaces torch.randn at ivae_ardae.py:761 and
 * Encoder.sample_noise, ivae/toy.py:61-65, which draws on the CPU and copies). */
ARDAE_API int ardae_randn(float* out, size_t n, uint64_t seed, uint32_t stream_id, void* stream);

/* out[i] = 1 if u_i < probs[i] else 0 (Philox uniform): dynamic binarisation on the device; replaces the torch.bernoulli
 * DataLoader transform of datasets/mnist.py:39-40,129. */
ARDAE_API int ardae_bernoulli(const float* probs, float* out, size_t n, uint64_t seed, void* stream);

/* CUDA-graph support (no reference counterpart: the reference launches eagerly).  While a device counter is set,
 * every launch issued by this library bakes the POINTER into its arguments: Philox seeds become
 * seed + counter * 64 * golden-ratio (the host numbers its draws base + (64 * iteration + k) * golden-ratio, k < 64, so
 * replayed and eager iterations never share a seed) and Adam's bias-correction step becomes step + counter, so a
 * captured step draws fresh
 * noise and advances Adam on every replay.  Set it before capturing, reset it to NULL afterwards (captured launches
 * keep the pointer; eager calls then run with offset 0); ardae_bump_replay_counter is the last launch of the graph. */
ARDAE_API int ardae_set_replay_counter(const unsigned long long* device_counter);
ARDAE_API int ardae_bump_replay_counter(unsigned long long* device_counter, void* stream);

/* ---- data parallel over NVLink peer memory (no reference counterpart: the reference is single-device) ----
 * One process per GPU.  Every rank allocates an exchange buffer of ardae_dp_xchg_bytes(n, world) ZEROED bytes (n = floats
 * of the largest arena), exports it with ardae_ipc_export (CUDA IPC handle of the underlying allocation + byte offset),
 * exchanges the 64-byte handles out of band (torch.distributed) and opens the peers' buffers with ardae_ipc_import.
 * ardae_dp_fused_step then replaces `allreduce(g); optimizer.step()` (ivae_ardae.py:779,846 under batch sharding) by
 * ONE kernel: gradient slices are pushed to their owner ranks, each rank reduces and updates its 1/world slice of the
 * arena (kind 0: utils.Adam, utils/optim.py:49-108; kind 1: torch RMSprop) and pushes the updated parameters back;
 * parameters stay replicated bit for bit, optimizer state is advanced on the owner rank only.  All ranks must call it
 * in the same order.  epoch_barrier_status: 4 zero-initialised uint64 of local device memory kept for the lifetime of
 * the communicator ([2] becomes non-zero if a peer never answered).  Safe to capture in a CUDA graph. */
ARDAE_API int ardae_ipc_export(const void* ptr, unsigned char* handle64, size_t* offset);
ARDAE_API int ardae_ipc_import(const unsigned char* handle64, size_t offset, void** out);
ARDAE_API int ardae_dp_xchg_bytes(size_t n, int world, size_t* bytes);
ARDAE_API int ardae_dp_fused_step(int kind, int rank, int world, float* p, const float* g, float* s1, float* s2, size_t n,
                                  void* const* peer_xchg, unsigned long long* epoch_barrier_status, float lr, float beta1,
                                  float beta2_or_alpha, float eps, float momentum, int step, float gscale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ARDAE_H_ */
