// Data-parallel gradient exchange fused with the optimizer update, over NVLink peer memory.
//
// The reference has no multi-GPU path; the B200 path shards the minibatch, so every iteration needs
//     g <- sum_r g_r ; p <- optimizer(p, g)          for the two flat arenas (CDAE: RMSprop, model: Adam).
// With NCCL that is an allreduce (45 us for 4 MB) followed by an optimizer launch (11 us) on the critical path, twice per
// iteration.  Here it is ONE kernel per arena: every rank owns a 1/world slice of the arena and
//   A. pushes its gradient contribution for slice s into rank s's exchange buffer (peer stores over NVLink),
//   B. after the flag of every peer arrived: sums the world contributions of its own slice, applies the optimizer to
//      that slice (parameters and optimizer state are replicated, each rank only advances its slice of the state) and
//      pushes the updated parameters of its slice into every peer's exchange buffer,
//   C. after the second flag: copies the other ranks' updated slices from its exchange buffer into its parameter arena.
// Flags are 64-bit epochs written with release / read with acquire semantics at system scope; the epoch lives in device
// memory and is advanced by the kernel, so CUDA-graph replays need no host involvement.  Pushes (posted stores) keep
// NVLink read latency off the path.  Optimizer state of slices a rank does not own is never read again by that rank, so
// it goes stale there: the host side gathers it before it is saved or inspected (ardae/dp.py: PeerComm.gather_state).
// Parameters stay replicated bit for bit: every rank receives the same updated values.
//
// Exchange buffer of one rank (device memory, opened by every peer through CUDA IPC):
//   [0, 1024)            flags A: world x uint64   (slot r written by rank r)
//   [1024, 2048)         flags B
//   [2048, 2048 + 4n)    gradient slots: world slots of q float4 (slot r = rank r's contribution to MY slice)
//   [.., + 4n)           parameter area: the arena layout; slice s written by rank s
#pragma once
#include "kernels.cuh"

namespace ardae {

constexpr int kDpMaxWorld = 16;
constexpr size_t kDpHeaderBytes = 2048;
constexpr int kDpGrid = 64, kDpBlock = 512;

struct DpFusedParams {
  float* p;               // [n] parameters (local arena)
  const float* g;         // [n] local gradient contribution
  float* s1;              // optimizer state (Adam: exp_avg, RMSprop: square_avg)
  float* s2;              // (Adam: exp_avg_sq, RMSprop: momentum buffer)
  size_t n;               // floats, n % 4 == 0
  int rank, world, kind;  // kind 0 = Adam, 1 = RMSprop
  float lr, b1, b2, eps, mu, alpha, gscale;
  int step0;
  const replay_ctr_t* ctr;
  unsigned long long* epoch;    // local device scalar: number of fused steps done so far
  unsigned long long* barrier;  // local device scalar: grid-barrier arrivals (monotonic)
  int* status;                  // local device scalar: set to 1 if a peer flag never arrived
  unsigned char* xchg[kDpMaxWorld];  // every rank's exchange buffer as seen from this device (own included)
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// all CTAs of this (co-resident) grid have arrived `target` times in total
__device__ __forceinline__ void dp_grid_sync(unsigned long long* ctr, unsigned long long target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    atomicAdd(ctr, 1ull);
    while (ld_acquire_gpu(ctr) < target) __nanosleep(20);
  }
  __syncthreads();
}

// wait until every peer wrote `epoch` into my flag array (bounded: ~2 s, then status = 1 and carry on)
__device__ __forceinline__ void dp_wait_flags(const unsigned long long* flags, int world, int rank, unsigned long long epoch,
                                              int* status) {
  if (threadIdx.x < world && static_cast<int>(threadIdx.x) != rank) {
    long long spins = 0;
    while (ld_acquire_sys(flags + threadIdx.x) < epoch) {
      __nanosleep(40);
      if (++spins > (1ll << 25)) {
        *status = 1;
        break;
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kDpBlock) dp_fused_step_kernel(DpFusedParams a) {
  const int rank = a.rank, world = a.world;
  const unsigned long long e = *a.epoch + 1ull;  // read before anyone advances it (advance happens after barrier 2)
  const unsigned long long bar_base = (e - 1ull) * 2ull * gridDim.x;
  const size_t n4 = a.n / 4;
  const size_t q = (n4 + world - 1) / world;  // float4 per slice
  const size_t tid = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  const size_t nth = static_cast<size_t>(gridDim.x) * blockDim.x;
  const float4* g4 = reinterpret_cast<const float4*>(a.g);
  auto flagsA = [&](int r) { return reinterpret_cast<unsigned long long*>(a.xchg[r]); };
  auto flagsB = [&](int r) { return reinterpret_cast<unsigned long long*>(a.xchg[r] + 1024); };
  auto slots = [&](int r) { return reinterpret_cast<float4*>(a.xchg[r] + kDpHeaderBytes); };
  auto parea = [&](int r) { return reinterpret_cast<float4*>(a.xchg[r] + kDpHeaderBytes + world * q * sizeof(float4)); };

  // ---- A: push my contribution to every other slice owner
  for (int k = 1; k < world; ++k) {
    const int s = (rank + k) % world;  // stagger the targets over the ranks
    const size_t lo = s * q, hi = (lo + q < n4) ? lo + q : n4;
    float4* dst = slots(s) + static_cast<size_t>(rank) * q;
    for (size_t i = lo + tid; i < hi; i += nth) dst[i - lo] = g4[i];
  }
  dp_grid_sync(a.barrier, bar_base + gridDim.x);
  if (blockIdx.x == 0 && threadIdx.x < world && static_cast<int>(threadIdx.x) != rank)
    st_release_sys(flagsA(threadIdx.x) + rank, e);
  dp_wait_flags(flagsA(rank), world, rank, e, a.status);

  // ---- B: reduce my slice, update it, push the new parameters
  {
    __shared__ float bc[2];
    if (threadIdx.x == 0) {
      const double t = static_cast<double>(a.step0) + (a.ctr != nullptr ? static_cast<double>(*a.ctr) : 0.0);
      const double bc1 = 1.0 - pow(static_cast<double>(a.b1), t), bc2 = 1.0 - pow(static_cast<double>(a.b2), t);
      bc[0] = static_cast<float>(static_cast<double>(a.lr) / bc1);
      bc[1] = static_cast<float>(1.0 / sqrt(bc2));
    }
    __syncthreads();
    const float lr_bc1 = bc[0], inv_sqrt_bc2 = bc[1];
    const size_t lo = rank * q, hi = (lo + q < n4) ? lo + q : n4;
    float4* p4 = reinterpret_cast<float4*>(a.p);
    float4* m4 = reinterpret_cast<float4*>(a.s1);
    float4* v4 = reinterpret_cast<float4*>(a.s2);
    const float4* mine = slots(rank);
    for (size_t i = lo + tid; i < hi; i += nth) {
      // fixed summation order (rank 0, 1, ...) so that every configuration of ranks reproduces itself
      float4 gg = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < world; ++r) {
        const float4 c = (r == rank) ? g4[i] : mine[static_cast<size_t>(r) * q + (i - lo)];
        gg.x += c.x; gg.y += c.y; gg.z += c.z; gg.w += c.w;
      }
      float4 pp = p4[i], mm = m4[i], vv = v4[i];
      float* pa = &pp.x; float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float gr = a.gscale * ga[j];
        if (a.kind == 0) {  // utils.Adam (utils/optim.py:49-108), as adam_kernel
          ma[j] = a.b1 * ma[j] + (1.0f - a.b1) * gr;
          va[j] = a.b2 * va[j] + (1.0f - a.b2) * gr * gr;
          const float denom = (sqrtf(va[j]) + a.eps) * inv_sqrt_bc2;
          pa[j] -= lr_bc1 * ma[j] / denom;
        } else {            // torch.optim.RMSprop, as rmsprop_kernel (s1 = square_avg, s2 = momentum buffer)
          ma[j] = a.alpha * ma[j] + (1.0f - a.alpha) * gr * gr;
          const float avg = sqrtf(ma[j]) + a.eps;
          if (a.mu > 0.0f) {
            va[j] = a.mu * va[j] + gr / avg;
            pa[j] -= a.lr * va[j];
          } else {
            pa[j] -= a.lr * gr / avg;
          }
        }
      }
      p4[i] = pp; m4[i] = mm; v4[i] = vv;
      for (int k = 1; k < world; ++k) parea((rank + k) % world)[i] = pp;
    }
  }
  dp_grid_sync(a.barrier, bar_base + 2ull * gridDim.x);
  if (blockIdx.x == 0 && threadIdx.x < world && static_cast<int>(threadIdx.x) != rank)
    st_release_sys(flagsB(threadIdx.x) + rank, e);
  if (blockIdx.x == 0 && threadIdx.x == 0) *a.epoch = e;  // every CTA read the epoch before barrier 1
  dp_wait_flags(flagsB(rank), world, rank, e, a.status);

  // ---- C: the other owners' updated slices arrived in my exchange buffer
  {
    float4* p4 = reinterpret_cast<float4*>(a.p);
    const float4* src = parea(rank);
    const size_t lo = rank * q, hi = (lo + q < n4) ? lo + q : n4;
    for (size_t i = tid; i < n4; i += nth)
      if (i < lo || i >= hi) p4[i] = src[i];
  }
}

inline size_t dp_xchg_bytes(size_t n, int world) {
  const size_t n4 = n / 4, q = (n4 + world - 1) / world;
  return kDpHeaderBytes + static_cast<size_t>(world) * q * 16 + n4 * 16 + 256;
}

}  // namespace ardae
