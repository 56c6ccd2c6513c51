// Gradient all-reduce over NVLink peer memory fused with the Adam update: ONE launch per step and rank, no NCCL on the
// data path, capturable in the step's CUDA graph.
//
// The only exchange of a data-parallel DGN step is the all-reduce of the flat fp32 gradient (~0.5 M floats = 2.2 MB):
// latency bound.  A library all-reduce costs a launch of its own plus the optimizer launch, both outside the captured
// graph.  Here every rank's gradient buffer lives in symmetric (peer-mapped) memory and one kernel does
// (two-shot form; with <= one_shot_max_world ranks every rank simply reads all buffers in full: one barrier fewer)
//   barrier A   every rank's backward has finished (per-CTA flags in the peers' signal pads, st.release.sys / ld.acquire.sys)
//   phase 1     reduce-scatter: rank r sums slice r of all ranks' gradients IN RANK ORDER (deterministic) over NVLink
//               loads and writes the sum back into its own buffer
//   barrier B
//   phase 2     all-gather fused with Adam: every rank reads each slice from its owner and updates its replica of the
//               parameters and moments (same arithmetic as adam_kernel; the 1 / world of the average is grad_scale)
//   barrier C   nobody still reads a peer's buffer when that peer starts the next step
// Every rank sees bit-identical sums, so the replicas never drift.  Traffic per rank: 2 (world-1)/world of the buffer.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/dgn_b200.h"
#include "dgn_launch.cuh"

extern thread_local cudaError_t g_dgn_last_cuda;

namespace dgn {

constexpr int kArBlocks = DGN_AR_BLOCKS, kArThreads = 1024;
constexpr long long kSpinLimit = 60000000000ll;                          // ~30 s of SM clocks

struct PeerArgs {
  int world, rank;
  const unsigned long long* grad_ptrs;    // [world] device pointers (peer mapped)
  const unsigned long long* flag_ptrs;    // [world] device pointers to uint32 [3][kArBlocks][world]
  unsigned* epoch;                        // local; [0] = launches so far, [1] = internal CTA counter, [2] = a barrier timed out
  long long n;
  float* p; float* m; float* v;
  float lr, b1, b2, eps, wd;
  const float* hyper;
  int* state;
  float* reduced;                         // optional: receives the all-reduced SUM
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer(const float4* p) {           // bypass L1: the data was written by another GPU
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// all ranks' CTA `blockIdx.x` meet: signal every peer, then wait for every peer's signal of this epoch.  FENCE: this
// CTA wrote data the peers read after the barrier.
template <bool FENCE>
__device__ __forceinline__ void peer_barrier(const PeerArgs& k, int phase, unsigned epoch) {
  if (FENCE) __threadfence_system();
  __syncthreads();
  const int t = threadIdx.x;
  if (t < k.world && t != k.rank) {
    unsigned* theirs = reinterpret_cast<unsigned*>(k.flag_ptrs[t]) + ((size_t)phase * kArBlocks + blockIdx.x) * k.world + k.rank;
    st_release_sys(theirs, epoch);
    const unsigned* mine = reinterpret_cast<const unsigned*>(k.flag_ptrs[k.rank]) +
                           ((size_t)phase * kArBlocks + blockIdx.x) * k.world + t;
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(mine) - epoch) < 0) {
      if (clock64() - t0 > kSpinLimit) { k.epoch[2] = 1u; break; }       // a peer never arrived: flag it, do not hang the GPU
    }
  }
  __syncthreads();
}

struct AdamCoef { float b1, b2, eps, wd, gs, step_size, inv_sqrt_c2; };

__device__ __forceinline__ void adam4(const PeerArgs& k, const AdamCoef& c, long long i, const float4 gg) {
  float4 pp = *reinterpret_cast<float4*>(k.p + 4 * i), mm = *reinterpret_cast<float4*>(k.m + 4 * i),
         vv = *reinterpret_cast<float4*>(k.v + 4 * i);
  float* pa = &pp.x; float* ma = &mm.x; float* va = &vv.x; const float* ga = &gg.x;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float gr = fmaf(c.wd, pa[j], c.gs * ga[j]);
    ma[j] = fmaf(c.b1, ma[j], (1.f - c.b1) * gr);
    va[j] = fmaf(c.b2, va[j], (1.f - c.b2) * gr * gr);
    pa[j] -= c.step_size * ma[j] / (sqrtf(va[j]) * c.inv_sqrt_c2 + c.eps);
  }
  *reinterpret_cast<float4*>(k.p + 4 * i) = pp;
  *reinterpret_cast<float4*>(k.m + 4 * i) = mm;
  *reinterpret_cast<float4*>(k.v + 4 * i) = vv;
  if (k.reduced) *reinterpret_cast<float4*>(k.reduced + 4 * i) = gg;
}

// rank-order sum of element i over all ranks' buffers (all loads in flight before the first add)
__device__ __forceinline__ float4 sum_ranks(const PeerArgs& k, long long i) {
  float4 g[DGN_AR_MAX_WORLD];
#pragma unroll
  for (int q = 0; q < DGN_AR_MAX_WORLD; ++q)
    if (q < k.world) g[q] = ld_peer(reinterpret_cast<const float4*>(k.grad_ptrs[q]) + i);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int q = 0; q < DGN_AR_MAX_WORLD; ++q)
    if (q < k.world) { acc.x += g[q].x; acc.y += g[q].y; acc.z += g[q].z; acc.w += g[q].w; }
  return acc;
}

// ONE_SHOT: every rank reads all the buffers in full (world - 1 buffers over NVLink, one barrier and one latency round
// fewer); two-shot: reduce-scatter + all-gather (2 (world - 1) / world buffers).
template <bool ONE_SHOT>
__global__ void __launch_bounds__(kArThreads) allreduce_adam_kernel(const __grid_constant__ PeerArgs k) {
  pdl_prologue();
  const unsigned epoch = k.epoch[0] + 1u;
  const int W = k.world;
  const long long n4 = k.n / 4;
  const long long tid = (long long)blockIdx.x * kArThreads + threadIdx.x, stride = (long long)gridDim.x * kArThreads;
  AdamCoef c;
  {
    float lr = k.lr;
    c.b1 = k.b1; c.b2 = k.b2; c.eps = k.eps; c.wd = k.wd; c.gs = 1.f;
    if (k.hyper) { lr = k.hyper[0]; c.wd = k.hyper[1]; c.gs = k.hyper[2]; }
    const int t = k.state[0] + 1;
    const float c1 = 1.f - powf(k.b1, (float)t), c2 = 1.f - powf(k.b2, (float)t);
    c.step_size = lr / c1; c.inv_sqrt_c2 = rsqrtf(c2);
  }
  peer_barrier<false>(k, 0, epoch);                                    // every rank's backward has finished
  if (ONE_SHOT) {
    for (long long i = tid; i < n4; i += stride) adam4(k, c, i, sum_ranks(k, i));
  } else {
    const long long per = (n4 + W - 1) / W;                            // float4 per rank slice
    float4* mine = reinterpret_cast<float4*>(k.grad_ptrs[k.rank]);
    const long long s0 = per * k.rank, s1 = min(s0 + per, n4);
    for (long long i = s0 + tid; i < s1; i += stride) mine[i] = sum_ranks(k, i);          // reduce-scatter
    peer_barrier<true>(k, 1, epoch);
    for (long long i = tid; i < n4; i += stride)                                          // all-gather + Adam
      adam4(k, c, i, ld_peer(reinterpret_cast<const float4*>(k.grad_ptrs[(int)(i / per)]) + i));
  }
  peer_barrier<false>(k, 2, epoch);                                    // nobody still reads my buffer when I move on
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(&k.epoch[1], 1u);
    if (done == gridDim.x - 1) { k.epoch[1] = 0u; k.epoch[0] = epoch; k.state[0] = k.state[0] + 1; }
  }
}

}  // namespace dgn

using namespace dgn;

extern "C" int dgn_allreduce_adam(const DgnPeerGroup* pg, int64_t n, float* param, float* exp_avg, float* exp_avg_sq,
                                  float lr, float beta1, float beta2, float eps, float weight_decay, const float* hyper,
                                  int32_t* state, void* stream) {
  if (!pg || !pg->grad_ptrs || !pg->flag_ptrs || !pg->epoch || pg->world < 1 || pg->world > DGN_AR_MAX_WORLD ||
      pg->rank < 0 || pg->rank >= pg->world || n < 0 || !param || !exp_avg || !exp_avg_sq || !state)
    return DGN_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15u)
    return DGN_ERR_ALIGNMENT;
  if (n % 4) return DGN_ERR_UNSUPPORTED;               // the flat buffers of the engine are padded to 16 B
  if (n == 0) return DGN_OK;
  PeerArgs k;
  k.world = pg->world; k.rank = pg->rank;
  k.grad_ptrs = reinterpret_cast<const unsigned long long*>(pg->grad_ptrs);
  k.flag_ptrs = reinterpret_cast<const unsigned long long*>(pg->flag_ptrs);
  k.epoch = reinterpret_cast<unsigned*>(pg->epoch);
  k.n = n; k.p = param; k.m = exp_avg; k.v = exp_avg_sq;
  k.lr = lr; k.b1 = beta1; k.b2 = beta2; k.eps = eps; k.wd = weight_decay; k.hyper = hyper; k.state = state;
  k.reduced = pg->reduced;
  // One CTA per SM.  Measured at N = 2 (tools/peer_bench.py, DGN_AR_GRID sweep): 148 CTAs 19.7 us, 64: 24.3, 34: 28.7,
  // 16: 42.9 - the kernel is bound by each thread's serial remote-load latency, not by the per-CTA barrier traffic, so
  // fewer CTAs (fewer flag stores, more iterations per thread) lose.  All ranks must use the same grid.
  static const int force_grid = [] { const char* e = getenv("DGN_AR_GRID"); return e ? atoi(e) : 0; }();
  const int want = force_grid > 0 ? force_grid : kArBlocks;
  const unsigned grid = (unsigned)(want < 1 ? 1 : (want > kArBlocks ? kArBlocks : want));
  if (pg->world <= pg->one_shot_max_world)
    launch_pdl(allreduce_adam_kernel<true>, dim3(grid), dim3(kArThreads), 0, (cudaStream_t)stream, k);
  else
    launch_pdl(allreduce_adam_kernel<false>, dim3(grid), dim3(kArThreads), 0, (cudaStream_t)stream, k);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_dgn_last_cuda = e; return DGN_ERR_CUDA; }
  return DGN_OK;
}
