// Graph-level prediction head and its loss in one launch per direction (sm_100a).
//
// MLPReadout (rb/nets/mlp_readout_layer.py:11-30, L = 2): y = W3 relu(W2 relu(W1 x + b1) + b2) + b3 on the B graph
// vectors of a batch, and nn.L1Loss (rb/nets/molecules_graph_regression/dgn_net.py:90-92).  Through the library
// this head is ~32 launches per step (3 GEMMs + epilogues + activations forward; sign / scale / 3 x (2 GEMMs + bias
// reduction) + gradient accumulation adds backward) for ~1 MFLOP of work - at the headline batch (128 graphs) that
// is a quarter of all launches of the step.  Here the batch is cut into tiles of 32 rows, every CTA keeps all three
// weight matrices in shared memory.  Forward: one CTA per tile.  Backward: a thread-block CLUSTER of 4 CTAs shares the
// tiles; every weight-gradient entry lives in a register of a fixed owner thread, rows are added in order, and the
// four partial sums meet in distributed shared memory: rank 0 reads its peers' partials (DSMEM) in rank order and
// writes or accumulates into the gradient buffers - deterministic, no atomics, no workspace, no second launch.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dgn_b200.h"
#include "dgn_launch.cuh"

namespace cg = cooperative_groups;

extern thread_local cudaError_t g_dgn_last_cuda;

namespace dgn {

// 1024 threads: the inner products are chains of dependent shared-memory loads + FMAs, so a single CTA is latency
// bound and the number of resident warps is its throughput (256 threads measured 3-4x slower than the ~32 library
// launches this replaces; profiles/README.md).
constexpr int HT = 1024;       // threads per CTA
constexpr int HR = 32;         // rows per tile
constexpr int HC = 4;          // CTAs of the backward's cluster
constexpr int HF = 32;         // at most this many CTAs in the forward
constexpr int HE1 = 4;         // weight-gradient entries per thread: d1*d0 <= HT*HE1, d2*d1 <= HT*HE2, ...
constexpr int HE2 = 1;
constexpr int HE3 = 1;

struct HeadSmem {
  float *w1t, *w2t, *w3t, *b1, *b2, *b3, *xs, *a1s, *a2s, *ys, *red;
};

__device__ __forceinline__ HeadSmem head_carve(float* sm, int d0, int d1, int d2, int dout) {
  HeadSmem s;
  s.w1t = sm;                  sm += d0 * d1;          // [d0][d1]  (transposed: consecutive threads, consecutive outputs)
  s.w2t = sm;                  sm += d1 * d2;          // [d1][d2]
  s.w3t = sm;                  sm += d2 * dout;        // [d2][dout]
  s.b1 = sm;                   sm += d1;
  s.b2 = sm;                   sm += d2;
  s.b3 = sm;                   sm += dout;
  s.xs = sm;                   sm += HR * d0;
  s.a1s = sm;                  sm += HR * d1;
  s.a2s = sm;                  sm += HR * d2;
  s.ys = sm;                   sm += HR * dout;
  s.red = sm;                                          // [HT / 32]
  return s;
}

static inline size_t head_smem_bytes(int d0, int d1, int d2, int dout) {
  return sizeof(float) * ((size_t)d0 * d1 + (size_t)d1 * d2 + (size_t)d2 * dout + d1 + d2 + dout +
                          (size_t)HR * (d0 + d1 + d2 + dout) + HT / 32 + 4);
}

__device__ __forceinline__ void head_load_weights(const DgnHeadArgs& a, const HeadSmem& s) {
  const int t = threadIdx.x;
  for (int i = t; i < a.d0 * a.d1; i += HT) { const int k = i / a.d1, j = i - k * a.d1; s.w1t[i] = __ldg(a.w1 + (size_t)j * a.d0 + k); }
  for (int i = t; i < a.d1 * a.d2; i += HT) { const int k = i / a.d2, j = i - k * a.d2; s.w2t[i] = __ldg(a.w2 + (size_t)j * a.d1 + k); }
  for (int i = t; i < a.d2 * a.d_out; i += HT) { const int k = i / a.d_out, j = i - k * a.d_out; s.w3t[i] = __ldg(a.w3 + (size_t)j * a.d2 + k); }
  for (int i = t; i < a.d1; i += HT) s.b1[i] = __ldg(a.b1 + i);
  for (int i = t; i < a.d2; i += HT) s.b2[i] = __ldg(a.b2 + i);
  for (int i = t; i < a.d_out; i += HT) s.b3[i] = __ldg(a.b3 + i);
}

// out[r][j] = act(bias[j] + sum_k in[r][k] * wt[k][j]) for the nr rows of the tile
template <bool RELU>
__device__ __forceinline__ void head_layer(const float* in, int din, const float* wt, const float* bias, int dn, int nr,
                                           float* out) {
  for (int i = threadIdx.x; i < nr * dn; i += HT) {
    const int r = i / dn, j = i - r * dn;
    float acc = bias[j];
    const float* x = in + r * din;
#pragma unroll 4
    for (int k = 0; k < din; ++k) acc = fmaf(x[k], wt[k * dn + j], acc);
    out[i] = RELU ? fmaxf(acc, 0.f) : acc;
  }
}

// deterministic block sum (fixed tree), result valid in thread 0
__device__ __forceinline__ float head_block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float tot = 0.f;
  if (threadIdx.x == 0)
    for (int w = 0; w < HT / 32; ++w) tot += red[w];
  return tot;
}

__global__ void __launch_bounds__(HT) head_fwd_kernel(const DgnHeadArgs a) {
  pdl_prologue();
  extern __shared__ __align__(16) float hsm[];
  const HeadSmem s = head_carve(hsm, a.d0, a.d1, a.d2, a.d_out);
  head_load_weights(a, s);
  for (int r0 = blockIdx.x * HR; r0 < a.n_rows; r0 += gridDim.x * HR) {
    const int nr = min(HR, a.n_rows - r0);
    __syncthreads();                                    // weights staged / previous tile consumed
    for (int i = threadIdx.x; i < nr * a.d0; i += HT) {
      const int r = i / a.d0, k = i - r * a.d0;
      s.xs[i] = __ldg(a.x + (size_t)(r0 + r) * a.ld_x + k);
    }
    __syncthreads();
    head_layer<true>(s.xs, a.d0, s.w1t, s.b1, a.d1, nr, s.a1s);
    __syncthreads();
    head_layer<true>(s.a1s, a.d1, s.w2t, s.b2, a.d2, nr, s.a2s);
    __syncthreads();
    head_layer<false>(s.a2s, a.d2, s.w3t, s.b3, a.d_out, nr, s.ys);
    __syncthreads();
    for (int i = threadIdx.x; i < nr * a.d1; i += HT) a.a1[(size_t)r0 * a.d1 + i] = s.a1s[i];
    for (int i = threadIdx.x; i < nr * a.d2; i += HT) a.a2[(size_t)r0 * a.d2 + i] = s.a2s[i];
    for (int i = threadIdx.x; i < nr * a.d_out; i += HT) {
      const int r = i / a.d_out, o = i - r * a.d_out;
      a.y[(size_t)(r0 + r) * a.ld_y + o] = s.ys[i];
    }
  }
}

// mean |y - target| over n elements: one CTA, fixed summation order
__global__ void __launch_bounds__(HT) l1_fwd_kernel(int n, const float* __restrict__ y, const float* __restrict__ t,
                                                    float* __restrict__ loss) {
  pdl_prologue();
  __shared__ float red[HT / 32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += HT) acc += fabsf(y[i] - t[i]);
  const float tot = head_block_sum(acc, red);
  if (threadIdx.x == 0) *loss = tot / (float)n;
}

// d_y = g_loss * sign(y - target) / n   (sign(0) = 0, as torch)
__global__ void __launch_bounds__(HT) l1_bwd_kernel(int n, const float* __restrict__ y, const float* __restrict__ t,
                                                    const float* __restrict__ g_loss, float* __restrict__ d_y) {
  pdl_prologue();
  const int i = blockIdx.x * HT + threadIdx.x;
  if (i >= n) return;
  const float d = y[i] - t[i], g = *g_loss / (float)n;
  d_y[i] = d > 0.f ? g : (d < 0.f ? -g : 0.f);
}

__global__ void __cluster_dims__(HC, 1, 1) __launch_bounds__(HT) head_bwd_kernel(const DgnHeadArgs a, const DgnHeadGrad g) {
  pdl_prologue();
  extern __shared__ __align__(16) float hsm[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int d0 = a.d0, d1 = a.d1, d2 = a.d2, dout = a.d_out, t = threadIdx.x;
  // shared memory: weights in their natural [out][in] layout (the backward multiplies by W, not W^T), the tile's
  // x / a1 / a2 and the three gradient tiles
  float* w1 = hsm;                 // [d1][d0]
  float* w2 = w1 + d1 * d0;        // [d2][d1]
  float* w3 = w2 + d2 * d1;        // [dout][d2]
  float* xs = w3 + dout * d2;      // [HR][d0]
  float* a1s = xs + HR * d0;       // [HR][d1]
  float* a2s = a1s + HR * d1;      // [HR][d2]
  float* dys = a2s + HR * d2;      // [HR][dout]
  float* dz2 = dys + HR * dout;    // [HR][d2]
  float* dz1 = dz2 + HR * d2;      // [HR][d1]
  float* part = dz1 + HR * d1;     // [d1*d0 + d2*d1 + dout*d2 + d1 + d2 + dout] this CTA's partial gradients
  for (int i = t; i < d1 * d0; i += HT) w1[i] = __ldg(a.w1 + i);
  for (int i = t; i < d2 * d1; i += HT) w2[i] = __ldg(a.w2 + i);
  for (int i = t; i < dout * d2; i += HT) w3[i] = __ldg(a.w3 + i);
  float gw1[HE1], gw2[HE2], gw3[HE3], gb = 0.f;         // gb: bias entry t of the concatenation [b1 | b2 | b3]
#pragma unroll
  for (int q = 0; q < HE1; ++q) gw1[q] = 0.f;
#pragma unroll
  for (int q = 0; q < HE2; ++q) gw2[q] = 0.f;
#pragma unroll
  for (int q = 0; q < HE3; ++q) gw3[q] = 0.f;

  for (int r0 = rank * HR; r0 < a.n_rows; r0 += HC * HR) {    // this CTA's tiles (possibly none: it still joins the syncs)
    const int nr = min(HR, a.n_rows - r0);
    __syncthreads();
    for (int i = t; i < nr * d0; i += HT) { const int r = i / d0, k = i - r * d0; xs[i] = __ldg(a.x + (size_t)(r0 + r) * a.ld_x + k); }
    for (int i = t; i < nr * d1; i += HT) a1s[i] = a.a1[(size_t)r0 * d1 + i];
    for (int i = t; i < nr * d2; i += HT) a2s[i] = a.a2[(size_t)r0 * d2 + i];
    for (int i = t; i < nr * dout; i += HT) { const int r = i / dout, o = i - r * dout; dys[i] = g.g_y[(size_t)(r0 + r) * g.ld_gy + o]; }
    __syncthreads();
    // dz2 = (dY W3) * relu'(a2)
    for (int i = t; i < nr * d2; i += HT) {
      const int r = i / d2, c = i - r * d2;
      float acc = 0.f;
      for (int o = 0; o < dout; ++o) acc = fmaf(dys[r * dout + o], w3[o * d2 + c], acc);
      dz2[i] = a2s[i] > 0.f ? acc : 0.f;
    }
    __syncthreads();
    // dz1 = (dz2 W2) * relu'(a1)
    for (int i = t; i < nr * d1; i += HT) {
      const int r = i / d1, c = i - r * d1;
      float acc = 0.f;
#pragma unroll 4
      for (int o = 0; o < d2; ++o) acc = fmaf(dz2[r * d2 + o], w2[o * d1 + c], acc);
      dz1[i] = a1s[i] > 0.f ? acc : 0.f;
    }
    __syncthreads();
    // d_x = dz1 W1
    if (g.d_x) {
      for (int i = t; i < nr * d0; i += HT) {
        const int r = i / d0, c = i - r * d0;
        float acc = 0.f;
#pragma unroll 4
        for (int o = 0; o < d1; ++o) acc = fmaf(dz1[r * d1 + o], w1[o * d0 + c], acc);
        g.d_x[(size_t)(r0 + r) * g.ld_dx + c] = acc;
      }
    }
    // weight / bias gradients: entry e of a matrix is owned by thread e % HT, rows added in order
#pragma unroll
    for (int q = 0; q < HE1; ++q) {
      const int e = t + q * HT;
      if (e < d1 * d0) {
        const int j = e / d0, k = e - j * d0;
        float acc = gw1[q];
        for (int r = 0; r < nr; ++r) acc = fmaf(dz1[r * d1 + j], xs[r * d0 + k], acc);
        gw1[q] = acc;
      }
    }
#pragma unroll
    for (int q = 0; q < HE2; ++q) {
      const int e = t + q * HT;
      if (e < d2 * d1) {
        const int j = e / d1, k = e - j * d1;
        float acc = gw2[q];
        for (int r = 0; r < nr; ++r) acc = fmaf(dz2[r * d2 + j], a1s[r * d1 + k], acc);
        gw2[q] = acc;
      }
    }
#pragma unroll
    for (int q = 0; q < HE3; ++q) {
      const int e = t + q * HT;
      if (e < dout * d2) {
        const int j = e / d2, k = e - j * d2;
        float acc = gw3[q];
        for (int r = 0; r < nr; ++r) acc = fmaf(dys[r * dout + j], a2s[r * d2 + k], acc);
        gw3[q] = acc;
      }
    }
    if (t < d1) { for (int r = 0; r < nr; ++r) gb += dz1[r * d1 + t]; }
    else if (t < d1 + d2) { for (int r = 0; r < nr; ++r) gb += dz2[r * d2 + (t - d1)]; }
    else if (t < d1 + d2 + dout) { for (int r = 0; r < nr; ++r) gb += dys[r * dout + (t - d1 - d2)]; }
  }
  // partial gradients -> this CTA's shared memory, then rank 0 adds the cluster's partials in rank order (DSMEM)
  const int o2 = d1 * d0, o3 = o2 + d2 * d1, ob = o3 + dout * d2, n_par = ob + d1 + d2 + dout;
#pragma unroll
  for (int q = 0; q < HE1; ++q) { const int e = t + q * HT; if (e < d1 * d0) part[e] = gw1[q]; }
#pragma unroll
  for (int q = 0; q < HE2; ++q) { const int e = t + q * HT; if (e < d2 * d1) part[o2 + e] = gw2[q]; }
#pragma unroll
  for (int q = 0; q < HE3; ++q) { const int e = t + q * HT; if (e < dout * d2) part[o3 + e] = gw3[q]; }
  if (t < d1 + d2 + dout) part[ob + t] = gb;
  cluster.sync();
  if (rank == 0) {
    const float* peer[HC];
#pragma unroll
    for (int r = 0; r < HC; ++r) peer[r] = cluster.map_shared_rank(part, r);
    for (int e = t; e < n_par; e += HT) {
      float v = 0.f;
#pragma unroll
      for (int r = 0; r < HC; ++r) v += peer[r][e];
      float* dst;
      if (e < o2) dst = g.d_w1 ? g.d_w1 + e : nullptr;
      else if (e < o3) dst = g.d_w2 ? g.d_w2 + (e - o2) : nullptr;
      else if (e < ob) dst = g.d_w3 ? g.d_w3 + (e - o3) : nullptr;
      else if (e < ob + d1) dst = g.d_b1 ? g.d_b1 + (e - ob) : nullptr;
      else if (e < ob + d1 + d2) dst = g.d_b2 ? g.d_b2 + (e - ob - d1) : nullptr;
      else dst = g.d_b3 ? g.d_b3 + (e - ob - d1 - d2) : nullptr;
      if (dst) *dst = g.accumulate ? *dst + v : v;
    }
  }
  cluster.sync();                                     // peers keep their shared memory alive until rank 0 has read it
}

static inline size_t head_bwd_smem_bytes(int d0, int d1, int d2, int dout) {
  return sizeof(float) * (2 * ((size_t)d1 * d0 + (size_t)d2 * d1 + (size_t)dout * d2) + d1 + d2 + dout +
                          (size_t)HR * (d0 + 2 * d1 + 2 * d2 + dout) + 4);
}

static int head_check(const DgnHeadArgs* a) {
  if (!a || a->n_rows < 0 || a->d0 <= 0 || a->d1 <= 0 || a->d2 <= 0 || a->d_out <= 0) return DGN_ERR_INVALID;
  if (!a->x || !a->w1 || !a->b1 || !a->w2 || !a->b2 || !a->w3 || !a->b3 || !a->a1 || !a->a2 || !a->y) return DGN_ERR_INVALID;
  if (a->n_rows > DGN_HEAD_MAX_ROWS) return DGN_ERR_UNSUPPORTED;
  if (a->d1 * a->d0 > HT * HE1 || a->d2 * a->d1 > HT * HE2 || a->d_out * a->d2 > HT * HE3 ||
      a->d1 + a->d2 + a->d_out > HT)
    return DGN_ERR_UNSUPPORTED;
  if (head_smem_bytes(a->d0, a->d1, a->d2, a->d_out) > 160 * 1024 ||
      head_bwd_smem_bytes(a->d0, a->d1, a->d2, a->d_out) > 160 * 1024)
    return DGN_ERR_UNSUPPORTED;
  return DGN_OK;
}

static int head_done() {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_dgn_last_cuda = e; return DGN_ERR_CUDA; }
  return DGN_OK;
}

}  // namespace dgn

using namespace dgn;

extern "C" int dgn_head_forward(const DgnHeadArgs* a, void* stream) {
  if (int rc = head_check(a)) return rc;
  if (a->n_rows == 0) return DGN_OK;
  const size_t smem = head_smem_bytes(a->d0, a->d1, a->d2, a->d_out);
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024) != cudaSuccess)
    return DGN_ERR_CUDA;
  const int tiles = (a->n_rows + HR - 1) / HR;
  launch_pdl(head_fwd_kernel, dim3((unsigned)(tiles < HF ? tiles : HF)), dim3(HT), smem, (cudaStream_t)stream, *a);
  return head_done();
}

extern "C" int dgn_head_backward(const DgnHeadArgs* a, const DgnHeadGrad* g, void* stream) {
  if (int rc = head_check(a)) return rc;
  if (!g || !g->g_y) return DGN_ERR_INVALID;
  if (a->n_rows == 0) return DGN_OK;
  const size_t smem = head_bwd_smem_bytes(a->d0, a->d1, a->d2, a->d_out);
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024) != cudaSuccess)
    return DGN_ERR_CUDA;
  launch_pdl(head_bwd_kernel, dim3(HC), dim3(HT), smem, (cudaStream_t)stream, *a, *g);      // one cluster
  return head_done();
}

extern "C" int dgn_l1_loss_forward(int32_t n, const float* y, const float* target, float* loss, void* stream) {
  if (n <= 0 || !y || !target || !loss) return DGN_ERR_INVALID;
  launch_pdl(l1_fwd_kernel, dim3(1), dim3(HT), 0, (cudaStream_t)stream, (int)n, y, target, loss);
  return head_done();
}

extern "C" int dgn_l1_loss_backward(int32_t n, const float* y, const float* target, const float* g_loss, float* d_y,
                                    void* stream) {
  if (n <= 0 || !y || !target || !g_loss || !d_y) return DGN_ERR_INVALID;
  launch_pdl(l1_bwd_kernel, dim3((unsigned)((n + HT - 1) / HT)), dim3(HT), 0, (cudaStream_t)stream, (int)n, y, target, g_loss,
             d_y);
  return head_done();
}
