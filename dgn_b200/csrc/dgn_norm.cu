// Layer epilogue and graph readout kernels (sm_100a).
//
// dgn_norm_forward / backward fuse rb/nets/dgn_layer.py:122-130
//     h = h * snorm_n ; h = BatchNorm1d(h) ; h = relu(h) ; h = h_in + h
// into two launches per direction: a column-statistics pass (Welford partials per row slab,
// merged in a fixed order - no atomics, so results are reproducible) and an apply pass.  The
// problem is a tall-skinny [N, C<=~300] matrix: threads map to columns (coalesced 128 B per
// warp row) and row slabs map to blocks.
//
// dgn_readout_* replace dgl.{sum,mean,max}_nodes: node rows of one graph are contiguous in the
// batched graph, so a readout is a segmented column reduction.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/dgn_b200.h"
#include "dgn_launch.cuh"

namespace dgn {

constexpr int kParts = 64;     // row slabs (fixed so that the workspace size only depends on C)
constexpr int kTX = 32, kTY = 8;

__device__ __forceinline__ int rows_of(const DgnNormArgs& a) { return a.n_rows_dev ? *a.n_rows_dev : a.n_rows; }

__device__ __forceinline__ void slab(int n, int part, int& r0, int& r1) {
  const int per = (n + kParts - 1) / kParts;
  r0 = min(part * per, n);
  r1 = min(r0 + per, n);
}

// Chan et al. merge of (n_a, mean_a, M2_a) with (n_b, mean_b, M2_b)
__device__ __forceinline__ void welford_merge(float& n, float& mean, float& m2, float nb, float mb, float m2b) {
  if (nb == 0.f) return;
  const float nt = n + nb;
  const float d = mb - mean;
  mean += d * (nb / nt);
  m2 += m2b + d * d * (n * nb / nt);
  n = nt;
}

// partial statistics of z = y * snorm for one (row slab, 32-column tile)
__global__ void __launch_bounds__(kTX * kTY) norm_stats_kernel(const DgnNormArgs a) {
  pdl_prologue();
  const int n = rows_of(a);
  const int col = blockIdx.y * kTX + threadIdx.x;
  int r0, r1;
  slab(n, blockIdx.x, r0, r1);
  float cnt = 0.f, mean = 0.f, m2 = 0.f;
  if (col < a.n_cols) {
    const float yb = a.y_bias ? a.y_bias[col] : 0.f;
    for (int r = r0 + threadIdx.y; r < r1; r += kTY) {
      float z = a.y[(size_t)r * a.ld_y + col] + yb;
      if (a.snorm) z *= a.snorm[r];
      cnt += 1.f;
      const float d = z - mean;
      mean += d / cnt;
      m2 += d * (z - mean);
    }
  }
  __shared__ float sn[kTY][kTX], sm[kTY][kTX], s2[kTY][kTX];
  sn[threadIdx.y][threadIdx.x] = cnt;
  sm[threadIdx.y][threadIdx.x] = mean;
  s2[threadIdx.y][threadIdx.x] = m2;
  __syncthreads();
  if (threadIdx.y == 0 && col < a.n_cols) {
    for (int j = 1; j < kTY; ++j) welford_merge(cnt, mean, m2, sn[j][threadIdx.x], sm[j][threadIdx.x], s2[j][threadIdx.x]);
    float* part = a.stats + 2 * a.n_cols + (size_t)blockIdx.x * 2 * a.n_cols;
    part[col] = mean;
    part[a.n_cols + col] = m2;
  }
}

// Merges the kParts slab partials of the block's 32 columns: every (tx, ty) thread first folds the
// partials ty, ty+kTY, ... (independent loads, issued together), then ty == 0 folds the kTY results
// in a fixed order.  Result: batch mean / biased variance in s_mean / s_var (valid for ty == 0 readers
// after the __syncthreads inside).
__device__ __forceinline__ void merged_stats(const DgnNormArgs& a, int n, int col, float (&s_cnt)[kTY][kTX],
                                             float (&s_mu)[kTY][kTX], float (&s_m2)[kTY][kTX], float& mean,
                                             float& var) {
  constexpr int PER = kParts / kTY;
  float cnt = 0.f, mu = 0.f, m2 = 0.f;
  if (a.stat_parts > 0) {
    // slabs written by dgn_post_forward (explicit counts): [cnt | mean | M2][C] each, merged in slab order
    for (int p0 = threadIdx.y; p0 < a.stat_parts; p0 += 4 * kTY) {
      float pn[4], pm[4], p2[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {                    // independent loads, issued together
        const int p = p0 + j * kTY;
        const bool ok = p < a.stat_parts && col < a.n_cols;
        const float* part = a.stats + 2 * a.n_cols + (size_t)p * 3 * a.n_cols;
        pn[j] = ok ? part[col] : 0.f;
        pm[j] = ok ? part[a.n_cols + col] : 0.f;
        p2[j] = ok ? part[2 * a.n_cols + col] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) welford_merge(cnt, mu, m2, pn[j], pm[j], p2[j]);
    }
  } else {
    float pm[PER], p2[PER], pn[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int p = threadIdx.y + j * kTY;
      int r0, r1;
      slab(n, p, r0, r1);
      const float* part = a.stats + 2 * a.n_cols + (size_t)p * 2 * a.n_cols;
      pn[j] = (float)(r1 - r0);
      pm[j] = (col < a.n_cols) ? part[col] : 0.f;
      p2[j] = (col < a.n_cols) ? part[a.n_cols + col] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < PER; ++j) welford_merge(cnt, mu, m2, pn[j], pm[j], p2[j]);
  }
  s_cnt[threadIdx.y][threadIdx.x] = cnt;
  s_mu[threadIdx.y][threadIdx.x] = mu;
  s_m2[threadIdx.y][threadIdx.x] = m2;
  __syncthreads();
  if (threadIdx.y == 0) {
    for (int j = 1; j < kTY; ++j) welford_merge(cnt, mu, m2, s_cnt[j][threadIdx.x], s_mu[j][threadIdx.x], s_m2[j][threadIdx.x]);
  }
  mean = mu;
  var = (cnt > 0.f) ? m2 / cnt : 0.f;
}

__global__ void __launch_bounds__(kTX * kTY) norm_apply_kernel(const DgnNormArgs a) {
  pdl_prologue();
  const int n = rows_of(a);
  const int col = blockIdx.y * kTX + threadIdx.x;
  __shared__ float s_mean[kTX], s_rstd[kTX];
  __shared__ float s_cnt[kTY][kTX], s_mu[kTY][kTX], s_m2[kTY][kTX];
  if (a.gamma) {
    float mean = 0.f, var = 1.f;
    if (a.training) merged_stats(a, n, col, s_cnt, s_mu, s_m2, mean, var);     // all threads (has a barrier)
    if (threadIdx.y == 0 && col < a.n_cols) {
      if (a.training) {
        if (blockIdx.x == 0 && a.running_mean) {      // nn.BatchNorm1d: unbiased variance in the running estimate
          const float unb = (n > 1) ? var * ((float)n / (float)(n - 1)) : var;
          a.running_mean[col] = (1.f - a.momentum) * a.running_mean[col] + a.momentum * mean;
          a.running_var[col] = (1.f - a.momentum) * a.running_var[col] + a.momentum * unb;
        }
      } else {
        mean = a.running_mean[col];
        var = a.running_var[col];
      }
      const float rstd = 1.f / sqrtf(var + a.eps);
      s_mean[threadIdx.x] = mean;
      s_rstd[threadIdx.x] = rstd;
      if (blockIdx.x == 0) {
        a.stats[col] = mean;
        a.stats[a.n_cols + col] = rstd;
      }
    }
  }
  __syncthreads();
  if (col >= a.n_cols) return;
  const float mean = a.gamma ? s_mean[threadIdx.x] : 0.f;
  const float rstd = a.gamma ? s_rstd[threadIdx.x] : 1.f;
  const float ga = a.gamma ? a.gamma[col] : 1.f, be = a.gamma ? a.beta[col] : 0.f;
  const float yb = a.y_bias ? a.y_bias[col] : 0.f;
  const int rows_cap = a.n_rows;
  for (int r = blockIdx.x * kTY + threadIdx.y; r < rows_cap; r += gridDim.x * kTY) {
    float o = 0.f;
    if (r < n) {
      float z = a.y[(size_t)r * a.ld_y + col] + yb;
      if (a.snorm) z *= a.snorm[r];
      o = (z - mean) * rstd * ga + be;
      if (a.relu) o = fmaxf(o, 0.f);
      if (a.residual) o += a.residual[(size_t)r * a.ld_res + col];
    }
    a.out[(size_t)r * a.ld_o + col] = o;
  }
}

// g1 = g_out * relu'(.) ; partial sums of g1 and g1 * xhat per (row slab, column)
__device__ __forceinline__ float masked_grad(const DgnNormArgs& a, const DgnNormGrad& g, int r, int col, float mean,
                                             float rstd, float ga, float be, float yb, float& xhat) {
  float z = a.y[(size_t)r * a.ld_y + col] + yb;
  if (a.snorm) z *= a.snorm[r];
  xhat = (z - mean) * rstd;
  float go = g.g_out[(size_t)r * g.ld_go + col];
  if (a.relu && !(xhat * ga + be > 0.f)) go = 0.f;
  return go;
}

constexpr int kBwdSums = 5;   // sum g1, sum g1*xhat, sum s*g1, sum s, sum s*xhat   (s = snorm_n)

// The five column sums are accumulated and merged in fp64: d_bias (and to a lesser degree d_gamma / d_beta)
// are small differences of large sums, and fp32 partial sums would leave ~1e-5 relative noise in them.
__global__ void __launch_bounds__(kTX * kTY) norm_bwd_reduce_kernel(const DgnNormArgs a, const DgnNormGrad g) {
  pdl_prologue();
  const int n = rows_of(a);
  const int col = blockIdx.y * kTX + threadIdx.x;
  int r0, r1;
  slab(n, blockIdx.x, r0, r1);
  double acc[kBwdSums] = {0.0, 0.0, 0.0, 0.0, 0.0};
  if (col < a.n_cols) {
    const float mean = a.gamma ? a.stats[col] : 0.f, rstd = a.gamma ? a.stats[a.n_cols + col] : 1.f;
    const float ga = a.gamma ? a.gamma[col] : 1.f, be = a.gamma ? a.beta[col] : 0.f;
    const float yb = a.y_bias ? a.y_bias[col] : 0.f;
    for (int r = r0 + threadIdx.y; r < r1; r += kTY) {
      float xhat;
      const float g1 = masked_grad(a, g, r, col, mean, rstd, ga, be, yb, xhat);
      const float sn = a.snorm ? a.snorm[r] : 1.f;
      acc[0] += (double)g1;
      acc[1] += (double)g1 * (double)xhat;
      acc[2] += (double)sn * (double)g1;
      acc[3] += (double)sn;
      acc[4] += (double)sn * (double)xhat;
    }
  }
  __shared__ double sh[kBwdSums][kTY][kTX];
#pragma unroll
  for (int q = 0; q < kBwdSums; ++q) sh[q][threadIdx.y][threadIdx.x] = acc[q];
  __syncthreads();
  if (threadIdx.y == 0 && col < a.n_cols) {
    double* part = reinterpret_cast<double*>(g.scratch) + (size_t)blockIdx.x * kBwdSums * a.n_cols;
#pragma unroll
    for (int q = 0; q < kBwdSums; ++q) {
      double t = acc[q];
      for (int j = 1; j < kTY; ++j) t += sh[q][j][threadIdx.x];
      part[q * a.n_cols + col] = t;
    }
  }
}

__global__ void __launch_bounds__(kTX * kTY) norm_bwd_apply_kernel(const DgnNormArgs a, const DgnNormGrad g) {
  pdl_prologue();
  const int n = rows_of(a);
  const int col = blockIdx.y * kTX + threadIdx.x;
  __shared__ float s_b[kTX], s_g[kTX];
  __shared__ double sh[kBwdSums][kTY][kTX];
  {
    constexpr int PER = kParts / kTY;
    double acc[kBwdSums] = {0.0, 0.0, 0.0, 0.0, 0.0};
    if (col < a.n_cols) {
      double v[PER][kBwdSums];
#pragma unroll
      for (int j = 0; j < PER; ++j) {                  // independent loads, fixed summation order
        const double* part = reinterpret_cast<const double*>(g.scratch) +
                             (size_t)(threadIdx.y + j * kTY) * kBwdSums * a.n_cols;
#pragma unroll
        for (int q = 0; q < kBwdSums; ++q) v[j][q] = part[q * a.n_cols + col];
      }
#pragma unroll
      for (int j = 0; j < PER; ++j)
#pragma unroll
        for (int q = 0; q < kBwdSums; ++q) acc[q] += v[j][q];
    }
#pragma unroll
    for (int q = 0; q < kBwdSums; ++q) sh[q][threadIdx.y][threadIdx.x] = acc[q];
    __syncthreads();
    if (threadIdx.y == 0 && col < a.n_cols) {
#pragma unroll
      for (int q = 0; q < kBwdSums; ++q)
        for (int j = 1; j < kTY; ++j) acc[q] += sh[q][j][threadIdx.x];
      const double inv_n = (n > 0) ? 1.0 / (double)n : 0.0;
      s_b[threadIdx.x] = (float)(acc[0] * inv_n);      // mean of g1
      s_g[threadIdx.x] = (float)(acc[1] * inv_n);      // mean of g1 * xhat
      if (blockIdx.x == 0) {
        // (select, not 0 * old: the destination of a non-accumulating call is uninitialised memory, possibly NaN)
        if (a.gamma) {
          if (g.d_beta) g.d_beta[col] = (g.accumulate ? g.d_beta[col] : 0.f) + (float)acc[0];
          if (g.d_gamma) g.d_gamma[col] = (g.accumulate ? g.d_gamma[col] : 0.f) + (float)acc[1];
        }
        if (g.d_bias) {
          // d_bias = sum_r d_y[r] with d_y = s * ga*rstd*(g1 - mb - xhat*mg) (training BN), s*ga*rstd*g1 (eval), s*g1 (no BN)
          double db = acc[2];
          if (a.gamma) {
            if (a.training) db = acc[2] - (acc[0] * inv_n) * acc[3] - (acc[1] * inv_n) * acc[4];
            db *= (double)a.gamma[col] * (double)a.stats[a.n_cols + col];
          }
          g.d_bias[col] = (g.accumulate ? g.d_bias[col] : 0.f) + (float)db;
        }
      }
    }
  }
  __syncthreads();
  if (col >= a.n_cols) return;
  const float mean = a.gamma ? a.stats[col] : 0.f, rstd = a.gamma ? a.stats[a.n_cols + col] : 1.f;
  const float ga = a.gamma ? a.gamma[col] : 1.f, be = a.gamma ? a.beta[col] : 0.f;
  const float yb = a.y_bias ? a.y_bias[col] : 0.f;
  const float mb = (a.gamma && a.training) ? s_b[threadIdx.x] : 0.f;
  const float mg = (a.gamma && a.training) ? s_g[threadIdx.x] : 0.f;
  for (int r = blockIdx.x * kTY + threadIdx.y; r < a.n_rows; r += gridDim.x * kTY) {
    float dy = 0.f, dres = 0.f;
    if (r < n) {
      float xhat;
      const float g1 = masked_grad(a, g, r, col, mean, rstd, ga, be, yb, xhat);
      float dz = a.gamma ? ga * rstd * (g1 - mb - xhat * mg) : g1;
      if (a.snorm) dz *= a.snorm[r];
      dy = dz;
      dres = g.g_out[(size_t)r * g.ld_go + col];
    }
    g.d_y[(size_t)r * g.ld_dy + col] = dy;
    if (g.d_residual) g.d_residual[(size_t)r * g.ld_dres + col] = dres;
  }
}

// ---- embedding gradient --------------------------------------------------------------------------
// grid = (column tiles of 32, kEmbSlabs row slabs); block = 32 columns x 8 row-lanes.  Every (row-lane, vocab
// entry, column) cell of the shared tile is owned by exactly one thread and the slab partials are merged in
// slab order by the last block of each column tile: the result does not depend on scheduling.
constexpr int kEmbSlabs = 32;

__global__ void __launch_bounds__(kTX * kTY) embedding_bwd_kernel(int n_rows, int C, int vocab,
                                                                  const long long* __restrict__ idx,
                                                                  const float* __restrict__ gsrc, int ld_g,
                                                                  float* __restrict__ dw, int ld_w,
                                                                  const int32_t* __restrict__ n_rows_dev,
                                                                  float* __restrict__ ws, unsigned* __restrict__ counters) {
  pdl_prologue();
  extern __shared__ float tile[];                      // [kTY][vocab][kTX]
  __shared__ bool is_last;
  const int n = n_rows_dev ? *n_rows_dev : n_rows;
  const int col = blockIdx.x * kTX + threadIdx.x;
  const int per = (n + kEmbSlabs - 1) / kEmbSlabs;
  const int r0 = min((int)blockIdx.y * per, n), r1 = min(r0 + per, n);
  float* mine = tile + (size_t)threadIdx.y * vocab * kTX + threadIdx.x;
  for (int t = 0; t < vocab; ++t) mine[t * kTX] = 0.f;
  if (col < C) {
    for (int r = r0 + threadIdx.y; r < r1; r += kTY) {
      const long long t = idx[r];
      if (t >= 0 && t < vocab) mine[(int)t * kTX] += gsrc[(size_t)r * ld_g + col];
    }
  }
  __syncthreads();
  float* part = ws + (size_t)blockIdx.y * vocab * C;
  if (col < C) {
    for (int t = threadIdx.y; t < vocab; t += kTY) {
      float acc = 0.f;
      for (int j = 0; j < kTY; ++j) acc += tile[((size_t)j * vocab + t) * kTX + threadIdx.x];
      part[(size_t)t * C + col] = acc;
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    const unsigned done = atomicAdd(&counters[blockIdx.x], 1u);
    is_last = (done == kEmbSlabs - 1);
    if (is_last) counters[blockIdx.x] = 0u;            // ready for the next launch
  }
  __syncthreads();
  if (!is_last || col >= C) return;
  __threadfence();
  for (int t = threadIdx.y; t < vocab; t += kTY) {
    float v[kEmbSlabs];
#pragma unroll
    for (int sl = 0; sl < kEmbSlabs; ++sl) v[sl] = __ldcg(ws + ((size_t)sl * vocab + t) * C + col);   // all in flight
    float acc = 0.f;
#pragma unroll
    for (int sl = 0; sl < kEmbSlabs; ++sl) acc += v[sl];                                              // slab order
    dw[(size_t)t * ld_w + col] += acc;
  }
}

// ---- readout -------------------------------------------------------------------------------------
__global__ void readout_fwd_kernel(int n_graphs, const int32_t* __restrict__ gp, int C, const float* __restrict__ h,
                                   int ld_h, int op, float* __restrict__ out, int ld_o) {
  pdl_prologue();
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int gi = blockIdx.y;
  if (col >= C || gi >= n_graphs) return;
  const int r0 = gp[gi], r1 = gp[gi + 1];
  float acc = (op == 2) ? -INFINITY : 0.f;
  for (int r = r0; r < r1; ++r) {
    const float v = h[(size_t)r * ld_h + col];
    acc = (op == 2) ? fmaxf(acc, v) : acc + v;
  }
  if (op == 1) acc = acc / (float)max(r1 - r0, 1);
  if (r1 == r0) acc = 0.f;
  out[(size_t)gi * ld_o + col] = acc;
}

__global__ void readout_bwd_kernel(int n_graphs, const int32_t* __restrict__ gp, int C, const float* __restrict__ h,
                                   int ld_h, const float* __restrict__ out, int ld_o, int op,
                                   const float* __restrict__ g_out, int ld_go, float* __restrict__ d_h, int ld_dh,
                                   int n_rows_total) {
  pdl_prologue();
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int gi = blockIdx.y;
  if (col >= C) return;
  if (gi == n_graphs) {                               // padding rows after the last graph get a zero gradient
    for (int r = gp[n_graphs]; r < n_rows_total; ++r) d_h[(size_t)r * ld_dh + col] = 0.f;
    return;
  }
  const int r0 = gp[gi], r1 = gp[gi + 1];
  float g = g_out[(size_t)gi * ld_go + col];
  if (op == 1) g = g / (float)max(r1 - r0, 1);
  const float top = (op == 2) ? out[(size_t)gi * ld_o + col] : 0.f;
  bool given = false;
  for (int r = r0; r < r1; ++r) {
    float d = g;
    if (op == 2) {                                   // first arg-max gets the gradient
      d = 0.f;
      if (!given && h[(size_t)r * ld_h + col] == top) { d = g; given = true; }
    }
    d_h[(size_t)r * ld_dh + col] = d;
  }
}

// ---- Adam over one flat parameter buffer ---------------------------------------------------------
// torch.optim.Adam semantics (L2 weight decay folded into the gradient, bias-corrected moments).  The step
// counter lives in device memory so the launch can be replayed from a CUDA graph: every block reads it on
// entry, the last block to finish increments it.
__global__ void __launch_bounds__(256) adam_kernel(long long n, float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, float lr, float b1,
                                                   float b2, float eps, float wd, const float* __restrict__ hyper,
                                                   int* __restrict__ state) {
  pdl_prologue();
  // hyper (device, optional) = {lr, weight_decay, grad_scale}: a captured CUDA graph freezes by-value arguments, so a
  // scheduler (ReduceLROnPlateau, rb/main_molecules.py:89-130) changes the step size by writing device memory
  float gs = 1.f;
  if (hyper) { lr = hyper[0]; wd = hyper[1]; gs = hyper[2]; }
  const int t = state[0] + 1;
  const float c1 = 1.f - powf(b1, (float)t), c2 = 1.f - powf(b2, (float)t);
  const float step_size = lr / c1, inv_sqrt_c2 = rsqrtf(c2);
  const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 + 3 < n) {
    float4 pp = *reinterpret_cast<float4*>(p + i4), mm = *reinterpret_cast<float4*>(m + i4),
           vv = *reinterpret_cast<float4*>(v + i4);
    const float4 gg = *reinterpret_cast<const float4*>(g + i4);
    float* pa = &pp.x; float* ma = &mm.x; float* va = &vv.x; const float* ga = &gg.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = fmaf(wd, pa[j], gs * ga[j]);
      ma[j] = fmaf(b1, ma[j], (1.f - b1) * gr);
      va[j] = fmaf(b2, va[j], (1.f - b2) * gr * gr);
      pa[j] -= step_size * ma[j] / (sqrtf(va[j]) * inv_sqrt_c2 + eps);
    }
    *reinterpret_cast<float4*>(p + i4) = pp;
    *reinterpret_cast<float4*>(m + i4) = mm;
    *reinterpret_cast<float4*>(v + i4) = vv;
  } else {
    for (long long i = i4; i < n; ++i) {
      const float gr = fmaf(wd, p[i], gs * g[i]);
      m[i] = fmaf(b1, m[i], (1.f - b1) * gr);
      v[i] = fmaf(b2, v[i], (1.f - b2) * gr * gr);
      p[i] -= step_size * m[i] / (sqrtf(v[i]) * inv_sqrt_c2 + eps);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int done = atomicAdd(&state[1], 1);
    if (done == (int)gridDim.x - 1) { state[1] = 0; state[0] = t; }
  }
}

}  // namespace dgn

using namespace dgn;

extern thread_local cudaError_t g_dgn_last_cuda;
static int check_launch() {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_dgn_last_cuda = e; return DGN_ERR_CUDA; }
  return DGN_OK;
}

static dim3 norm_grid_apply(const DgnNormArgs* a) {
  // every block pays a fixed prologue (merging the 64 slab partials of its 32 columns), so the grid is
  // sized to ~2 blocks per SM and the rows are covered by a grid-stride loop
  const int col_tiles = (a->n_cols + kTX - 1) / kTX;
  int gx = (a->n_rows + kTY - 1) / kTY;
  const int cap = (296 + col_tiles - 1) / col_tiles;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  return dim3((unsigned)gx, (unsigned)col_tiles);
}

extern "C" int dgn_norm_forward(const DgnNormArgs* a, void* stream) {
  if (!a || !a->y || !a->out || a->n_rows < 0 || a->n_cols <= 0 || a->stat_parts < 0) return DGN_ERR_INVALID;
  if (a->gamma && (!a->beta || !a->stats)) return DGN_ERR_INVALID;
  if (a->gamma && !a->training && (!a->running_mean || !a->running_var)) return DGN_ERR_INVALID;
  if (a->n_rows == 0) return DGN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 block(kTX, kTY);
  if (a->gamma && a->training && a->stat_parts == 0) {
    launch_pdl(norm_stats_kernel, dim3(kParts, (a->n_cols + kTX - 1) / kTX), block, 0, st, *a);
    if (int rc = check_launch()) return rc;
  }
  launch_pdl(norm_apply_kernel, norm_grid_apply(a), block, 0, st, *a);
  return check_launch();
}

extern "C" int dgn_norm_backward(const DgnNormArgs* a, const DgnNormGrad* g, void* stream) {
  if (!a || !g || !a->y || !g->g_out || !g->d_y || !g->scratch || a->n_cols <= 0) return DGN_ERR_INVALID;
  if (a->gamma && !a->stats) return DGN_ERR_INVALID;
  if (a->n_rows == 0) return DGN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 block(kTX, kTY);
  launch_pdl(norm_bwd_reduce_kernel, dim3(kParts, (a->n_cols + kTX - 1) / kTX), block, 0, st, *a, *g);
  if (int rc = check_launch()) return rc;
  launch_pdl(norm_bwd_apply_kernel, norm_grid_apply(a), block, 0, st, *a, *g);
  return check_launch();
}

extern "C" int dgn_embedding_backward(int32_t n_rows, int32_t n_cols, int32_t vocab, const int64_t* idx,
                                      const float* g, int32_t ld_g, float* d_weight, int32_t ld_w,
                                      const int32_t* n_rows_dev, float* ws, void* stream) {
  if (n_rows < 0 || n_cols <= 0 || vocab <= 0 || !idx || !g || !d_weight || !ws) return DGN_ERR_INVALID;
  const size_t smem = (size_t)kTY * vocab * kTX * sizeof(float);
  if (smem > 200 * 1024) return DGN_ERR_UNSUPPORTED;
  if (n_rows == 0) return DGN_OK;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(embedding_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  unsigned* counters = reinterpret_cast<unsigned*>(ws + (size_t)kEmbSlabs * vocab * n_cols);
  launch_pdl(embedding_bwd_kernel, dim3((n_cols + kTX - 1) / kTX, kEmbSlabs), dim3(kTX, kTY), smem, (cudaStream_t)stream, n_rows, n_cols, vocab, reinterpret_cast<const long long*>(idx), g, ld_g, d_weight, ld_w, n_rows_dev, ws, counters);
  return check_launch();
}

extern "C" int dgn_readout_forward(int32_t n_graphs, const int32_t* graph_ptr, int32_t n_cols, const float* h,
                                   int32_t ld_h, int32_t op, float* out, int32_t ld_o, void* stream) {
  if (n_graphs < 0 || n_cols <= 0 || !graph_ptr || !h || !out || op < 0 || op > 2) return DGN_ERR_INVALID;
  if (n_graphs == 0) return DGN_OK;
  const int block = 64;
  launch_pdl(readout_fwd_kernel, dim3((n_cols + block - 1) / block, n_graphs), block, 0, (cudaStream_t)stream, n_graphs, graph_ptr, n_cols, h, ld_h, op, out, ld_o);
  return check_launch();
}

extern "C" int dgn_readout_backward(int32_t n_graphs, const int32_t* graph_ptr, int32_t n_cols, const float* h,
                                    int32_t ld_h, const float* out, int32_t ld_o, int32_t op, const float* g_out,
                                    int32_t ld_go, float* d_h, int32_t ld_dh, int32_t n_rows_total, void* stream) {
  if (n_graphs < 0 || n_cols <= 0 || !graph_ptr || !g_out || !d_h || op < 0 || op > 2) return DGN_ERR_INVALID;
  if (op == 2 && (!h || !out)) return DGN_ERR_INVALID;
  if (n_graphs == 0) return DGN_OK;
  const int block = 64;
  launch_pdl(readout_bwd_kernel, dim3((n_cols + block - 1) / block, n_graphs + 1), block, 0, (cudaStream_t)stream, n_graphs, graph_ptr, n_cols, h, ld_h, out, ld_o, op, g_out, ld_go, d_h, ld_dh, n_rows_total);
  return check_launch();
}

extern "C" int dgn_adam_step(int64_t n, float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float lr,
                             float beta1, float beta2, float eps, float weight_decay, const float* hyper,
                             int32_t* state, void* stream) {
  if (n < 0 || !param || !grad || !exp_avg || !exp_avg_sq || !state) return DGN_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
       reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15u)
    return DGN_ERR_ALIGNMENT;
  if (n == 0) return DGN_OK;
  const long long threads = (n + 3) / 4;
  launch_pdl(adam_kernel, dim3((unsigned)((threads + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, n, param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, hyper, state);
  return check_launch();
}
