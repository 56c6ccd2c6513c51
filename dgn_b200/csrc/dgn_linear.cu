// Node-level halves of the 1-layer pretrans (rb/nets/dgn_layer.py:75-80), fp32 on CUDA cores.
//
// pretrans(cat(h_u, h_v)) = W_src h_u + W_dst h_v + b splits per node into P = h W_src^T and Q = h W_dst^T
// with W = [W_src | W_dst] stored as one [F_out, 2 F_in (+edge)] parameter.  These K = F_in (<= 128) products are
// launch-latency bound; doing both halves in ONE launch straight from the parameter's layout (no slicing, no
// second GEMM) halves their cost.  Same for the backward d_h += d_P W_src + d_Q W_dst.
// Tiles: 16 rows x 32 columns per 128-thread block (8 column groups x 16 rows: 1 x 4 outputs of each product per
// thread).  Round 1 used 32 x 64 tiles with 4 x 4 outputs per thread: 94 CTAs of 4 warps for the bench batch, i.e. less
// than one warp per scheduler on 2/3 of the SMs - ncu: 6 % warps active, 17 % issue slots used, 14 us for 49 MFLOP.
// These products are latency bound, so the tile is sized for warps in flight (376 CTAs = 10 warps per SM), not reuse.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dgn_b200.h"
#include "dgn_launch.cuh"

namespace dgn {

constexpr int LR = 16, LC = 32, LT = 128;     // rows / columns per block, threads
constexpr int TPAD = 4;                       // row padding of the row-major operand tiles (keeps rows 16 B aligned)

// Asynchronous global -> shared copies (LDGSTS): the staging loops issue all their copies back to back instead of
// dependent load -> store round trips - with 4 warps per CTA nothing else hides that latency.
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// W = [W_src | W_dst] is stored [F_out, 2 F_in (+ edge columns)].  The forward products keep the block
// W[c0 : c0 + LC, 0 : 2 F_in] in its stored orientation: row o at o * wld, wld = 2 F_in + 4, copied with 16-byte
// cp.async (8 per thread; round 1 transposed it with 4-byte copies - 64 per thread plus their index arithmetic were
// most of the kernel's instructions).  A thread owns the output columns tx, tx + 8, tx + 16, tx + 24 and walks k four
// at a time: one LDS.128 of its h row and one per column and half; wld / 4 is odd, so the 8 column groups of a warp hit
// 8 different 4-bank groups (conflict free), the 4 rows of a warp broadcast.
__device__ __forceinline__ void stage_w_block(float* wb, int wld, const float* __restrict__ W, int ld_w, int w_vec, int Fi,
                                              int Fo, int c0, int t) {
  const int q4 = (2 * Fi) / 4;
  for (int idx = t; idx < LC * q4; idx += LT) {
    const int o = idx / q4, i = (idx - o * q4) * 4;
    float* dst = wb + o * wld + i;
    if (c0 + o < Fo) {
      const float* src = W + (size_t)(c0 + o) * ld_w + i;
      if (w_vec) cp_async16(dst, src);
      else { cp_async4(dst, src); cp_async4(dst + 1, src + 1); cp_async4(dst + 2, src + 2); cp_async4(dst + 3, src + 3); }
    } else {
      dst[0] = 0.f; dst[1] = 0.f; dst[2] = 0.f; dst[3] = 0.f;
    }
  }
}

// both products of one thread: row `hrow` (shared memory, Fi floats) against the staged W block, columns tx + 8 b.
// Every accumulator is one fma chain in ascending k, like the reference's GEMM row.
__device__ __forceinline__ void pair_products(const float* hrow, const float* wb, int wld, int Fi, int tx,
                                              float (&ap)[4], float (&aq)[4]) {
#pragma unroll 2
  for (int i = 0; i < Fi; i += 4) {
    const float4 hv = *reinterpret_cast<const float4*>(hrow + i);
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const float* wr = wb + (tx + 8 * b) * wld + i;
      const float4 pv = *reinterpret_cast<const float4*>(wr);
      const float4 qv = *reinterpret_cast<const float4*>(wr + Fi);
      ap[b] = fmaf(hv.w, pv.w, fmaf(hv.z, pv.z, fmaf(hv.y, pv.y, fmaf(hv.x, pv.x, ap[b]))));
      aq[b] = fmaf(hv.w, qv.w, fmaf(hv.z, qv.z, fmaf(hv.y, qv.y, fmaf(hv.x, qv.x, aq[b]))));
    }
  }
}

// rows of P / Q: this thread's four columns are 8 apart
__device__ __forceinline__ void store_pq(float* __restrict__ P, int ld_p, float* __restrict__ Q, int ld_q, int r, int c0,
                                         int tx, int Fo, const float (&ap)[4], const float (&aq)[4]) {
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const int c = c0 + tx + 8 * b;
    if (c < Fo) {
      P[(size_t)r * ld_p + c] = ap[b];
      Q[(size_t)r * ld_q + c] = aq[b];
    }
  }
}

// P[n,o] = sum_i h[n,i] W[o,i] ;  Q[n,o] = sum_i h[n,i] W[o,Fi+i]
__global__ void __launch_bounds__(LT) pair_linear_fwd_kernel(int N, int Fi, int Fo, const float* __restrict__ h, int ld_h,
                                                             int h_vec, const float* __restrict__ W, int ld_w,
                                                             int w_vec, float* __restrict__ P, int ld_p,
                                                             float* __restrict__ Q, int ld_q) {
  pdl_prologue();
  extern __shared__ __align__(16) float sm[];
  const int hld = Fi + TPAD, wld = 2 * Fi + TPAD;
  float* hs = sm;                              // [LR][Fi + 4]    h tile, row-major
  float* wb = hs + LR * hld;                   // [LC][2 Fi + 4]  W block, rows = output columns
  const int t = threadIdx.x, r0 = blockIdx.x * LR, c0 = blockIdx.y * LC;
  const int q4 = Fi / 4;
  for (int idx = t; idx < LR * q4; idx += LT) {
    const int r = idx / q4, i = (idx - r * q4) * 4;
    float* dst = hs + r * hld + i;
    if (r0 + r < N) {
      const float* src = h + (size_t)(r0 + r) * ld_h + i;
      if (h_vec) cp_async16(dst, src);
      else { cp_async4(dst, src); cp_async4(dst + 1, src + 1); cp_async4(dst + 2, src + 2); cp_async4(dst + 3, src + 3); }
    } else {
      dst[0] = 0.f; dst[1] = 0.f; dst[2] = 0.f; dst[3] = 0.f;
    }
  }
  stage_w_block(wb, wld, W, ld_w, w_vec, Fi, Fo, c0, t);
  cp_async_wait_all();
  __syncthreads();
  const int tx = t & 7, ty = t >> 3;           // 8 column groups x 16 rows
  float ap[4] = {}, aq[4] = {};
  pair_products(hs + ty * hld, wb, wld, Fi, tx, ap, aq);
  if (r0 + ty < N) store_pq(P, ld_p, Q, ld_q, r0 + ty, c0, tx, Fo, ap, aq);
}

// Layer epilogue of layer l fused with the node-level pretrans halves of layer l+1 (one launch instead of
// norm_apply_kernel + pair_linear_fwd_kernel):
//   out = relu(BN((y + b) * snorm)) + residual      rb/nets/dgn_layer.py:122-130, statistics final in a.stats
//   P = out W_src^T, Q = out W_dst^T                rb/nets/dgn_layer.py:75-80 of the next layer
// The row tile is produced with coalesced 128-bit accesses (16 threads per row), written to `out` and kept in shared
// memory for the two products.
__global__ void __launch_bounds__(LT) norm_pair_fwd_kernel(const DgnNormArgs a, int Fo, const float* __restrict__ W, int ld_w,
                                                           int w_vec, float* __restrict__ P, int ld_p,
                                                           float* __restrict__ Q, int ld_q) {
  pdl_prologue();
  extern __shared__ __align__(16) float sm[];
  const int Fi = a.n_cols, hld = Fi + TPAD, wld = 2 * Fi + TPAD;
  float* hs = sm;                              // [LR][Fi + 4]    epilogue output tile, row-major
  float* wb = hs + LR * hld;                   // [LC][2 Fi + 4]  W block, rows = output columns
  float* cst = wb + LC * wld;                  // [5][Fi]         mean, rstd, gamma, beta, bias
  float* part = cst + 5 * Fi;                  // [4][3][Fi]      statistics merge scratch
  const int t = threadIdx.x, r0 = blockIdx.x * LR, c0 = blockIdx.y * LC;
  const int n = a.n_rows_dev ? *a.n_rows_dev : a.n_rows;
  stage_w_block(wb, wld, W, ld_w, w_vec, Fi, Fo, c0, t);
  // this thread's row operands are requested BEFORE the statistics merge (a chain of dependent loads and two barriers)
  constexpr int kItems = (LR * 32 + LT - 1) / LT;               // Fi <= 128: at most LR * 32 float4 items per tile
  const int q4 = Fi / 4;
  float4 yv[kItems], rv[kItems];
  float sn[kItems];
#pragma unroll
  for (int it = 0; it < kItems; ++it) {
    const int idx = t + it * LT, r = idx / q4, i = (idx - r * q4) * 4, row = r0 + r;
    yv[it] = rv[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    sn[it] = 1.f;
    if (idx < LR * q4 && row < n) {
      yv[it] = *reinterpret_cast<const float4*>(a.y + (size_t)row * a.ld_y + i);
      if (a.snorm) sn[it] = a.snorm[row];
      if (a.residual) rv[it] = *reinterpret_cast<const float4*>(a.residual + (size_t)row * a.ld_res + i);
    }
  }
  // per-column constants; training-mode BatchNorm: merge the statistics slabs dgn_post_forward left in a.stats (slab
  // order -> every CTA gets the same bits), block (0, 0) also updates the running statistics and the [mean | rstd] header
  const bool bn = a.gamma != nullptr, merge = bn && a.training;
  if (merge) {
    const int nsub = (Fi <= LT) ? ((LT / Fi) < 4 ? (LT / Fi) : 4) : 1;   // threads per column (Fi = 64: 2)
    for (int c0 = 0; c0 < Fi; c0 += LT) {
      const int col = c0 + t % (Fi < LT ? Fi : LT), sub = (Fi < LT) ? t / Fi : 0;
      float cnt = 0.f, mu = 0.f, m2 = 0.f;
      if (col < Fi && sub < nsub) {
        for (int p0 = sub; p0 < a.stat_parts; p0 += 8 * nsub) {
          float pn[8], pm[8], p2[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {                          // independent loads first
            const int p = p0 + j * nsub;
            const float* sl = a.stats + 2 * Fi + (size_t)p * 3 * Fi;
            const bool ok = p < a.stat_parts;
            pn[j] = ok ? sl[col] : 0.f;
            pm[j] = ok ? sl[Fi + col] : 0.f;
            p2[j] = ok ? sl[2 * Fi + col] : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (pn[j] > 0.f) {
              const float nt = cnt + pn[j], d = pm[j] - mu;
              mu += d * (pn[j] / nt);
              m2 += p2[j] + d * d * (cnt * pn[j] / nt);
              cnt = nt;
            }
          }
        }
        part[(sub * 3 + 0) * Fi + col] = cnt;
        part[(sub * 3 + 1) * Fi + col] = mu;
        part[(sub * 3 + 2) * Fi + col] = m2;
      }
      __syncthreads();
      if (col < Fi && sub == 0) {
        for (int j = 1; j < nsub; ++j) {
          const float nb = part[(j * 3 + 0) * Fi + col], mb = part[(j * 3 + 1) * Fi + col], m2b = part[(j * 3 + 2) * Fi + col];
          if (nb > 0.f) {
            const float nt = cnt + nb, d = mb - mu;
            mu += d * (nb / nt);
            m2 += m2b + d * d * (cnt * nb / nt);
            cnt = nt;
          }
        }
        const float var = cnt > 0.f ? m2 / cnt : 0.f, rstd = 1.f / sqrtf(var + a.eps);
        cst[col] = mu;
        cst[Fi + col] = rstd;
        if (blockIdx.x == 0 && blockIdx.y == 0) {
          a.stats[col] = mu;                                    // the backward reads the header
          a.stats[Fi + col] = rstd;
          if (a.running_mean) {                                 // nn.BatchNorm1d: unbiased variance in the running estimate
            const float unb = cnt > 1.f ? var * (cnt / (cnt - 1.f)) : var;
            a.running_mean[col] = (1.f - a.momentum) * a.running_mean[col] + a.momentum * mu;
            a.running_var[col] = (1.f - a.momentum) * a.running_var[col] + a.momentum * unb;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int i = t; i < Fi; i += LT) {
    if (!merge) {
      cst[i] = bn ? a.running_mean[i] : 0.f;
      cst[Fi + i] = bn ? 1.f / sqrtf(a.running_var[i] + a.eps) : 1.f;
    }
    cst[2 * Fi + i] = bn ? a.gamma[i] : 1.f;
    cst[3 * Fi + i] = bn ? a.beta[i] : 0.f;
    cst[4 * Fi + i] = a.y_bias ? a.y_bias[i] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < kItems; ++it) {
    const int idx = t + it * LT;
    if (idx >= LR * q4) break;
    const int r = idx / q4, i = (idx - r * q4) * 4, row = r0 + r;
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    if (row < n) {
      const float yy[4] = {yv[it].x, yv[it].y, yv[it].z, yv[it].w};
      const float rs[4] = {rv[it].x, rv[it].y, rv[it].z, rv[it].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float z = (yy[j] + cst[4 * Fi + i + j]) * sn[it];
        float v = (z - cst[i + j]) * cst[Fi + i + j] * cst[2 * Fi + i + j] + cst[3 * Fi + i + j];
        if (a.relu) v = fmaxf(v, 0.f);
        o[j] = v + rs[j];
      }
    }
    if (row < a.n_rows && blockIdx.y == 0)
      *reinterpret_cast<float4*>(a.out + (size_t)row * a.ld_o + i) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(hs + r * hld + i) = make_float4(o[0], o[1], o[2], o[3]);
  }
  cp_async_wait_all();
  __syncthreads();
  const int tx = t & 7, ty = t >> 3;           // 8 column groups x 16 rows
  float ap[4] = {}, aq[4] = {};
  pair_products(hs + ty * hld, wb, wld, Fi, tx, ap, aq);
  if (r0 + ty < a.n_rows) store_pq(P, ld_p, Q, ld_q, r0 + ty, c0, tx, Fo, ap, aq);
}

// The backward products read W in its stored orientation (k = output row): the column block W[:, c0 : c0 + LC] of each
// half goes to shared memory with 16-byte copies.
__device__ __forceinline__ void stage_w_rows(float* wsr, float* wds, const float* __restrict__ W, int ld_w, int w_vec, int Fi,
                                             int Fo, int c0, int t) {
  constexpr int C4 = LC / 4;
  for (int idx = t; idx < Fo * C4; idx += LT) {
    const int o = idx / C4, i = (idx - o * C4) * 4;
    float* d1 = wsr + o * LC + i;
    float* d2 = wds + o * LC + i;
    if (c0 + i < Fi) {                          // Fi % 4 == 0: the 4 columns are in or out together
      const float* s1 = W + (size_t)o * ld_w + c0 + i;
      const float* s2 = s1 + Fi;
      if (w_vec) {
        cp_async16(d1, s1);
        cp_async16(d2, s2);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) { cp_async4(d1 + j, s1 + j); cp_async4(d2 + j, s2 + j); }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) { d1[j] = 0.f; d2[j] = 0.f; }
    }
  }
}

// d_h[r, c0 + tx*4 .. +3] += sum_o dP[r,o] W[o,c] + dQ[r,o] W[o,Fi+c]   (one row, four columns per thread)
__device__ __forceinline__ void pair_products_bwd(const float* prow, const float* qrow, const float* wsr, const float* wds,
                                                  int Fo, int tx, float (&acc)[4]) {
#pragma unroll 8
  for (int o = 0; o < Fo; ++o) {
    const float pr = prow[o], qr = qrow[o];
    const float4 sv = *reinterpret_cast<const float4*>(wsr + o * LC + tx * 4);
    const float4 dv = *reinterpret_cast<const float4*>(wds + o * LC + tx * 4);
    acc[0] = fmaf(pr, sv.x, fmaf(qr, dv.x, acc[0]));
    acc[1] = fmaf(pr, sv.y, fmaf(qr, dv.y, acc[1]));
    acc[2] = fmaf(pr, sv.z, fmaf(qr, dv.z, acc[2]));
    acc[3] = fmaf(pr, sv.w, fmaf(qr, dv.w, acc[3]));
  }
}

__device__ __forceinline__ void add_row4(float* d_h, int ld_dh, int r, int c, const float (&acc)[4]) {
  float4* dst = reinterpret_cast<float4*>(d_h + (size_t)r * ld_dh + c);
  float4 v = *dst;
  v.x += acc[0]; v.y += acc[1]; v.z += acc[2]; v.w += acc[3];
  *dst = v;
}

// d_h[n,i] += sum_o dP[n,o] W[o,i] + dQ[n,o] W[o,Fi+i]
__global__ void __launch_bounds__(LT) pair_linear_bwd_kernel(int N, int Fi, int Fo, const float* __restrict__ dP, int ld_p,
                                                             const float* __restrict__ dQ, int ld_q,
                                                             const float* __restrict__ W, int ld_w, int w_vec,
                                                             float* __restrict__ d_h, int ld_dh) {
  pdl_prologue();
  extern __shared__ __align__(16) float sm[];
  const int pld = Fo + TPAD;
  float* ps = sm;                              // [LR][Fo + 4]  dP tile, row-major
  float* qs = ps + LR * pld;                   // [LR][Fo + 4]  dQ tile
  float* wsr = qs + LR * pld;                  // [Fo][LC]      W[:, c0:c0+LC]
  float* wds = wsr + Fo * LC;                  // [Fo][LC]      W[:, Fi+c0 : Fi+c0+LC]
  const int t = threadIdx.x, r0 = blockIdx.x * LR, c0 = blockIdx.y * LC;
  stage_w_rows(wsr, wds, W, ld_w, w_vec, Fi, Fo, c0, t);
  const int q4 = Fo / 4;
  for (int idx = t; idx < LR * q4; idx += LT) {          // dP / dQ rows are 16 B aligned (lin_ok)
    const int r = idx / q4, o = (idx - r * q4) * 4;
    if (r0 + r < N) {
      cp_async16(ps + r * pld + o, dP + (size_t)(r0 + r) * ld_p + o);
      cp_async16(qs + r * pld + o, dQ + (size_t)(r0 + r) * ld_q + o);
    } else {
      *reinterpret_cast<float4*>(ps + r * pld + o) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(qs + r * pld + o) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  cp_async_wait_all();
  __syncthreads();
  const int tx = t & 7, ty = t >> 3;
  float acc[4] = {};
  pair_products_bwd(ps + ty * pld, qs + ty * pld, wsr, wds, Fo, tx, acc);
  const int r = r0 + ty, c = c0 + tx * 4;
  if (r < N && c < Fi) add_row4(d_h, ld_dh, r, c, acc);
}

// pair_linear_bwd with the source-side reduction of the aggregation backward folded in (one launch instead of
// agg_bwd_src_kernel + pair_linear_bwd_kernel):
//   d_P[u]  = sum over the out-edges j of u of ws[out_slot[j]]      (deterministic gather, edge-slot order)
//   d_h[u] += d_P[u] W_src + d_Q[u] W_dst                           and d_P is written out for the weight gradient
// The row tile is produced with 128-bit accesses (16 threads per row) and kept row-major in shared memory.
__global__ void __launch_bounds__(LT) pair_gather_bwd_kernel(int N, int Fi, int Fo, const int32_t* __restrict__ out_ptr,
                                                             const int32_t* __restrict__ out_slot,
                                                             const float* __restrict__ ws, int ld_ws,
                                                             const float* __restrict__ dQ, int ld_q,
                                                             const float* __restrict__ W, int ld_w, int w_vec,
                                                             float* __restrict__ d_h, int ld_dh,
                                                             float* __restrict__ dP, int ld_p) {
  pdl_prologue();
  extern __shared__ __align__(16) float sm[];
  const int pld = Fo + TPAD;
  float* ps = sm;                              // [LR][Fo + 4]  gathered dP tile, row-major
  float* qs = ps + LR * pld;                   // [LR][Fo + 4]  dQ tile
  float* wsr = qs + LR * pld;                  // [Fo][LC]      W[:, c0:c0+LC]
  float* wds = wsr + Fo * LC;                  // [Fo][LC]      W[:, Fi+c0 : Fi+c0+LC]
  const int t = threadIdx.x, r0 = blockIdx.x * LR, c0 = blockIdx.y * LC;
  const int tx = t & 7, ty = t >> 3;
  const int orow = r0 + ty, ocol = c0 + tx * 4;
  float4 old = make_float4(0.f, 0.f, 0.f, 0.f);                 // d_h is accumulated into: request it first
  if (orow < N && ocol < Fi) old = *reinterpret_cast<const float4*>(d_h + (size_t)orow * ld_dh + ocol);
  stage_w_rows(wsr, wds, W, ld_w, w_vec, Fi, Fo, c0, t);
  const int q4 = Fo / 4;
  for (int idx = t; idx < LR * q4; idx += LT) {
    const int r = idx / q4, o = (idx - r * q4) * 4, u = r0 + r;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (u < N) {
      cp_async16(qs + r * pld + o, dQ + (size_t)u * ld_q + o);
      const int j0 = __ldg(out_ptr + u), j1 = __ldg(out_ptr + u + 1);
      for (int j = j0; j < j1; j += 4) {                         // up to 4 edge rows in flight
        float4 v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          v[e] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (j + e < j1) v[e] = __ldcs(reinterpret_cast<const float4*>(ws + (size_t)__ldg(out_slot + j + e) * ld_ws + o));
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) { acc.x += v[e].x; acc.y += v[e].y; acc.z += v[e].z; acc.w += v[e].w; }
      }
      if (blockIdx.y == 0) *reinterpret_cast<float4*>(dP + (size_t)u * ld_p + o) = acc;
    } else {
      *reinterpret_cast<float4*>(qs + r * pld + o) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    *reinterpret_cast<float4*>(ps + r * pld + o) = acc;
  }
  cp_async_wait_all();
  __syncthreads();
  float acc[4] = {};
  pair_products_bwd(ps + ty * pld, qs + ty * pld, wsr, wds, Fo, tx, acc);
  if (orow < N && ocol < Fi)
    *reinterpret_cast<float4*>(d_h + (size_t)orow * ld_dh + ocol) =
        make_float4(old.x + acc[0], old.y + acc[1], old.z + acc[2], old.w + acc[3]);
}

}  // namespace dgn

using namespace dgn;
extern thread_local cudaError_t g_dgn_last_cuda;

static bool lin_ok(int Fi, int Fo, const void* a, const void* b, int l1, int l2) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  return Fi > 0 && Fo > 0 && Fi <= 128 && Fo <= 128 && Fi % 4 == 0 && Fo % 4 == 0 && al(a) && al(b) && l1 % 4 == 0 &&
         l2 % 4 == 0;
}

template <typename K>
static int lin_attr(K kern, size_t smem) {
  if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
    return DGN_ERR_CUDA;
  return DGN_OK;
}

extern "C" int dgn_pair_linear_forward(int32_t N, int32_t Fi, int32_t Fo, const float* h, int32_t ld_h, const float* W,
                                       int32_t ld_w, float* P, int32_t ld_p, float* Q, int32_t ld_q, void* stream) {
  if (N < 0 || !h || !W || !P || !Q) return DGN_ERR_INVALID;
  if (!lin_ok(Fi, Fo, P, Q, ld_p, ld_q)) return DGN_ERR_UNSUPPORTED;
  if (N == 0) return DGN_OK;
  const size_t smem = (size_t)(LR * (Fi + TPAD) + LC * (2 * Fi + TPAD)) * sizeof(float);
  if (int rc = lin_attr(pair_linear_fwd_kernel, smem)) return rc;
  const int h_vec = (reinterpret_cast<uintptr_t>(h) & 15u) == 0 && ld_h % 4 == 0;
  const int w_vec = (reinterpret_cast<uintptr_t>(W) & 15u) == 0 && ld_w % 4 == 0;
  launch_pdl(pair_linear_fwd_kernel, dim3((N + LR - 1) / LR, (Fo + LC - 1) / LC), dim3(LT), smem, (cudaStream_t)stream, N, Fi, Fo, h, ld_h, h_vec, W, ld_w, w_vec, P, ld_p, Q, ld_q);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_dgn_last_cuda = e; return DGN_ERR_CUDA; }
  return DGN_OK;
}

extern "C" int dgn_norm_pair_forward(const DgnNormArgs* a, int32_t f_out, const float* w, int32_t ld_w, float* p,
                                     int32_t ld_p, float* q, int32_t ld_q, void* stream) {
  if (!a || !a->y || !a->out || !w || !p || !q || a->n_rows < 0 || a->n_cols <= 0) return DGN_ERR_INVALID;
  if (a->gamma && (!a->beta || !a->stats)) return DGN_ERR_INVALID;
  if (a->gamma && a->training && a->stat_parts <= 0) return DGN_ERR_INVALID;          // needs the statistics slabs
  if (a->gamma && !a->training && (!a->running_mean || !a->running_var)) return DGN_ERR_INVALID;
  auto al = [](const void* x) { return (reinterpret_cast<uintptr_t>(x) & 15u) == 0; };
  if (!lin_ok(a->n_cols, f_out, p, q, ld_p, ld_q) || !al(a->y) || !al(a->out) || a->ld_y % 4 || a->ld_o % 4 ||
      (a->residual && (!al(a->residual) || a->ld_res % 4)))
    return DGN_ERR_UNSUPPORTED;
  if (a->n_rows == 0) return DGN_OK;
  const int Fi = a->n_cols;
  const size_t smem = (size_t)(LR * (Fi + TPAD) + LC * (2 * Fi + TPAD) + (5 + 12) * Fi) * sizeof(float);
  if (int rc = lin_attr(norm_pair_fwd_kernel, smem)) return rc;
  const int w_vec = al(w) && ld_w % 4 == 0;
  launch_pdl(norm_pair_fwd_kernel, dim3((a->n_rows + LR - 1) / LR, (f_out + LC - 1) / LC), dim3(LT), smem,
             (cudaStream_t)stream, *a, f_out, w, ld_w, w_vec, p, ld_p, q, ld_q);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_dgn_last_cuda = e; return DGN_ERR_CUDA; }
  return DGN_OK;
}

extern "C" int dgn_pair_linear_backward(int32_t N, int32_t Fi, int32_t Fo, const float* dP, int32_t ld_p, const float* dQ,
                                        int32_t ld_q, const float* W, int32_t ld_w, float* d_h, int32_t ld_dh,
                                        void* stream) {
  if (N < 0 || !dP || !dQ || !W || !d_h) return DGN_ERR_INVALID;
  if (!lin_ok(Fi, Fo, d_h, d_h, ld_dh, ld_dh)) return DGN_ERR_UNSUPPORTED;
  if (N == 0) return DGN_OK;
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  if (!al(dP) || !al(dQ) || ld_p % 4 || ld_q % 4) return DGN_ERR_UNSUPPORTED;
  const size_t smem = (size_t)(2 * LR * (Fo + TPAD) + 2 * Fo * LC) * sizeof(float);
  if (int rc = lin_attr(pair_linear_bwd_kernel, smem)) return rc;
  const int w_vec = al(W) && ld_w % 4 == 0;
  launch_pdl(pair_linear_bwd_kernel, dim3((N + LR - 1) / LR, (Fi + LC - 1) / LC), dim3(LT), smem, (cudaStream_t)stream, N, Fi, Fo, dP, ld_p, dQ, ld_q, W, ld_w, w_vec, d_h, ld_dh);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_dgn_last_cuda = e; return DGN_ERR_CUDA; }
  return DGN_OK;
}

extern "C" int dgn_pair_gather_backward(int32_t N, int32_t Fi, int32_t Fo, const int32_t* out_ptr, const int32_t* out_slot,
                                        const float* edge_ws, int32_t ld_ws, const float* dQ, int32_t ld_q, const float* W,
                                        int32_t ld_w, float* d_h, int32_t ld_dh, float* dP, int32_t ld_p, void* stream) {
  if (N < 0 || !out_ptr || !edge_ws || !dQ || !W || !d_h || !dP) return DGN_ERR_INVALID;
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  if (!lin_ok(Fi, Fo, d_h, dP, ld_dh, ld_p) || !al(edge_ws) || !al(dQ) || ld_ws % 4 || ld_q % 4) return DGN_ERR_UNSUPPORTED;
  if (N == 0) return DGN_OK;
  const size_t smem = (size_t)(2 * LR * (Fo + TPAD) + 2 * Fo * LC) * sizeof(float);
  if (int rc = lin_attr(pair_gather_bwd_kernel, smem)) return rc;
  const int w_vec = al(W) && ld_w % 4 == 0;
  launch_pdl(pair_gather_bwd_kernel, dim3((N + LR - 1) / LR, (Fi + LC - 1) / LC), dim3(LT), smem, (cudaStream_t)stream, N, Fi,
             Fo, out_ptr, out_slot, edge_ws, ld_ws, dQ, ld_q, W, ld_w, w_vec, d_h, ld_dh, dP, ld_p);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_dgn_last_cuda = e; return DGN_ERR_CUDA; }
  return DGN_OK;
}
