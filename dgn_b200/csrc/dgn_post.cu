// Scaler-folded posttrans products of a DGN layer on the tcgen05 tensor cores (sm_100a).
//
// The reference concatenates the aggregates once per scaler, cat = [h | c_0 agg | c_1 agg | c_2 agg] ([N, (1 + S A) F],
// rb/nets/dgn_layer.py:94-96, 116) and multiplies by W_post^T (:119).  The scaler coefficient c_s(v) (rb/nets/scalers.py)
// is a per-NODE factor, so
//       y[v] = h[v] W_h^T + sum_s c_s(v) * (agg[v] W_s^T)            W_post = [W_h | W_0 | ... | W_{S-1}]
// and the [N, S A F] tensor (24 MB per layer and direction at the bench workload) never has to exist: the aggregation
// kernel writes the raw aggregates once ([N, A F]), the kernels here keep one fp32 accumulator PER SCALER in tensor
// memory and fold them with c_s(v) in the epilogue.  Three products per layer:
//
//   post_fwd_kernel   y      = fold_s( [h | agg] , W_post )                          K = F + A F, split over a cluster
//   post_bwd_kernel   d_cat1 = [ d_y W_h | sum_s c_s(v) (d_y W_s) ]                  K = F_out
//   wgrad_kernel      d_W_h  = d_y^T h ,  d_W_s = (c_s * d_y)^T agg                  K = N nodes, split over a cluster
//                     (also d_W_pre = [d_P^T h | d_Q^T h] and d_b_pre = column sums of d_Q, one launch)
//
// All of them: 128-row tiles, 3xTF32 split (fp32 accuracy, see dgn_gemm.cu), operands converted on the way into the
// swizzled shared-memory tiles by 8 loader warps, one MMA-issuing thread, accumulators in TMEM.  Split-K partial tiles are
// NOT written to global memory: the CTAs of a thread-block cluster share one output tile, stage their folded partial in
// shared memory and every rank sums its slice of the tile over its peers' shared memory (DSMEM) in rank order -
// deterministic, no workspace, no second launch.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "../../include/dgn_b200.h"
#include "dgn_launch.cuh"
#include "dgn_umma.cuh"

extern thread_local cudaError_t g_dgn_last_cuda;

namespace dgn {
using namespace umma;

constexpr int PM = 128, PN = 64;                    // output tile
constexpr int kLoad = 256, kPostThreads = kLoad + 32;   // 8 loader / epilogue warps + 1 MMA warp
constexpr int A_BYTES = PM * BK * 4, B_BYTES = PN * BK * 4;
constexpr int kMaxTerms = DGN_MAX_SCALERS;
constexpr int kMaxSplit = 16;
constexpr int PATCH_LD = PN + 1;                    // padded row of the epilogue staging patch [PM][PATCH_LD]
constexpr int PATCH_T = PM + 1;                     // ... and of the transposed patch [PN][PATCH_T] (weight gradients)

__device__ __forceinline__ float scaler_coef(int kind, float ld, float avg) {
  if (kind == DGN_SCALE_AMPLIFICATION) return __fdiv_rn(ld, avg);
  if (kind == DGN_SCALE_ATTENUATION) return ld > 0.f ? __fdiv_rn(avg, ld) : 0.f;   // D = 0: the aggregates are 0 anyway
  return 1.f;
}

struct PostK {
  int N, F, Ka, Fo, S, fold;              // S = accumulators of the aggregate segment; fold: coefficients from log_deg
  int skind[kMaxTerms];
  float avg_log;
  const float* log_deg;
  const float* cat; int ld_cat;
  const float* W; int ld_w;
  float* y; int ld_y;
  const float* dy; int ld_dy;
  float* dcat; int ld_dcat;
  int ksplit, n_lead_kb, n_agg_kb;
  int kb_start[kMaxSplit + 1];
  int n_lead_tiles, n_agg_tiles;
  // optional BatchNorm partial statistics of z = (y + y_bias) * snorm, one slab per row tile: [cnt | mean | M2][Fo]
  float* stat_parts; const float* y_bias; const float* snorm; const int* n_rows_dev;
  unsigned long long* dbg;                 // optional [n_ctas][8] %globaltimer stamps of the forward kernel's phases (tools/)
};

__device__ __forceinline__ void stamp(const PostK& g, int slot) {
  if (g.dbg) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    g.dbg[(size_t)cta * 8 + slot] = t;
  }
}

struct Smem {
  unsigned char* base;
  int stage_bytes;
  uint64_t *full, *empty, *accum_full;
  uint32_t* tmem_slot;
  __device__ __forceinline__ unsigned char* stage(int s) const { return base + s * stage_bytes; }
};

__device__ __forceinline__ Smem carve(unsigned char* raw, int n_btiles) {
  Smem sm;
  sm.base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  sm.stage_bytes = 2 * A_BYTES + n_btiles * 2 * B_BYTES;
  sm.full = reinterpret_cast<uint64_t*>(sm.base + 2 * sm.stage_bytes);
  sm.empty = sm.full + 2;
  sm.accum_full = sm.empty + 2;
  sm.tmem_slot = reinterpret_cast<uint32_t*>(sm.accum_full + 1);
  return sm;
}
static int smem_bytes(int n_btiles) { return 2 * (2 * A_BYTES + n_btiles * 2 * B_BYTES) + 1024 + 256; }

// barrier / TMEM set-up common to the three kernels; returns the TMEM base address
__device__ __forceinline__ uint32_t prologue(const Smem& sm, int tmem_cols) {
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) { mb_init(&sm.full[s], kLoad); mb_init(&sm.empty[s], 1); }
    mb_init(sm.accum_full, 1);
    fence_barrier_init();
  }
  if (warp == kLoad / 32) tmem_alloc_n(sm.tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *sm.tmem_slot;
}

// ------------------------------------------------------------------------------------------------------------
// forward: y[m, n] = sum_k h[m,k] W[n,k] + sum_s c_s(m) sum_k agg[m,k] W[n, F + s Ka + k]
// grid (m tiles, ksplit, n tiles), cluster (1, ksplit, 1): rank r takes the k-blocks [kb_start[r], kb_start[r+1])
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPostThreads, 1) post_fwd_kernel(const __grid_constant__ PostK g) {
  pdl_prologue();
  extern __shared__ unsigned char raw[];
  const Smem sm = carve(raw, g.S);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * PM, n0 = blockIdx.z * PN;
  const int rank = g.ksplit > 1 ? (int)cluster_ctarank() : 0;
  const int kb0 = g.kb_start[rank], kb1 = g.kb_start[rank + 1], nkb = kb1 - kb0;
  const int tmem_cols = (g.S + 1) * PN;
  if (tid == 0) stamp(g, 0);
  const uint32_t tmem_d = prologue(sm, tmem_cols);
  if (tid == 0) stamp(g, 1);

  if (warp < kLoad / 32) {
    // ------------------------------ loaders ------------------------------
    float4 va[PM * 8 / kLoad], vb[kMaxTerms][PN * 8 / kLoad];
    auto fetch = [&](int kb, float4 (&a)[PM * 8 / kLoad], float4 (&b)[kMaxTerms][PN * 8 / kLoad]) {
      const bool lead = kb < g.n_lead_kb;
      const int kcol = (lead ? kb : kb - g.n_lead_kb) * BK;
      if (lead) {
        fetch_k<PM, kLoad>(g.cat, g.ld_cat, m0, g.N, kcol, g.F, a, tid);
        fetch_k<PN, kLoad>(g.W, g.ld_w, n0, g.Fo, kcol, g.F, b[0], tid);
      } else {
        fetch_k<PM, kLoad>(g.cat + g.F, g.ld_cat, m0, g.N, kcol, g.Ka, a, tid);
#pragma unroll
        for (int t = 0; t < kMaxTerms; ++t)
          if (t < g.S) fetch_k<PN, kLoad>(g.W + g.F + (size_t)t * g.Ka, g.ld_w, n0, g.Fo, kcol, g.Ka, b[t], tid);
      }
    };
    // Two register sets in ping-pong (no copies: a copy would wait for the loads it copies): block i+1 is in flight
    // while block i is split and stored.
    float4 wa[PM * 8 / kLoad], wb[kMaxTerms][PN * 8 / kLoad];
    auto step = [&](int i, float4 (&ca)[PM * 8 / kLoad], float4 (&cb)[kMaxTerms][PN * 8 / kLoad],
                    float4 (&xa)[PM * 8 / kLoad], float4 (&xb)[kMaxTerms][PN * 8 / kLoad]) {
      const int s = i & 1;
      if (i + 1 < nkb) fetch(kb0 + i + 1, xa, xb);
      if (i >= 2) mb_wait(&sm.empty[s], ((i >> 1) - 1) & 1);
      unsigned char* st = sm.stage(s);
      const bool lead = (kb0 + i) < g.n_lead_kb;
      store_k<PM, kLoad>(ca, st, st + A_BYTES, tid);
#pragma unroll
      for (int t = 0; t < kMaxTerms; ++t)
        if (t < (lead ? 1 : g.S))
          store_k<PN, kLoad>(cb[t], st + 2 * A_BYTES + t * B_BYTES, st + 2 * A_BYTES + (g.S + t) * B_BYTES, tid);
      fence_async_smem();
      mb_arrive(&sm.full[s]);
    };
    if (nkb > 0) fetch(kb0, va, vb);
    for (int i = 0; i < nkb; i += 2) {
      step(i, va, vb, wa, wb);
      if (i + 1 < nkb) step(i + 1, wa, wb, va, vb);
    }
    if (tid == 0) stamp(g, 2);
  } else if (lane == 0) {
    // ------------------------------ MMA issuer ------------------------------
    // The S scaler terms of an aggregate k-block are ONE wide operand (their B tiles are stacked along N): one
    // N = 64 S instruction per k-step and split term instead of S narrow ones (which re-read A from shared memory
    // S times and are bound by that).  The lead block goes to its own 64 accumulator columns.
    constexpr uint32_t idesc_lead = instr_desc_tf32(PM, PN, false, false);
    const uint32_t idesc_agg = instr_desc_tf32(PM, g.S * PN, false, false);
    uint32_t used = 0;
    for (int i = 0; i < nkb; ++i) {
      const int s = i & 1;
      mb_wait(&sm.full[s], (i >> 1) & 1);
      tc_fence_after();
      const uint32_t a_hi = s32(sm.stage(s)), a_lo = a_hi + A_BYTES;
      const bool lead = (kb0 + i) < g.n_lead_kb;
      const uint32_t b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + g.S * B_BYTES;
      const uint32_t dcol = lead ? g.S * PN : 0, bit = lead ? 2u : 1u;
#pragma unroll
      for (int kk = 0; kk < BK / 8; ++kk)
        umma_tf32x3(tmem_d + dcol, tile_desc<PM, true>(a_hi, kk), tile_desc<PM, true>(a_lo, kk),
                    tile_desc<PN, true>(b_hi, kk), tile_desc<PN, true>(b_lo, kk), lead ? idesc_lead : idesc_agg,
                    (((used & bit) ? 1u : 0u) | (kk > 0 ? 1u : 0u)));
      used |= bit;
      umma_commit(&sm.empty[s]);
    }
    umma_commit(sm.accum_full);
  }

  // ------------------------------ epilogue: fold, stage, cluster reduce ------------------------------
  float* patch = reinterpret_cast<float*>(sm.base);                  // [PM][PATCH_LD]; the stages are idle by now
  if (warp < kLoad / 32) {
    mb_wait(sm.accum_full, 0);
    tc_fence_after();
    if (tid == 0) stamp(g, 3);
    const int q = warp & 3, hf = warp >> 2, row = q * 32 + lane, gm = m0 + row;
    const bool has_lead = nkb > 0 && kb0 < g.n_lead_kb, has_agg = nkb > 0 && kb1 > g.n_lead_kb;
    const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + hf * 32;
    float o[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) o[j] = 0.f;
    if (has_agg) {
      const float ld = (g.fold && gm < g.N) ? __ldg(g.log_deg + gm) : 0.f;
      for (int t = 0; t < g.S; ++t) {
        float v[32];
        tmem_ld32(taddr + t * PN, v);
        const float c = g.fold ? scaler_coef(g.skind[t], ld, g.avg_log) : 1.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = fmaf(c, v[j], o[j]);
      }
    }
    if (has_lead) {
      float v[32];
      tmem_ld32(taddr + g.S * PN, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) o[j] += v[j];
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) patch[row * PATCH_LD + hf * 32 + j] = o[j];
  }
  tc_fence_before();
  if (g.ksplit > 1) cluster_sync_all(); else __syncthreads();
  if (tid == 0) stamp(g, 4);
  if (warp < kLoad / 32) {
    // Every thread's elements share one column (c = tid % PN) and step 4 rows: it keeps the running batch statistics
    // (Welford) of z = (y + bias) * snorm over its rows for the BatchNorm that follows.
    const int n_real = g.n_rows_dev ? *g.n_rows_dev : g.N;
    const int my_c = tid % PN;
    const float yb = (g.stat_parts && g.y_bias && n0 + my_c < g.Fo) ? __ldg(g.y_bias + n0 + my_c) : 0.f;
    float w_cnt = 0.f, w_mean = 0.f, w_m2 = 0.f;
    auto emit_y = [&](int r, int c, float acc) {
      const int gm = m0 + r, gn = n0 + c;
      if (gm < g.N && gn < g.Fo) {
        g.y[(size_t)gm * g.ld_y + gn] = acc;
        if (g.stat_parts && gm < n_real) {
          float z = acc + yb;
          if (g.snorm) z *= __ldg(g.snorm + gm);
          w_cnt += 1.f;
          const float d = z - w_mean;
          w_mean += __fdividef(d, w_cnt);
          w_m2 = fmaf(d, z - w_mean, w_m2);
        }
      }
    };
    // rank r owns the rows [r * PM / ksplit, ...) of the tile: 32 / ksplit elements x ksplit remote loads per thread
    auto reduce_rows = [&](auto nr) {
      constexpr int NR = decltype(nr)::value;
      const int r_base = rank * (PM / NR);
      auto pos = [&](int e, int& r, int& c) { const int idx = tid + e * kLoad; r = r_base + idx / PN; c = idx % PN; };
      dsmem_reduce32<NR>(
          [&](int e) { int r, c; pos(e, r, c); return s32(patch + r * PATCH_LD + c); },
          [&](int e, float acc) { int r, c; pos(e, r, c); emit_y(r, c, acc); });
    };
    if (g.ksplit == 1) {
      for (int idx = tid; idx < PM * PN; idx += kLoad) emit_y(idx / PN, idx % PN, patch[(idx / PN) * PATCH_LD + idx % PN]);
    } else if (g.ksplit == 2) reduce_rows(std::integral_constant<int, 2>());
    else if (g.ksplit == 4) reduce_rows(std::integral_constant<int, 4>());
    else reduce_rows(std::integral_constant<int, 8>());
    if (g.stat_parts) {
      // merge the kLoad / PN = 4 threads of every column in a fixed order, one slab per (row tile, rank)
      float* sred = reinterpret_cast<float*>(sm.base + 40 * 1024);       // behind the patch; [3][4][PN]
      const int sub = tid / PN;
      sred[(0 * 4 + sub) * PN + my_c] = w_cnt;
      sred[(1 * 4 + sub) * PN + my_c] = w_mean;
      sred[(2 * 4 + sub) * PN + my_c] = w_m2;
      asm volatile("bar.sync 1, %0;" ::"n"(kLoad) : "memory");           // the 8 loader / epilogue warps only
      if (tid < PN && n0 + tid < g.Fo) {
        float cnt = sred[tid], mean = sred[4 * PN + tid], m2 = sred[8 * PN + tid];
        for (int j = 1; j < 4; ++j) {
          const float nb = sred[j * PN + tid], mb = sred[(4 + j) * PN + tid], m2b = sred[(8 + j) * PN + tid];
          if (nb > 0.f) {
            const float nt_ = cnt + nb, d = mb - mean;
            mean += d * (nb / nt_);
            m2 += m2b + d * d * (cnt * nb / nt_);
            cnt = nt_;
          }
        }
        if (g.ksplit > 1) {                                   // combined over the cluster below: one slab per row tile
          float* sfin = sred + 12 * PN;
          sfin[tid] = cnt; sfin[PN + tid] = mean; sfin[2 * PN + tid] = m2;
        } else {
          float* part = g.stat_parts + (size_t)blockIdx.x * 3 * g.Fo;
          part[n0 + tid] = cnt; part[g.Fo + n0 + tid] = mean; part[2 * g.Fo + n0 + tid] = m2;
        }
      }
    }
  }
  if (tid == 0) stamp(g, 5);
  if (g.ksplit > 1) cluster_sync_all();                               // peers may still be reading this CTA's patch
  if (g.stat_parts && g.ksplit > 1) {
    // rank 0 merges the per-rank statistics of the row tile over DSMEM (rank order) into ONE slab
    if (rank == 0 && tid < PN && n0 + tid < g.Fo) {
      const uint32_t local = s32(reinterpret_cast<float*>(sm.base + 40 * 1024) + 12 * PN + tid);
      float pn[8], pm[8], p2[8];
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const bool ok = p < g.ksplit;
        const uint32_t a = ok ? dsmem_addr(local, (uint32_t)p) : 0u;
        pn[p] = ok ? dsmem_ld(a) : 0.f;
        pm[p] = ok ? dsmem_ld(a + PN * 4) : 0.f;
        p2[p] = ok ? dsmem_ld(a + 2 * PN * 4) : 0.f;
      }
      float cnt = 0.f, mean = 0.f, m2 = 0.f;
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        if (pn[p] > 0.f) {
          const float nt_ = cnt + pn[p], d = pm[p] - mean;
          mean += d * (pn[p] / nt_);
          m2 += p2[p] + d * d * (cnt * pn[p] / nt_);
          cnt = nt_;
        }
      }
      float* part = g.stat_parts + (size_t)blockIdx.x * 3 * g.Fo;
      part[n0 + tid] = cnt; part[g.Fo + n0 + tid] = mean; part[2 * g.Fo + n0 + tid] = m2;
    }
    cluster_sync_all();                                               // rank 0 has read its peers' shared memory
  }
  __syncwarp();
  if (warp == kLoad / 32) tmem_dealloc_n(tmem_d, tmem_cols);
  if (tid == 0) stamp(g, 6);
}

// ------------------------------------------------------------------------------------------------------------
// backward: d_cat1[m, :F] = d_y[m] W_h ; d_cat1[m, F + c] = sum_s c_s(m) (d_y[m] W_s)[c]
// grid (m tiles, lead tiles + aggregate tiles); K = F_out
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPostThreads, 1) post_bwd_kernel(const __grid_constant__ PostK g) {
  pdl_prologue();
  extern __shared__ unsigned char raw[];
  const Smem sm = carve(raw, g.S);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * PM;
  const bool lead = (int)blockIdx.y < g.n_lead_tiles;
  const int c0 = (lead ? blockIdx.y : blockIdx.y - g.n_lead_tiles) * PN;   // first column inside the segment
  const int seg = lead ? g.F : g.Ka, nt = lead ? 1 : g.S;
  const int nkb = (g.Fo + BK - 1) / BK;
  const int tmem_cols = nt * PN;
  const uint32_t tmem_d = prologue(sm, tmem_cols);

  if (warp < kLoad / 32) {
    float4 va[PM * 8 / kLoad], vb[kMaxTerms][PN * 8 / kLoad];
    auto fetch = [&](int kb, float4 (&a)[PM * 8 / kLoad], float4 (&b)[kMaxTerms][PN * 8 / kLoad]) {
      fetch_k<PM, kLoad>(g.dy, g.ld_dy, m0, g.N, kb * BK, g.Fo, a, tid);
#pragma unroll
      for (int t = 0; t < kMaxTerms; ++t)
        if (t < nt)
          fetch_mn<PN, kLoad>(g.W + (lead ? 0 : g.F + (size_t)t * g.Ka), g.ld_w, c0, seg, kb * BK, g.Fo, -1, b[t], tid);
    };
    float4 wa[PM * 8 / kLoad], wb[kMaxTerms][PN * 8 / kLoad];
    auto step = [&](int i, float4 (&ca)[PM * 8 / kLoad], float4 (&cb)[kMaxTerms][PN * 8 / kLoad],
                    float4 (&xa)[PM * 8 / kLoad], float4 (&xb)[kMaxTerms][PN * 8 / kLoad]) {
      const int s = i & 1;
      if (i + 1 < nkb) fetch(i + 1, xa, xb);
      if (i >= 2) mb_wait(&sm.empty[s], ((i >> 1) - 1) & 1);
      unsigned char* st = sm.stage(s);
      store_k<PM, kLoad>(ca, st, st + A_BYTES, tid);
#pragma unroll
      for (int t = 0; t < kMaxTerms; ++t)
        if (t < nt)
          store_mn_wide<kLoad>(cb[t], st + 2 * A_BYTES, st + 2 * A_BYTES + g.S * B_BYTES, tid, nt * PN, t);
      fence_async_smem();
      mb_arrive(&sm.full[s]);
    };
    fetch(0, va, vb);
    for (int i = 0; i < nkb; i += 2) {
      step(i, va, vb, wa, wb);
      if (i + 1 < nkb) step(i + 1, wa, wb, va, vb);
    }
  } else if (lane == 0) {
    const uint32_t idesc = instr_desc_tf32(PM, nt * PN, false, true);      // all terms in one wide-N instruction
    for (int i = 0; i < nkb; ++i) {
      const int s = i & 1;
      mb_wait(&sm.full[s], (i >> 1) & 1);
      tc_fence_after();
      const uint32_t a_hi = s32(sm.stage(s)), a_lo = a_hi + A_BYTES;
      const uint32_t b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + g.S * B_BYTES;
#pragma unroll
      for (int kk = 0; kk < BK / 8; ++kk)
        umma_tf32x3(tmem_d, tile_desc<PM, true>(a_hi, kk), tile_desc<PM, true>(a_lo, kk), tile_desc_mn(b_hi, kk, nt * PN),
                    tile_desc_mn(b_lo, kk, nt * PN), idesc, (i > 0 || kk > 0) ? 1u : 0u);
      umma_commit(&sm.empty[s]);
    }
    umma_commit(sm.accum_full);
  }

  float* patch = reinterpret_cast<float*>(sm.base);
  if (warp < kLoad / 32) {
    mb_wait(sm.accum_full, 0);
    tc_fence_after();
    const int q = warp & 3, hf = warp >> 2, row = q * 32 + lane, gm = m0 + row;
    const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + hf * 32;
    const bool fold = !lead && g.fold;
    const float ld = (fold && gm < g.N) ? __ldg(g.log_deg + gm) : 0.f;
    float o[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) o[j] = 0.f;
    for (int t = 0; t < nt; ++t) {
      float v[32];
      tmem_ld32(taddr + t * PN, v);
      const float c = fold ? scaler_coef(g.skind[t], ld, g.avg_log) : 1.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) o[j] = fmaf(c, v[j], o[j]);
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) patch[row * PATCH_LD + hf * 32 + j] = o[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp < kLoad / 32) {
    float* out = g.dcat + (lead ? 0 : g.F) + c0;
    const bool v4 = (c0 + PN <= seg);
    if (v4) {                                                       // full tile: 128-bit stores, 16 threads per row
      for (int idx = tid; idx < PM * (PN / 4); idx += kLoad) {
        const int r = idx / (PN / 4), c = (idx % (PN / 4)) * 4, gm = m0 + r;
        if (gm < g.N) {
          const float* p = patch + r * PATCH_LD + c;
          *reinterpret_cast<float4*>(out + (size_t)gm * g.ld_dcat + c) = make_float4(p[0], p[1], p[2], p[3]);
        }
      }
    } else {
      for (int idx = tid; idx < PM * PN; idx += kLoad) {
        const int r = idx / PN, c = idx % PN, gm = m0 + r;
        if (gm < g.N && c0 + c < seg) out[(size_t)gm * g.ld_dcat + c] = patch[r * PATCH_LD + c];
      }
    }
  }
  __syncwarp();
  if (warp == kLoad / 32) tmem_dealloc_n(tmem_d, tmem_cols);
}

// ------------------------------------------------------------------------------------------------------------
// weight gradients: for up to two problems p (blockIdx.x < n_tiles[0] -> problem 0)
//     C_t[n, c_off[t] + m] (+)= sum_v A[v, m] * s_t(v) * B_t[v, n]           m < Ma (+ ones column), n < Nb, t < nterm
// A, B_t are [nodes, cols] row-major (MN-major operands, K = nodes).  grid (m tiles, ksplit, n tiles), cluster over y.
// ------------------------------------------------------------------------------------------------------------
struct WgProb {
  int Ma, nterm, ones_col;                 // ones_col = Ma (append a column of ones to A) or -1
  const float* A; int ld_a;
  const float* B[kMaxTerms]; int ld_b;
  int skind[kMaxTerms];                    // -1: no scale, else DgnScalerKind applied per node
  int c_off[kMaxTerms];
  float* bias; int bias_term;              // row m == ones_col of term bias_term -> bias[n] (column sums of B_t)
};
struct WgK {
  int N, Nb;
  WgProb p[2];
  int n_tiles0;
  const float* log_deg; float avg_log;
  float* C; int ld_c; int accumulate;
  int ksplit; int kb_start[kMaxSplit + 1];
};

struct CoefScale {
  static constexpr bool on = true;
  const float* log_deg; float avg; int kind;
  __device__ __forceinline__ float operator()(int v) const { return scaler_coef(kind, __ldg(log_deg + v), avg); }
};

__global__ void __launch_bounds__(kPostThreads, 1) wgrad_kernel(const __grid_constant__ WgK g) {
  pdl_prologue();
  extern __shared__ unsigned char raw[];
  const int pi = (int)blockIdx.x < g.n_tiles0 ? 0 : 1;
  const WgProb& P = g.p[pi];
  const Smem sm = carve(raw, kMaxTerms);                                // sized for the larger problem
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = (pi == 0 ? blockIdx.x : blockIdx.x - g.n_tiles0) * PM, n0 = blockIdx.z * PN;
  const int rank = g.ksplit > 1 ? (int)cluster_ctarank() : 0;
  const int kb0 = g.kb_start[rank], kb1 = g.kb_start[rank + 1], nkb = kb1 - kb0;
  const int nt = P.nterm;
  const int tmem_cols = nt * PN;
  const uint32_t tmem_d = prologue(sm, tmem_cols);

  if (warp < kLoad / 32) {
    float4 va[PM * 8 / kLoad], vb[kMaxTerms][PN * 8 / kLoad];
    auto fetch = [&](int kb, float4 (&a)[PM * 8 / kLoad], float4 (&b)[kMaxTerms][PN * 8 / kLoad]) {
      fetch_mn<PM, kLoad>(P.A, P.ld_a, m0, P.Ma, kb * BK, g.N, P.ones_col, a, tid);
#pragma unroll
      for (int t = 0; t < kMaxTerms; ++t) {
        if (t < nt) {
          if (P.skind[t] >= 0)
            fetch_mn<PN, kLoad, CoefScale>(P.B[t], P.ld_b, n0, g.Nb, kb * BK, g.N, -1, b[t], tid,
                                           CoefScale{g.log_deg, g.avg_log, P.skind[t]});
          else
            fetch_mn<PN, kLoad>(P.B[t], P.ld_b, n0, g.Nb, kb * BK, g.N, -1, b[t], tid);
        }
      }
    };
    float4 wa[PM * 8 / kLoad], wb[kMaxTerms][PN * 8 / kLoad];
    auto step = [&](int i, float4 (&ca)[PM * 8 / kLoad], float4 (&cb)[kMaxTerms][PN * 8 / kLoad],
                    float4 (&xa)[PM * 8 / kLoad], float4 (&xb)[kMaxTerms][PN * 8 / kLoad]) {
      const int s = i & 1;
      if (i + 1 < nkb) fetch(kb0 + i + 1, xa, xb);
      if (i >= 2) mb_wait(&sm.empty[s], ((i >> 1) - 1) & 1);
      unsigned char* st = sm.stage(s);
      store_mn<PM, kLoad>(ca, st, st + A_BYTES, tid);
#pragma unroll
      for (int t = 0; t < kMaxTerms; ++t)
        if (t < nt)
          store_mn_wide<kLoad>(cb[t], st + 2 * A_BYTES, st + 2 * A_BYTES + kMaxTerms * B_BYTES, tid, nt * PN, t);
      fence_async_smem();
      mb_arrive(&sm.full[s]);
    };
    if (nkb > 0) fetch(kb0, va, vb);
    for (int i = 0; i < nkb; i += 2) {
      step(i, va, vb, wa, wb);
      if (i + 1 < nkb) step(i + 1, wa, wb, va, vb);
    }
  } else if (lane == 0) {
    const uint32_t idesc = instr_desc_tf32(PM, nt * PN, true, true);       // all terms in one wide-N instruction
    for (int i = 0; i < nkb; ++i) {
      const int s = i & 1;
      mb_wait(&sm.full[s], (i >> 1) & 1);
      tc_fence_after();
      const uint32_t a_hi = s32(sm.stage(s)), a_lo = a_hi + A_BYTES;
      const uint32_t b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + kMaxTerms * B_BYTES;
#pragma unroll
      for (int kk = 0; kk < BK / 8; ++kk)
        umma_tf32x3(tmem_d, tile_desc<PM, false>(a_hi, kk), tile_desc<PM, false>(a_lo, kk), tile_desc_mn(b_hi, kk, nt * PN),
                    tile_desc_mn(b_lo, kk, nt * PN), idesc, (i > 0 || kk > 0) ? 1u : 0u);
      umma_commit(&sm.empty[s]);
    }
    umma_commit(sm.accum_full);
  }

  // epilogue: one TRANSPOSED patch per term ([n][m]: C is written n-major), then every rank reduces its slice of the
  // N columns over the cluster; consecutive threads take consecutive m = contiguous remote words (DSMEM wants
  // coalesced accesses like global memory) and contiguous words of C
  float* patch = reinterpret_cast<float*>(sm.base);                    // [nt][PN][PATCH_T]
  if (warp < kLoad / 32) {
    mb_wait(sm.accum_full, 0);
    tc_fence_after();
    const int q = warp & 3, hf = warp >> 2, row = q * 32 + lane;
    const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + hf * 32;
    for (int t = 0; t < nt; ++t) {
      float v[32];
      if (nkb > 0) {
        tmem_ld32(taddr + t * PN, v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      float* pt = patch + (size_t)t * PN * PATCH_T;
#pragma unroll
      for (int j = 0; j < 32; ++j) pt[(hf * 32 + j) * PATCH_T + row] = v[j];
    }
  }
  tc_fence_before();
  if (g.ksplit > 1) cluster_sync_all(); else __syncthreads();
  if (warp < kLoad / 32) {
    const int m_hi = P.ones_col >= 0 ? P.Ma + 1 : P.Ma;
    auto emit = [&](int t, int m, int n, float acc) {
      const int gm = m0 + m, gn = n0 + n;
      if (gm >= m_hi || gn >= g.Nb) return;
      if (gm == P.ones_col) {
        if (P.bias && t == P.bias_term) P.bias[gn] = g.accumulate ? P.bias[gn] + acc : acc;
      } else {
        float* c = g.C + (size_t)gn * g.ld_c + P.c_off[t] + gm;
        *c = g.accumulate ? *c + acc : acc;
      }
    };
    // rank r owns the output columns [r * PN / ksplit, ...): consecutive threads -> consecutive m (contiguous in C)
    auto reduce_cols = [&](auto nr) {
      constexpr int NR = decltype(nr)::value;
      constexpr int PER = 32 / NR;
      const int n_base = rank * (PN / NR);
      for (int t = 0; t < nt; ++t) {
        const float* pt = patch + (size_t)t * PN * PATCH_T;
        for (int base = 0; base < (PN / NR) * PM; base += PER * kLoad) {
          auto pos = [&](int e, int& m, int& n) { const int idx = base + tid + e * kLoad; n = n_base + idx / PM; m = idx % PM; };
          dsmem_reduce32<NR>([&](int e) { int m, n; pos(e, m, n); return s32(pt + n * PATCH_T + m); },
                             [&](int e, float acc) { int m, n; pos(e, m, n); emit(t, m, n, acc); });
        }
      }
    };
    if (g.ksplit == 1) {
      for (int t = 0; t < nt; ++t) {
        const float* pt = patch + (size_t)t * PN * PATCH_T;
        for (int idx = tid; idx < PN * PM; idx += kLoad) emit(t, idx % PM, idx / PM, pt[(idx / PM) * PATCH_T + idx % PM]);
      }
    } else if (g.ksplit == 2) reduce_cols(std::integral_constant<int, 2>());
    else if (g.ksplit == 4) reduce_cols(std::integral_constant<int, 4>());
    else if (g.ksplit == 8) reduce_cols(std::integral_constant<int, 8>());
    else reduce_cols(std::integral_constant<int, 16>());
  }
  if (g.ksplit > 1) cluster_sync_all();
  __syncwarp();
  if (warp == kLoad / 32) tmem_dealloc_n(tmem_d, tmem_cols);
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
template <typename K, typename Arg>
static cudaError_t launch_cluster(K kern, dim3 grid, int cluster_y, int smem, cudaStream_t st, const Arg& a) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kPostThreads);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = (unsigned)cluster_y;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = cluster_y > 1 ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, a);
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// splits `costs` (one per k-block) into `parts` contiguous ranges of roughly equal cost
static void split_by_cost(const int* cost, int n, int parts, int* start) {
  long long total = 0;
  for (int i = 0; i < n; ++i) total += cost[i];
  int kb = 0;
  long long acc = 0;
  start[0] = 0;
  for (int r = 1; r < parts; ++r) {
    const long long target = total * r / parts;
    while (kb < n - (parts - r) && acc + cost[kb] / 2 < target) acc += cost[kb++];
    if (kb < r) { acc += cost[kb]; ++kb; }                              // every rank gets at least one block
    start[r] = kb;
  }
  start[parts] = n;
}

static int fill_post(const DgnPostArgs* a, PostK& k) {
  if (!a || !a->cat || !a->w || a->n_rows < 0 || a->n_lead < 0 || a->n_agg <= 0 || a->n_out <= 0 || a->n_scalers <= 0 ||
      a->n_scalers > DGN_MAX_SCALERS)
    return DGN_ERR_INVALID;
  memset(&k, 0, sizeof(k));
  k.N = a->n_rows; k.F = a->n_lead; k.Ka = a->n_agg; k.Fo = a->n_out;
  k.S = a->n_scalers > 1 ? a->n_scalers : 1;                            // rb/nets/dgn_layer.py:95
  k.fold = a->n_scalers > 1 ? 1 : 0;
  for (int s = 0; s < k.S; ++s) {
    if (a->scaler_kind[s] > DGN_SCALE_ATTENUATION) return DGN_ERR_INVALID;
    k.skind[s] = a->scaler_kind[s];
  }
  if (k.fold && !a->log_deg) return DGN_ERR_INVALID;
  k.avg_log = a->avg_log; k.log_deg = a->log_deg;
  k.cat = a->cat; k.ld_cat = a->ld_cat; k.W = a->w; k.ld_w = a->ld_w;
  if ((k.F | k.Ka | k.Fo | k.ld_cat | k.ld_w) % 4 || !al16(k.cat) || !al16(k.W)) return DGN_ERR_UNSUPPORTED;
  k.n_lead_kb = (k.F + BK - 1) / BK; k.n_agg_kb = (k.Ka + BK - 1) / BK;
  k.n_lead_tiles = (k.F + PN - 1) / PN; k.n_agg_tiles = (k.Ka + PN - 1) / PN;
  return DGN_OK;
}

template <typename K>
static cudaError_t set_smem(K kern, int bytes, bool nonportable) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && nonportable) e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  return e;
}

// cluster size for a split-K launch with `tiles` output tiles and `nkb` k-blocks: fill ~148 SMs, power of two <= limit
static int pick_ksplit(int tiles, int nkb, int limit) {
  int ks = 1;
  while (ks * 2 <= limit && tiles * ks * 2 <= 160 && ks * 2 <= nkb) ks *= 2;
  return ks;
}

}  // namespace dgn

using namespace dgn;

static int done(cudaError_t e) {
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { g_dgn_last_cuda = e; return DGN_ERR_CUDA; }
  return DGN_OK;
}

static int launch_post_bwd(PostK& k, cudaStream_t st) {
  k.ksplit = 1;
  const int mt = (k.N + PM - 1) / PM;
  const int smem = smem_bytes(k.S);
  static int smem_set = 0;
  cudaError_t e = cudaSuccess;
  if (smem > 220 * 1024) return DGN_ERR_UNSUPPORTED;
  if (smem_set < smem) { e = set_smem(post_bwd_kernel, smem, false); smem_set = smem; }
  if (e == cudaSuccess) e = launch_cluster(post_bwd_kernel, dim3(mt, k.n_lead_tiles + k.n_agg_tiles, 1), 1, smem, st, k);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { g_dgn_last_cuda = e; return DGN_ERR_CUDA; }
  return DGN_OK;
}


extern "C" int dgn_post_forward(const DgnPostArgs* a, float* y, int32_t ld_y, const DgnPostStats* st, int32_t* stat_parts,
                                void* stream) {
  PostK k;
  if (stat_parts) *stat_parts = 0;
  if (int rc = fill_post(a, k)) return rc;
  if (!y) return DGN_ERR_INVALID;
  if (k.N == 0) return DGN_OK;
  k.y = y; k.ld_y = ld_y;
  {   // DGN_POST_DBG=<device pointer, hex>: phase time stamps of every CTA (tools/post_phases.py)
    static const unsigned long long dbg = [] { const char* e = getenv("DGN_POST_DBG"); return e ? strtoull(e, nullptr, 16) : 0ull; }();
    k.dbg = reinterpret_cast<unsigned long long*>(dbg);
  }
  const int mt = (k.N + PM - 1) / PM, ntl = (k.Fo + PN - 1) / PN, nkb = k.n_lead_kb + k.n_agg_kb;
  k.ksplit = pick_ksplit(mt * ntl, nkb, 8);
  // statistics slabs live behind the [mean | rstd] header of the norm workspace: 2 C + parts * 3 C <= DGN_NORM_WS_FLOATS(C)
  if (st && st->stats && stat_parts && 2 + 3 * (long long)mt <= DGN_NORM_WS_FLOATS(1)) {
    k.stat_parts = st->stats + 2 * k.Fo;
    k.y_bias = st->y_bias; k.snorm = st->snorm; k.n_rows_dev = st->n_rows_dev;
    *stat_parts = mt;
  }
  int cost[4096];
  if (nkb > 4096) return DGN_ERR_UNSUPPORTED;
  for (int i = 0; i < nkb; ++i) cost[i] = i < k.n_lead_kb ? 2 : 2 * k.S;
  split_by_cost(cost, nkb, k.ksplit, k.kb_start);
  const int smem = smem_bytes(k.S);
  static int smem_set = 0;
  cudaError_t e = cudaSuccess;
  if (smem_set < smem) { e = set_smem(post_fwd_kernel, smem, false); smem_set = smem; }
  if (e == cudaSuccess) e = launch_cluster(post_fwd_kernel, dim3(mt, k.ksplit, ntl), k.ksplit, smem, (cudaStream_t)stream, k);
  return done(e);
}

extern "C" int dgn_post_backward(const DgnPostArgs* a, const float* d_y, int32_t ld_dy, float* d_cat, int32_t ld_dcat,
                                 void* stream) {
  PostK k;
  if (int rc = fill_post(a, k)) return rc;
  if (!d_y || !d_cat) return DGN_ERR_INVALID;
  if (ld_dy % 4 || ld_dcat % 4 || !al16(d_y) || !al16(d_cat)) return DGN_ERR_UNSUPPORTED;
  if (k.N == 0) return DGN_OK;
  k.dy = d_y; k.ld_dy = ld_dy; k.dcat = d_cat; k.ld_dcat = ld_dcat;
  return launch_post_bwd(k, (cudaStream_t)stream);
}

static int launch_wgrad(WgK& k, int tiles, cudaStream_t st) {
  const int nkb = (k.N + BK - 1) / BK, ntl = (k.Nb + PN - 1) / PN;
  static const int wg_limit = [] { const char* e = getenv("DGN_WG_SPLIT"); const int v = e ? atoi(e) : kMaxSplit; return v < 1 ? 1 : (v > kMaxSplit ? kMaxSplit : v); }();
  k.ksplit = pick_ksplit(tiles * ntl, nkb, wg_limit);
  const int per = (nkb + k.ksplit - 1) / k.ksplit;
  for (int r = 0; r <= k.ksplit; ++r) k.kb_start[r] = r * per < nkb ? r * per : nkb;
  const int smem = smem_bytes(kMaxTerms);
  static bool attr = false;
  cudaError_t e = cudaSuccess;
  if (!attr) { e = set_smem(wgrad_kernel, smem, true); attr = true; }
  if (e == cudaSuccess) e = launch_cluster(wgrad_kernel, dim3(tiles, k.ksplit, ntl), k.ksplit, smem, st, k);
  return done(e);
}

extern "C" int dgn_post_wgrad(const DgnPostArgs* a, const float* d_y, int32_t ld_dy, float* d_w, int32_t ld_dw,
                              int32_t accumulate, void* stream) {
  PostK pk;
  if (int rc = fill_post(a, pk)) return rc;
  if (!d_y || !d_w) return DGN_ERR_INVALID;
  if (ld_dy % 4 || !al16(d_y)) return DGN_ERR_UNSUPPORTED;
  if (pk.N == 0) return DGN_OK;
  WgK k;
  memset(&k, 0, sizeof(k));
  k.N = pk.N; k.Nb = pk.Fo; k.log_deg = pk.log_deg; k.avg_log = pk.avg_log;
  k.C = d_w; k.ld_c = ld_dw; k.accumulate = accumulate;
  int tiles = 0;
  WgProb* agg = &k.p[0];
  if (pk.F > 0) {                                                       // problem 0: d_W_h = d_y^T h
    WgProb& L = k.p[0];
    L.Ma = pk.F; L.nterm = 1; L.ones_col = -1; L.A = pk.cat; L.ld_a = pk.ld_cat; L.B[0] = d_y; L.ld_b = ld_dy;
    L.skind[0] = -1; L.c_off[0] = 0;
    k.n_tiles0 = (pk.F + PM - 1) / PM;
    tiles += k.n_tiles0;
    agg = &k.p[1];
  } else {
    k.n_tiles0 = 0;
    agg = &k.p[1];
  }
  agg->Ma = pk.Ka; agg->nterm = pk.S; agg->ones_col = -1; agg->A = pk.cat + pk.F; agg->ld_a = pk.ld_cat; agg->ld_b = ld_dy;
  for (int t = 0; t < pk.S; ++t) {
    agg->B[t] = d_y;
    agg->skind[t] = pk.fold ? pk.skind[t] : -1;
    agg->c_off[t] = pk.F + t * pk.Ka;
  }
  tiles += (pk.Ka + PM - 1) / PM;
  return launch_wgrad(k, tiles, (cudaStream_t)stream);
}

extern "C" int dgn_pre_wgrad(int32_t n_rows, int32_t f_in, int32_t f_out, const float* h, int32_t ld_h, const float* d_p,
                             int32_t ld_dp, const float* d_q, int32_t ld_dq, float* d_w, int32_t ld_dw, float* d_b,
                             int32_t accumulate, void* stream) {
  if (n_rows < 0 || f_in <= 0 || f_out <= 0 || !h || !d_p || !d_q || !d_w) return DGN_ERR_INVALID;
  if ((f_in | f_out | ld_h | ld_dp | ld_dq) % 4 || ld_dp != ld_dq || !al16(h) || !al16(d_p) || !al16(d_q))
    return DGN_ERR_UNSUPPORTED;
  if (n_rows == 0) return DGN_OK;
  WgK k;
  memset(&k, 0, sizeof(k));
  k.N = n_rows; k.Nb = f_out; k.C = d_w; k.ld_c = ld_dw; k.accumulate = accumulate;
  WgProb& P = k.p[0];
  P.Ma = f_in; P.nterm = 2; P.ones_col = d_b ? f_in : -1; P.A = h; P.ld_a = ld_h;
  P.B[0] = d_p; P.B[1] = d_q; P.ld_b = ld_dp;
  P.skind[0] = P.skind[1] = -1;
  P.c_off[0] = 0; P.c_off[1] = f_in;
  P.bias = d_b; P.bias_term = 1;
  k.n_tiles0 = (f_in + (d_b ? 1 : 0) + PM - 1) / PM;
  k.p[1] = P;
  return launch_wgrad(k, k.n_tiles0, (cudaStream_t)stream);
}
