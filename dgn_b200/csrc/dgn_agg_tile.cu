// Fused DGN aggregation, tile kernels (sm_100a) - the fast path of dgn_agg_forward / dgn_agg_backward.
//
// A CTA owns a tile of TN consecutive destination nodes; thread = (local node, VEC-column chunk).
//
// (1) Everything that does not depend on the feature column is computed ONCE per node / per edge and
//     staged in shared memory instead of once per (node, chunk) thread: the tile's slice of the CSR, the k
//     eigenvector components of the destination nodes, the per-edge eigen-weights w_s(delta_uv), the
//     per-node normalisers (sum |delta|, sum delta ...) with the factors derived from them (1/Z, W) and the
//     degree-scaler coefficients.  The column threads only do vector work.
//
// (2) The gather is latency bound (dependent in_ptr -> in_src -> row chain), so what buys bandwidth is
//     memory-level parallelism.  All bulky operands reach shared memory through the TMA engine
//     (cp.async.bulk + mbarrier complete_tx), i.e. without holding registers or issue slots:
//       * the gathered message rows of a batch of edge slots: one row-sized bulk copy per edge, all rows of
//         the batch in flight at once while the threads compute the eigen-weights;
//       * when the tile's sources fall into a narrow node window that is re-used enough (block-diagonal
//         batches with in-degree >> 1: CIFAR kNN, SBM PATTERN), the whole window once per tile instead, so
//         the per-edge gathers become shared-memory reads;
//       * backward: the tile's S*A gradient slabs (the dominant traffic), one bulk copy per node row.
//
// Edges are processed in rounds: in round r every node of the tile contributes its in-edge slots
// [r*EBN, (r+1)*EBN), so all node threads stay busy whatever the degree and shared memory stays bounded.
// The softmax aggregators (W_EXP) and F/VEC > 256 fall back to the generic kernels in dgn_agg_fwd/bwd.cu.
#include <limits.h>

#include "dgn_plan.cuh"

namespace dgn {

constexpr int kMaxTileThreads = 256;

struct TileCfg {
  int threads;                       // block size (128 or 256)
  int TN;                            // destination nodes per tile
  int EBN;                           // edge slots per node and round
  int EB;                            // edge slots per round = TN * EBN
  int win_rows;                      // capacity of the staged source window in rows (0 = never stage)
  int use_msg;                       // gathered message rows are bulk-copied into shared memory per batch
  int use_gt;                        // backward: the tile's gradient slabs are bulk-copied into shared memory
  int gt_row;                        // floats per node in the gradient tile (S*A*F)
  int off_ev, off_zw, off_zabs, off_f0, off_f1, off_coef, off_src, off_w, off_bar, off_red, off_win, off_msg, off_gt, off_g;
  int total;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}

// Shared-memory view of one tile + the state every thread needs.
struct Tile {
  int* s_ptr; float* s_ev; float* s_zw; float* s_zabs; float* s_f0; float* s_f1; float* s_coef;
  int* s_src; float* s_w; uint64_t* s_bar; int* s_red; float* s_win; float* s_msg; float* s_gt;
  int v0, nv, E0, E1, ln, c, v, ne0, ne1, umin, TN, EB, EBN;
  uint32_t msg_parity;
  bool active, staged, use_msg;
};
// s_bar[0]: source window, s_bar[1]: message rows of the current batch, s_bar[2]: gradient tile

template <int MODE, int VEC, int NS>
__device__ __forceinline__ void tile_prologue(const KernelArgs& k, const TileCfg& tc, unsigned char* smem, Tile& T) {
  const AggPlan& P = k.plan;
  const int tid = threadIdx.x, TN = tc.TN, nthr = blockDim.x;
  T.s_ptr = reinterpret_cast<int*>(smem);
  T.s_ev = reinterpret_cast<float*>(smem + tc.off_ev);
  T.s_zw = reinterpret_cast<float*>(smem + tc.off_zw);
  T.s_zabs = reinterpret_cast<float*>(smem + tc.off_zabs);
  T.s_f0 = reinterpret_cast<float*>(smem + tc.off_f0);
  T.s_f1 = reinterpret_cast<float*>(smem + tc.off_f1);
  T.s_coef = reinterpret_cast<float*>(smem + tc.off_coef);
  T.s_src = reinterpret_cast<int*>(smem + tc.off_src);
  T.s_w = reinterpret_cast<float*>(smem + tc.off_w);
  T.s_bar = reinterpret_cast<uint64_t*>(smem + tc.off_bar);
  T.s_red = reinterpret_cast<int*>(smem + tc.off_red);
  T.s_win = reinterpret_cast<float*>(smem + tc.off_win);
  T.s_msg = reinterpret_cast<float*>(smem + tc.off_msg);
  T.s_gt = reinterpret_cast<float*>(smem + tc.off_gt);
  T.TN = TN;
  T.EB = tc.EB;
  T.EBN = tc.EBN;
  T.v0 = blockIdx.x * TN;
  T.nv = min(TN, k.N - T.v0);
  T.ln = tid / P.chunks;
  T.c = (tid - T.ln * P.chunks) * VEC;
  T.active = T.ln < T.nv;
  T.v = T.v0 + T.ln;
  T.msg_parity = 0;
  if (tid == 0) {
    mbar_init(&T.s_bar[0], 1);
    mbar_init(&T.s_bar[1], 1);
    mbar_init(&T.s_bar[2], 1);
    mbar_fence_init();
  }
  for (int i = tid; i <= T.nv; i += nthr) T.s_ptr[i] = __ldg(k.in_ptr + T.v0 + i);
  for (int i = tid; i < NS * TN; i += nthr) {
    const int s = i / TN, l = i - s * TN;
    T.s_ev[i] = (s < P.n_slots && l < T.nv) ? __ldg(k.eig + (size_t)(T.v0 + l) * k.ld_eig + P.slot_eig[s]) : 0.f;
    T.s_zw[i] = 0.f;
    T.s_zabs[i] = 0.f;
  }
  __syncthreads();
  T.E0 = T.s_ptr[0];
  T.E1 = T.s_ptr[T.nv];
  T.ne0 = T.active ? T.s_ptr[T.ln] : 0;
  T.ne1 = T.active ? T.s_ptr[T.ln + 1] : 0;
  T.staged = false;
  T.umin = 0;
  T.use_msg = tc.use_msg != 0;
  if constexpr (MODE != DGN_MSG_DENSE && VEC == 4) {
    if (tc.win_rows > 0 && T.E1 > T.E0) {
      // source window of the tile: block-wide min / max of in_src over the tile's slots
      int lo = INT_MAX, hi = -1;
      for (int e = T.E0 + tid; e < T.E1; e += nthr) {
        const int u = __ldg(k.in_src + e);
        lo = min(lo, u);
        hi = max(hi, u);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
      }
      if ((tid & 31) == 0) { T.s_red[tid >> 5] = lo; T.s_red[8 + (tid >> 5)] = hi; }
      __syncthreads();
      lo = T.s_red[0]; hi = T.s_red[8];
      for (int w = 1; w < nthr / 32; ++w) { lo = min(lo, T.s_red[w]); hi = max(hi, T.s_red[8 + w]); }
      const int rows = hi - lo + 1;
      // stage only when the window fits and every staged row is re-used on average at least twice
      T.staged = rows <= tc.win_rows && (T.E1 - T.E0) >= 2 * rows;
      if (T.staged) {
        T.umin = lo;
        T.use_msg = false;
        const uint32_t row_bytes = (uint32_t)P.F * 4u;
        if (tid == 0) mbar_expect_tx(&T.s_bar[0], row_bytes * (uint32_t)rows);
        if (k.ld_x == P.F) {                      // rows are contiguous: one bulk copy
          if (tid == 0) bulk_g2s(T.s_win, k.x + (size_t)lo * k.ld_x, row_bytes * (uint32_t)rows, &T.s_bar[0]);
        } else {
          for (int r = tid; r < rows; r += nthr)
            bulk_g2s(T.s_win + (size_t)r * P.F, k.x + (size_t)(lo + r) * k.ld_x, row_bytes, &T.s_bar[0]);
        }
      }
    }
  }
}

// Phase A of round r: slot j of local node l is global in-edge slot s_ptr[l] + r*EBN + j (if it exists) and lives at
// index l*EBN + j of the round buffers.  Kick off the message-row copies, compute the eigen-weights into
// s_src / s_w, then (optionally) add this round to the per-(node, slot) sums.
template <int MODE, int VEC, int NS>
__device__ __forceinline__ void tile_round_begin(const KernelArgs& k, Tile& T, int r, bool accumulate_z) {
  const AggPlan& P = k.plan;
  const int tid = threadIdx.x, TN = T.TN, nthr = blockDim.x, EBN = T.EBN;
  if (T.use_msg && tid == 0) {
    int cnt = 0;                                // valid slots of this round
    for (int l = 0; l < T.nv; ++l) cnt += min(EBN, max(0, T.s_ptr[l + 1] - T.s_ptr[l] - r * EBN));
    mbar_expect_tx(&T.s_bar[1], (uint32_t)cnt * (uint32_t)P.F * 4u);
  }
  for (int i = tid; i < T.nv * EBN; i += nthr) {
    const int l = i / EBN, j = i - l * EBN;
    const int e = T.s_ptr[l] + r * EBN + j;
    if (e >= T.s_ptr[l + 1]) continue;
    const int u = __ldg(k.in_src + e);
    T.s_src[i] = u;
    if (T.use_msg) {                            // whole message row -> shared memory, asynchronously
      const float* row;
      if constexpr (MODE == DGN_MSG_DENSE) row = k.r + (size_t)(k.in_eid ? __ldg(k.in_eid + e) : e) * k.ld_r;
      else row = k.x + (size_t)u * k.ld_x;
      bulk_g2s(T.s_msg + (size_t)i * P.F, row, (uint32_t)P.F * 4u, &T.s_bar[1]);
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      if (s < P.n_slots) {
        const float d = __ldg(k.eig + (size_t)u * k.ld_eig + P.slot_eig[s]) - T.s_ev[s * TN + l];
        T.s_w[s * T.EB + i] = edge_weight(P.slot_w[s], d, 0.f, 0.f);
      }
    }
  }
  __syncthreads();
  if constexpr (NS > 0) {
    if (accumulate_z) {                         // one thread per (node, slot): sequential, edge-id order
      for (int q = tid; q < T.nv * NS; q += nthr) {
        const int l = q / NS, s = q - l * NS;
        if (s < P.n_slots) {
          const int cnt = min(EBN, max(0, T.s_ptr[l + 1] - T.s_ptr[l] - r * EBN));
          float zw = T.s_zw[s * TN + l], za = T.s_zabs[s * TN + l];
          for (int j = 0; j < cnt; ++j) {
            const float w = T.s_w[s * T.EB + l * EBN + j];
            zw += w;
            za += fabsf(w);
          }
          T.s_zw[s * TN + l] = zw;
          T.s_zabs[s * TN + l] = za;
        }
      }
    }
  }
  if (T.use_msg) {                              // every thread observes the completed copies of this round
    mbar_wait(&T.s_bar[1], T.msg_parity);
    T.msg_parity ^= 1u;
  }
}

// message of batch slot i (global slot e, source u) for this thread's columns
template <int MODE, int VEC>
__device__ __forceinline__ Vec<VEC> tile_message(const KernelArgs& k, const Tile& T, int i, int e, const Vec<VEC>& qv) {
  const int u = T.s_src[i];
  if constexpr (VEC == 4) {
    const bool from_win = (MODE != DGN_MSG_DENSE) && T.staged;
    if (from_win || T.use_msg) {
      const float* row = from_win ? T.s_win + (size_t)(u - T.umin) * k.plan.F : T.s_msg + (size_t)i * k.plan.F;
      const float4 t = *reinterpret_cast<const float4*>(row + T.c);
      Vec<VEC> m;
      m.a[0] = t.x; m.a[1] = t.y; m.a[2] = t.z; m.a[3] = t.w;
      if constexpr (MODE == DGN_MSG_AFFINE) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) m.a[j] += qv.a[j];
        if (k.r) {
          const int id = k.in_eid ? __ldg(k.in_eid + e) : e;
          const Vec<VEC> rv = vload<VEC>(k.r + (size_t)id * k.ld_r + T.c);
#pragma unroll
          for (int j = 0; j < VEC; ++j) m.a[j] += rv.a[j];
        }
      }
      return m;
    }
  }
  return load_message<MODE, VEC>(k, u, e, T.c, qv);
}

// per-node scaler coefficients (rb/nets/scalers.py): 1, log(D+1)/avg, avg/log(D+1)
__device__ __forceinline__ void tile_scaler_coefs(const KernelArgs& k, Tile& T) {
  const AggPlan& P = k.plan;
  for (int j = threadIdx.x; j < T.nv * DGN_MAX_SCALERS; j += blockDim.x) {
    const int l = j / DGN_MAX_SCALERS, s = j - l * DGN_MAX_SCALERS;
    float cf = 1.f;
    if (P.S > 1 && s < P.S) {
      const float ld = __ldg(k.log_deg + T.v0 + l);
      const int kind = P.scaler_kind[s];
      cf = (kind == DGN_SCALE_AMPLIFICATION) ? __fdiv_rn(ld, P.avg_log)
           : (kind == DGN_SCALE_ATTENUATION) ? __fdiv_rn(P.avg_log, ld) : 1.f;
    }
    T.s_coef[s * T.TN + l] = cf;
  }
}

// per-(node, slot) factors derived from the completed sums
template <int NS>
__device__ __forceinline__ void tile_factors(const KernelArgs& k, Tile& T) {
  const AggPlan& P = k.plan;
  const int TN = T.TN;
  for (int j = threadIdx.x; j < T.nv * NS; j += blockDim.x) {
    const int l = j / NS, s = j - l * NS;
    if (s < P.n_slots) {
      const int kind = P.slot_w[s];
      const float zw = T.s_zw[s * TN + l], za = T.s_zabs[s * TN + l];
      float f0;
      if (kind == W_POS || kind == W_NEG) f0 = 0.5f * __frcp_rn(zw + DGN_EPS);   // balanced halves
      else f0 = __frcp_rn(za + DGN_EPS);                                          // av / dx: 1 / (sum |delta| + eps)
      T.s_f0[s * TN + l] = f0;
      T.s_f1[s * TN + l] = zw * f0;                                               // W = sum(w) / Z
    }
  }
}

template <int VEC, int NS, bool ISO>
__device__ __forceinline__ void acc_init(RowAcc<VEC, NS, ISO>& R) {
  R.sum = vfill<VEC>(0.f);
  if constexpr (ISO) {
    R.sq = vfill<VEC>(0.f);
    R.mx = vfill<VEC>(-INFINITY);
    R.mn = vfill<VEC>(INFINITY);
  }
#pragma unroll
  for (int s = 0; s < NS; ++s) R.acc[s] = vfill<VEC>(0.f);
}

template <int VEC, int NS, bool ISO>
__device__ __forceinline__ void acc_edge(const AggPlan& P, RowAcc<VEC, NS, ISO>& R, const Vec<VEC>& m, const float* s_w,
                                         int EB, int i) {
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    R.sum.a[j] += m.a[j];
    if constexpr (ISO) {
      R.sq.a[j] = __fadd_rn(R.sq.a[j], __fmul_rn(m.a[j], m.a[j]));   // square, round, then sum (as the reference)
      R.mx.a[j] = fmaxf(R.mx.a[j], m.a[j]);
      R.mn.a[j] = fminf(R.mn.a[j], m.a[j]);
    }
  }
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    if (s < P.n_slots) {
      const float w = s_w[s * EB + i];
#pragma unroll
      for (int j = 0; j < VEC; ++j) R.acc[s].a[j] = fmaf(w, m.a[j], R.acc[s].a[j]);
    }
  }
}

template <int VEC, bool ISO>
__device__ __forceinline__ void mean_var(const Vec<VEC>& sum, const Vec<VEC>& sq, float fD, Vec<VEC>& mean, Vec<VEC>& var) {
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    // IEEE division: keeps var == 0 exactly for constant mailboxes (the reference's relu / sqrt gradient jumps there)
    mean.a[i] = __fdiv_rn(sum.a[i], fD);
    if constexpr (ISO) {
      const float msq = __fdiv_rn(sq.a[i], fD);
      var.a[i] = fmaxf(__fsub_rn(msq, __fmul_rn(mean.a[i], mean.a[i])), 0.f);
    } else {
      var.a[i] = 0.f;
    }
  }
}

// ======================================================================================================
// forward
// ======================================================================================================
template <int MODE, int VEC, int NS, bool ISO>
__global__ void __launch_bounds__(kMaxTileThreads) agg_fwd_tile_kernel(const __grid_constant__ KernelArgs k,
                                                                       const __grid_constant__ TileCfg tc) {
  extern __shared__ __align__(128) unsigned char smem[];
  const AggPlan& P = k.plan;
  Tile T;
  tile_prologue<MODE, VEC, NS>(k, tc, smem, T);
  const int TN = T.TN;

  Vec<VEC> hv = vfill<VEC>(0.f), qv = vfill<VEC>(0.f);
  int tower = 0, cg = 0;
  if (T.active) {
    tower = T.c / P.Fg;
    cg = T.c - tower * P.Fg;
    hv = vload<VEC>(k.h_in + (size_t)T.v * k.ld_h + T.c);
    if (k.h_copy) vstore<VEC>(k.h_copy + (size_t)T.v * k.ld_hc + (size_t)tower * k.hc_gs + cg, hv);
    if constexpr (MODE == DGN_MSG_AFFINE) {
      qv = vload<VEC>(k.q + (size_t)T.v * k.ld_q + T.c);
      if (k.q_bias) {
        const Vec<VEC> bv = vload<VEC>(k.q_bias + T.c);
#pragma unroll
        for (int i = 0; i < VEC; ++i) qv.a[i] += bv.a[i];
      }
    }
  }

  RowAcc<VEC, NS, ISO> R;
  acc_init(R);
  bool win_waited = false;
  const int Dn = T.ne1 - T.ne0;
  for (int r = 0; __syncthreads_or(T.active && r * T.EBN < Dn); ++r) {   // barrier + "any node has slots left"
    tile_round_begin<MODE, VEC, NS>(k, T, r, true);
    if (T.staged && !win_waited) { mbar_wait(&T.s_bar[0], 0); win_waited = true; }
    if (T.active) {
      const int cnt = min(T.EBN, max(0, Dn - r * T.EBN)), base = T.ln * T.EBN, e0 = T.ne0 + r * T.EBN;
#pragma unroll 4
      for (int j = 0; j < cnt; ++j) {
        const Vec<VEC> m = tile_message<MODE, VEC>(k, T, base + j, e0 + j, qv);
        acc_edge<VEC, NS, ISO>(P, R, m, T.s_w, T.EB, base + j);
      }
    }
  }
  tile_factors<NS>(k, T);
  tile_scaler_coefs(k, T);
  __syncthreads();
  if (!T.active) return;

  float* orow = k.out + (size_t)T.v * k.ld_out + (size_t)tower * k.out_gs + cg;
  const int D = T.ne1 - T.ne0;
  if (D == 0) {                                        // DGL: zero rows for isolated nodes
    const Vec<VEC> z = vfill<VEC>(0.f);
    for (int j = 0; j < P.S * P.A; ++j) vstore_stream<VEC>(orow + (size_t)j * P.Fg, z);
    return;
  }
  float coef[DGN_MAX_SCALERS];
#pragma unroll
  for (int s = 0; s < DGN_MAX_SCALERS; ++s) coef[s] = T.s_coef[s * TN + T.ln];
  Vec<VEC> mean, var;
  mean_var<VEC, ISO>(R.sum, R.sq, (float)D, mean, var);

  const int scaler_stride = P.A * P.Fg;
  auto store_scaled = [&](int a, const Vec<VEC>& y) {
    float* dst = orow + a * P.Fg;
#pragma unroll
    for (int s = 0; s < DGN_MAX_SCALERS; ++s) {
      if (s < P.S) {
        Vec<VEC> o;
#pragma unroll
        for (int i = 0; i < VEC; ++i) o.a[i] = y.a[i] * coef[s];
        vstore_stream<VEC>(dst + s * scaler_stride, o);
      }
    }
  };
  for (int a = 0; a < P.A; ++a) {                      // isotropic aggregators
    const int kind = P.agg_kind[a];
    if (kind >= DGN_AGG_DIR_AV) continue;
    Vec<VEC> y;
    switch (kind) {
      case DGN_AGG_MEAN: y = mean; break;
      case DGN_AGG_SUM: y = R.sum; break;
      case DGN_AGG_MAX: if constexpr (ISO) y = R.mx; else y = mean; break;
      case DGN_AGG_MIN: if constexpr (ISO) y = R.mn; else y = mean; break;
      case DGN_AGG_VAR: y = var; break;
      default:
#pragma unroll
        for (int i = 0; i < VEC; ++i) y.a[i] = sqrtf(var.a[i] + DGN_EPS);
    }
    store_scaled(a, y);
  }
#pragma unroll
  for (int s = 0; s < NS; ++s) {                       // directional aggregators, static slot index
    if (s >= P.n_slots) break;
    const float f0 = T.s_f0[s * TN + T.ln], f1 = T.s_f1[s * TN + T.ln];
    unsigned todo = P.slot_aggs[s];
    while (todo) {
      const int a = __ffs(todo) - 1;
      todo &= todo - 1;
      const int kind = P.agg_kind[a];
      const Vec<VEC>& A1 = R.acc[s];
      Vec<VEC> y;
      if (kind == DGN_AGG_DIR_AV) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) y.a[i] = A1.a[i] * f0;
      } else if (kind == DGN_AGG_DIR_DX_BALANCED) {
        if constexpr (NS >= 2) {                       // the W_NEG half always sits in slot s+1
          constexpr int LAST = NS - 1;
          const int s2 = (s + 1 <= LAST) ? s + 1 : LAST;
          const float g0 = T.s_f0[s2 * TN + T.ln], wsum = f1 + T.s_f1[s2 * TN + T.ln];
          const Vec<VEC>& A2 = R.acc[s2];
#pragma unroll
          for (int i = 0; i < VEC; ++i) y.a[i] = fabsf(A1.a[i] * f0 + A2.a[i] * g0 - wsum * hv.a[i]);
        } else {
          y = vfill<VEC>(0.f);
        }
      } else {                                         // dx / dx-no-abs
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const float sv = A1.a[i] * f0 - f1 * hv.a[i];
          y.a[i] = (kind == DGN_AGG_DIR_DX) ? fabsf(sv) : sv;
        }
      }
      store_scaled(a, y);
    }
  }
}

// ======================================================================================================
// backward, destination side (the source-side gather is agg_bwd_src_kernel in dgn_agg_bwd.cu)
// ======================================================================================================
template <int MODE, int VEC, int NS, bool ISO>
__global__ void __launch_bounds__(kMaxTileThreads) agg_bwd_tile_kernel(const __grid_constant__ KernelArgs k,
                                                                       const __grid_constant__ TileCfg tc) {
  extern __shared__ __align__(128) unsigned char smem[];
  const AggPlan& P = k.plan;
  const int tid = threadIdx.x;
  Tile T;
  tile_prologue<MODE, VEC, NS>(k, tc, smem, T);
  const int TN = T.TN;

  // the tile's gradient slabs: one bulk copy per node row (S*A*F contiguous floats), issued first so they are
  // in flight during everything that follows
  const bool use_gt = tc.use_gt != 0;
  if (use_gt) {
    if (tid == 0) mbar_expect_tx(&T.s_bar[2], (uint32_t)T.nv * (uint32_t)tc.gt_row * 4u);
    for (int l = tid; l < T.nv; l += blockDim.x)
      bulk_g2s(T.s_gt + (size_t)l * tc.gt_row, k.g_out + (size_t)(T.v0 + l) * k.ld_out, (uint32_t)tc.gt_row * 4u,
               &T.s_bar[2]);
  }
  tile_scaler_coefs(k, T);

  int tower = 0, cg = 0;
  Vec<VEC> hv = vfill<VEC>(0.f), qv = vfill<VEC>(0.f), dh = vfill<VEC>(0.f);
  const int D = T.ne1 - T.ne0;
  if (T.active) {
    tower = T.c / P.Fg;
    cg = T.c - tower * P.Fg;
    if (k.g_hcopy) dh = vload_stream<VEC>(k.g_hcopy + (size_t)T.v * k.ld_hc + (size_t)tower * k.hc_gs + cg);
    if (k.d_h_add) {
      const Vec<VEC> t = vload_stream<VEC>(k.d_h_add + (size_t)T.v * k.ld_dha + T.c);
#pragma unroll
      for (int i = 0; i < VEC; ++i) dh.a[i] += t.a[i];
    }
    if (D > 0) {
      hv = vload<VEC>(k.h_in + (size_t)T.v * k.ld_h + T.c);
      if constexpr (MODE == DGN_MSG_AFFINE) {
        qv = vload<VEC>(k.q + (size_t)T.v * k.ld_q + T.c);
        if (k.q_bias) {
          const Vec<VEC> bv = vload<VEC>(k.q_bias + T.c);
#pragma unroll
          for (int i = 0; i < VEC; ++i) qv.a[i] += bv.a[i];
        }
      }
    }
  }
  Vec<VEC>* sG = reinterpret_cast<Vec<VEC>*>(smem + tc.off_g);
  if (!use_gt) {
    __syncthreads();                                   // s_coef visible
    if (T.active && D > 0) {
      // G_a = sum_s coef_s * g_out[v, s, a, :] for every aggregator, staged per thread in shared memory: the S*A
      // slab loads dominate this kernel's traffic and are issued back to back (unroll 4 => 4*S loads in flight)
      float coef[DGN_MAX_SCALERS];
#pragma unroll
      for (int s = 0; s < DGN_MAX_SCALERS; ++s) coef[s] = T.s_coef[s * TN + T.ln];
      const float* grow = k.g_out + (size_t)T.v * k.ld_out + (size_t)tower * k.out_gs + cg;
      const int scaler_stride = P.A * P.Fg;
#pragma unroll 4
      for (int a = 0; a < P.A; ++a) {
        Vec<VEC> G = vfill<VEC>(0.f);
        const float* src = grow + a * P.Fg;
#pragma unroll
        for (int s = 0; s < DGN_MAX_SCALERS; ++s) {
          if (s < P.S) {
            const Vec<VEC> gs = vload_stream<VEC>(src + s * scaler_stride);
#pragma unroll
            for (int i = 0; i < VEC; ++i) G.a[i] = fmaf(coef[s], gs.a[i], G.a[i]);
          }
        }
        sG[a * (int)blockDim.x + tid] = G;
      }
    }
  }

  // ---- pass 1: recompute the row statistics ------------------------------------------------------------
  RowAcc<VEC, NS, ISO> R;
  acc_init(R);
  bool win_waited = false;
  int rounds = 0;
  for (int r = 0; __syncthreads_or(T.active && r * T.EBN < D); ++r) {    // barrier + "any node has slots left"
    tile_round_begin<MODE, VEC, NS>(k, T, r, true);
    if (T.staged && !win_waited) { mbar_wait(&T.s_bar[0], 0); win_waited = true; }
    if (T.active) {
      const int cnt = min(T.EBN, max(0, D - r * T.EBN)), base = T.ln * T.EBN, e0 = T.ne0 + r * T.EBN;
#pragma unroll 4
      for (int j = 0; j < cnt; ++j) {
        const Vec<VEC> m = tile_message<MODE, VEC>(k, T, base + j, e0 + j, qv);
        acc_edge<VEC, NS, ISO>(P, R, m, T.s_w, T.EB, base + j);
      }
    }
    ++rounds;
  }
  const bool single_batch = rounds <= 1;               // round buffers still hold every slot of the tile
  __syncthreads();                                     // sums complete (and s_coef visible) before the factors
  tile_factors<NS>(k, T);
  __syncthreads();
  if (use_gt) mbar_wait(&T.s_bar[2], 0);

  // ---- fold the S*A gradient slabs into per-column coefficients ---------------------------------------
  Vec<VEC> c0 = vfill<VEC>(0.f), c1 = vfill<VEC>(0.f), gmx = vfill<VEC>(0.f), gmn = vfill<VEC>(0.f);
  Vec<VEC> cs[NS > 0 ? NS : 1];
#pragma unroll
  for (int s = 0; s < NS; ++s) cs[s] = vfill<VEC>(0.f);
  if (T.active && D > 0) {
    float coef[DGN_MAX_SCALERS];
#pragma unroll
    for (int s = 0; s < DGN_MAX_SCALERS; ++s) coef[s] = T.s_coef[s * TN + T.ln];
    const int scaler_stride = P.A * P.Fg;
    const float* grow_s = T.s_gt + (size_t)T.ln * tc.gt_row + cg;          // single tower when use_gt
    auto slab_grad = [&](int a) {               // G_a = sum_s coef_s * g_out[v, s, a, :]
      if (!use_gt) return sG[a * (int)blockDim.x + tid];
      Vec<VEC> G = vfill<VEC>(0.f);
#pragma unroll
      for (int s = 0; s < DGN_MAX_SCALERS; ++s) {
        if (s < P.S) {
          Vec<VEC> gs;
          if constexpr (VEC == 4) {
            const float4 t = *reinterpret_cast<const float4*>(grow_s + a * P.Fg + s * scaler_stride);
            gs.a[0] = t.x; gs.a[1] = t.y; gs.a[2] = t.z; gs.a[3] = t.w;
          } else {
#pragma unroll
            for (int i = 0; i < VEC; ++i) gs.a[i] = grow_s[a * P.Fg + s * scaler_stride + i];
          }
#pragma unroll
          for (int i = 0; i < VEC; ++i) G.a[i] = fmaf(coef[s], gs.a[i], G.a[i]);
        }
      }
      return G;
    };
    const float fD = (float)D, rD = __frcp_rn(fD);
    Vec<VEC> mean, var;
    mean_var<VEC, ISO>(R.sum, R.sq, fD, mean, var);
    for (int a = 0; a < P.A; ++a) {                    // isotropic aggregators
      const int kind = P.agg_kind[a];
      if (kind >= DGN_AGG_DIR_AV) continue;
      const Vec<VEC> G = slab_grad(a);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        if (kind == DGN_AGG_MEAN) c0.a[i] += G.a[i] * rD;
        else if (kind == DGN_AGG_SUM) c0.a[i] += G.a[i];
        else if (kind == DGN_AGG_MAX) gmx.a[i] += G.a[i];
        else if (kind == DGN_AGG_MIN) gmn.a[i] += G.a[i];
        else {
          // var = relu(t), t = E[m^2] - E[m]^2 ; dt/dm_u = 2 (m_u - mean) / D ; relu'(0) = 0
          float gv = (var.a[i] > 0.f) ? G.a[i] : 0.f;
          if (kind == DGN_AGG_STD) gv *= 0.5f * rsqrtf(var.a[i] + DGN_EPS);
          const float two_over_d = 2.f * gv * rD;
          c1.a[i] += two_over_d;
          c0.a[i] -= two_over_d * mean.a[i];
        }
      }
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) {                     // directional aggregators, static slot index
      if (s >= P.n_slots) break;
      const float f0 = T.s_f0[s * TN + T.ln], f1 = T.s_f1[s * TN + T.ln];
      unsigned todo = P.slot_aggs[s];
      while (todo) {
        const int a = __ffs(todo) - 1;
        todo &= todo - 1;
        const int kind = P.agg_kind[a];
        const Vec<VEC> G = slab_grad(a);
        const Vec<VEC>& A1 = R.acc[s];
        if (kind == DGN_AGG_DIR_AV) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) cs[s].a[i] = fmaf(G.a[i], f0, cs[s].a[i]);
        } else if (kind == DGN_AGG_DIR_DX_BALANCED) {
          if constexpr (NS >= 2) {
            constexpr int LAST = NS - 1;
            const int s2 = (s + 1 <= LAST) ? s + 1 : LAST;
            const float g0 = T.s_f0[s2 * TN + T.ln], wsum = f1 + T.s_f1[s2 * TN + T.ln];
            const Vec<VEC>& A2 = R.acc[s2];
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
              const float sv = A1.a[i] * f0 + A2.a[i] * g0 - wsum * hv.a[i];
              const float g = G.a[i] * sign0(sv);
              cs[s].a[i] = fmaf(g, f0, cs[s].a[i]);
              cs[s2].a[i] = fmaf(g, g0, cs[s2].a[i]);
              dh.a[i] -= wsum * g;
            }
          }
        } else {                                       // dx / dx-no-abs
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            float g = G.a[i];
            if (kind == DGN_AGG_DIR_DX) g *= sign0(A1.a[i] * f0 - f1 * hv.a[i]);   // same expression as the forward
            cs[s].a[i] = fmaf(g, f0, cs[s].a[i]);
            dh.a[i] -= f1 * g;
          }
        }
      }
    }
  }

  // ---- pass 2: per-edge message gradients ----------------------------------------------------------------
  Vec<VEC> dq = vfill<VEC>(0.f);
  unsigned given = 0u;          // bit i: max gradient of column i already routed; bit 4+i: min
  for (int r = 0; r < rounds; ++r) {
    if (!single_batch) {
      __syncthreads();                                 // previous round's buffers are no longer read
      tile_round_begin<MODE, VEC, NS>(k, T, r, false); // this round's rows / weights again
    }
    if (T.active) {
      const int cnt = min(T.EBN, max(0, D - r * T.EBN)), base = T.ln * T.EBN, e0 = T.ne0 + r * T.EBN;
#pragma unroll 2
      for (int j = 0; j < cnt; ++j) {
        const int e = e0 + j, i = base + j;
        const Vec<VEC> m = tile_message<MODE, VEC>(k, T, i, e, qv);
        Vec<VEC> dm;
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
          float g = fmaf(c1.a[q], m.a[q], c0.a[q]);
          if constexpr (ISO) {
            // torch.max / torch.min send the whole gradient to the FIRST extremal mailbox entry
            if (m.a[q] == R.mx.a[q] && !(given & (1u << q))) { g += gmx.a[q]; given |= 1u << q; }
            if (m.a[q] == R.mn.a[q] && !(given & (16u << q))) { g += gmn.a[q]; given |= 16u << q; }
          }
          dm.a[q] = g;
        }
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          if (s < P.n_slots) {
            const float w = T.s_w[s * T.EB + i];
#pragma unroll
            for (int q = 0; q < VEC; ++q) dm.a[q] = fmaf(w, cs[s].a[q], dm.a[q]);
          }
        }
#pragma unroll
        for (int q = 0; q < VEC; ++q) dq.a[q] += dm.a[q];
        if (k.edge_ws) vstore<VEC>(k.edge_ws + (size_t)e * P.F + T.c, dm);
        if (k.d_r) {
          const int id = k.in_eid ? __ldg(k.in_eid + e) : e;
          vstore<VEC>(k.d_r + (size_t)id * k.ld_dr + T.c, dm);
        }
      }
    }
  }
  if (T.active) {
    if (k.d_q) vstore<VEC>(k.d_q + (size_t)T.v * k.ld_dq + T.c, dq);
    if (k.d_h) vstore<VEC>(k.d_h + (size_t)T.v * k.ld_dh + T.c, dh);
  }
}

// ======================================================================================================
// host side
// ======================================================================================================
static inline int align_up(int x, int a) { return (x + a - 1) / a * a; }

static bool make_tile_cfg(const KernelArgs& k, int vec, int NS, bool backward, TileCfg& tc) {
  const AggPlan& P = k.plan;
  if (P.has_exp || P.chunks > kMaxTileThreads || P.chunks <= 0) return false;
  const int T_towers = P.F / P.Fg;
  const int row_bytes = P.F * 4;
  const bool rows_bulk_ok = vec == 4 && (row_bytes % 16 == 0);
  // backward: gradient tile through TMA when a node's slabs are one contiguous, 16 B aligned run
  tc.gt_row = P.S * P.A * P.F;
  // (measured on B200: the smaller CTAs it forces cost more than the async copy gains for the 7.7 KB rows of
  //  cfg2, so it is only taken when a 256-thread tile of gradient rows fits in 48 KB)
  tc.use_gt = backward && vec == 4 && T_towers == 1 && (k.ld_out % 4 == 0) &&
              ((reinterpret_cast<uintptr_t>(k.g_out) & 15u) == 0) &&
              (kMaxTileThreads / P.chunks) * tc.gt_row * 4 <= 48 * 1024;
  // block size: 256 threads unless that makes the gradient tile too large for >= 2 CTAs per SM
  tc.threads = kMaxTileThreads;
  if (tc.use_gt && (kMaxTileThreads / P.chunks) * tc.gt_row * 4 > 72 * 1024 && P.chunks <= 128) tc.threads = 128;
  tc.TN = tc.threads / P.chunks;
  if (tc.TN > 64) tc.TN = 64;
  if (tc.use_gt && tc.TN * tc.gt_row * 4 > 96 * 1024) tc.use_gt = 0;
  const int TN = tc.TN, ns = NS > 0 ? NS : 1;
  // messages of a batch in shared memory (row-sized bulk copies): SOURCE / AFFINE rows of x, DENSE rows of r
  // High-degree graphs (SBM PATTERN) take the source-window path instead: no per-edge copies, large batches.
  const double avg_deg = k.N > 0 ? (double)k.E / (double)k.N : 0.0;
  const bool dense_graph = avg_deg >= 16.0 && k.mode != DGN_MSG_DENSE;
  tc.use_msg = rows_bulk_ok && !dense_graph && ((k.mode == DGN_MSG_DENSE) ? (k.ld_r % 4 == 0) : (k.ld_x % 4 == 0));
  // slots per node and round: ~2.5x the average in-degree (power of two, 4..64) so that almost every tile is done
  // in one round, bounded so that a round of message rows stays within 32 KB (64 KB for denser graphs) of smem
  int ebn = 4;
  while (ebn < 64 && ebn < 2.5 * avg_deg) ebn <<= 1;
  if (tc.use_msg) {
    const int msg_cap = (avg_deg >= 6.0 ? 64 : 32) * 1024;
    while (ebn > 2 && TN * ebn * row_bytes > msg_cap) ebn >>= 1;
  }
  tc.EBN = ebn;
  tc.EB = TN * ebn;
  int off = align_up((TN + 1) * 4, 16);
  tc.off_ev = off;   off += align_up(ns * TN * 4, 16);
  tc.off_zw = off;   off += align_up(ns * TN * 4, 16);
  tc.off_zabs = off; off += align_up(ns * TN * 4, 16);
  tc.off_f0 = off;   off += align_up(ns * TN * 4, 16);
  tc.off_f1 = off;   off += align_up(ns * TN * 4, 16);
  tc.off_coef = off; off += align_up(DGN_MAX_SCALERS * TN * 4, 16);
  tc.off_src = off;  off += tc.EB * 4;
  tc.off_w = off;    off += ns * tc.EB * 4;
  tc.off_bar = off;  off += 32;
  tc.off_red = off;  off += 64;
  off = align_up(off, 128);
  tc.off_gt = off;
  if (tc.use_gt) off += TN * tc.gt_row * 4;
  off = align_up(off, 128);
  tc.off_g = off;                                      // per-thread staged slab sums G_a (backward without gradient tile)
  if (backward && !tc.use_gt) off += P.A * tc.threads * vec * 4;
  off = align_up(off, 128);
  tc.off_msg = off;
  if (tc.use_msg) off += tc.EB * row_bytes;
  off = align_up(off, 128);
  tc.off_win = off;
  // Source-window staging pays when rows are re-used (average in-degree well above 1).  Its shared-memory
  // reservation costs occupancy, so it is only offered when the graph is dense enough to use it.
  tc.win_rows = 0;
  if (rows_bulk_ok && dense_graph) {
    const int budget = 100 * 1024 - off;
    const int rows = budget > 0 ? budget / row_bytes : 0;
    if (rows >= 32) tc.win_rows = rows;
  }
  off += tc.win_rows * row_bytes;
  tc.total = off;
  return tc.total <= 200 * 1024;
}

template <typename Kern>
static int launch_tile(Kern kern, const KernelArgs& k, const TileCfg& tc, cudaStream_t st) {
  if (k.N == 0) return DGN_OK;
  if (tc.total > 48 * 1024) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return DGN_ERR_CUDA;
  }
  const unsigned grid = (unsigned)((k.N + tc.TN - 1) / tc.TN);
  kern<<<grid, tc.threads, tc.total, st>>>(k, tc);
  return cudaGetLastError() == cudaSuccess ? DGN_OK : DGN_ERR_CUDA;
}

static bool plan_needs_iso(const AggPlan& P) {
  for (int a = 0; a < P.A; ++a) {
    const int kd = P.agg_kind[a];
    if (kd == DGN_AGG_MAX || kd == DGN_AGG_MIN || kd == DGN_AGG_STD || kd == DGN_AGG_VAR) return true;
  }
  return false;
}

template <bool BWD, int MODE, int VEC, int NS>
static int launch_iso(const KernelArgs& k, cudaStream_t st) {
  TileCfg tc;
  if (!make_tile_cfg(k, VEC, NS, BWD, tc)) return DGN_ERR_UNSUPPORTED;
  const bool iso = plan_needs_iso(k.plan);
  if constexpr (BWD) {
    return iso ? launch_tile(agg_bwd_tile_kernel<MODE, VEC, NS, true>, k, tc, st)
               : launch_tile(agg_bwd_tile_kernel<MODE, VEC, NS, false>, k, tc, st);
  } else {
    return iso ? launch_tile(agg_fwd_tile_kernel<MODE, VEC, NS, true>, k, tc, st)
               : launch_tile(agg_fwd_tile_kernel<MODE, VEC, NS, false>, k, tc, st);
  }
}

template <bool BWD, int MODE, int VEC>
static int launch_slots(const KernelArgs& k, cudaStream_t st) {
  const int ns = k.plan.n_slots;
  if (ns == 0) return launch_iso<BWD, MODE, VEC, 0>(k, st);
  if (ns <= 2) return launch_iso<BWD, MODE, VEC, 2>(k, st);
  if (ns <= 4) return launch_iso<BWD, MODE, VEC, 4>(k, st);
  return launch_iso<BWD, MODE, VEC, 8>(k, st);
}

template <bool BWD, int VEC>
static int launch_mode(const KernelArgs& k, cudaStream_t st) {
  if (k.mode == DGN_MSG_SOURCE) return launch_slots<BWD, DGN_MSG_SOURCE, VEC>(k, st);
  if (k.mode == DGN_MSG_AFFINE) return launch_slots<BWD, DGN_MSG_AFFINE, VEC>(k, st);
  return launch_slots<BWD, DGN_MSG_DENSE, VEC>(k, st);
}

// Returns DGN_ERR_UNSUPPORTED when the tile kernels do not cover the request (caller falls back).
int launch_forward_tile(const KernelArgs& k, int vec, cudaStream_t st) {
  if (k.plan.has_exp) return DGN_ERR_UNSUPPORTED;
  if (vec == 4) return launch_mode<false, 4>(k, st);
  if (vec == 2) return launch_mode<false, 2>(k, st);
  return launch_mode<false, 1>(k, st);
}

int launch_backward_tile_dst(const KernelArgs& k, int vec, cudaStream_t st) {
  if (k.plan.has_exp) return DGN_ERR_UNSUPPORTED;
  if (vec == 4) return launch_mode<true, 4>(k, st);
  if (vec == 2) return launch_mode<true, 2>(k, st);
  return launch_mode<true, 1>(k, st);
}

}  // namespace dgn
