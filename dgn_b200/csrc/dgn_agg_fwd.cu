// Fused DGN aggregation, forward (sm_100a).
//
// One thread owns one (destination node, VEC-column chunk) pair and walks that node's in-edge
// slots in edge-id order - the order of the reference's mailbox (rb/nets/dgn_layer.py:86-98 on
// DGL 0.4.2 degree bucketing).  In one pass it accumulates everything all requested aggregators
// need (sum, sum of squares, max, min and up to NS eigen-weighted sums), then writes all
// S*A output slabs of the node, each slab store being a contiguous 16 B per lane
// (F*4 B contiguous per node).  The eigenvector weights depend only on (u, v), never on the
// feature column, so every aggregator and every tower is served by the same pass.
//
// HBM roofline: compulsory bytes per node are  4*(r*F + K_u + 1) read  +  4*S*A*F written
// (SURVEY.md 8(d)); the output stores dominate, so they are issued as streaming (evict-first)
// 128-bit stores while the gathered operand rows stay L2-resident.
#include <stdlib.h>

#include "dgn_plan.cuh"

namespace dgn {

template <int MODE, int VEC, int NS, bool ISO, bool EXP>
__global__ void __launch_bounds__(256) agg_fwd_kernel(const __grid_constant__ KernelArgs k) {
  const AggPlan& P = k.plan;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int v = (int)(tid / P.chunks);
  if (v >= k.N) return;
  const int c = (int)(tid - (long long)v * P.chunks) * VEC;      // first column of this thread
  const int tower = c / P.Fg;
  const int cg = c - tower * P.Fg;                               // column inside the tower

  const int e0 = __ldg(k.in_ptr + v), e1 = __ldg(k.in_ptr + v + 1);
  const int D = e1 - e0;

  const Vec<VEC> hv = vload<VEC>(k.h_in + (size_t)v * k.ld_h + c);
  if (k.h_copy) vstore<VEC>(k.h_copy + (size_t)v * k.ld_hc + (size_t)tower * k.hc_gs + cg, hv);

  float* orow = k.out + (size_t)v * k.ld_out + (size_t)tower * k.out_gs + cg;
  if (D == 0) {                                                  // DGL: zero rows for isolated nodes
    const Vec<VEC> z = vfill<VEC>(0.f);
    for (int j = 0; j < P.S * P.A; ++j) vstore_stream<VEC>(orow + (size_t)j * P.Fg, z);
    return;
  }

  Vec<VEC> qv = vfill<VEC>(0.f);
  if constexpr (MODE == DGN_MSG_AFFINE) {
    qv = vload<VEC>(k.q + (size_t)v * k.ld_q + c);
    if (k.q_bias) {
      const Vec<VEC> bv = vload<VEC>(k.q_bias + c);
#pragma unroll
      for (int i = 0; i < VEC; ++i) qv.a[i] += bv.a[i];
    }
  }

  float ev[NS > 0 ? NS : 1], shift[NS > 0 ? NS : 1];
#pragma unroll
  for (int s = 0; s < NS; ++s) ev[s] = (s < P.n_slots) ? __ldg(k.eig + (size_t)v * k.ld_eig + P.slot_eig[s]) : 0.f;

  RowAcc<VEC, NS, ISO> R;
  accumulate_row<MODE, VEC, NS, ISO, EXP>(k, v, c, e0, e1, qv, ev, shift, R);

  float coef[DGN_MAX_SCALERS];
  scaler_coefs(k, v, coef);

  const float fD = (float)D;
  Vec<VEC> mean, var;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    // IEEE division (not a reciprocal multiply): keeps var == 0 exactly for constant mailboxes, where
    // the reference's relu / sqrt gradient is discontinuous
    mean.a[i] = __fdiv_rn(R.sum.a[i], fD);
    if constexpr (ISO) {
      const float msq = __fdiv_rn(R.sq.a[i], fD);
      var.a[i] = fmaxf(__fsub_rn(msq, __fmul_rn(mean.a[i], mean.a[i])), 0.f);
    } else {
      var.a[i] = 0.f;
    }
  }

  const int scaler_stride = P.A * P.Fg;            // distance between the slabs of two scalers
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

  // isotropic aggregators: only statically named registers
  for (int a = 0; a < P.A; ++a) {
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

  // directional aggregators: the slot loop is unrolled so that every accumulator keeps a static
  // register name (a runtime slot index would push the accumulators to local memory)
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    if (s >= P.n_slots) break;
    unsigned todo = P.slot_aggs[s];
    while (todo) {
      const int a = __ffs(todo) - 1;
      todo &= todo - 1;
      const int kind = P.agg_kind[a];
      const Vec<VEC>& A1 = R.acc[s];
      const float zw1 = R.zw[s], zabs1 = R.zabs[s];
      Vec<VEC> y;
      if (kind == DGN_AGG_DIR_AV) {
        const float rz = __frcp_rn(zabs1 + DGN_EPS);
#pragma unroll
        for (int i = 0; i < VEC; ++i) y.a[i] = A1.a[i] * rz;
      } else if (kind == DGN_AGG_DIR_DX || kind == DGN_AGG_DIR_DX_NO_ABS) {
        const float rz = __frcp_rn(zabs1 + DGN_EPS);
        const float wsum = zw1 * rz;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const float sv = A1.a[i] * rz - wsum * hv.a[i];
          y.a[i] = (kind == DGN_AGG_DIR_DX) ? fabsf(sv) : sv;
        }
      } else if (kind == DGN_AGG_DIR_DX_BALANCED) {
        if constexpr (NS >= 2) {                       // the W_NEG half always sits in slot s+1
          constexpr int LAST = NS - 1;
          const int s2 = (s + 1 <= LAST) ? s + 1 : LAST;
          const Vec<VEC>& A2 = R.acc[s2];
          const float zw2 = R.zw[s2];
          const float rp = 0.5f * __frcp_rn(zw1 + DGN_EPS), rn = 0.5f * __frcp_rn(zw2 + DGN_EPS);
          const float wsum = zw1 * rp + zw2 * rn;
#pragma unroll
          for (int i = 0; i < VEC; ++i) y.a[i] = fabsf(A1.a[i] * rp + A2.a[i] * rn - wsum * hv.a[i]);
        } else {
          y = vfill<VEC>(0.f);
        }
      } else {                                         // DGN_AGG_DIR_SOFTMAX
        const float rz = __frcp_rn(zw1);
#pragma unroll
        for (int i = 0; i < VEC; ++i) y.a[i] = A1.a[i] * rz;
      }
      store_scaled(a, y);
    }
  }
}

template <int MODE, int VEC, int NS, bool ISO, bool EXP>
static int launch_one(const KernelArgs& k, cudaStream_t st) {
  const long long threads = (long long)k.N * k.plan.chunks;
  if (threads == 0) return DGN_OK;
  const int block = 256;
  const long long grid = (threads + block - 1) / block;
  agg_fwd_kernel<MODE, VEC, NS, ISO, EXP><<<(unsigned)grid, block, 0, st>>>(k);
  return cudaGetLastError() == cudaSuccess ? DGN_OK : DGN_ERR_CUDA;
}

template <int MODE, int VEC>
static int dispatch_slots(const KernelArgs& k, bool iso, cudaStream_t st) {
  const int ns = k.plan.n_slots;
  if (k.plan.has_exp) return launch_one<MODE, VEC, 8, true, true>(k, st);
  if (iso) {
    if (ns == 0) return launch_one<MODE, VEC, 0, true, false>(k, st);
    if (ns <= 2) return launch_one<MODE, VEC, 2, true, false>(k, st);
    if (ns <= 4) return launch_one<MODE, VEC, 4, true, false>(k, st);
    return launch_one<MODE, VEC, 8, true, false>(k, st);
  }
  if (ns == 0) return launch_one<MODE, VEC, 0, false, false>(k, st);
  if (ns <= 2) return launch_one<MODE, VEC, 2, false, false>(k, st);
  if (ns <= 4) return launch_one<MODE, VEC, 4, false, false>(k, st);
  return launch_one<MODE, VEC, 8, false, false>(k, st);
}

static bool needs_iso(const AggPlan& P) {
  for (int a = 0; a < P.A; ++a) {
    const int kd = P.agg_kind[a];
    if (kd == DGN_AGG_MAX || kd == DGN_AGG_MIN || kd == DGN_AGG_STD || kd == DGN_AGG_VAR) return true;
  }
  return false;
}

template <int VEC>
static int dispatch_mode(const KernelArgs& k, bool iso, cudaStream_t st) {
  if (k.mode == DGN_MSG_SOURCE) return dispatch_slots<DGN_MSG_SOURCE, VEC>(k, iso, st);
  if (k.mode == DGN_MSG_AFFINE) return dispatch_slots<DGN_MSG_AFFINE, VEC>(k, iso, st);
  return dispatch_slots<DGN_MSG_DENSE, VEC>(k, iso, st);
}

int launch_forward(const KernelArgs& k, int vec, cudaStream_t st) {
  const bool iso = needs_iso(k.plan);
  if (vec == 4) return dispatch_mode<4>(k, iso, st);
  if (vec == 2) return dispatch_mode<2>(k, iso, st);
  return dispatch_mode<1>(k, iso, st);
}

int choose_vec(int max_vec, long long n_nodes, int n_feat, bool narrow_small) {
  static const int forced = [] { const char* e = getenv("DGN_FORCE_VEC"); return e ? atoi(e) : 0; }();
  if (forced == 1 || forced == 2 || forced == 4) return forced <= max_vec ? forced : max_vec;
  // Measured on B200 (profiles/README.md, cfg2 N*F = 190 k): both directions are faster with 8 B lanes when
  // 16 B lanes would leave most SMs with a single block (fwd 10.9 -> 9.5 us, bwd 19.4 -> 17.0 us); larger launches
  // always prefer the widest lanes.
  int vec = max_vec;
  if (narrow_small && vec == 4 && n_nodes * n_feat / 4 < 148LL * 512) vec = 2;
  return vec;
}

}  // namespace dgn
