// Internal: launch plan derived from DgnAggSpec, and small device helpers shared by the
// forward and backward aggregation kernels.  Not part of the C ABI.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dgn_b200.h"

namespace dgn {

// weight w(delta) multiplying the message inside one eigen-weighted feature sum ("slot")
enum WKind : int {
  W_ABS = 0,  // |delta|            dir-av
  W_SGN = 1,  // delta              dir-dx, dir-dx-no-abs
  W_POS = 2,  // relu(delta)        dir-dx-balanced, forward half
  W_NEG = 3,  // relu(-delta)       dir-dx-balanced, backward half
  W_EXP = 4   // exp(alpha*|delta| - max)   dir softmax
};

struct AggPlan {
  int F, Fg, K, A, S;        // S = scalers actually applied (1 when the reference skips them)
  int chunks;                // ceil(F / VEC)
  int n_slots;
  int has_exp;
  int slot_eig[DGN_MAX_SLOTS];
  int slot_w[DGN_MAX_SLOTS];
  float slot_alpha[DGN_MAX_SLOTS];
  uint8_t agg_kind[DGN_MAX_AGG];
  unsigned slot_aggs[DGN_MAX_SLOTS];  // bit a set: aggregator a reads slot s (balanced: its W_POS slot; W_NEG is s+1)
  uint8_t scaler_kind[DGN_MAX_SCALERS];
  float avg_log;
};

struct KernelArgs {
  AggPlan plan;
  int N, E;
  int mode;                  // DgnMsgMode
  const int32_t* in_ptr;
  const int32_t* in_src;
  const int32_t* in_eid;
  const int32_t* out_ptr;
  const int32_t* out_slot;
  const float* log_deg;
  // forward operands
  const float* x; int ld_x;
  const float* q; int ld_q;
  const float* q_bias;
  const float* r; int ld_r;
  const float* h_in; int ld_h;
  const float* eig; int ld_eig;
  float* out; int ld_out; int out_gs;
  float* h_copy; int ld_hc; int hc_gs;
  // backward operands
  const float* g_out;
  const float* g_hcopy;
  float* d_q; int ld_dq;
  float* d_r; int ld_dr;
  float* d_h; int ld_dh;
  const float* d_h_add; int ld_dha;
  float* edge_ws;
  const int32_t* graph_ptr; int n_graphs; int max_graph_nodes;   // optional graph boundaries (tile kernels)
};

// ---- VEC-wide register vector -------------------------------------------------------------------
template <int VEC> struct Vec { float a[VEC]; };

template <int VEC> __device__ __forceinline__ Vec<VEC> vfill(float s) {
  Vec<VEC> r;
#pragma unroll
  for (int i = 0; i < VEC; ++i) r.a[i] = s;
  return r;
}

// read-only (non-coherent) load; 128-bit when VEC == 4 (callers guarantee 16 B alignment)
template <int VEC> __device__ __forceinline__ Vec<VEC> vload(const float* __restrict__ p) {
  Vec<VEC> r;
  if constexpr (VEC == 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    r.a[0] = t.x; r.a[1] = t.y; r.a[2] = t.z; r.a[3] = t.w;
  } else if constexpr (VEC == 2) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(p));
    r.a[0] = t.x; r.a[1] = t.y;
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) r.a[i] = __ldg(p + i);
  }
  return r;
}

// streaming load for data read exactly once by the grid (gradient of the wide output)
template <int VEC> __device__ __forceinline__ Vec<VEC> vload_stream(const float* __restrict__ p) {
  Vec<VEC> r;
  if constexpr (VEC == 4) {
    const float4 t = __ldcs(reinterpret_cast<const float4*>(p));
    r.a[0] = t.x; r.a[1] = t.y; r.a[2] = t.z; r.a[3] = t.w;
  } else if constexpr (VEC == 2) {
    const float2 t = __ldcs(reinterpret_cast<const float2*>(p));
    r.a[0] = t.x; r.a[1] = t.y;
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) r.a[i] = __ldcs(p + i);
  }
  return r;
}

template <int VEC> __device__ __forceinline__ void vstore(float* __restrict__ p, const Vec<VEC>& v) {
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(v.a[0], v.a[1], v.a[2], v.a[3]);
  } else if constexpr (VEC == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(v.a[0], v.a[1]);
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) p[i] = v.a[i];
  }
}

// streaming store (evict-first): the wide [N, S*A*F] output is not re-read by this kernel
template <int VEC> __device__ __forceinline__ void vstore_stream(float* __restrict__ p, const Vec<VEC>& v) {
  if constexpr (VEC == 4) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(v.a[0], v.a[1], v.a[2], v.a[3]));
  } else if constexpr (VEC == 2) {
    __stcs(reinterpret_cast<float2*>(p), make_float2(v.a[0], v.a[1]));
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) __stcs(p + i, v.a[i]);
  }
}

__device__ __forceinline__ float edge_weight(int kind, float d, float alpha, float shift) {
  const float a = fabsf(d);
  float w = a;                                   // W_ABS
  w = (kind == W_SGN) ? d : w;
  w = (kind == W_POS) ? fmaxf(d, 0.f) : w;
  w = (kind == W_NEG) ? fmaxf(-d, 0.f) : w;
  if (kind == W_EXP) w = expf(alpha * a - shift);
  return w;
}

__device__ __forceinline__ float sign0(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }

// Per-(node, column chunk) accumulators of one pass over the in-edges.
template <int VEC, int NS, bool ISO>
struct RowAcc {
  Vec<VEC> sum;
  Vec<VEC> sq, mx, mn;       // only maintained when ISO
  Vec<VEC> acc[NS > 0 ? NS : 1];
  float zw[NS > 0 ? NS : 1];   // sum_u w(delta_u)
  float zabs[NS > 0 ? NS : 1]; // sum_u |delta_u|  (the L1 normaliser of av / dx)
};

// message of in-edge slot e (source u) for columns [c, c+VEC)
template <int MODE, int VEC>
__device__ __forceinline__ Vec<VEC> load_message(const KernelArgs& k, int u, int e, int c, const Vec<VEC>& qv) {
  Vec<VEC> m;
  if constexpr (MODE == DGN_MSG_DENSE) {
    const int id = k.in_eid ? __ldg(k.in_eid + e) : e;
    m = vload<VEC>(k.r + (size_t)id * k.ld_r + c);
  } else {
    m = vload<VEC>(k.x + (size_t)u * k.ld_x + c);
    if constexpr (MODE == DGN_MSG_AFFINE) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) m.a[i] += qv.a[i];
      if (k.r) {
        const int id = k.in_eid ? __ldg(k.in_eid + e) : e;
        const Vec<VEC> rv = vload<VEC>(k.r + (size_t)id * k.ld_r + c);
#pragma unroll
        for (int i = 0; i < VEC; ++i) m.a[i] += rv.a[i];
      }
    }
  }
  return m;
}

// One sequential pass over the in-edges [e0, e1) of node v (edge-id order, like the mailbox).
template <int MODE, int VEC, int NS, bool ISO, bool EXP>
__device__ __forceinline__ void accumulate_row(const KernelArgs& k, int v, int c, int e0, int e1,
                                               const Vec<VEC>& qv, const float (&ev)[NS > 0 ? NS : 1],
                                               float (&shift)[NS > 0 ? NS : 1], RowAcc<VEC, NS, ISO>& R) {
  const AggPlan& P = k.plan;
  R.sum = vfill<VEC>(0.f);
  if constexpr (ISO) {
    R.sq = vfill<VEC>(0.f);
    R.mx = vfill<VEC>(-INFINITY);
    R.mn = vfill<VEC>(INFINITY);
  }
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    R.acc[s] = vfill<VEC>(0.f);
    R.zw[s] = 0.f;
    R.zabs[s] = 0.f;
    shift[s] = 0.f;
  }
  if constexpr (EXP) {
    // softmax max-shift pre-pass (scalar work only): max_u alpha*|delta_u| per W_EXP slot
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      if (s < P.n_slots && P.slot_w[s] == W_EXP) {
        float m = -INFINITY;
        for (int e = e0; e < e1; ++e) {
          const int u = __ldg(k.in_src + e);
          const float d = __ldg(k.eig + (size_t)u * k.ld_eig + P.slot_eig[s]) - ev[s];
          m = fmaxf(m, P.slot_alpha[s] * fabsf(d));
        }
        shift[s] = m;
      }
    }
  }
#pragma unroll 4
  for (int e = e0; e < e1; ++e) {
    const int u = __ldg(k.in_src + e);
    const Vec<VEC> m = load_message<MODE, VEC>(k, u, e, c, qv);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      R.sum.a[i] += m.a[i];
      if constexpr (ISO) {
        // separate multiply and add (no FMA): the reference squares, rounds, then sums
        R.sq.a[i] = __fadd_rn(R.sq.a[i], __fmul_rn(m.a[i], m.a[i]));
        R.mx.a[i] = fmaxf(R.mx.a[i], m.a[i]);
        R.mn.a[i] = fminf(R.mn.a[i], m.a[i]);
      }
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      if (s < P.n_slots) {
        const float d = __ldg(k.eig + (size_t)u * k.ld_eig + P.slot_eig[s]) - ev[s];
        const float w = edge_weight(P.slot_w[s], d, P.slot_alpha[s], shift[s]);
        R.zw[s] += w;
        R.zabs[s] += fabsf(d);
#pragma unroll
        for (int i = 0; i < VEC; ++i) R.acc[s].a[i] = fmaf(w, m.a[i], R.acc[s].a[i]);
      }
    }
  }
}

// scaler coefficients of a node (rb/nets/scalers.py): 1, log(D+1)/avg, avg/log(D+1)
__device__ __forceinline__ void scaler_coefs(const KernelArgs& k, int v, float (&coef)[DGN_MAX_SCALERS]) {
  const AggPlan& P = k.plan;
  const float ld = (P.S > 1) ? __ldg(k.log_deg + v) : 1.f;
#pragma unroll
  for (int s = 0; s < DGN_MAX_SCALERS; ++s) {
    float cf = 1.f;
    if (s < P.S && P.S > 1) {
      const int kind = P.scaler_kind[s];
      cf = (kind == DGN_SCALE_AMPLIFICATION) ? __fdiv_rn(ld, P.avg_log)
           : (kind == DGN_SCALE_ATTENUATION) ? __fdiv_rn(P.avg_log, ld) : 1.f;
    }
    coef[s] = cf;
  }
}

// host side: kernels.cu
// vec = columns per thread (4, 2 or 1): 4 needs 16 B alignment of every row/slab start, 2 needs 8 B
int launch_forward(const KernelArgs& k, int vec, cudaStream_t st);
int launch_backward(const KernelArgs& k, int vec, float* d_x, int ld_dx, const float* addend, int ld_add,
                    cudaStream_t st);

int launch_backward_src(const KernelArgs& k, int vec, float* d_x, int ld_dx, const float* addend, int ld_add,
                        cudaStream_t st);

// Row kernels over a precomputed eigen-field (dgn_agg_row.cu): the default path when DgnAggIO.field is given.
int launch_forward_row(const KernelArgs& k, const DgnAggSpec* spec, const DgnField* f, int vec, cudaStream_t st);
int launch_backward_row_dst(const KernelArgs& k, const DgnAggSpec* spec, const DgnField* f, int vec, cudaStream_t st);

// Picks the vector width: the widest the operands' alignment allows; `narrow_small` lets small launches
// (that would not fill the 148 SMs) use 8 B lanes for more, shorter threads.  DGN_FORCE_VEC overrides.
int choose_vec(int max_vec, long long n_nodes, int n_feat, bool narrow_small);

}  // namespace dgn
