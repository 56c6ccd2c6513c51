// Fused DGN aggregation over a precomputed eigen-field (sm_100a): dgn_field_build + the "row" kernels.
//
// The eigenvector weights of an edge, w_s(u->v) (rb/nets/aggregators.py:35-71), depend on the graph and on
// ndata['eig'] only - not on the layer, not on the feature column.  A DGN net calls its L layers (forward and
// backward: 2L launches) on the SAME batch, so the weights are computed once per batch by dgn_field_build,
// already divided by their per-node normaliser, and laid out for 128-bit loads:
//
//   group g = 4 consecutive in-edge slots of one destination node:  [ int32 src[4] | float w_0[4] | ... | w_{ns-1}[4] ]
//   node v owns group v (its first 4 in-edges; src = -1 pads) - no pointer chase for molecules (degree <= 4) -
//   and the overflow groups N + ovf_ptr[v] .. N + ovf_ptr[v+1] (degree > 4: superpixel kNN, SBM PATTERN).
//   wsum[v][s] = sum_u w_s(u->v), the factor of h_in in the dx aggregators (row stride 4 * ceil(ns / 4)).
//
// With that, a (node, column chunk) thread's dependent chain is group -> message rows -> stores: no in_ptr ->
// in_src -> eig[u] chain, no per-launch abs / relu / exp / division, no shared memory, no block barrier.  The
// kernels keep everything in registers, issue the 4 row gathers of a group back to back and write / read the
// wide [N, S*A*F] operand with streaming 128-bit accesses.  The aggregator list is compiled into a short
// op table (isotropic first, then slot by slot) so that every accumulator keeps a static register name.
#include <stdlib.h>

#include "dgn_launch.cuh"
#include "dgn_plan.cuh"

extern thread_local cudaError_t g_dgn_last_cuda;

#ifndef ROW_MINB_FWD
#define ROW_MINB_FWD 8
#endif
#ifndef ROW_MINB_BWD
#define ROW_MINB_BWD 6
#endif
#ifndef ROW_THREADS
#define ROW_THREADS 128
#endif

namespace dgn {

// ------------------------------------------------------------------------------------------------------
// plans
// ------------------------------------------------------------------------------------------------------
enum FieldKind : int {
  FW_ABS = 0,   // |d| / (sum |d| + eps)                         dir-av
  FW_SGN = 1,   // d / (sum |d| + eps)                           dir-dx, dir-dx-no-abs
  FW_BAL = 2,   // (relu(d)/(sum relu(d)+eps) + relu(-d)/(sum relu(-d)+eps)) / 2     dir-dx-balanced
  FW_SOFT = 3   // softmax_u(alpha |d|)                          dirK-0.1 / dirK-neg-0.1
};

enum RowOp : int { OP_WSUM = 0, OP_DX = 1, OP_DX_ABS = 2 };   // directional: sum w m ; sum w m - W h ; |sum w m - W h|

struct FieldPlan {
  int n_slots;
  int eig[DGN_MAX_SLOTS];
  int kind[DGN_MAX_SLOTS];
  float alpha[DGN_MAX_SLOTS];
};

struct RowPlan {
  int F, Fg, A, S, n_slots;
  // forward: static op layout - the output block (aggregator index) of every isotropic kind and of every op of
  // every slot, -1 = absent.  A repeated directional aggregator gets a slot of its own; repeated isotropic names
  // (never used by the reference configs) go to the `extra` list.
  int16_t iso_pos[6];                 // indexed by DgnAggKind MEAN .. VAR
  int16_t slot_pos[DGN_MAX_SLOTS][3]; // indexed by RowOp
  int n_extra;
  uint8_t extra_kind[DGN_MAX_AGG], extra_pos[DGN_MAX_AGG];
  // backward: the same aggregators as a table walked with a software-pipelined loop
  int n_iso;                          // order[0, n_iso): isotropic aggregators
  int slot_end[DGN_MAX_SLOTS];        // order[slot_end[s-1], slot_end[s]) read slot s
  uint8_t order[DGN_MAX_AGG];         // aggregator index (output block) of every position
  uint8_t op[DGN_MAX_AGG];            // by position: DgnAggKind (isotropic) or RowOp (directional)
  uint8_t scaler_kind[DGN_MAX_SCALERS];
  float avg_log;
};

struct RowArgs {
  RowPlan rp;
  int N, mode, gstride;               // gstride = 1 + n_slots (float4 per group)
  const int32_t* in_ptr;
  const int32_t* in_eid;
  const int32_t* ovf_ptr;
  const float4* groups;
  const float* wsum;
  const float* log_deg;
  const float* x; int ld_x;
  const float* q; int ld_q;
  const float* q_bias;
  const float* r; int ld_r;
  const float* h_in; int ld_h;
  float* out; int ld_out; int out_gs;
  float* h_copy; int ld_hc; int hc_gs;
  const float* g_out;
  const float* g_hcopy;
  int pf_bytes;                       // > 0: bytes of a node's gradient row to prefetch into L2 (16 B multiple)
  float* d_q; int ld_dq;
  float* d_r; int ld_dr;
  float* d_h; int ld_dh;
  const float* d_h_add; int ld_dha;
  float* edge_ws;
  const int32_t* graph_ptr; int n_graphs; int tile_rows;   // tile kernels: node range per graph, rows of the smem tile
  int tile_split;                                          // CTAs per graph (each stages the block, takes 1 / split of the nodes)
  long long n_edges;
};

// Where a thread finds the source rows X[u] it gathers: global memory, or the shared-memory copy of its graph's node
// block (tile kernels: every in-edge of a node comes from the node's own graph - the batched adjacency is block diagonal).
struct RowSrc {
  const float* base;       // row u at base + (u - first) * ld
  int ld, first;
  bool smem;
};
template <int VEC> __device__ __forceinline__ Vec<VEC> row_load(const RowSrc& s, int u, int c) {
  if (s.smem) {
    const float* p = s.base + (u >= 0 ? u - s.first : 0) * s.ld + c;
    Vec<VEC> r;
    if constexpr (VEC == 4) { const float4 t = *reinterpret_cast<const float4*>(p); r.a[0] = t.x; r.a[1] = t.y; r.a[2] = t.z; r.a[3] = t.w; }
    else if constexpr (VEC == 2) { const float2 t = *reinterpret_cast<const float2*>(p); r.a[0] = t.x; r.a[1] = t.y; }
    else { r.a[0] = p[0]; }
    return r;
  }
  return vload<VEC>(s.base + (size_t)max(u, 0) * s.ld + c);
}

// slot with these weights whose op `op` is still free (a repeated aggregator opens a new slot with equal weights)
static int field_slot(FieldPlan& f, int eig, int kind, float alpha, int op, unsigned (&used)[DGN_MAX_SLOTS]) {
  for (int s = 0; s < f.n_slots; ++s)
    if (f.eig[s] == eig && f.kind[s] == kind && (kind != FW_SOFT || f.alpha[s] == alpha) && !(used[s] & (1u << op))) {
      used[s] |= 1u << op;
      return s;
    }
  if (f.n_slots >= DGN_MAX_SLOTS) return -1;
  used[f.n_slots] = 1u << op;
  const int s = f.n_slots++;
  f.eig[s] = eig; f.kind[s] = kind; f.alpha[s] = alpha;
  return s;
}

// spec -> slots (FieldPlan) and op table (RowPlan).  Both the field builder and the kernels derive the slot
// numbering from here, so a field built for one spec serves every spec with the same directional aggregators.
int make_row_plan(const DgnAggSpec* spec, FieldPlan& fp, RowPlan* rp) {
  memset(&fp, 0, sizeof(fp));
  if (spec->n_feat <= 0 || spec->n_agg <= 0 || spec->n_agg > DGN_MAX_AGG || spec->n_scalers <= 0 ||
      spec->n_scalers > DGN_MAX_SCALERS || spec->n_eig < 0)
    return DGN_ERR_INVALID;
  int slot_of[DGN_MAX_AGG], op_of[DGN_MAX_AGG];
  unsigned used[DGN_MAX_SLOTS] = {0};
  for (int a = 0; a < spec->n_agg; ++a) {
    const int kind = spec->agg_kind[a], eig = spec->agg_eig[a];
    slot_of[a] = -1; op_of[a] = kind;
    if (kind > DGN_AGG_DIR_SOFTMAX) return DGN_ERR_INVALID;
    if (kind < DGN_AGG_DIR_AV) continue;
    if (eig >= spec->n_eig) return DGN_ERR_INVALID;
    switch (kind) {
      case DGN_AGG_DIR_AV: op_of[a] = OP_WSUM; slot_of[a] = field_slot(fp, eig, FW_ABS, 0.f, OP_WSUM, used); break;
      case DGN_AGG_DIR_DX: op_of[a] = OP_DX_ABS; slot_of[a] = field_slot(fp, eig, FW_SGN, 0.f, OP_DX_ABS, used); break;
      case DGN_AGG_DIR_DX_NO_ABS: op_of[a] = OP_DX; slot_of[a] = field_slot(fp, eig, FW_SGN, 0.f, OP_DX, used); break;
      case DGN_AGG_DIR_DX_BALANCED: op_of[a] = OP_DX_ABS; slot_of[a] = field_slot(fp, eig, FW_BAL, 0.f, OP_DX_ABS, used); break;
      default: op_of[a] = OP_WSUM; slot_of[a] = field_slot(fp, eig, FW_SOFT, spec->agg_alpha[a], OP_WSUM, used); break;
    }
    if (slot_of[a] < 0) return DGN_ERR_UNSUPPORTED;
  }
  if (!rp) return DGN_OK;
  memset(rp, 0, sizeof(*rp));
  rp->F = spec->n_feat;
  rp->Fg = spec->group_feat > 0 ? spec->group_feat : spec->n_feat;
  if (rp->F % rp->Fg != 0) return DGN_ERR_INVALID;
  rp->A = spec->n_agg;
  rp->S = spec->n_scalers > 1 ? spec->n_scalers : 1;     // rb/nets/dgn_layer.py:95
  rp->avg_log = spec->avg_log;
  rp->n_slots = fp.n_slots;
  for (int s = 0; s < spec->n_scalers; ++s) {
    if (spec->scaler_kind[s] > DGN_SCALE_ATTENUATION) return DGN_ERR_INVALID;
    rp->scaler_kind[s] = spec->scaler_kind[s];
  }
  for (int i = 0; i < 6; ++i) rp->iso_pos[i] = -1;
  for (int t = 0; t < DGN_MAX_SLOTS; ++t) rp->slot_pos[t][0] = rp->slot_pos[t][1] = rp->slot_pos[t][2] = -1;
  for (int a = 0; a < rp->A; ++a) {
    if (slot_of[a] >= 0) rp->slot_pos[slot_of[a]][op_of[a]] = (int16_t)a;
    else if (rp->iso_pos[op_of[a]] < 0) rp->iso_pos[op_of[a]] = (int16_t)a;
    else { rp->extra_kind[rp->n_extra] = (uint8_t)op_of[a]; rp->extra_pos[rp->n_extra] = (uint8_t)a; ++rp->n_extra; }
  }
  int n = 0;
  for (int a = 0; a < rp->A; ++a)
    if (slot_of[a] < 0) { rp->order[n] = (uint8_t)a; rp->op[n] = (uint8_t)op_of[a]; ++n; }
  rp->n_iso = n;
  for (int s = 0; s < DGN_MAX_SLOTS; ++s) {
    for (int a = 0; a < rp->A; ++a)
      if (slot_of[a] == s) { rp->order[n] = (uint8_t)a; rp->op[n] = (uint8_t)op_of[a]; ++n; }
    rp->slot_end[s] = n;
  }
  return DGN_OK;
}

// ------------------------------------------------------------------------------------------------------
// dgn_field_build: thread = (node, slot)
// ------------------------------------------------------------------------------------------------------
struct FieldArgs {
  FieldPlan fp;
  int N;
  const int32_t* in_ptr;
  const int32_t* in_src;
  const int32_t* ovf_ptr;
  const float* eig; int ld_eig;
  float* groups;                       // viewed as floats: group g, row r (0 = sources), lane l -> (g*gstride + r)*4 + l
  float* wsum;
};

__global__ void __launch_bounds__(256) field_build_kernel(const __grid_constant__ FieldArgs k) {
  pdl_prologue();
  const int ns = k.fp.n_slots, lanes = ns > 0 ? ns : 1, gstride = 1 + ns;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int v = (int)(t / lanes), s = (int)(t - (long long)v * lanes);
  if (v >= k.N) return;
  const int e0 = __ldg(k.in_ptr + v), e1 = __ldg(k.in_ptr + v + 1), D = e1 - e0;
  const int ovf0 = __ldg(k.ovf_ptr + v);
  auto group_of = [&](int j) { return j < 4 ? v : k.N + ovf0 + ((j - 4) >> 2); };
  const int n_groups_v = D <= 4 ? 1 : 1 + ((D - 4 + 3) >> 2);
  if (s == 0) {                                   // padded source lists
    for (int j = 0; j < 4 * n_groups_v; ++j) {
      const int u = j < D ? __ldg(k.in_src + e0 + j) : -1;
      reinterpret_cast<int*>(k.groups)[((size_t)group_of(j) * gstride) * 4 + (j & 3)] = u;
    }
  }
  if (ns == 0) return;
  const int col = k.fp.eig[s], kind = k.fp.kind[s];
  const float alpha = k.fp.alpha[s];
  const float ev = __ldg(k.eig + (size_t)v * k.ld_eig + col);
  // pass 1: the per-node normalisers, summed in mailbox (edge-id) order
  float z0 = 0.f, z1 = 0.f, mx = -INFINITY;
  for (int e = e0; e < e1; ++e) {
    const float d = __ldg(k.eig + (size_t)__ldg(k.in_src + e) * k.ld_eig + col) - ev;
    if (kind == FW_BAL) { z0 += fmaxf(d, 0.f); z1 += fmaxf(-d, 0.f); }
    else if (kind == FW_SOFT) mx = fmaxf(mx, alpha * fabsf(d));
    else z0 += fabsf(d);
  }
  if (kind == FW_SOFT) {
    for (int e = e0; e < e1; ++e) {
      const float d = __ldg(k.eig + (size_t)__ldg(k.in_src + e) * k.ld_eig + col) - ev;
      z0 += expf(alpha * fabsf(d) - mx);
    }
  } else {
    z0 += DGN_EPS; z1 += DGN_EPS;
  }
  // pass 2: normalised weights into the group lanes (IEEE division, like the reference's tensor division)
  float ws = 0.f;
  for (int j = 0; j < 4 * n_groups_v; ++j) {
    float w = 0.f;
    if (j < D) {
      const float d = __ldg(k.eig + (size_t)__ldg(k.in_src + e0 + j) * k.ld_eig + col) - ev;
      switch (kind) {
        case FW_ABS: w = __fdiv_rn(fabsf(d), z0); break;
        case FW_SGN: w = __fdiv_rn(d, z0); break;
        case FW_BAL: w = __fdiv_rn(__fadd_rn(__fdiv_rn(fmaxf(d, 0.f), z0), __fdiv_rn(fmaxf(-d, 0.f), z1)), 2.f); break;
        default: w = __fdiv_rn(expf(alpha * fabsf(d) - mx), z0); break;
      }
      ws += w;
    }
    k.groups[((size_t)group_of(j) * gstride + 1 + s) * 4 + (j & 3)] = w;
  }
  k.wsum[(size_t)v * ((ns + 3) & ~3) + s] = ws;
}

// ------------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------------
template <int VEC, int NS, bool ISO>
struct Acc {
  Vec<VEC> sum;
  Vec<VEC> sq, mx, mn;                 // only maintained when ISO
  Vec<VEC> w[NS > 0 ? NS : 1];
};

template <int VEC, int NS, bool ISO>
__device__ __forceinline__ void acc_clear(Acc<VEC, NS, ISO>& R) {
  R.sum = vfill<VEC>(0.f);
  R.sq = vfill<VEC>(0.f);
  R.mx = vfill<VEC>(-INFINITY);
  R.mn = vfill<VEC>(INFINITY);
#pragma unroll
  for (int s = 0; s < (NS > 0 ? NS : 1); ++s) R.w[s] = vfill<VEC>(0.f);
}

__device__ __forceinline__ float lane_of(const float4& f, int j) { return j == 0 ? f.x : j == 1 ? f.y : j == 2 ? f.z : f.w; }
__device__ __forceinline__ int lane_of(const int4& f, int j) { return j == 0 ? f.x : j == 1 ? f.y : j == 2 ? f.z : f.w; }

// x / D for a small positive integer D given rD = RN(1/D): one Newton step on the residual gives the correctly
// rounded quotient (Markstein), i.e. the value of the reference's true division, in 3 instructions.
__device__ __forceinline__ float div_by(float x, float fD, float rD) {
  const float q = x * rD;
  return fmaf(fmaf(-q, fD, x), rD, q);
}

// Walks the in-edge groups of node v: the node-aligned first group, then its overflow groups.  fn(jg, m, wts) is
// called for every real in-edge in mailbox order with its message (this thread's columns) and slot weights.
// The message mode is a launch-uniform runtime branch: SOURCE is AFFINE with q = 0 (qv stays zero), DENSE reads R[eid].
// WIDE: all 4 gathers of a group in flight at once (latency-optimised variant for single-wave launches).
template <int VEC, int NS, bool WIDE, typename Fn>
__device__ __forceinline__ void walk_groups(const RowArgs& k, const RowSrc& xs, int v, int c, const Vec<VEC>& qv, int e0,
                                            int ovf0, int ovf1, Fn&& fn) {
  constexpr int NSA = NS > 0 ? NS : 1;
  const bool dense = k.mode == DGN_MSG_DENSE;
  const bool edge_term = !dense && k.r != nullptr;
  int g = v, o = ovf0, jbase = 0;
  while (true) {
    const float4* gp = k.groups + (size_t)g * k.gstride;
    const int4 su4 = __ldg(reinterpret_cast<const int4*>(gp));
    if constexpr (WIDE) {
      float4 wv[NSA];
#pragma unroll
      for (int s = 0; s < NS; ++s) wv[s] = (s < k.rp.n_slots) ? __ldg(gp + 1 + s) : make_float4(0.f, 0.f, 0.f, 0.f);
      Vec<VEC> m[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int u = lane_of(su4, j);
        if (dense) {
          int id = e0 + jbase + j;
          if (k.in_eid && u >= 0) id = __ldg(k.in_eid + id);
          m[j] = vload<VEC>(k.r + (size_t)(u >= 0 ? id : 0) * k.ld_r + c);
        } else {
          m[j] = row_load<VEC>(xs, u, c);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (lane_of(su4, j) >= 0) {
          Vec<VEC> mm = m[j];
#pragma unroll
          for (int i = 0; i < VEC; ++i) mm.a[i] += qv.a[i];
          if (edge_term) {
            const int e = e0 + jbase + j;
            const int id = k.in_eid ? __ldg(k.in_eid + e) : e;
            const Vec<VEC> rv = vload<VEC>(k.r + (size_t)id * k.ld_r + c);
#pragma unroll
            for (int i = 0; i < VEC; ++i) mm.a[i] += rv.a[i];
          }
          float wj[NSA];
#pragma unroll
          for (int s = 0; s < NSA; ++s) wj[s] = lane_of(wv[s], j);
          fn(jbase + j, mm, wj);
        }
      }
    } else {
    // the group is consumed as two pairs of in-edge slots: half the staging registers of a 4-wide pass (occupancy is
    // what hides the gather latency), and no gathers at all for the padding pair of a degree <= 2 node
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int u0 = h == 0 ? su4.x : su4.z, u1 = h == 0 ? su4.y : su4.w;
      if (u0 < 0) break;
      float2 wv[NSA];
#pragma unroll
      for (int s = 0; s < NS; ++s)
        wv[s] = (s < k.rp.n_slots) ? __ldg(reinterpret_cast<const float2*>(gp + 1 + s) + h) : make_float2(0.f, 0.f);
      Vec<VEC> m[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {               // both gathers of the pair are issued back to back
        const int u = j == 0 ? u0 : u1;
        if (dense) {
          int id = e0 + jbase + 2 * h + j;
          if (k.in_eid && u >= 0) id = __ldg(k.in_eid + id);
          m[j] = vload<VEC>(k.r + (size_t)(u >= 0 ? id : 0) * k.ld_r + c);
        } else {
          m[j] = row_load<VEC>(xs, u, c);
        }
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if ((j == 0 ? u0 : u1) >= 0) {
          Vec<VEC> mm = m[j];
#pragma unroll
          for (int i = 0; i < VEC; ++i) mm.a[i] += qv.a[i];
          if (edge_term) {
            const int e = e0 + jbase + 2 * h + j;
            const int id = k.in_eid ? __ldg(k.in_eid + e) : e;
            const Vec<VEC> rv = vload<VEC>(k.r + (size_t)id * k.ld_r + c);
#pragma unroll
            for (int i = 0; i < VEC; ++i) mm.a[i] += rv.a[i];
          }
          float wj[NSA];
#pragma unroll
          for (int s = 0; s < NSA; ++s) wj[s] = j == 0 ? wv[s].x : wv[s].y;
          fn(jbase + 2 * h + j, mm, wj);
        }
      }
    }
    }
    if (o >= ovf1) break;
    g = k.N + o;
    ++o;
    jbase += 4;
  }
}

// Tile-kernel walk: all 4 gathers of a group at once (they are shared-memory reads) and the NEXT group's sources and
// weights requested before the current group is consumed - at D ~ 51 a node walks 13 groups, and a dependent
// group-load -> gather chain per group (one L2 round trip each) was what bounded the row kernels there.
template <int VEC, int NS, typename Fn>
__device__ __forceinline__ void walk_groups_pf(const RowArgs& k, const RowSrc& xs, int v, int c, const Vec<VEC>& qv, int e0,
                                               int ovf0, int ovf1, Fn&& fn) {
  constexpr int NSA = NS > 0 ? NS : 1;
  const bool edge_term = k.r != nullptr;
  int o = ovf0, jbase = 0;
  const float4* gp = k.groups + (size_t)v * k.gstride;
  int4 su4 = __ldg(reinterpret_cast<const int4*>(gp));
  float4 wv[NSA];
#pragma unroll
  for (int s = 0; s < NS; ++s) wv[s] = (s < k.rp.n_slots) ? __ldg(gp + 1 + s) : make_float4(0.f, 0.f, 0.f, 0.f);
  while (true) {
    const bool more = o < ovf1;
    int4 su4n = su4;
    float4 wvn[NSA];
#pragma unroll
    for (int s = 0; s < NSA; ++s) wvn[s] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (more) {
      const float4* gn = k.groups + (size_t)(k.N + o) * k.gstride;
      su4n = __ldg(reinterpret_cast<const int4*>(gn));
#pragma unroll
      for (int s = 0; s < NS; ++s) if (s < k.rp.n_slots) wvn[s] = __ldg(gn + 1 + s);
    }
    Vec<VEC> m[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) m[j] = row_load<VEC>(xs, lane_of(su4, j), c);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (lane_of(su4, j) >= 0) {
        Vec<VEC> mm = m[j];
#pragma unroll
        for (int i = 0; i < VEC; ++i) mm.a[i] += qv.a[i];
        if (edge_term) {
          const int e = e0 + jbase + j;
          const int id = k.in_eid ? __ldg(k.in_eid + e) : e;
          const Vec<VEC> rv = vload<VEC>(k.r + (size_t)id * k.ld_r + c);
#pragma unroll
          for (int i = 0; i < VEC; ++i) mm.a[i] += rv.a[i];
        }
        float wj[NSA];
#pragma unroll
        for (int s = 0; s < NSA; ++s) wj[s] = lane_of(wv[s], j);
        fn(jbase + j, mm, wj);
      }
    }
    if (!more) break;
    su4 = su4n;
#pragma unroll
    for (int s = 0; s < NSA; ++s) wv[s] = wvn[s];
    ++o;
    jbase += 4;
  }
}

// wsum row of node v: one 128-bit load per 4 slots
template <int NS>
__device__ __forceinline__ void load_wsum(const RowArgs& k, int v, float (&w)[NS > 0 ? NS : 1]) {
  if constexpr (NS == 0) {
    w[0] = 0.f;
  } else if constexpr (NS == 2) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(k.wsum + (size_t)v * 4));
    w[0] = t.x; w[1] = t.y;
  } else {
    const int nsp = (k.rp.n_slots + 3) & ~3;
#pragma unroll
    for (int q = 0; q < NS / 4; ++q) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (4 * q < nsp) t = __ldg(reinterpret_cast<const float4*>(k.wsum + (size_t)v * nsp) + q);
      w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
    }
  }
}

template <int VEC, int NS, bool ISO>
__device__ __forceinline__ void acc_add(Acc<VEC, NS, ISO>& R, const Vec<VEC>& m, const float* wj) {
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    R.sum.a[i] += m.a[i];
    if constexpr (ISO) {
      R.sq.a[i] = __fadd_rn(R.sq.a[i], __fmul_rn(m.a[i], m.a[i]));     // square, round, then sum (as the reference)
      R.mx.a[i] = fmaxf(R.mx.a[i], m.a[i]);
      R.mn.a[i] = fminf(R.mn.a[i], m.a[i]);
    }
  }
#pragma unroll
  for (int s = 0; s < NS; ++s) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) R.w[s].a[i] = fmaf(wj[s], m.a[i], R.w[s].a[i]);
  }
}

__device__ __forceinline__ void row_scaler_coefs(const RowPlan& P, float ld, float (&coef)[DGN_MAX_SCALERS]) {
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

template <int VEC, bool ISO>
__device__ __forceinline__ void row_mean_var(const Vec<VEC>& sum, const Vec<VEC>& sq, float fD, float rD, Vec<VEC>& mean,
                                             Vec<VEC>& var) {
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    mean.a[i] = div_by(sum.a[i], fD, rD);
    if constexpr (ISO) {
      const float msq = div_by(sq.a[i], fD, rD);
      var.a[i] = fmaxf(__fsub_rn(msq, __fmul_rn(mean.a[i], mean.a[i])), 0.f);
    } else {
      var.a[i] = 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sqrt_approx(float x) {      // 1 ulp; the IEEE sqrtf costs ~10 instructions per column
  float y;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// SS = number of scalers known at compile time (0 = runtime P.S).  Straight-line code: every possible op is a
// launch-uniform "is it requested" test on the plan followed by S scaled 128-bit streaming stores; the S slab base
// pointers are formed once, every store address is base + (block * Fg) in one IMAD.WIDE.
template <int VEC, int NS, bool ISO, int SS>
__device__ __forceinline__ void fwd_epilogue(const RowPlan& P, const Acc<VEC, NS, ISO>& R, const Vec<VEC>& hv,
                                             const float* wsumv, int D, float ld, float* __restrict__ orow) {
  const int S = SS > 0 ? SS : P.S;
  float coef[DGN_MAX_SCALERS];
  row_scaler_coefs(P, ld, coef);
  // one opaque 64-bit row base + 32-bit element offsets: each store address is a single IMAD.WIDE (left to itself the
  // compiler re-derives every address from k.out with 64-bit index arithmetic, 4 instructions per store)
  asm volatile("" : "+l"(orow));
  const unsigned scaler_stride = (unsigned)(P.A * P.Fg);
  auto store_scaled = [&](int a, const Vec<VEC>& y) {
    const unsigned off = (unsigned)a * (unsigned)P.Fg;
#pragma unroll
    for (int s = 0; s < DGN_MAX_SCALERS; ++s) {
      if (s < S) {
        Vec<VEC> o;
#pragma unroll
        for (int i = 0; i < VEC; ++i) o.a[i] = y.a[i] * coef[s];
        vstore_stream<VEC>(orow + (off + (unsigned)s * scaler_stride), o);
      }
    }
  };
  const float fD = (float)D, rD = __frcp_rn(fD);
  Vec<VEC> mean, var;
  row_mean_var<VEC, ISO>(R.sum, R.sq, fD, rD, mean, var);
  if (P.iso_pos[DGN_AGG_MEAN] >= 0) store_scaled(P.iso_pos[DGN_AGG_MEAN], mean);
  if (P.iso_pos[DGN_AGG_SUM] >= 0) store_scaled(P.iso_pos[DGN_AGG_SUM], R.sum);
  if constexpr (ISO) {
    if (P.iso_pos[DGN_AGG_MAX] >= 0) store_scaled(P.iso_pos[DGN_AGG_MAX], R.mx);
    if (P.iso_pos[DGN_AGG_MIN] >= 0) store_scaled(P.iso_pos[DGN_AGG_MIN], R.mn);
    if (P.iso_pos[DGN_AGG_VAR] >= 0) store_scaled(P.iso_pos[DGN_AGG_VAR], var);
    if (P.iso_pos[DGN_AGG_STD] >= 0) {
      Vec<VEC> y;
#pragma unroll
      for (int i = 0; i < VEC; ++i) y.a[i] = sqrt_approx(var.a[i] + DGN_EPS);
      store_scaled(P.iso_pos[DGN_AGG_STD], y);
    }
  }
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    if (P.slot_pos[s][OP_WSUM] >= 0) store_scaled(P.slot_pos[s][OP_WSUM], R.w[s]);
    if (P.slot_pos[s][OP_DX] >= 0 || P.slot_pos[s][OP_DX_ABS] >= 0) {
      Vec<VEC> sv;
#pragma unroll
      for (int i = 0; i < VEC; ++i) sv.a[i] = R.w[s].a[i] - wsumv[s] * hv.a[i];
      if (P.slot_pos[s][OP_DX] >= 0) store_scaled(P.slot_pos[s][OP_DX], sv);
      if (P.slot_pos[s][OP_DX_ABS] >= 0) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) sv.a[i] = fabsf(sv.a[i]);
        store_scaled(P.slot_pos[s][OP_DX_ABS], sv);
      }
    }
  }
  for (int x = 0; x < P.n_extra; ++x) {                // repeated isotropic names
    Vec<VEC> y;
    switch (P.extra_kind[x]) {
      case DGN_AGG_MEAN: y = mean; break;
      case DGN_AGG_SUM: y = R.sum; break;
      case DGN_AGG_MAX: y = ISO ? R.mx : mean; break;
      case DGN_AGG_MIN: y = ISO ? R.mn : mean; break;
      case DGN_AGG_VAR: y = var; break;
      default:
#pragma unroll
        for (int i = 0; i < VEC; ++i) y.a[i] = sqrt_approx(var.a[i] + DGN_EPS);
    }
    store_scaled(P.extra_pos[x], y);
  }
}

// One (node, column chunk) work item of the forward.
// LAT = latency-optimised variant for launches that fit in one wave: wide gathers, no register cap.
template <int VEC, int NS, bool ISO, bool LAT, bool TILE = false>
__device__ __forceinline__ void fwd_node(const RowArgs& k, const RowSrc& xs, int v, int c) {
  constexpr int NSA = NS > 0 ? NS : 1;
  const RowPlan& P = k.rp;
  int tower = 0, cg = c;
  if (P.Fg != P.F) { tower = c / P.Fg; cg = c - tower * P.Fg; }

  // everything that does not depend on the in-edges is requested first
  const int ovf0 = __ldg(k.ovf_ptr + v), ovf1 = __ldg(k.ovf_ptr + v + 1);
  const int e0 = (k.r != nullptr) ? __ldg(k.in_ptr + v) : 0;      // edge ids are only needed for R[eid]
  const Vec<VEC> hv = vload<VEC>(k.h_in + (size_t)v * k.ld_h + c);
  Vec<VEC> qv = vfill<VEC>(0.f);
  if (k.q) {
    qv = vload<VEC>(k.q + (size_t)v * k.ld_q + c);
    if (k.q_bias) {
      const Vec<VEC> bv = vload<VEC>(k.q_bias + c);
#pragma unroll
      for (int i = 0; i < VEC; ++i) qv.a[i] += bv.a[i];
    }
  }
  const float ld = (P.S > 1) ? __ldg(k.log_deg + v) : 1.f;
  float wsumv[NSA];
  load_wsum<NS>(k, v, wsumv);
  if (k.h_copy) vstore<VEC>(k.h_copy + (size_t)v * k.ld_hc + (size_t)tower * k.hc_gs + cg, hv);

  Acc<VEC, NS, ISO> R;
  acc_clear(R);
  int D = 0;
  auto add = [&](int, const Vec<VEC>& m, const float* wj) {
    ++D;
    acc_add<VEC, NS, ISO>(R, m, wj);
  };
  if constexpr (TILE) walk_groups_pf<VEC, NS>(k, xs, v, c, qv, e0, ovf0, ovf1, add);
  else walk_groups<VEC, NS, LAT>(k, xs, v, c, qv, e0, ovf0, ovf1, add);

  float* orow = k.out + (size_t)v * k.ld_out + (size_t)tower * k.out_gs + cg;
  if (D == 0) {                                        // DGL: zero rows for isolated nodes
    const Vec<VEC> z = vfill<VEC>(0.f);
    for (int j = 0; j < P.S * P.A; ++j) vstore_stream<VEC>(orow + (size_t)j * P.Fg, z);
    return;
  }
  if (P.S == 3) fwd_epilogue<VEC, NS, ISO, 3>(P, R, hv, wsumv, D, ld, orow);
  else if (P.S == 1) fwd_epilogue<VEC, NS, ISO, 1>(P, R, hv, wsumv, D, ld, orow);
  else fwd_epilogue<VEC, NS, ISO, 0>(P, R, hv, wsumv, D, ld, orow);
}

template <int VEC, int NS, bool ISO, bool LAT>
__global__ void __launch_bounds__(ROW_THREADS, LAT ? 4 : ROW_MINB_FWD) agg_fwd_row_kernel(const __grid_constant__ RowArgs k) {
  pdl_prologue();
  const int v = blockIdx.x * blockDim.y + threadIdx.y;
  if (v >= k.N) return;
  const RowSrc xs{k.x, k.ld_x, 0, false};
  fwd_node<VEC, NS, ISO, LAT>(k, xs, v, threadIdx.x * VEC);
}

// ------------------------------------------------------------------------------------------------------
// tile kernels (high-degree graphs: superpixel kNN, SBM PATTERN): one CTA per graph.  A (node, chunk) thread of the row
// kernels gathers D source rows from L2 - at D ~ 51 that is E*F*4 bytes of L2 -> SM traffic per launch (2.8x the
// algorithmic bytes on PATTERN) and the launch is L2-gather bound.  The batched adjacency is block diagonal, so all
// sources of a graph's nodes are that graph's own node block: the CTA copies the block [n_g, F] (<= ~36 KB) into
// shared memory ONCE with coalesced 128-bit loads and every gather becomes a shared-memory read.
// ------------------------------------------------------------------------------------------------------
#ifndef TILE_THREADS
#define TILE_THREADS 256
#endif

__device__ __forceinline__ RowSrc stage_graph_rows(const RowArgs& k, float* tile, int v0, int v1) {
  const int n = v1 - v0, F = k.rp.F;
  if (k.mode == DGN_MSG_DENSE || n > k.tile_rows || (F & 3) || (k.ld_x & 3)) return RowSrc{k.x, k.ld_x, 0, false};
  const int q4 = F >> 2, tid = threadIdx.y * blockDim.x + threadIdx.x, nt = blockDim.x * blockDim.y;
  for (int i = tid; i < n * q4; i += nt) {
    const int r = i / q4, cc = (i - r * q4) * 4;
    *reinterpret_cast<float4*>(tile + r * F + cc) = __ldg(reinterpret_cast<const float4*>(k.x + (size_t)(v0 + r) * k.ld_x + cc));
  }
  __syncthreads();
  return RowSrc{tile, F, v0, true};
}

template <int VEC, int NS, bool ISO>
__global__ void __launch_bounds__(TILE_THREADS, 2) agg_fwd_tile_kernel(const __grid_constant__ RowArgs k) {
  pdl_prologue();
  extern __shared__ __align__(16) float tile[];
  // tile_split CTAs share a graph: each stages the graph's block and takes one slice of its nodes (enough CTAs to keep
  // the SMs' warp slots full; the redundant staging reads are N*F*4*split bytes of L2 traffic, still far below E*F*4).
  // The last CTA takes the padding rows behind the last graph (no in-edges: zero rows).
  const int g = blockIdx.x / k.tile_split, part = blockIdx.x - g * k.tile_split;
  const int v0 = g < k.n_graphs ? __ldg(k.graph_ptr + g) : __ldg(k.graph_ptr + k.n_graphs);
  const int v1 = g < k.n_graphs ? __ldg(k.graph_ptr + g + 1) : k.N;
  if (v1 <= v0 || (g >= k.n_graphs && part > 0)) return;
  const int per = g < k.n_graphs ? (v1 - v0 + k.tile_split - 1) / k.tile_split : v1 - v0;
  const int a0 = v0 + part * per, a1 = min(a0 + per, v1);
  if (a1 <= a0) return;
  const RowSrc xs = g < k.n_graphs ? stage_graph_rows(k, tile, v0, v1) : RowSrc{k.x, k.ld_x, 0, false};
  if (xs.smem) {
    for (int v = a0 + threadIdx.y; v < a1; v += blockDim.y) fwd_node<VEC, NS, ISO, false, true>(k, xs, v, threadIdx.x * VEC);
  } else {
    for (int v = a0 + threadIdx.y; v < a1; v += blockDim.y) fwd_node<VEC, NS, ISO, false>(k, xs, v, threadIdx.x * VEC);
  }
}

// ------------------------------------------------------------------------------------------------------
// backward, destination side (the source-side gather stays agg_bwd_src_kernel)
// ------------------------------------------------------------------------------------------------------
template <int VEC, int SS>
__device__ __forceinline__ void load_slabs(const float* __restrict__ grow, int a, int Fg, int scaler_stride, int S,
                                           Vec<VEC> (&g)[DGN_MAX_SCALERS]) {
  const unsigned off = (unsigned)a * (unsigned)Fg;
#pragma unroll
  for (int s = 0; s < DGN_MAX_SCALERS; ++s)
    if (s < (SS > 0 ? SS : S)) g[s] = vload_stream<VEC>(grow + (off + (unsigned)s * (unsigned)scaler_stride));
}

// PIPE: the slab loads of the next kPipeDepth aggregators are in flight while this one is folded.  Costs
// 12 registers per stage (occupancy), so it is used for launches that fit in one wave anyway, where the serial load
// latency (one DRAM round trip per aggregator otherwise) is what is left.
#ifndef ROW_PIPE_DEPTH
#define ROW_PIPE_DEPTH 3
#endif
constexpr int kPipeDepth = ROW_PIPE_DEPTH;

template <int VEC, int NS, bool ISO, int SS, bool PIPE>
__device__ __forceinline__ void bwd_fold(const RowPlan& P, const Acc<VEC, NS, ISO>& R, const Vec<VEC>& hv, const float* wsumv,
                                         int D, float ld, const float* __restrict__ grow, Vec<VEC>& c0, Vec<VEC>& c1,
                                         Vec<VEC>& gmx, Vec<VEC>& gmn, Vec<VEC> (&cs)[NS > 0 ? NS : 1], Vec<VEC>& dh) {
  const int S = SS > 0 ? SS : P.S;
  float coef[DGN_MAX_SCALERS];
  row_scaler_coefs(P, ld, coef);
  const float fD = (float)D, rD = __frcp_rn(fD);
  Vec<VEC> mean, var;
  row_mean_var<VEC, ISO>(R.sum, R.sq, fD, rD, mean, var);
  const int scaler_stride = P.A * P.Fg;
  // sign(sum w m - W h) of every (slot, column), packed: bit s*VEC+i of `pos_m` / `neg_m`.  Frees the NS
  // accumulators before the fold starts (registers decide the occupancy of this kernel).
  unsigned pos_m = 0u, neg_m = 0u;
#pragma unroll
  for (int s = 0; s < NS; ++s) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const float sv = R.w[s].a[i] - wsumv[s] * hv.a[i];        // same expression as the forward
      pos_m |= (sv > 0.f ? 1u : 0u) << (s * VEC + i);
      neg_m |= (sv < 0.f ? 1u : 0u) << (s * VEC + i);
    }
  }
  Vec<VEC> ring[PIPE ? kPipeDepth : 1][DGN_MAX_SCALERS];      // static indices only: stays in registers
  if constexpr (PIPE) {
#pragma unroll
    for (int d = 0; d < kPipeDepth; ++d)
      if (d < P.A) load_slabs<VEC, SS>(grow, P.order[d], P.Fg, scaler_stride, S, ring[d]);
  }
  auto next_G = [&](int pos) {
    Vec<VEC> gcur[DGN_MAX_SCALERS];
    if constexpr (PIPE) {
#pragma unroll
      for (int s = 0; s < DGN_MAX_SCALERS; ++s) gcur[s] = ring[0][s];
#pragma unroll
      for (int d = 0; d + 1 < kPipeDepth; ++d) {
#pragma unroll
        for (int s = 0; s < DGN_MAX_SCALERS; ++s) ring[d][s] = ring[d + 1][s];
      }
      if (pos + kPipeDepth < P.A)
        load_slabs<VEC, SS>(grow, P.order[pos + kPipeDepth], P.Fg, scaler_stride, S, ring[kPipeDepth - 1]);
    } else {
      load_slabs<VEC, SS>(grow, P.order[pos], P.Fg, scaler_stride, S, gcur);    // L2 hits: the row was prefetched
    }
    Vec<VEC> G = vfill<VEC>(0.f);
#pragma unroll
    for (int s = 0; s < DGN_MAX_SCALERS; ++s) {
      if (s < S) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) G.a[i] = fmaf(coef[s], gcur[s].a[i], G.a[i]);
      }
    }
    return G;
  };
  int pos = 0;
  for (; pos < P.n_iso; ++pos) {
    const int kind = P.op[pos];
    const Vec<VEC> G = next_G(pos);
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
  for (int s = 0; s < NS; ++s) {
    const int end = P.slot_end[s];
    for (; pos < end; ++pos) {
      const int op = P.op[pos];
      Vec<VEC> G = next_G(pos);
      if (op != OP_WSUM) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          if (op == OP_DX_ABS) {
            const unsigned b = 1u << (s * VEC + i);
            G.a[i] = (pos_m & b) ? G.a[i] : ((neg_m & b) ? -G.a[i] : 0.f);
          }
          dh.a[i] -= wsumv[s] * G.a[i];
        }
      }
#pragma unroll
      for (int i = 0; i < VEC; ++i) cs[s].a[i] += G.a[i];
    }
  }
}

template <int VEC, int NS, bool ISO, bool LAT, bool TILE = false>
__device__ __forceinline__ void bwd_node(const RowArgs& k, const RowSrc& xs, int v, int c, bool prefetch_lane) {
  constexpr int NSA = NS > 0 ? NS : 1;
  const RowPlan& P = k.rp;
  int tower = 0, cg = c;
  if (P.Fg != P.F) { tower = c / P.Fg; cg = c - tower * P.Fg; }

  // The node's gradient row (S*A slabs per tower, contiguous) dominates this kernel's traffic and its address is
  // known up front: one thread per node asks the L2 for it now, so the slab loads of the fold hit L2 instead of
  // paying one DRAM round trip per aggregator.
  if (k.pf_bytes > 0 && prefetch_lane) {
    const float* grow0 = k.g_out + (size_t)v * k.ld_out;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(grow0), "r"(k.pf_bytes) : "memory");
  }
  const int ovf0 = __ldg(k.ovf_ptr + v), ovf1 = __ldg(k.ovf_ptr + v + 1);
  const int e0 = __ldg(k.in_ptr + v);
  const Vec<VEC> hv = vload<VEC>(k.h_in + (size_t)v * k.ld_h + c);
  Vec<VEC> qv = vfill<VEC>(0.f);
  if (k.q) {
    qv = vload<VEC>(k.q + (size_t)v * k.ld_q + c);
    if (k.q_bias) {
      const Vec<VEC> bv = vload<VEC>(k.q_bias + c);
#pragma unroll
      for (int i = 0; i < VEC; ++i) qv.a[i] += bv.a[i];
    }
  }
  const float ld = (P.S > 1) ? __ldg(k.log_deg + v) : 1.f;
  float wsumv[NSA];
  load_wsum<NS>(k, v, wsumv);
  Vec<VEC> dh = vfill<VEC>(0.f);
  if (k.g_hcopy) dh = vload_stream<VEC>(k.g_hcopy + (size_t)v * k.ld_hc + (size_t)tower * k.hc_gs + cg);
  if (k.d_h_add) {
    const Vec<VEC> t = vload_stream<VEC>(k.d_h_add + (size_t)v * k.ld_dha + c);
#pragma unroll
    for (int i = 0; i < VEC; ++i) dh.a[i] += t.a[i];
  }

  // ---- pass 1: recompute the row statistics ------------------------------------------------------------
  Acc<VEC, NS, ISO> R;
  acc_clear(R);
  int D = 0;
  auto add = [&](int, const Vec<VEC>& m, const float* wj) {
    ++D;
    acc_add<VEC, NS, ISO>(R, m, wj);
  };
  if constexpr (TILE) walk_groups_pf<VEC, NS>(k, xs, v, c, qv, e0, ovf0, ovf1, add);
  else walk_groups<VEC, NS, LAT>(k, xs, v, c, qv, e0, ovf0, ovf1, add);

  Vec<VEC> dq = vfill<VEC>(0.f);
  if (D > 0) {
    // ---- fold the S*A gradient slabs into per-column coefficients -------------------------------------
    Vec<VEC> c0 = vfill<VEC>(0.f), c1 = vfill<VEC>(0.f), gmx = vfill<VEC>(0.f), gmn = vfill<VEC>(0.f);
    Vec<VEC> cs[NSA];
#pragma unroll
    for (int s = 0; s < NSA; ++s) cs[s] = vfill<VEC>(0.f);
    const float* grow = k.g_out + (size_t)v * k.ld_out + (size_t)tower * k.out_gs + cg;
    asm volatile("" : "+l"(grow));                    // opaque base: slab addresses are base + 32-bit offset
    if (P.S == 3) bwd_fold<VEC, NS, ISO, 3, LAT>(P, R, hv, wsumv, D, ld, grow, c0, c1, gmx, gmn, cs, dh);
    else if (P.S == 1) bwd_fold<VEC, NS, ISO, 1, LAT>(P, R, hv, wsumv, D, ld, grow, c0, c1, gmx, gmn, cs, dh);
    else bwd_fold<VEC, NS, ISO, 0, LAT>(P, R, hv, wsumv, D, ld, grow, c0, c1, gmx, gmn, cs, dh);

    // ---- pass 2: per-edge message gradients (rows and weights come back from L1 / L2) -------------------
    unsigned given = 0u;          // bit i: max gradient of column i already routed; bit VEC+i: min
    auto emit = [&](int jg, const Vec<VEC>& m, const float* wj) {
      Vec<VEC> dm;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        float g = fmaf(c1.a[i], m.a[i], c0.a[i]);
        if constexpr (ISO) {
          // torch.max / torch.min send the whole gradient to the FIRST extremal mailbox entry
          if (m.a[i] == R.mx.a[i] && !(given & (1u << i))) { g += gmx.a[i]; given |= 1u << i; }
          if (m.a[i] == R.mn.a[i] && !(given & (16u << i))) { g += gmn.a[i]; given |= 16u << i; }
        }
        dm.a[i] = g;
      }
#pragma unroll
      for (int s = 0; s < NS; ++s) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) dm.a[i] = fmaf(wj[s], cs[s].a[i], dm.a[i]);
      }
#pragma unroll
      for (int i = 0; i < VEC; ++i) dq.a[i] += dm.a[i];
      const int e = e0 + jg;
      if (k.edge_ws) vstore<VEC>(k.edge_ws + (size_t)e * P.F + c, dm);
      if (k.d_r) {
        const int id = k.in_eid ? __ldg(k.in_eid + e) : e;
        vstore<VEC>(k.d_r + (size_t)id * k.ld_dr + c, dm);
      }
    };
    if constexpr (TILE) walk_groups_pf<VEC, NS>(k, xs, v, c, qv, e0, ovf0, ovf1, emit);
    else walk_groups<VEC, NS, LAT>(k, xs, v, c, qv, e0, ovf0, ovf1, emit);
  }
  if (k.d_q) vstore<VEC>(k.d_q + (size_t)v * k.ld_dq + c, dq);
  if (k.d_h) vstore<VEC>(k.d_h + (size_t)v * k.ld_dh + c, dh);
}

template <int VEC, int NS, bool ISO, bool LAT>
__global__ void __launch_bounds__(ROW_THREADS, LAT ? 4 : ROW_MINB_BWD) agg_bwd_row_kernel(const __grid_constant__ RowArgs k) {
  pdl_prologue();
  const int v = blockIdx.x * blockDim.y + threadIdx.y;
  if (v >= k.N) return;
  const RowSrc xs{k.x, k.ld_x, 0, false};
  bwd_node<VEC, NS, ISO, LAT>(k, xs, v, threadIdx.x * VEC, threadIdx.x == 0);
}

template <int VEC, int NS, bool ISO>
__global__ void __launch_bounds__(TILE_THREADS, 2) agg_bwd_tile_kernel(const __grid_constant__ RowArgs k) {
  pdl_prologue();
  extern __shared__ __align__(16) float tile[];
  const int g = blockIdx.x / k.tile_split, part = blockIdx.x - g * k.tile_split;
  const int v0 = g < k.n_graphs ? __ldg(k.graph_ptr + g) : __ldg(k.graph_ptr + k.n_graphs);
  const int v1 = g < k.n_graphs ? __ldg(k.graph_ptr + g + 1) : k.N;
  if (v1 <= v0 || (g >= k.n_graphs && part > 0)) return;
  const int per = g < k.n_graphs ? (v1 - v0 + k.tile_split - 1) / k.tile_split : v1 - v0;
  const int a0 = v0 + part * per, a1 = min(a0 + per, v1);
  if (a1 <= a0) return;
  const RowSrc xs = g < k.n_graphs ? stage_graph_rows(k, tile, v0, v1) : RowSrc{k.x, k.ld_x, 0, false};
  if (xs.smem) {
    for (int v = a0 + threadIdx.y; v < a1; v += blockDim.y)
      bwd_node<VEC, NS, ISO, false, true>(k, xs, v, threadIdx.x * VEC, threadIdx.x == 0);
  } else {
    for (int v = a0 + threadIdx.y; v < a1; v += blockDim.y)
      bwd_node<VEC, NS, ISO, false>(k, xs, v, threadIdx.x * VEC, threadIdx.x == 0);
  }
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
static bool row_needs_iso(const RowPlan& P) {
  for (int i = 0; i < P.n_iso; ++i) {
    const int kd = P.op[i];
    if (kd == DGN_AGG_MAX || kd == DGN_AGG_MIN || kd == DGN_AGG_STD || kd == DGN_AGG_VAR) return true;
  }
  return false;
}

template <bool BWD, int VEC, int NS>
static int launch_row_iso(const RowArgs& k, cudaStream_t st) {
  if (k.N == 0) return DGN_OK;
  const int chunks = k.rp.F / VEC;
  if (chunks <= 0 || chunks > ROW_THREADS) return DGN_ERR_UNSUPPORTED;
  const dim3 block((unsigned)chunks, (unsigned)(ROW_THREADS / chunks));
  const unsigned grid = (unsigned)((k.N + block.y - 1) / block.y);
  const bool iso = row_needs_iso(k.rp);
  // A launch that fits in one wave at the lean kernels' occupancy is bound by the latency of one thread's dependent
  // chain, not by occupancy: it takes the LAT variants (wide gathers, pipelined fold).  DGN_ROW_LAT=0/1 overrides.
  static const int force_lat = [] { const char* e = getenv("DGN_ROW_LAT"); return e ? atoi(e) : -1; }();
  const bool lat = force_lat >= 0 ? force_lat != 0 : grid <= 148u * 6u;
  // Tile kernels (one CTA slice per graph, the graph's source rows staged in shared memory, group loads prefetched):
  // OPT-IN with DGN_TILE=1.  Measured on B200 (profiles/README.md, round 2): PATTERN b=256 forward 119 us vs 94 us for
  // the row kernels, backward 375 vs 287 us; CIFAR 22 vs 12 us - removing the L2 row gathers did NOT help, i.e. round
  // 1's diagnosis ("L2 -> SM gather bound") was wrong: at D ~ 51 the walk is bound by the ~40 instructions per
  // (edge, 4-column chunk) it issues (37 M warp instructions ~ 33 us at peak issue for PATTERN) and by divergence of the
  // 12-lane node groups inside a warp, which the row kernels' higher occupancy hides better.
  static const int force_tile = [] { const char* e = getenv("DGN_TILE"); return e ? atoi(e) : 0; }();
  const size_t tile_bytes = (size_t)k.tile_rows * k.rp.F * sizeof(float);
  const bool tile_ok = k.graph_ptr && k.n_graphs > 0 && k.tile_rows > 0 && k.mode != DGN_MSG_DENSE && VEC == 4 &&
                       tile_bytes <= 100 * 1024 && (k.ld_x % 4) == 0;
  const bool tile = tile_ok && force_tile != 0 && k.n_edges >= 6LL * k.N;
  if (tile) {
    const dim3 tb((unsigned)chunks, (unsigned)(TILE_THREADS / chunks));
    static const int passes = [] { const char* e = getenv("DGN_TILE_PASSES"); const int v = e ? atoi(e) : 1; return v < 1 ? 1 : v; }();
    RowArgs kt = k;
    kt.pf_bytes = 0;
    kt.tile_split = (k.tile_rows + (int)tb.y * passes - 1) / ((int)tb.y * passes);      // <= `passes` nodes per thread row
    const unsigned tg = (unsigned)(k.n_graphs + 1) * (unsigned)kt.tile_split;
    cudaError_t e = cudaSuccess;
    if constexpr (BWD) {
      auto kern = iso ? agg_bwd_tile_kernel<VEC, NS, true> : agg_bwd_tile_kernel<VEC, NS, false>;
      if (tile_bytes > 48 * 1024) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      if (e == cudaSuccess) launch_pdl(kern, dim3(tg), tb, tile_bytes, st, kt);
    } else {
      auto kern = iso ? agg_fwd_tile_kernel<VEC, NS, true> : agg_fwd_tile_kernel<VEC, NS, false>;
      if (tile_bytes > 48 * 1024) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      if (e == cudaSuccess) launch_pdl(kern, dim3(tg), tb, tile_bytes, st, kt);
    }
    return (e == cudaSuccess && cudaGetLastError() == cudaSuccess) ? DGN_OK : DGN_ERR_CUDA;
  }
  if constexpr (BWD) {
    if (lat) {
      RowArgs kl = k;
      kl.pf_bytes = 0;                 // single wave: the pipelined fold already overlaps the loads (prefetch measured slower)
      if (iso) launch_pdl(agg_bwd_row_kernel<VEC, NS, true, true>, grid, block, 0, st, kl);
      else launch_pdl(agg_bwd_row_kernel<VEC, NS, false, true>, grid, block, 0, st, kl);
    } else {
      if (iso) launch_pdl(agg_bwd_row_kernel<VEC, NS, true, false>, grid, block, 0, st, k);
      else launch_pdl(agg_bwd_row_kernel<VEC, NS, false, false>, grid, block, 0, st, k);
    }
  } else {
    if (lat) {
      if (iso) launch_pdl(agg_fwd_row_kernel<VEC, NS, true, true>, grid, block, 0, st, k);
      else launch_pdl(agg_fwd_row_kernel<VEC, NS, false, true>, grid, block, 0, st, k);
    } else {
      if (iso) launch_pdl(agg_fwd_row_kernel<VEC, NS, true, false>, grid, block, 0, st, k);
      else launch_pdl(agg_fwd_row_kernel<VEC, NS, false, false>, grid, block, 0, st, k);
    }
  }
  return cudaGetLastError() == cudaSuccess ? DGN_OK : DGN_ERR_CUDA;
}

template <bool BWD, int VEC>
static int launch_row_slots(const RowArgs& k, cudaStream_t st) {
  const int ns = k.rp.n_slots;
  if (ns == 0) return launch_row_iso<BWD, VEC, 0>(k, st);
  if (ns <= 2) return launch_row_iso<BWD, VEC, 2>(k, st);
  if (ns <= 4) return launch_row_iso<BWD, VEC, 4>(k, st);
  return launch_row_iso<BWD, VEC, 8>(k, st);
}

static int fill_row_args(const KernelArgs& ka, const DgnAggSpec* spec, const DgnField* f, RowArgs& k) {
  memset(&k, 0, sizeof(k));
  FieldPlan fp;
  if (int rc = make_row_plan(spec, fp, &k.rp)) return rc;
  if (!f->groups || !f->ovf_ptr || f->n_slots != fp.n_slots || (fp.n_slots > 0 && !f->wsum)) return DGN_ERR_INVALID;
  if (reinterpret_cast<uintptr_t>(f->groups) % 16 != 0 || reinterpret_cast<uintptr_t>(f->wsum) % 16 != 0)
    return DGN_ERR_ALIGNMENT;
  k.N = ka.N; k.mode = ka.mode; k.gstride = 1 + fp.n_slots;
  k.in_ptr = ka.in_ptr; k.in_eid = ka.in_eid; k.ovf_ptr = f->ovf_ptr;
  k.groups = reinterpret_cast<const float4*>(f->groups); k.wsum = f->wsum; k.log_deg = ka.log_deg;
  k.x = ka.x; k.ld_x = ka.ld_x; k.q = ka.q; k.ld_q = ka.ld_q; k.q_bias = ka.q_bias; k.r = ka.r; k.ld_r = ka.ld_r;
  k.h_in = ka.h_in; k.ld_h = ka.ld_h;
  k.out = ka.out; k.ld_out = ka.ld_out; k.out_gs = ka.out_gs;
  k.h_copy = ka.h_copy; k.ld_hc = ka.ld_hc; k.hc_gs = ka.hc_gs;
  k.g_out = ka.g_out; k.g_hcopy = ka.g_hcopy;
  k.d_q = ka.d_q; k.ld_dq = ka.ld_dq; k.d_r = ka.d_r; k.ld_dr = ka.ld_dr; k.d_h = ka.d_h; k.ld_dh = ka.ld_dh;
  k.d_h_add = ka.d_h_add; k.ld_dha = ka.ld_dha; k.edge_ws = ka.edge_ws;
  k.graph_ptr = ka.graph_ptr; k.n_graphs = ka.n_graphs; k.tile_rows = ka.max_graph_nodes; k.n_edges = ka.E;
  if (k.g_out) {
    static const bool pf_on = [] { const char* e = getenv("DGN_ROW_NO_PREFETCH"); return !(e && atoi(e) != 0); }();
    const long long row = ((long long)(k.rp.F / k.rp.Fg - 1) * k.out_gs + (long long)k.rp.S * k.rp.A * k.rp.Fg) * 4;
    if (pf_on && row % 16 == 0 && (k.ld_out % 4) == 0 && reinterpret_cast<uintptr_t>(k.g_out) % 16 == 0 && row < (1 << 20))
      k.pf_bytes = (int)row;
  }
  return DGN_OK;
}

int launch_forward_row(const KernelArgs& ka, const DgnAggSpec* spec, const DgnField* f, int vec, cudaStream_t st) {
  RowArgs k;
  if (int rc = fill_row_args(ka, spec, f, k)) return rc;
  if (vec == 4) return launch_row_slots<false, 4>(k, st);
  if (vec == 2) return launch_row_slots<false, 2>(k, st);
  return launch_row_slots<false, 1>(k, st);
}

int launch_backward_row_dst(const KernelArgs& ka, const DgnAggSpec* spec, const DgnField* f, int vec, cudaStream_t st) {
  RowArgs k;
  if (int rc = fill_row_args(ka, spec, f, k)) return rc;
  if (vec == 4) return launch_row_slots<true, 4>(k, st);
  if (vec == 2) return launch_row_slots<true, 2>(k, st);
  return launch_row_slots<true, 1>(k, st);
}

}  // namespace dgn

using namespace dgn;

extern "C" int dgn_field_slots(const DgnAggSpec* spec) {
  if (!spec) return DGN_ERR_INVALID;
  FieldPlan fp;
  if (int rc = make_row_plan(spec, fp, nullptr)) return rc;
  return fp.n_slots;
}

// HOST: overflow-group offsets of the eigen-field layout: node v has max(0, ceil((D_v - 4) / 4)) overflow groups.
extern "C" int dgn_build_groups_host(int32_t n_nodes, const int32_t* in_ptr, int32_t* ovf_ptr) {
  if (n_nodes < 0 || !in_ptr || !ovf_ptr) return DGN_ERR_INVALID;
  int32_t n = 0;
  for (int v = 0; v < n_nodes; ++v) {
    ovf_ptr[v] = n;
    const int D = in_ptr[v + 1] - in_ptr[v];
    if (D < 0) return DGN_ERR_INVALID;
    if (D > 4) n += (D - 4 + 3) / 4;
  }
  ovf_ptr[n_nodes] = n;
  return n;
}

extern "C" int dgn_field_build(const DgnGraph* g, const DgnAggSpec* spec, const float* eig, int32_t ld_eig,
                               const DgnField* f, void* stream) {
  if (!g || !spec || !f || !g->in_ptr || (g->n_edges > 0 && !g->in_src)) return DGN_ERR_INVALID;
  FieldArgs k;
  memset(&k, 0, sizeof(k));
  if (int rc = make_row_plan(spec, k.fp, nullptr)) return rc;
  if (!f->groups || !f->ovf_ptr || f->n_slots != k.fp.n_slots) return DGN_ERR_INVALID;
  if (k.fp.n_slots > 0 && (!eig || !f->wsum)) return DGN_ERR_INVALID;
  if (reinterpret_cast<uintptr_t>(f->groups) % 16 != 0) return DGN_ERR_ALIGNMENT;
  if (f->n_groups < g->n_nodes) return DGN_ERR_INVALID;
  k.N = g->n_nodes; k.in_ptr = g->in_ptr; k.in_src = g->in_src; k.ovf_ptr = f->ovf_ptr;
  k.eig = eig; k.ld_eig = ld_eig; k.groups = f->groups; k.wsum = f->wsum;
  if (k.N == 0) return DGN_OK;
  const long long threads = (long long)k.N * (k.fp.n_slots > 0 ? k.fp.n_slots : 1);
  launch_pdl(field_build_kernel, dim3((unsigned)((threads + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, k);
  if (cudaGetLastError() != cudaSuccess) { g_dgn_last_cuda = cudaPeekAtLastError(); return DGN_ERR_CUDA; }
  return DGN_OK;
}
