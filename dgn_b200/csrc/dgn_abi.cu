// C ABI entry points (include/dgn_b200.h): argument validation, spec -> launch plan, dispatch.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "dgn_plan.cuh"

thread_local cudaError_t g_dgn_last_cuda = cudaSuccess;

namespace dgn {

static int find_or_add_slot(AggPlan& P, int eig, int w, float alpha) {
  for (int s = 0; s < P.n_slots; ++s)
    if (P.slot_eig[s] == eig && P.slot_w[s] == w && (w != W_EXP || P.slot_alpha[s] == alpha)) return s;
  if (P.n_slots >= DGN_MAX_SLOTS) return -1;
  const int s = P.n_slots++;
  P.slot_eig[s] = eig;
  P.slot_w[s] = w;
  P.slot_alpha[s] = alpha;
  if (w == W_EXP) P.has_exp = 1;
  return s;
}

static int make_plan(const DgnAggSpec* spec, AggPlan& P) {
  memset(&P, 0, sizeof(P));
  if (spec->n_feat <= 0 || spec->n_agg <= 0 || spec->n_agg > DGN_MAX_AGG || spec->n_scalers <= 0 ||
      spec->n_scalers > DGN_MAX_SCALERS || spec->n_eig < 0)
    return DGN_ERR_INVALID;
  P.F = spec->n_feat;
  P.Fg = spec->group_feat > 0 ? spec->group_feat : spec->n_feat;
  if (P.F % P.Fg != 0) return DGN_ERR_INVALID;
  P.K = spec->n_eig;
  P.A = spec->n_agg;
  P.S = spec->n_scalers > 1 ? spec->n_scalers : 1;     // rb/nets/dgn_layer.py:95
  P.avg_log = spec->avg_log;
  for (int s = 0; s < spec->n_scalers; ++s) {
    if (spec->scaler_kind[s] > DGN_SCALE_ATTENUATION) return DGN_ERR_INVALID;
    P.scaler_kind[s] = spec->scaler_kind[s];
  }
  for (int a = 0; a < P.A; ++a) {
    const int kind = spec->agg_kind[a];
    const int eig = spec->agg_eig[a];
    P.agg_kind[a] = (uint8_t)kind;
    if (kind > DGN_AGG_DIR_SOFTMAX) return DGN_ERR_INVALID;
    if (kind < DGN_AGG_DIR_AV) continue;
    if (eig >= P.K) return DGN_ERR_INVALID;
    int s1 = -1;
    switch (kind) {
      case DGN_AGG_DIR_AV: s1 = find_or_add_slot(P, eig, W_ABS, 0.f); break;
      case DGN_AGG_DIR_DX:
      case DGN_AGG_DIR_DX_NO_ABS: s1 = find_or_add_slot(P, eig, W_SGN, 0.f); break;
      case DGN_AGG_DIR_DX_BALANCED: {
        s1 = find_or_add_slot(P, eig, W_POS, 0.f);      // the pair is always allocated together,
        const int s2 = find_or_add_slot(P, eig, W_NEG, 0.f);  // so W_NEG sits in slot s1 + 1
        if (s1 < 0 || s2 != s1 + 1) return DGN_ERR_UNSUPPORTED;
      } break;
      case DGN_AGG_DIR_SOFTMAX: s1 = find_or_add_slot(P, eig, W_EXP, spec->agg_alpha[a]); break;
    }
    if (s1 < 0) return DGN_ERR_UNSUPPORTED;
    P.slot_aggs[s1] |= 1u << a;
  }
  return DGN_OK;
}

// widest vector (in floats) that divides n / keeps pointer p aligned
static inline int vw(long long n) { return (n % 4 == 0) ? 4 : (n % 2 == 0) ? 2 : 1; }
static inline int vwp(const void* p) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  return (a % 16 == 0) ? 4 : (a % 8 == 0) ? 2 : 1;
}
static inline int imin(int a, int b) { return a < b ? a : b; }

static int fill_args(const DgnGraph* g, const DgnAggSpec* spec, const DgnAggIO* io, KernelArgs& k, int& vec) {
  if (!g || !spec || !io) return DGN_ERR_INVALID;
  if (g->n_nodes < 0 || g->n_edges < 0 || !g->in_ptr || (g->n_edges > 0 && !g->in_src)) return DGN_ERR_INVALID;
  memset(&k, 0, sizeof(k));
  if (int rc = make_plan(spec, k.plan)) return rc;
  const AggPlan& P = k.plan;
  k.N = g->n_nodes; k.E = g->n_edges; k.mode = io->msg_mode;
  k.in_ptr = g->in_ptr; k.in_src = g->in_src; k.in_eid = g->in_eid;
  k.out_ptr = g->out_ptr; k.out_slot = g->out_slot; k.log_deg = g->log_deg;
  k.graph_ptr = g->graph_ptr; k.n_graphs = g->graph_ptr ? g->n_graphs : 0; k.max_graph_nodes = g->max_graph_nodes;
  if (P.S > 1 && !g->log_deg) return DGN_ERR_INVALID;
  k.x = io->x; k.ld_x = io->ld_x; k.q = io->q; k.ld_q = io->ld_q; k.r = io->r; k.ld_r = io->ld_r;
  k.q_bias = (io->msg_mode == DGN_MSG_AFFINE) ? io->q_bias : nullptr;
  k.h_in = io->h_in; k.ld_h = io->ld_h; k.eig = io->eig; k.ld_eig = io->ld_eig;
  k.out = io->out; k.ld_out = io->ld_out; k.out_gs = io->out_group_stride;
  k.h_copy = io->h_copy; k.ld_hc = io->ld_hcopy; k.hc_gs = io->hcopy_group_stride;
  if (!k.h_in || !k.out) return DGN_ERR_INVALID;
  if (P.n_slots > 0 && !k.eig) return DGN_ERR_INVALID;
  switch (k.mode) {
    case DGN_MSG_SOURCE: if (!k.x) return DGN_ERR_INVALID; k.q = nullptr; k.r = nullptr; break;
    case DGN_MSG_AFFINE: if (!k.x || !k.q) return DGN_ERR_INVALID; break;
    case DGN_MSG_DENSE: if (!k.r) return DGN_ERR_INVALID; k.x = nullptr; k.q = nullptr; break;
    default: return DGN_ERR_INVALID;
  }
  // vector width: every row start and every slab start must be aligned to it
  vec = imin(imin(vw(P.F), vw(P.Fg)), imin(imin(vw(k.ld_h), vw(k.ld_out)), vw(k.out_gs)));
  vec = imin(vec, imin(vwp(k.h_in), vwp(k.out)));
  if (k.x) vec = imin(vec, imin(vw(k.ld_x), vwp(k.x)));
  if (k.q) vec = imin(vec, imin(vw(k.ld_q), vwp(k.q)));
  if (k.q_bias) vec = imin(vec, vwp(k.q_bias));
  if (k.r) vec = imin(vec, imin(vw(k.ld_r), vwp(k.r)));
  if (k.h_copy) vec = imin(vec, imin(imin(vw(k.ld_hc), vw(k.hc_gs)), vwp(k.h_copy)));
  return DGN_OK;
}

}  // namespace dgn

using namespace dgn;

extern "C" int dgn_abi_version(void) { return DGN_ABI_VERSION; }

extern "C" const char* dgn_status_string(int status) {
  switch (status) {
    case DGN_OK: return "ok";
    case DGN_ERR_INVALID: return "invalid argument";
    case DGN_ERR_UNSUPPORTED: return "unsupported configuration (more than DGN_MAX_SLOTS eigen-weighted sums)";
    case DGN_ERR_ALIGNMENT: return "misaligned operand";
    case DGN_ERR_CUDA: return "CUDA error";
    default: return "unknown status";
  }
}

extern "C" const char* dgn_last_cuda_error(void) { return cudaGetErrorString(g_dgn_last_cuda); }

extern "C" int dgn_agg_forward(const DgnGraph* g, const DgnAggSpec* spec, const DgnAggIO* io, void* stream) {
  KernelArgs k;
  int vec = 1;
  if (int rc = fill_args(g, spec, io, k, vec)) return rc;
  vec = choose_vec(vec, k.N, k.plan.F, true);
  k.plan.chunks = k.plan.F / vec;
  int rc = io->field ? launch_forward_row(k, spec, io->field, vec, (cudaStream_t)stream) : DGN_ERR_UNSUPPORTED;
  if (rc == DGN_ERR_UNSUPPORTED)           // no field, or a row wider than the row kernels' CTA: in-kernel weights
    rc = launch_forward(k, vec, (cudaStream_t)stream);
  if (rc == DGN_ERR_CUDA) g_dgn_last_cuda = cudaPeekAtLastError();
  return rc;
}

extern "C" int dgn_agg_backward(const DgnGraph* g, const DgnAggSpec* spec, const DgnAggIO* io, const DgnAggGrad* grad,
                                void* stream) {
  KernelArgs k;
  int vec = 1;
  if (!grad || !grad->g_out) return DGN_ERR_INVALID;
  if (int rc = fill_args(g, spec, io, k, vec)) return rc;
  k.g_out = grad->g_out;
  k.g_hcopy = grad->g_hcopy;
  k.d_q = grad->d_q; k.ld_dq = grad->ld_dq;
  k.d_r = grad->d_r; k.ld_dr = grad->ld_dr;
  k.d_h = grad->d_h_in; k.ld_dh = grad->ld_dh;
  k.d_h_add = grad->d_h_in ? grad->d_h_addend : nullptr; k.ld_dha = grad->ld_dha;
  k.edge_ws = grad->edge_ws;               // with d_x == NULL: spill only, the caller reduces over the out-edges
  if (grad->d_x && (!grad->edge_ws || !g->out_ptr || (g->n_edges > 0 && !g->out_slot))) return DGN_ERR_INVALID;
  if (grad->fold_h_in && (!grad->d_x || !grad->d_h_in)) return DGN_ERR_INVALID;
  if (k.g_hcopy && !k.h_copy) { k.ld_hc = io->ld_hcopy; k.hc_gs = io->hcopy_group_stride; }
  vec = imin(vec, vwp(k.g_out));
  if (k.g_hcopy) vec = imin(vec, imin(vwp(k.g_hcopy), imin(vw(k.ld_hc), vw(k.hc_gs))));
  if (k.d_q) vec = imin(vec, imin(vwp(k.d_q), vw(k.ld_dq)));
  if (k.d_r) vec = imin(vec, imin(vwp(k.d_r), vw(k.ld_dr)));
  if (k.d_h) vec = imin(vec, imin(vwp(k.d_h), vw(k.ld_dh)));
  if (k.d_h_add) vec = imin(vec, imin(vwp(k.d_h_add), vw(k.ld_dha)));
  if (grad->d_x) vec = imin(vec, imin(vwp(grad->d_x), vw(grad->ld_dx)));
  if (grad->edge_ws) vec = imin(vec, vwp(grad->edge_ws));
  // measured (profiles/README.md): the row backward is fastest with 16 B lanes at every size; the kernels that derive
  // the weights in the launch prefer 8 B lanes for small launches
  vec = choose_vec(vec, k.N, k.plan.F, io->field == nullptr);
  k.plan.chunks = k.plan.F / vec;
  int rc = DGN_ERR_UNSUPPORTED;
  if (io->field) {
    rc = launch_backward_row_dst(k, spec, io->field, vec, (cudaStream_t)stream);
    if (rc == DGN_OK && grad->d_x)
      rc = launch_backward_src(k, vec, grad->d_x, grad->ld_dx, grad->fold_h_in ? grad->d_h_in : nullptr, grad->ld_dh,
                               (cudaStream_t)stream);
  }
  if (rc == DGN_ERR_UNSUPPORTED)           // no field, or a row wider than the row kernels' CTA: in-kernel weights
    rc = launch_backward(k, vec, grad->d_x, grad->ld_dx, grad->fold_h_in ? grad->d_h_in : nullptr, grad->ld_dh,
                         (cudaStream_t)stream);
  if (rc == DGN_ERR_CUDA) g_dgn_last_cuda = cudaPeekAtLastError();
  return rc;
}

// Stable counting sort by destination, then by source over the slot order.  O(N + E), host.
extern "C" int dgn_build_csr_host(int32_t n_nodes, int32_t n_edges, const int32_t* src, const int32_t* dst,
                                  int32_t* in_ptr, int32_t* in_src, int32_t* in_eid, int32_t* out_ptr,
                                  int32_t* out_slot, float* log_deg) {
  if (n_nodes < 0 || n_edges < 0 || !in_ptr || (n_edges > 0 && (!src || !dst || !in_src))) return DGN_ERR_INVALID;
  for (int v = 0; v <= n_nodes; ++v) in_ptr[v] = 0;
  for (int e = 0; e < n_edges; ++e) {
    if (dst[e] < 0 || dst[e] >= n_nodes || src[e] < 0 || src[e] >= n_nodes) return DGN_ERR_INVALID;
    ++in_ptr[dst[e] + 1];
  }
  for (int v = 0; v < n_nodes; ++v) in_ptr[v + 1] += in_ptr[v];
  std::vector<int32_t> cur(in_ptr, in_ptr + n_nodes);
  for (int e = 0; e < n_edges; ++e) {                 // ascending edge id => mailbox order inside a row
    const int slot = cur[dst[e]]++;
    in_src[slot] = src[e];
    if (in_eid) in_eid[slot] = e;
  }
  if (log_deg)
    for (int v = 0; v < n_nodes; ++v) log_deg[v] = (float)log((double)(in_ptr[v + 1] - in_ptr[v]) + 1.0);
  if (out_ptr && (out_slot || n_edges == 0)) {
    for (int v = 0; v <= n_nodes; ++v) out_ptr[v] = 0;
    for (int e = 0; e < n_edges; ++e) ++out_ptr[src[e] + 1];
    for (int v = 0; v < n_nodes; ++v) out_ptr[v + 1] += out_ptr[v];
    std::vector<int32_t> oc(out_ptr, out_ptr + n_nodes);
    for (int slot = 0; slot < n_edges; ++slot) out_slot[oc[in_src[slot]]++] = slot;
  }
  return DGN_OK;
}
