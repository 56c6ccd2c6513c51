// Fused DGN aggregation, backward (sm_100a).
//
// Destination-side kernel: same (node, VEC-column chunk) ownership as the forward.  Pass 1
// re-runs the forward accumulation (recomputing mean / var / max / min / eigen-weighted sums
// costs D gathered rows that are L2-resident, far less than saving and re-reading per-node
// statistics next to the S*A*F-wide gradient).  The S*A gradient slabs of the node are then
// folded into a handful of per-column coefficients so that the gradient of every in-edge
// message is
//     dm_u = c0 + c1*m_u + [u first argmax]*g_max + [u first argmin]*g_min + sum_s w_s(delta_u)*cs_s
// (SURVEY.md appendix A).  Pass 2 walks the in-edges again, emits dm_u into an [E,F]
// slot-ordered workspace (and/or the per-edge gradient d_r) and reduces d_q[v] = sum_u dm_u.
//
// Source-side kernel: d_x[u] = sum over out-edges of dm (by-source CSR), a deterministic gather
// instead of float atomics, so training is bit-reproducible run to run.
#include "dgn_launch.cuh"
#include "dgn_plan.cuh"

namespace dgn {

template <int MODE, int VEC, int NS, bool ISO, bool EXP>
__global__ void __launch_bounds__(256) agg_bwd_dst_kernel(const __grid_constant__ KernelArgs k) {
  const AggPlan& P = k.plan;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int v = (int)(tid / P.chunks);
  if (v >= k.N) return;
  const int c = (int)(tid - (long long)v * P.chunks) * VEC;
  const int tower = c / P.Fg;
  const int cg = c - tower * P.Fg;

  const int e0 = __ldg(k.in_ptr + v), e1 = __ldg(k.in_ptr + v + 1);
  const int D = e1 - e0;

  Vec<VEC> dh = vfill<VEC>(0.f);
  if (k.g_hcopy) dh = vload_stream<VEC>(k.g_hcopy + (size_t)v * k.ld_hc + (size_t)tower * k.hc_gs + cg);
  if (k.d_h_add) {
    const Vec<VEC> t = vload_stream<VEC>(k.d_h_add + (size_t)v * k.ld_dha + c);
#pragma unroll
    for (int i = 0; i < VEC; ++i) dh.a[i] += t.a[i];
  }

  if (D == 0) {                      // isolated node: the forward wrote zeros, no gradient flows
    if (k.d_h) vstore<VEC>(k.d_h + (size_t)v * k.ld_dh + c, dh);
    if (k.d_q) vstore<VEC>(k.d_q + (size_t)v * k.ld_dq + c, vfill<VEC>(0.f));
    return;
  }

  // ---- phase 1: G_a = sum_s coef_s * g_out[v, s, a, :] for every aggregator, staged in shared memory.
  // The S*A slab loads are the dominant traffic of this kernel and independent of everything else, so they
  // are issued first and back to back (unroll 4 => 4*S loads in flight per thread) instead of three at a
  // time inside the per-aggregator switch.  Each thread only ever reads its own entries: no barrier needed.
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Vec<VEC>* sG = reinterpret_cast<Vec<VEC>*>(smem_raw);
  float coef[DGN_MAX_SCALERS];
  scaler_coefs(k, v, coef);
  {
    const float* grow = k.g_out + (size_t)v * k.ld_out + (size_t)tower * k.out_gs + cg;
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
      sG[a * (int)blockDim.x + (int)threadIdx.x] = G;
    }
  }

  const Vec<VEC> hv = vload<VEC>(k.h_in + (size_t)v * k.ld_h + c);
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

  const float fD = (float)D;
  Vec<VEC> mean, var;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    mean.a[i] = __fdiv_rn(R.sum.a[i], fD);
    if constexpr (ISO) {
      const float msq = __fdiv_rn(R.sq.a[i], fD);
      var.a[i] = fmaxf(__fsub_rn(msq, __fmul_rn(mean.a[i], mean.a[i])), 0.f);
    } else {
      var.a[i] = 0.f;
    }
  }

  // ---- fold the S*A gradient slabs into per-column coefficients -----------------------------------
  const float rD = __frcp_rn(fD);
  Vec<VEC> c0 = vfill<VEC>(0.f), c1 = vfill<VEC>(0.f), gmx = vfill<VEC>(0.f), gmn = vfill<VEC>(0.f);
  Vec<VEC> cs[NS > 0 ? NS : 1];
#pragma unroll
  for (int s = 0; s < NS; ++s) cs[s] = vfill<VEC>(0.f);

  auto slab_grad = [&](int a) { return sG[a * (int)blockDim.x + (int)threadIdx.x]; };

  for (int a = 0; a < P.A; ++a) {             // isotropic aggregators
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
  for (int s = 0; s < NS; ++s) {              // directional aggregators, static slot index
    if (s >= P.n_slots) break;
    unsigned todo = P.slot_aggs[s];
    while (todo) {
      const int a = __ffs(todo) - 1;
      todo &= todo - 1;
      const int kind = P.agg_kind[a];
      const Vec<VEC> G = slab_grad(a);
      const Vec<VEC>& A1 = R.acc[s];
      const float zw1 = R.zw[s], zabs1 = R.zabs[s];
      if (kind == DGN_AGG_DIR_AV) {
        const float rz = __frcp_rn(zabs1 + DGN_EPS);
#pragma unroll
        for (int i = 0; i < VEC; ++i) cs[s].a[i] = fmaf(G.a[i], rz, cs[s].a[i]);
      } else if (kind == DGN_AGG_DIR_DX || kind == DGN_AGG_DIR_DX_NO_ABS) {
        const float rz = __frcp_rn(zabs1 + DGN_EPS);
        const float wsum = zw1 * rz;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          float g = G.a[i];
          if (kind == DGN_AGG_DIR_DX) g *= sign0(A1.a[i] * rz - wsum * hv.a[i]);   // same expression as the forward
          cs[s].a[i] = fmaf(g, rz, cs[s].a[i]);
          dh.a[i] -= wsum * g;
        }
      } else if (kind == DGN_AGG_DIR_DX_BALANCED) {
        if constexpr (NS >= 2) {
          constexpr int LAST = NS - 1;
          const int s2 = (s + 1 <= LAST) ? s + 1 : LAST;
          const Vec<VEC>& A2 = R.acc[s2];
          const float zw2 = R.zw[s2];
          const float rp = 0.5f * __frcp_rn(zw1 + DGN_EPS), rn = 0.5f * __frcp_rn(zw2 + DGN_EPS);
          const float wsum = zw1 * rp + zw2 * rn;
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            const float sv = A1.a[i] * rp + A2.a[i] * rn - wsum * hv.a[i];
            const float g = G.a[i] * sign0(sv);
            cs[s].a[i] = fmaf(g, rp, cs[s].a[i]);
            cs[s2].a[i] = fmaf(g, rn, cs[s2].a[i]);
            dh.a[i] -= wsum * g;
          }
        }
      } else {                                 // DGN_AGG_DIR_SOFTMAX
        const float rz = __frcp_rn(zw1);
#pragma unroll
        for (int i = 0; i < VEC; ++i) cs[s].a[i] = fmaf(G.a[i], rz, cs[s].a[i]);
      }
    }
  }

  // ---- pass 2: per-edge message gradients -------------------------------------------------------
  Vec<VEC> dq = vfill<VEC>(0.f);
  unsigned given = 0u;      // bit i: max gradient of column i already routed; bit 4+i: min
#pragma unroll 2
  for (int e = e0; e < e1; ++e) {
    const int u = __ldg(k.in_src + e);
    const Vec<VEC> m = load_message<MODE, VEC>(k, u, e, c, qv);
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
      if (s < P.n_slots) {
        const float d = __ldg(k.eig + (size_t)u * k.ld_eig + P.slot_eig[s]) - ev[s];
        const float w = edge_weight(P.slot_w[s], d, P.slot_alpha[s], shift[s]);
#pragma unroll
        for (int i = 0; i < VEC; ++i) dm.a[i] = fmaf(w, cs[s].a[i], dm.a[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) dq.a[i] += dm.a[i];
    if (k.edge_ws) vstore<VEC>(k.edge_ws + (size_t)e * P.F + c, dm);
    if (k.d_r) {
      const int id = k.in_eid ? __ldg(k.in_eid + e) : e;
      vstore<VEC>(k.d_r + (size_t)id * k.ld_dr + c, dm);
    }
  }
  if (k.d_q) vstore<VEC>(k.d_q + (size_t)v * k.ld_dq + c, dq);
  if (k.d_h) vstore<VEC>(k.d_h + (size_t)v * k.ld_dh + c, dh);
}

// d_x[u] = (addend[u]) + sum over out-edges j of u of ws[out_slot[j]]
template <int VEC>
__global__ void __launch_bounds__(256) agg_bwd_src_kernel(int N, int F, int chunks, const int32_t* __restrict__ out_ptr,
                                                          const int32_t* __restrict__ out_slot,
                                                          const float* __restrict__ ws, float* __restrict__ d_x,
                                                          int ld_dx, const float* __restrict__ addend, int ld_add) {
  pdl_prologue();
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int u = (int)(tid / chunks);
  if (u >= N) return;
  const int c = (int)(tid - (long long)u * chunks) * VEC;
  Vec<VEC> acc = addend ? vload<VEC>(addend + (size_t)u * ld_add + c) : vfill<VEC>(0.f);
  const int j0 = __ldg(out_ptr + u), j1 = __ldg(out_ptr + u + 1);
#pragma unroll 4
  for (int j = j0; j < j1; ++j) {
    const int slot = __ldg(out_slot + j);
    const Vec<VEC> t = vload_stream<VEC>(ws + (size_t)slot * F + c);
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc.a[i] += t.a[i];
  }
  vstore<VEC>(d_x + (size_t)u * ld_dx + c, acc);
}

template <int MODE, int VEC, int NS, bool ISO, bool EXP>
static int launch_dst(const KernelArgs& k, cudaStream_t st) {
  const long long threads = (long long)k.N * k.plan.chunks;
  if (threads == 0) return DGN_OK;
  const int block = 256;
  const long long grid = (threads + block - 1) / block;
  const size_t smem = (size_t)k.plan.A * block * VEC * sizeof(float);     // staged G_a, <= 128 KB (A <= 32)
  static bool big_smem_enabled = false;                                    // per template instantiation
  if (smem > 48 * 1024 && !big_smem_enabled) {
    if (cudaFuncSetAttribute(agg_bwd_dst_kernel<MODE, VEC, NS, ISO, EXP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             DGN_MAX_AGG * block * VEC * (int)sizeof(float)) != cudaSuccess)
      return DGN_ERR_CUDA;
    big_smem_enabled = true;
  }
  agg_bwd_dst_kernel<MODE, VEC, NS, ISO, EXP><<<(unsigned)grid, block, smem, st>>>(k);
  return cudaGetLastError() == cudaSuccess ? DGN_OK : DGN_ERR_CUDA;
}

template <int MODE, int VEC>
static int dispatch_slots(const KernelArgs& k, bool iso, cudaStream_t st) {
  const int ns = k.plan.n_slots;
  if (k.plan.has_exp) return launch_dst<MODE, VEC, 8, true, true>(k, st);
  if (iso) {
    if (ns == 0) return launch_dst<MODE, VEC, 0, true, false>(k, st);
    if (ns <= 2) return launch_dst<MODE, VEC, 2, true, false>(k, st);
    if (ns <= 4) return launch_dst<MODE, VEC, 4, true, false>(k, st);
    return launch_dst<MODE, VEC, 8, true, false>(k, st);
  }
  if (ns == 0) return launch_dst<MODE, VEC, 0, false, false>(k, st);
  if (ns <= 2) return launch_dst<MODE, VEC, 2, false, false>(k, st);
  if (ns <= 4) return launch_dst<MODE, VEC, 4, false, false>(k, st);
  return launch_dst<MODE, VEC, 8, false, false>(k, st);
}

template <int VEC>
static int dispatch_mode(const KernelArgs& k, bool iso, cudaStream_t st) {
  if (k.mode == DGN_MSG_SOURCE) return dispatch_slots<DGN_MSG_SOURCE, VEC>(k, iso, st);
  if (k.mode == DGN_MSG_AFFINE) return dispatch_slots<DGN_MSG_AFFINE, VEC>(k, iso, st);
  return dispatch_slots<DGN_MSG_DENSE, VEC>(k, iso, st);
}

int launch_backward(const KernelArgs& k, int vec, float* d_x, int ld_dx, const float* addend, int ld_add,
                    cudaStream_t st) {
  bool iso = false;
  for (int a = 0; a < k.plan.A; ++a) {
    const int kd = k.plan.agg_kind[a];
    iso |= (kd == DGN_AGG_MAX || kd == DGN_AGG_MIN || kd == DGN_AGG_STD || kd == DGN_AGG_VAR);
  }
  int rc = vec == 4 ? dispatch_mode<4>(k, iso, st) : vec == 2 ? dispatch_mode<2>(k, iso, st) : dispatch_mode<1>(k, iso, st);
  if (rc != DGN_OK || !d_x) return rc;
  return launch_backward_src(k, vec, d_x, ld_dx, addend, ld_add, st);
}

// d_x = (addend) + source-side gather of the per-edge message gradients in edge_ws
int launch_backward_src(const KernelArgs& k, int vec, float* d_x, int ld_dx, const float* addend, int ld_add,
                        cudaStream_t st) {
  const long long threads = (long long)k.N * k.plan.chunks;
  if (threads == 0) return DGN_OK;
  const int block = 256;
  const unsigned grid = (unsigned)((threads + block - 1) / block);
  if (vec == 4)
    launch_pdl(agg_bwd_src_kernel<4>, dim3(grid), dim3(block), 0, st, k.N, k.plan.F, k.plan.chunks, k.out_ptr, k.out_slot, k.edge_ws, d_x, ld_dx, addend, ld_add);
  else if (vec == 2)
    launch_pdl(agg_bwd_src_kernel<2>, dim3(grid), dim3(block), 0, st, k.N, k.plan.F, k.plan.chunks, k.out_ptr, k.out_slot, k.edge_ws, d_x, ld_dx, addend, ld_add);
  else
    launch_pdl(agg_bwd_src_kernel<1>, dim3(grid), dim3(block), 0, st, k.N, k.plan.F, k.plan.chunks, k.out_ptr, k.out_slot, k.edge_ws, d_x, ld_dx, addend, ld_add);
  return cudaGetLastError() == cudaSuccess ? DGN_OK : DGN_ERR_CUDA;
}

}  // namespace dgn
