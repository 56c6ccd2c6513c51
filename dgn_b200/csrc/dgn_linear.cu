// Node-level halves of the 1-layer pretrans (rb/nets/dgn_layer.py:75-80), fp32 on CUDA cores.
//
// pretrans(cat(h_u, h_v)) = W_src h_u + W_dst h_v + b splits per node into P = h W_src^T and Q = h W_dst^T
// with W = [W_src | W_dst] stored as one [F_out, 2 F_in (+edge)] parameter.  These K = F_in (<= 128) products are
// launch-latency bound; doing both halves in ONE launch straight from the parameter's layout (no slicing, no
// second GEMM) halves their cost.  Same for the backward d_h += d_P W_src + d_Q W_dst.
// Tiles: 32 rows x 64 columns per 128-thread block, 4 x 4 outputs per thread, operands staged in shared
// memory so that every inner-loop read is a conflict-free 128-bit load.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dgn_b200.h"
#include "dgn_launch.cuh"

namespace dgn {

constexpr int LR = 32, LC = 64, LT = 128;     // rows / columns per block, threads

// Asynchronous 4-byte global -> shared copies (LDGSTS): the staging loops issue all their copies back to back instead
// of 80 dependent load -> store round trips per thread - with 4 warps per CTA nothing else hides that latency.
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// P[n,o] = sum_i h[n,i] W[o,i] ;  Q[n,o] = sum_i h[n,i] W[o,Fi+i]
__global__ void __launch_bounds__(LT) pair_linear_fwd_kernel(int N, int Fi, int Fo, const float* __restrict__ h, int ld_h,
                                                             const float* __restrict__ W, int ld_w,
                                                             float* __restrict__ P, int ld_p, float* __restrict__ Q,
                                                             int ld_q) {
  pdl_prologue();
  extern __shared__ __align__(16) float sm[];
  float* ht = sm;                              // [Fi][LR]   h tile, transposed
  float* wp = ht + Fi * LR;                    // [Fi][LC]   W_src block, transposed
  float* wq = wp + Fi * LC;                    // [Fi][LC]   W_dst block, transposed
  const int t = threadIdx.x, r0 = blockIdx.x * LR, c0 = blockIdx.y * LC;
  // Staging transposes the tiles.  Consecutive threads take consecutive rows / output columns, so the shared-memory
  // stores are conflict free (the other order - coalesced global reads, stride-LR stores - serialises every store
  // instruction 32 ways and was most of this kernel's time); the strided global reads hit L1 / L2.
  for (int idx = t; idx < LR * Fi; idx += LT) {
    const int i = idx / LR, r = idx - i * LR;
    if (r0 + r < N) cp_async4(ht + idx, h + (size_t)(r0 + r) * ld_h + i);
    else ht[idx] = 0.f;
  }
  for (int idx = t; idx < LC * Fi; idx += LT) {
    const int i = idx / LC, o = idx - i * LC;
    if (c0 + o < Fo) {
      cp_async4(wp + idx, W + (size_t)(c0 + o) * ld_w + i);
      cp_async4(wq + idx, W + (size_t)(c0 + o) * ld_w + Fi + i);
    } else {
      wp[idx] = 0.f;
      wq[idx] = 0.f;
    }
  }
  cp_async_wait_all();
  __syncthreads();
  const int tx = t & 15, ty = t >> 4;          // 16 column groups x 8 row groups
  float ap[4][4] = {}, aq[4][4] = {};
#pragma unroll 4
  for (int i = 0; i < Fi; ++i) {
    const float4 hv = *reinterpret_cast<const float4*>(ht + i * LR + ty * 4);
    const float4 pv = *reinterpret_cast<const float4*>(wp + i * LC + tx * 4);
    const float4 qv = *reinterpret_cast<const float4*>(wq + i * LC + tx * 4);
    const float hr[4] = {hv.x, hv.y, hv.z, hv.w}, pw[4] = {pv.x, pv.y, pv.z, pv.w}, qw[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        ap[a][b] = fmaf(hr[a], pw[b], ap[a][b]);
        aq[a][b] = fmaf(hr[a], qw[b], aq[a][b]);
      }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int r = r0 + ty * 4 + a, c = c0 + tx * 4;
    if (r < N && c < Fo) {                      // Fo % 4 == 0: the 4 columns are in or out together
      *reinterpret_cast<float4*>(P + (size_t)r * ld_p + c) = make_float4(ap[a][0], ap[a][1], ap[a][2], ap[a][3]);
      *reinterpret_cast<float4*>(Q + (size_t)r * ld_q + c) = make_float4(aq[a][0], aq[a][1], aq[a][2], aq[a][3]);
    }
  }
}

// d_h[n,i] += sum_o dP[n,o] W[o,i] + dQ[n,o] W[o,Fi+i]
__global__ void __launch_bounds__(LT) pair_linear_bwd_kernel(int N, int Fi, int Fo, const float* __restrict__ dP, int ld_p,
                                                             const float* __restrict__ dQ, int ld_q,
                                                             const float* __restrict__ W, int ld_w,
                                                             float* __restrict__ d_h, int ld_dh) {
  pdl_prologue();
  extern __shared__ __align__(16) float sm[];
  float* pt = sm;                              // [Fo][LR]  dP tile, transposed
  float* qt = pt + Fo * LR;                    // [Fo][LR]  dQ tile, transposed
  float* ws = qt + Fo * LR;                    // [Fo][LC]  W[:, c0:c0+LC]
  float* wd = ws + Fo * LC;                    // [Fo][LC]  W[:, Fi+c0 : Fi+c0+LC]
  const int t = threadIdx.x, r0 = blockIdx.x * LR, c0 = blockIdx.y * LC;
  for (int idx = t; idx < LR * Fo; idx += LT) {          // conflict-free transposing stores, see the forward
    const int o = idx / LR, r = idx - o * LR;
    if (r0 + r < N) {
      cp_async4(pt + idx, dP + (size_t)(r0 + r) * ld_p + o);
      cp_async4(qt + idx, dQ + (size_t)(r0 + r) * ld_q + o);
    } else {
      pt[idx] = 0.f;
      qt[idx] = 0.f;
    }
  }
  for (int idx = t; idx < Fo * LC; idx += LT) {
    const int o = idx / LC, i = idx - o * LC;
    if (c0 + i < Fi) {
      cp_async4(ws + idx, W + (size_t)o * ld_w + c0 + i);
      cp_async4(wd + idx, W + (size_t)o * ld_w + Fi + c0 + i);
    } else {
      ws[idx] = 0.f;
      wd[idx] = 0.f;
    }
  }
  cp_async_wait_all();
  __syncthreads();
  const int tx = t & 15, ty = t >> 4;
  float acc[4][4] = {};
#pragma unroll 4
  for (int o = 0; o < Fo; ++o) {
    const float4 pv = *reinterpret_cast<const float4*>(pt + o * LR + ty * 4);
    const float4 qv = *reinterpret_cast<const float4*>(qt + o * LR + ty * 4);
    const float4 sv = *reinterpret_cast<const float4*>(ws + o * LC + tx * 4);
    const float4 dv = *reinterpret_cast<const float4*>(wd + o * LC + tx * 4);
    const float pr[4] = {pv.x, pv.y, pv.z, pv.w}, qr[4] = {qv.x, qv.y, qv.z, qv.w};
    const float sw[4] = {sv.x, sv.y, sv.z, sv.w}, dw[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(pr[a], sw[b], fmaf(qr[a], dw[b], acc[a][b]));
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int r = r0 + ty * 4 + a, c = c0 + tx * 4;
    if (r < N && c < Fi) {
      float4* dst = reinterpret_cast<float4*>(d_h + (size_t)r * ld_dh + c);
      float4 v = *dst;
      v.x += acc[a][0]; v.y += acc[a][1]; v.z += acc[a][2]; v.w += acc[a][3];
      *dst = v;
    }
  }
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
  const size_t smem = (size_t)(Fi * LR + 2 * Fi * LC) * sizeof(float);
  if (int rc = lin_attr(pair_linear_fwd_kernel, smem)) return rc;
  launch_pdl(pair_linear_fwd_kernel, dim3((N + LR - 1) / LR, (Fo + LC - 1) / LC), dim3(LT), smem, (cudaStream_t)stream, N, Fi, Fo, h, ld_h, W, ld_w, P, ld_p, Q, ld_q);
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
  const size_t smem = (size_t)(2 * Fo * LR + 2 * Fo * LC) * sizeof(float);
  if (int rc = lin_attr(pair_linear_bwd_kernel, smem)) return rc;
  launch_pdl(pair_linear_bwd_kernel, dim3((N + LR - 1) / LR, (Fi + LC - 1) / LC), dim3(LT), smem, (cudaStream_t)stream, N, Fi, Fo, dP, ld_p, dQ, ld_q, W, ld_w, d_h, ld_dh);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_dgn_last_cuda = e; return DGN_ERR_CUDA; }
  return DGN_OK;
}
