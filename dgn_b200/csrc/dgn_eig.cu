// Laplacian eigenvectors of every graph of a dataset / batch on the device (sm_100a), and the training-time sign-flip
// augmentation.  Replaces the per-graph host loop of the reference's loaders
//     L = diag(clip(deg, 1)) - A   ('none')   |   I - D^-1/2 A D^-1/2   ('sym')   |   I - D^-1 A   ('walk')
//     EigVal, EigVec = scipy.sparse.linalg.eigs(L, k, which='SR', tol=...) ; sort ; ndata['eig'] = real(EigVec[:, :k])
// (rb/data/molecules.py:100-116, rb/data/SBMs.py:110-139, rb/data/HIV.py:17-46) and
//     sign_flip = rand(eig.size()) >= 0.5 ? +1 : -1 ; eig *= sign_flip      (rb/train/train_molecules_graph_regression.py:29-33)
//
// One CTA per graph, the whole (shifted) Laplacian in shared memory (graphs of the benchmarks have <= ~220 nodes:
// 220^2 floats = 190 KB).  One-sided (Hestenes) Jacobi: the columns of G = L + I are rotated pairwise until they are
// mutually orthogonal; then G = V diag(lambda + 1), i.e. the eigenvectors are the normalised columns and the
// eigenvalues their norms - no second n x n matrix is needed, a warp owns one column pair (coalesced, conflict-free
// column accesses, warp-shuffle reductions), and n / 2 disjoint pairs rotate concurrently in a round-robin schedule.
// The shift by I makes every column norm >= 1 (the Laplacian is positive semi-definite), so the constant
// eigenvector (lambda = 0) is as well conditioned as the others.  ARPACK with tol = 5e-1 is not reproducible - sign
// and the basis of degenerate eigenspaces are arbitrary there too; here they are deterministic (largest-magnitude
// entry positive), parity is on the eigenpairs (residual, orthonormality, eigenvalues vs a dense fp64 solve).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/dgn_b200.h"
#include "dgn_launch.cuh"

extern thread_local cudaError_t g_dgn_last_cuda;

namespace dgn {

struct EigArgs {
  int n_graphs, norm, k, ld_eig, max_sweeps;
  const int32_t* node_off;
  const int32_t* in_ptr;
  const int32_t* in_src;
  float* eig;
  float* eigval;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void eig_jacobi_kernel(const __grid_constant__ EigArgs k) {
  pdl_prologue();
  extern __shared__ float sm[];
  __shared__ unsigned s_off;
  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int n0 = __ldg(k.node_off + g), n = __ldg(k.node_off + g + 1) - n0;
  float* G = sm;                       // [n][n] column-major: column j at G + j * n
  float* dg = sm + n * n;              // [n] column norms (eigenvalue + 1)
  float* deg = dg + n;                 // [n] clip(degree, 1) of the symmetrised adjacency
  if (n <= 0) return;
  // ---- symmetrised adjacency marks, degrees, shifted Laplacian ----------------------------------------------------
  for (int i = tid; i < n * n; i += blockDim.x) G[i] = 0.f;
  __syncthreads();
  for (int v = warp; v < n; v += nwarps) {
    const int e0 = __ldg(k.in_ptr + n0 + v), e1 = __ldg(k.in_ptr + n0 + v + 1);
    for (int e = e0 + lane; e < e1; e += 32) {
      const int u = __ldg(k.in_src + e) - n0;
      if (u >= 0 && u < n && u != v) { G[u * n + v] = 1.f; G[v * n + u] = 1.f; }
    }
  }
  __syncthreads();
  for (int j = warp; j < n; j += nwarps) {
    float d = 0.f;
    for (int i = lane; i < n; i += 32) d += G[j * n + i];
    d = warp_sum(d);
    if (lane == 0) deg[j] = fmaxf(d, 1.f);                 // clip(deg, 1)
  }
  __syncthreads();
  for (int idx = tid; idx < n * n; idx += blockDim.x) {
    const int j = idx / n, i = idx - j * n;
    float v;
    if (i == j) v = (k.norm == 0 ? deg[j] : 1.f) + 1.f;    // + I: the shift
    else v = -G[idx] * (k.norm == 0 ? 1.f : rsqrtf(deg[i] * deg[j]));   // 'walk' shares the 'sym' eigenproblem
    G[idx] = v;
  }
  __syncthreads();
  // ---- Hestenes sweeps ---------------------------------------------------------------------------------------------
  const int m = n + (n & 1);
  for (int sweep = 0; sweep < k.max_sweeps && n > 1; ++sweep) {
    if (tid == 0) s_off = 0u;
    __syncthreads();
    for (int s = 0; s < m - 1; ++s) {
      for (int kk = warp; kk < m / 2; kk += nwarps) {
        int p = (kk == 0) ? m - 1 : (s + kk) % (m - 1);
        int q = (kk == 0) ? s : (s - kk + (m - 1)) % (m - 1);
        if (p >= n || q >= n) continue;                    // the bye of an odd-sized tournament
        if (p > q) { const int t = p; p = q; q = t; }
        float* gp = G + p * n;
        float* gq = G + q * n;
        float a = 0.f, b = 0.f, c = 0.f;
        for (int i = lane; i < n; i += 32) {
          const float x = gp[i], y = gq[i];
          a = fmaf(x, x, a); b = fmaf(y, y, b); c = fmaf(x, y, c);
        }
        a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
        const float rel = fabsf(c) * rsqrtf(a * b);
        if (rel > 1e-7f) {
          const float zeta = (b - a) / (2.f * c);
          const float t = copysignf(1.f, zeta) / (fabsf(zeta) + sqrtf(1.f + zeta * zeta));
          const float cs = rsqrtf(1.f + t * t), sn = cs * t;
          for (int i = lane; i < n; i += 32) {
            const float x = gp[i], y = gq[i];
            gp[i] = cs * x - sn * y;
            gq[i] = sn * x + cs * y;
          }
          if (lane == 0) atomicMax(&s_off, __float_as_uint(rel));
        }
      }
      __syncthreads();
    }
    if (__uint_as_float(s_off) < 2e-6f) break;
    __syncthreads();
  }
  // ---- eigenvalues = column norms - 1; rank them; write the k smallest ------------------------------------------------
  for (int j = warp; j < n; j += nwarps) {
    float a = 0.f;
    for (int i = lane; i < n; i += 32) a = fmaf(G[j * n + i], G[j * n + i], a);
    a = warp_sum(a);
    if (lane == 0) dg[j] = sqrtf(a);
  }
  __syncthreads();
  for (int j = warp; j < n; j += nwarps) {
    const float lj = dg[j];
    int rank = 0;
    for (int i = lane; i < n; i += 32) rank += (dg[i] < lj || (dg[i] == lj && i < j)) ? 1 : 0;
    rank = (int)warp_sum((float)rank);
    if (rank >= k.k) continue;
    // deterministic sign: the largest-magnitude entry is positive
    float best = 0.f;
    for (int i = lane; i < n; i += 32) { const float v = G[j * n + i]; if (fabsf(v) > fabsf(best)) best = v; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float other = __shfl_xor_sync(0xffffffffu, best, o);
      if (fabsf(other) > fabsf(best) || (fabsf(other) == fabsf(best) && other > best)) best = other;
    }
    float scale = (best < 0.f ? -1.f : 1.f) / lj;
    if (k.norm == 2) {                                      // 'walk': right eigenvectors of I - D^-1 A are D^-1/2 u, unit norm
      float nn = 0.f;
      for (int i = lane; i < n; i += 32) { const float v = G[j * n + i]; nn += v * v / deg[i]; }
      nn = warp_sum(nn);
      scale = (best < 0.f ? -1.f : 1.f) * rsqrtf(nn);
      for (int i = lane; i < n; i += 32) k.eig[(size_t)(n0 + i) * k.ld_eig + rank] = G[j * n + i] * rsqrtf(deg[i]) * scale;
    } else
    for (int i = lane; i < n; i += 32) k.eig[(size_t)(n0 + i) * k.ld_eig + rank] = G[j * n + i] * scale;
    if (lane == 0 && k.eigval) k.eigval[(size_t)g * k.k + rank] = lj - 1.f;
  }
  for (int idx = tid; idx < n * k.k; idx += blockDim.x) {   // graphs with fewer than k nodes: zero columns
    const int i = idx / k.k, r = idx - i * k.k;
    if (r >= n) k.eig[(size_t)(n0 + i) * k.ld_eig + r] = 0.f;
  }
}

// entry-wise random sign flip (the reference draws one uniform per ENTRY, not per column)
__global__ void __launch_bounds__(256) eig_flip_kernel(float* __restrict__ eig, long long n, unsigned long long seed,
                                                       unsigned long long step) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long x = (unsigned long long)i + 0x9E3779B97F4A7C15ull * (step + 1) + seed * 0xD1B54A32D192ED03ull;
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;   // splitmix64
  if (x & 0x8000000000000000ull) eig[i] = -eig[i];
}

}  // namespace dgn

using namespace dgn;

extern "C" int dgn_eig_precompute(int32_t n_graphs, const int32_t* node_off, const int32_t* in_ptr, const int32_t* in_src,
                                  int32_t max_nodes, int32_t norm, int32_t k, float* eig, int32_t ld_eig, float* eigval,
                                  void* stream) {
  if (n_graphs < 0 || !node_off || !in_ptr || !eig || k <= 0 || ld_eig < k || norm < 0 || norm > 2 || max_nodes < 0)
    return DGN_ERR_INVALID;
  if (n_graphs == 0 || max_nodes == 0) return DGN_OK;
  const size_t smem = ((size_t)max_nodes * max_nodes + 2 * (size_t)max_nodes) * sizeof(float);
  if (smem > 225 * 1024) return DGN_ERR_UNSUPPORTED;       // > ~238 nodes: does not fit one CTA's shared memory
  EigArgs a;
  a.n_graphs = n_graphs; a.norm = norm; a.k = k; a.ld_eig = ld_eig; a.max_sweeps = 40;
  a.node_off = node_off; a.in_ptr = in_ptr; a.in_src = in_src; a.eig = eig; a.eigval = eigval;
  static size_t attr_set = 0;
  if (smem > 48 * 1024 && attr_set < smem) {
    if (cudaFuncSetAttribute(eig_jacobi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024) != cudaSuccess) {
      g_dgn_last_cuda = cudaGetLastError();
      return DGN_ERR_CUDA;
    }
    attr_set = 225 * 1024;
  }
  int pairs = (max_nodes + 1) / 2;
  int threads = 32 * (pairs < 2 ? 2 : pairs);
  threads = threads > 1024 ? 1024 : threads;
  launch_pdl(eig_jacobi_kernel, dim3((unsigned)n_graphs), dim3((unsigned)threads), smem, (cudaStream_t)stream, a);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_dgn_last_cuda = e; return DGN_ERR_CUDA; }
  return DGN_OK;
}

extern "C" int dgn_eig_flip(float* eig, int64_t n_elems, uint64_t seed, uint64_t step, void* stream) {
  if (!eig || n_elems < 0) return DGN_ERR_INVALID;
  if (n_elems == 0) return DGN_OK;
  launch_pdl(eig_flip_kernel, dim3((unsigned)((n_elems + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, eig,
             (long long)n_elems, (unsigned long long)seed, (unsigned long long)step);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_dgn_last_cuda = e; return DGN_ERR_CUDA; }
  return DGN_OK;
}
