// Device-side collation (dgn_collate_device): a mini-batch is assembled in HBM from the dataset-resident per-graph
// fragments by ONE launch - no host collate(), no per-step H2D of the batch (the step's host input is the index list).
// Replaces dgl.batch + MoleculeDataset.collate (rb/data/molecules.py:219-230) and the degree bucketing set-up of
// DGL 0.4.2's update_all for the batch.
//
// grid = n_ids + kTailCtas CTAs of 256 threads.  CTA b < n_ids copies graph ids[b]: it first derives its node / edge /
// overflow-group base as the sum of the sizes of the graphs before it (b <= a few hundred terms, one block
// reduction), then streams the graph's node, edge and payload rows with the offsets added.  The tail CTAs write
// graph_ptr[B], meta, the CSR padding (empty in / out ranges for the padding nodes) and zero the padding rows of
// every payload, so a batch never sees leftovers of the previous one.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dgn_b200.h"
#include "dgn_launch.cuh"

extern thread_local cudaError_t g_dgn_last_cuda;

namespace dgn {

constexpr int kColThreads = 256, kTailCtas = 8;

struct CollateArgs {
  DgnDataset ds;
  DgnBatchOut out;
  const int32_t* ids;
  int n_ids;
};

__device__ __forceinline__ int block_sum3(int a, int b, int c, int& sb, int& sc) {
  __shared__ int red[3][kColThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { red[0][w] = a; red[1][w] = b; red[2][w] = c; }
  __syncthreads();
  int ra = 0, rb = 0, rc = 0;
  for (int i = 0; i < kColThreads / 32; ++i) { ra += red[0][i]; rb += red[1][i]; rc += red[2][i]; }
  __syncthreads();
  sb = rb; sc = rc;
  return ra;
}

// sizes of the first `upto` selected graphs: (nodes, edges, overflow groups)
__device__ __forceinline__ void prefix_sizes(const CollateArgs& k, int upto, int& nb, int& eb, int& ob) {
  int a = 0, b = 0, c = 0;
  for (int i = threadIdx.x; i < upto; i += kColThreads) {
    const int g = __ldg(k.ids + i);
    a += __ldg(k.ds.node_off + g + 1) - __ldg(k.ds.node_off + g);
    b += __ldg(k.ds.edge_off + g + 1) - __ldg(k.ds.edge_off + g);
    c += __ldg(k.ds.ovf_off + g + 1) - __ldg(k.ds.ovf_off + g);
  }
  nb = block_sum3(a, b, c, eb, ob);
}

__global__ void __launch_bounds__(kColThreads) collate_kernel(const __grid_constant__ CollateArgs k) {
  pdl_prologue();
  const int t = threadIdx.x;
  const DgnBatchOut& o = k.out;
  if ((int)blockIdx.x < k.n_ids) {
    const int b = blockIdx.x, g = __ldg(k.ids + b);
    int nb, eb, ob;
    prefix_sizes(k, b, nb, eb, ob);
    const int n0 = __ldg(k.ds.node_off + g), n = __ldg(k.ds.node_off + g + 1) - n0;
    const int e0 = __ldg(k.ds.edge_off + g), e = __ldg(k.ds.edge_off + g + 1) - e0;
    const int o0 = __ldg(k.ds.ovf_off + g);
    if (nb + n > o.n_cap || eb + e > o.e_cap) {            // does not fit: skip the graph, flag the batch
      if (t == 0) o.meta[3] = 1;
      return;
    }
    if (t == 0) o.graph_ptr[b] = nb;
    const float sn = sqrtf(1.0f / (float)n);                // collate(): snorm_n = sqrt(1 / n_g), rb/data/molecules.py:222-224
    for (int v = t; v < n; v += kColThreads) {
      o.in_ptr[nb + v] = __ldg(k.ds.in_ptr + n0 + v) - e0 + eb;
      o.out_ptr[nb + v] = __ldg(k.ds.out_ptr + n0 + v) - e0 + eb;
      o.ovf_ptr[nb + v] = __ldg(k.ds.ovf_ptr + n0 + v) - o0 + ob;
      o.log_deg[nb + v] = __ldg(k.ds.log_deg + n0 + v);
      o.snorm_n[nb + v] = sn;
    }
    for (int s = t; s < e; s += kColThreads) {
      o.in_src[eb + s] = __ldg(k.ds.in_src + e0 + s) - n0 + nb;
      o.in_eid[eb + s] = __ldg(k.ds.in_eid + e0 + s) - e0 + eb;
      o.out_slot[eb + s] = __ldg(k.ds.out_slot + e0 + s) - e0 + eb;
      o.src[eb + s] = __ldg(k.ds.src + e0 + s) - n0 + nb;
      o.dst[eb + s] = __ldg(k.ds.dst + e0 + s) - n0 + nb;
    }
    for (int p = 0; p < o.n_payloads; ++p) {
      const DgnPayload& pl = o.payload[p];
      const int words = pl.row_bytes >> 2;
      const int rows = pl.per == 0 ? n : (pl.per == 1 ? e : 1);
      const long long s0 = (long long)(pl.per == 0 ? n0 : (pl.per == 1 ? e0 : g)) * words;
      const long long d0 = (long long)(pl.per == 0 ? nb : (pl.per == 1 ? eb : b)) * words;
      const uint32_t* sp = reinterpret_cast<const uint32_t*>(pl.src) + s0;
      uint32_t* dp = reinterpret_cast<uint32_t*>(pl.dst) + d0;
      for (int i = t; i < rows * words; i += kColThreads) dp[i] = __ldg(sp + i);
    }
    return;
  }
  // ---- tail: totals, CSR padding, zero padding rows ------------------------------------------------------------
  const int tail = blockIdx.x - k.n_ids, stride = kTailCtas * kColThreads, tt = tail * kColThreads + t;
  int nr, er, orr;
  prefix_sizes(k, k.n_ids, nr, er, orr);
  if (nr > o.n_cap || er > o.e_cap) {                      // overflowing batch (flagged above): clamp the totals
    nr = nr > o.n_cap ? o.n_cap : nr;
    er = er > o.e_cap ? o.e_cap : er;
  }
  if (tail == 0 && t == 0) {
    o.graph_ptr[k.n_ids] = nr;
    o.meta[0] = nr; o.meta[1] = er; o.meta[2] = k.n_ids;
  }
  if (tail == 0)
    for (int b = k.n_ids + 1 + t; b <= o.b_cap; b += kColThreads) o.graph_ptr[b] = nr;
  for (int v = nr + tt; v <= o.n_cap; v += stride) {       // padding nodes: empty in / out ranges, no overflow groups
    o.in_ptr[v] = er;
    o.out_ptr[v] = er;
    o.ovf_ptr[v] = orr;
    if (v < o.n_cap) { o.log_deg[v] = 0.f; o.snorm_n[v] = 0.f; }
  }
  for (int s = er + tt; s < o.e_cap; s += stride) {
    o.in_src[s] = 0; o.in_eid[s] = 0; o.out_slot[s] = 0; o.src[s] = 0; o.dst[s] = 0;
  }
  for (int p = 0; p < o.n_payloads; ++p) {
    const DgnPayload& pl = o.payload[p];
    const int words = pl.row_bytes >> 2;
    const long long from = (long long)(pl.per == 0 ? nr : (pl.per == 1 ? er : k.n_ids)) * words;
    const long long to = (long long)(pl.per == 0 ? o.n_cap : (pl.per == 1 ? o.e_cap : o.b_cap)) * words;
    uint32_t* dp = reinterpret_cast<uint32_t*>(pl.dst);
    for (long long i = from + tt; i < to; i += stride) dp[i] = 0u;
  }
}

}  // namespace dgn

using namespace dgn;

extern "C" int dgn_collate_device(const DgnDataset* ds, const int32_t* ids, int32_t n_ids, const DgnBatchOut* out,
                                  void* stream) {
  if (!ds || !ids || !out || n_ids < 0 || out->n_payloads < 0 || out->n_payloads > DGN_MAX_PAYLOADS) return DGN_ERR_INVALID;
  if (!ds->node_off || !ds->edge_off || !ds->ovf_off || !ds->in_ptr || !ds->out_ptr || !ds->log_deg || !ds->ovf_ptr)
    return DGN_ERR_INVALID;
  if (!out->in_ptr || !out->in_src || !out->in_eid || !out->out_ptr || !out->out_slot || !out->src || !out->dst ||
      !out->graph_ptr || !out->ovf_ptr || !out->meta || !out->log_deg || !out->snorm_n)
    return DGN_ERR_INVALID;
  if (n_ids > out->b_cap) return DGN_ERR_INVALID;
  for (int p = 0; p < out->n_payloads; ++p)
    if (!out->payload[p].src || !out->payload[p].dst || out->payload[p].row_bytes <= 0 || out->payload[p].row_bytes % 4 ||
        out->payload[p].per < 0 || out->payload[p].per > 2)
      return DGN_ERR_INVALID;
  CollateArgs k;
  k.ds = *ds; k.out = *out; k.ids = ids; k.n_ids = n_ids;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(out->meta, 0, 4 * sizeof(int32_t), st);
  if (e == cudaSuccess) {
    launch_pdl(collate_kernel, dim3((unsigned)(n_ids + kTailCtas)), dim3(kColThreads), 0, st, k);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) { g_dgn_last_cuda = e; return DGN_ERR_CUDA; }
  return DGN_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Block scatter / gather between the per-tower parameters of a DGNLayerTower (rb/nets/dgn_layer.py:279-307) and the dense
// block-structured operands the single-launch tower path runs on: one launch moves up to a few hundred rectangular
// segments described by a device table (built once per layer).
// ---------------------------------------------------------------------------------------------------------------
namespace dgn {

struct Seg {                       // 32 bytes
  float* a;                        // tower-side tensor element (0, 0) of the segment
  float* b;                        // dense-side tensor element (0, 0) of the segment
  int32_t rows, cols, ld_a, ld_b;
};

// dir 0: b = a (pack); 1: a = b (unpack, overwrite); 2: a += b (unpack, accumulate)
__global__ void __launch_bounds__(256) seg_copy_kernel(const Seg* __restrict__ segs, int dir) {
  pdl_prologue();
  const Seg s = segs[blockIdx.x];
  const int n = s.rows * s.cols;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int r = i / s.cols, c = i - r * s.cols;
    float* pa = s.a + (size_t)r * s.ld_a + c;
    float* pb = s.b + (size_t)r * s.ld_b + c;
    if (dir == 0) *pb = *pa;
    else if (dir == 1) *pa = *pb;
    else *pa += *pb;
  }
}

}  // namespace dgn

extern "C" int dgn_segment_copy(const void* seg_table, int32_t n_segments, int32_t direction, void* stream) {
  if (!seg_table || n_segments < 0 || direction < 0 || direction > 2) return DGN_ERR_INVALID;
  if (n_segments == 0) return DGN_OK;
  launch_pdl(dgn::seg_copy_kernel, dim3((unsigned)n_segments), dim3(256), 0, (cudaStream_t)stream,
             reinterpret_cast<const dgn::Seg*>(seg_table), (int)direction);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_dgn_last_cuda = e; return DGN_ERR_CUDA; }
  return DGN_OK;
}
