// Memory-path microbenchmark (B200): what the store / load instruction variants reach on a buffer >> L2.
// Answers "what is the ceiling of a write-dominated kernel" for the aggregation forward (profiles/README.md).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o membw tools/membw.cu && ./membw
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__global__ void fill_v4(float4* p, size_t n) {
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void fill_v4_cs(float4* p, size_t n) {
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) __stcs(p + i, v);
}
__global__ void fill_v4_wt(float4* p, size_t n) {
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) __stwt(p + i, v);
}
// 256-bit stores (sm_100: st.global.v8.f32)
__global__ void fill_v8(float4* p, size_t n) {
  const size_t n8 = n / 2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    asm volatile("st.global.v8.f32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p + 2 * i), "f"(1.f) : "memory");
  }
}
// one-shot: every thread writes exactly UN float4 (no grid-stride loop), like the aggregation epilogue
template <int UN>
__global__ void fill_oneshot(float4* p, size_t n) {
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  size_t base = ((size_t)blockIdx.x * blockDim.x) * UN + threadIdx.x;
#pragma unroll
  for (int j = 0; j < UN; ++j) {
    const size_t i = base + (size_t)j * blockDim.x;
    if (i < n) __stcs(p + i, v);
  }
}
// TMA bulk store: each CTA fills a CHUNK-byte shared buffer once and streams it out repeatedly
template <int CHUNK>
__global__ void fill_tma(float4* p, size_t n) {
  extern __shared__ __align__(128) unsigned char sm[];
  float4* s = reinterpret_cast<float4*>(sm);
  for (int i = threadIdx.x; i < CHUNK / 16; i += blockDim.x) s[i] = make_float4(1.f, 2.f, 3.f, 4.f);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    const size_t chunks = n * 16 / CHUNK;
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(s);
    for (size_t c = blockIdx.x; c < chunks; c += gridDim.x) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"((char*)p + c * CHUNK), "r"(sa), "r"(CHUNK) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}
__global__ void read_v4(const float4* p, size_t n, float* out) {
  float acc = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = __ldcs(p + i);
    acc += v.x + v.y + v.z + v.w;
  }
  if (acc == 12345.678f) *out = acc;
}
__global__ void copy_v4(const float4* a, float4* b, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) __stcs(b + i, __ldcs(a + i));
}
// read 3 slabs -> write 1 (the backward fold pattern), and read 1 -> write 3 (the forward scaler pattern)
__global__ void r3w1(const float4* a, float4* b, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float4 x = __ldcs(a + i), y = __ldcs(a + n + i), z = __ldcs(a + 2 * n + i);
    __stcs(b + i, make_float4(x.x + y.x + z.x, x.y + y.y + z.y, x.z + y.z + z.z, x.w + y.w + z.w));
  }
}


// ---- the aggregation forward's store pattern: thread = (node, float4 chunk of F=64), S*A = 30 slabs of 256 B per node row
// (row pitch 31 * 256 B as in the layer's cat buffer).  ORDER 0: a-major / scaler inner (as the kernel), 1: sequential slabs.
template <int ORDER, bool CS>
__global__ void rowpat(float4* out, int N) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int v = t >> 4, c = t & 15;
  if (v >= N) return;
  float4* row = out + (size_t)v * (31 * 16) + 16 + c;
  const float4 y = make_float4((float)v, (float)c, 1.f, 2.f);
#pragma unroll
  for (int a = 0; a < 10; ++a) {
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int slab = ORDER == 0 ? (s * 10 + a) : (a * 3 + s);
      if (CS) __stcs(row + slab * 16, y); else row[slab * 16] = y;
    }
  }
}
// same bytes, but a warp writes ONE node row as 15 fully contiguous 512 B stores
__global__ void rowpat_warp(float4* out, int N) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= N) return;
  float4* row = out + (size_t)w * (31 * 16) + 16 + lane;
  const float4 y = make_float4((float)w, (float)lane, 1.f, 2.f);
#pragma unroll
  for (int j = 0; j < 15; ++j) __stcs(row + j * 32, y);
}
// CTA tile of 16 node rows staged in shared memory, written with one TMA bulk store per row (7680 B)
__global__ void rowpat_tma(float4* out, int N) {
  extern __shared__ __align__(128) unsigned char sm[];
  float4* s = reinterpret_cast<float4*>(sm);
  const int v0 = blockIdx.x * 16;
  const int ln = threadIdx.x >> 4, c = threadIdx.x & 15;
  const float4 y = make_float4((float)(v0 + ln), (float)c, 1.f, 2.f);
#pragma unroll
  for (int j = 0; j < 30; ++j) s[ln * 480 + j * 16 + c] = y;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 16 && v0 + threadIdx.x < N) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(s + threadIdx.x * 480);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + (size_t)(v0 + threadIdx.x) * (31 * 16) + 16), "r"(sa), "r"(7680) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

template <typename F>
static float time_ms(F f, int reps) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(); f();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) f();
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  CK(cudaGetLastError());
  return ms / reps;
}

int main() {
  const size_t bytes = (size_t)1 << 30;         // 1 GiB per buffer
  const size_t n = bytes / 16;
  float4 *A, *B; float* out;
  CK(cudaMalloc(&A, 3 * bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&out, 4));
  CK(cudaMemset(A, 0, 3 * bytes)); CK(cudaMemset(B, 0, bytes));
  const int reps = 10;
  auto rep = [&](const char* name, double moved, float ms) { printf("%-44s %8.3f ms  %8.1f GB/s\n", name, ms, moved / ms * 1e-6); };
  for (int ctas_per_sm : {2, 4, 8, 16}) {
    const int grid = 148 * ctas_per_sm, blk = 256;
    printf("--- grid = 148 x %d, block %d\n", ctas_per_sm, blk);
    rep("fill st.v4", bytes, time_ms([&] { fill_v4<<<grid, blk>>>(B, n); }, reps));
    rep("fill st.cs.v4", bytes, time_ms([&] { fill_v4_cs<<<grid, blk>>>(B, n); }, reps));
    rep("fill st.wt.v4", bytes, time_ms([&] { fill_v4_wt<<<grid, blk>>>(B, n); }, reps));
    rep("fill st.v8 (256-bit)", bytes, time_ms([&] { fill_v8<<<grid, blk>>>(B, n); }, reps));
    rep("read ld.cs.v4", bytes, time_ms([&] { read_v4<<<grid, blk>>>(A, n, out); }, reps));
    rep("copy ld.cs/st.cs v4 (r+w bytes)", 2.0 * bytes, time_ms([&] { copy_v4<<<grid, blk>>>(A, B, n); }, reps));
    rep("read 3 slabs, write 1 (r+w bytes)", 4.0 * bytes, time_ms([&] { r3w1<<<grid, blk>>>(A, B, n); }, reps));
  }
  printf("--- one-shot grids (no grid-stride loop)\n");
  rep("fill one-shot, 8 x st.cs.v4 / thread", bytes, time_ms([&] { fill_oneshot<8><<<(unsigned)((n + 256 * 8 - 1) / (256 * 8)), 256>>>(B, n); }, reps));
  rep("fill one-shot, 32 x st.cs.v4 / thread", bytes, time_ms([&] { fill_oneshot<32><<<(unsigned)((n + 256 * 32 - 1) / (256 * 32)), 256>>>(B, n); }, reps));
  printf("--- TMA bulk stores from shared memory (one issuing thread per CTA)\n");
  CK(cudaFuncSetAttribute(fill_tma<65536>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  for (int ctas_per_sm : {1, 2, 3}) {
    const int grid = 148 * ctas_per_sm;
    char nm[64];
    snprintf(nm, sizeof nm, "fill TMA 16 KB chunks, %d CTA/SM", ctas_per_sm);
    rep(nm, bytes, time_ms([&] { fill_tma<16384><<<grid, 128, 16384>>>(B, n); }, reps));
    snprintf(nm, sizeof nm, "fill TMA 64 KB chunks, %d CTA/SM", ctas_per_sm);
    rep(nm, bytes, time_ms([&] { fill_tma<65536><<<grid, 128, 65536>>>(B, n); }, reps));
  }
  rep("cudaMemsetAsync", bytes, time_ms([&] { CK(cudaMemsetAsync(B, 1, bytes)); }, reps));
  rep("cudaMemcpyAsync D2D (r+w bytes)", 2.0 * bytes, time_ms([&] { CK(cudaMemcpyAsync(B, A, bytes, cudaMemcpyDeviceToDevice)); }, reps));
  {
    printf("--- aggregation forward store pattern (rows of 31 x 256 B, 30 slabs written)\n");
    const int N = (int)(bytes / (31 * 256));
    const double moved = (double)N * 30 * 256;
    rep("rowpat a-major/s-inner st.cs (kernel order)", moved, time_ms([&] { rowpat<0, true><<<(N * 16 + 255) / 256, 256>>>(B, N); }, reps));
    rep("rowpat a-major/s-inner st", moved, time_ms([&] { rowpat<0, false><<<(N * 16 + 255) / 256, 256>>>(B, N); }, reps));
    rep("rowpat sequential slabs st.cs", moved, time_ms([&] { rowpat<1, true><<<(N * 16 + 255) / 256, 256>>>(B, N); }, reps));
    rep("rowpat sequential slabs st", moved, time_ms([&] { rowpat<1, false><<<(N * 16 + 255) / 256, 256>>>(B, N); }, reps));
    rep("rowpat warp-per-row 512 B stores", moved, time_ms([&] { rowpat_warp<<<(N * 32 + 255) / 256, 256>>>(B, N); }, reps));
    CK(cudaFuncSetAttribute(rowpat_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 7680));
    rep("rowpat smem tile + TMA bulk store per row", moved, time_ms([&] { rowpat_tma<<<(N + 15) / 16, 256, 16 * 7680>>>(B, N); }, reps));
  }
  return 0;
}
