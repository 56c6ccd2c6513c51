// fp32-accurate GEMM on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a.
//
// The DGN layer's pre/post-transform MLPs are small dense GEMMs with awkward shapes: tall-skinny
// [N_nodes x 1984] x [1984 x 64], its transpose products for the backward, and K = N_nodes products with a
// 64 x 64 output.  Parity with the reference needs fp32 accuracy (1e-5), which rules out plain TF32.  This
// kernel computes C (+)= A * B with the 3xTF32 split
//     a = a_hi + a_lo,  a_hi = a rounded to TF32,  a_lo = a - a_hi (exact in fp32, then rounded to TF32)
//     a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo                         (error ~2^-21 per product)
// on `tcgen05.mma.kind::tf32` with the fp32 accumulator in tensor memory.
//
// Structure of one CTA (output tile 128 x 64, K consumed in blocks of 32 floats, optional split-K):
//   warps 0-3  loaders: global -> registers -> {hi, lo} split -> shared memory in the UMMA canonical
//              SWIZZLE_128B layout (K-major or MN-major, so all four transpose combinations are served without
//              transposed copies), `fence.proxy.async`, arrive on the stage's `full` mbarrier; afterwards the
//              epilogue: tcgen05.ld the accumulator (one row per thread), store / accumulate / split-K reduce;
//   warp 4     MMA issuer: one elected lane waits on `full`, issues 4 k-steps x 3 tcgen05.mma per stage and
//              hands the stage back with tcgen05.commit -> `empty` mbarrier; the last commit signals the epilogue.
// Split-K partial tiles go to a workspace and a small second kernel adds them in split order, so the result does
// not depend on scheduling.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dgn_b200.h"
#include "dgn_launch.cuh"
#include "dgn_umma.cuh"                            // mbarrier / tcgen05 / descriptor wrappers, 3xTF32 split (shared with dgn_post.cu)

namespace dgn {
using namespace umma;

constexpr int BM = 128, BN = 64;                   // output tile; BK = 32 floats of K (one 128 B swizzle row) from dgn_umma.cuh
constexpr int kStages = 2;                         // 96 KB per CTA -> two CTAs per SM hide each other's load latency
constexpr int kLoaderThreads = 128;
constexpr int kGemmThreads = 160;                   // 4 loader/epilogue warps + 1 MMA warp
constexpr int A_TILE = BM * BK * 4, B_TILE = BN * BK * 4;                 // bytes
constexpr int STAGE_BYTES = 2 * A_TILE + 2 * B_TILE;                      // hi + lo of both operands
constexpr int GEMM_SMEM = kStages * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

// instruction descriptor for kind::tf32, fp32 accumulate, M x N = 128 x 64
__host__ __device__ constexpr uint32_t instr_desc(bool a_mn_major, bool b_mn_major) {
  return instr_desc_tf32(BM, BN, a_mn_major, b_mn_major);
}

struct GemmArgs {
  int M, N, K;
  const float* A; int lda;
  const float* B; int ldb;
  float* C; int ldc;
  int accumulate, c_transposed, splits, kb_per_split;
  float* ws; unsigned* counters;
};

// One operand tile: ROWS (M or N extent) x BK, K-major source [ROWS][K] or MN-major source [K][ROWS].
// fetch_tile issues the global loads into registers (the next k-block is fetched before the current one is
// converted and stored, so its latency overlaps that work); store_tile does the split + swizzled stores.
template <int ROWS, bool KMAJOR>
__device__ __forceinline__ void fetch_tile(const float* __restrict__ src, int ld, int r0, int k0, int r_max, int k_max,
                                           float4 (&v)[ROWS * BK / 4 / kLoaderThreads], int t) {
  constexpr int CHUNKS = ROWS * BK / 4;
#pragma unroll
  for (int i = 0; i < CHUNKS / kLoaderThreads; ++i) {
    const int id = t + i * kLoaderThreads;
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (KMAJOR) {
      const int row = id >> 3, ch = id & 7;
      const int gr = r0 + row, gk = k0 + ch * 4;
      if (gr < r_max && gk < k_max) v[i] = __ldg(reinterpret_cast<const float4*>(src + (size_t)gr * ld + gk));
    } else {
      constexpr int CPR = ROWS / 4;
      const int krow = id / CPR, ch = id % CPR;
      const int gk = k0 + krow, gr = r0 + ch * 4;
      if (gk < k_max && gr < r_max) v[i] = __ldg(reinterpret_cast<const float4*>(src + (size_t)gk * ld + gr));
    }
  }
}

template <int ROWS, bool KMAJOR>
__device__ __forceinline__ void store_tile(const float4 (&vv)[ROWS * BK / 4 / kLoaderThreads], unsigned char* hi_tile,
                                           unsigned char* lo_tile, int t) {
  constexpr int CHUNKS = ROWS * BK / 4;
#pragma unroll
  for (int i = 0; i < CHUNKS / kLoaderThreads; ++i) {
    const int id = t + i * kLoaderThreads;
    const float4 v = vv[i];
    uint32_t off;
    if constexpr (KMAJOR) {
      const int row = id >> 3, ch = id & 7;                      // 8 chunks (32 floats of K) per row
      off = (uint32_t)(row >> 3) * 1024u + sw128(row & 7, ch);
    } else {
      constexpr int CPR = ROWS / 4;                              // 16 B chunks per K-row
      const int krow = id / CPR, ch = id % CPR;
      // SW128_32B atoms: 4 K-rows x 32 MN-floats (512 B); atoms contiguous along MN, then along K
      const uint32_t kl = krow & 3, c16 = ch & 7;
      off = (uint32_t)(krow >> 2) * (uint32_t)(ROWS / 32) * 512u + (uint32_t)(ch >> 3) * 512u + kl * 128u +
            ((((c16 >> 1) ^ kl)) << 5) + ((c16 & 1u) << 4);
    }
    split_store(hi_tile, lo_tile, off, v);
  }
}

template <bool A_K, bool B_K>
__global__ void __launch_bounds__(kGemmThreads, 2) gemm_tf32x3_kernel(const GemmArgs g) {
  pdl_prologue();
  extern __shared__ unsigned char raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * STAGE_BYTES);
  uint64_t* empty = full + kStages;
  uint64_t* accum_full = empty + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN, split = blockIdx.z;
  const int kb_total = (g.K + BK - 1) / BK;
  const int kb0 = split * g.kb_per_split;
  const int kb1 = min(kb0 + g.kb_per_split, kb_total);
  const int nkb = kb1 - kb0;                                       // >= 1 by construction of the grid

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mb_init(&full[s], kLoaderThreads); mb_init(&empty[s], 1); }
    mb_init(accum_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {                                                 // TMEM: 64 fp32 accumulator columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "n"(BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;

  if (warp < 4) {
    // ------------------------------ loaders ------------------------------
    float4 va[BM * BK / 4 / kLoaderThreads], vb[BN * BK / 4 / kLoaderThreads];
    fetch_tile<BM, A_K>(g.A, g.lda, m0, kb0 * BK, g.M, g.K, va, tid);
    fetch_tile<BN, B_K>(g.B, g.ldb, n0, kb0 * BK, g.N, g.K, vb, tid);
    for (int i = 0; i < nkb; ++i) {
      const int s = i % kStages;
      float4 na[BM * BK / 4 / kLoaderThreads], nb_[BN * BK / 4 / kLoaderThreads];
      if (i + 1 < nkb) {                                             // next k-block in flight during this one's stores
        fetch_tile<BM, A_K>(g.A, g.lda, m0, (kb0 + i + 1) * BK, g.M, g.K, na, tid);
        fetch_tile<BN, B_K>(g.B, g.ldb, n0, (kb0 + i + 1) * BK, g.N, g.K, nb_, tid);
      }
      if (i >= kStages) mb_wait(&empty[s], ((i / kStages) - 1) & 1);
      unsigned char* st = smem + s * STAGE_BYTES;
      store_tile<BM, A_K>(va, st, st + A_TILE, tid);
      store_tile<BN, B_K>(vb, st + 2 * A_TILE, st + 2 * A_TILE + B_TILE, tid);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
      mb_arrive(&full[s]);
      if (i + 1 < nkb) {
#pragma unroll
        for (int q = 0; q < BM * BK / 4 / kLoaderThreads; ++q) va[q] = na[q];
#pragma unroll
        for (int q = 0; q < BN * BK / 4 / kLoaderThreads; ++q) vb[q] = nb_[q];
      }
    }
  } else if (lane == 0) {
    // ------------------------------ MMA issuer ------------------------------
    constexpr uint32_t idesc = instr_desc(!A_K, !B_K);
    for (int i = 0; i < nkb; ++i) {
      const int s = i % kStages;
      mb_wait(&full[s], (i / kStages) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_hi = s32(smem + s * STAGE_BYTES), a_lo = a_hi + A_TILE;
      const uint32_t b_hi = a_hi + 2 * A_TILE, b_lo = b_hi + B_TILE;
#pragma unroll
      for (int kk = 0; kk < BK / 8; ++kk) {                          // UMMA_K = 8 for tf32
        uint64_t da_hi, da_lo, db_hi, db_lo;
        if constexpr (A_K) {                                         // K-major: advance 32 B inside the swizzled row
          da_hi = smem_desc(a_hi + kk * 32, 0, 1024, kLayoutSW128);
          da_lo = smem_desc(a_lo + kk * 32, 0, 1024, kLayoutSW128);
        } else {                                                     // MN-major: two 4-row K-atoms per k-step
          constexpr uint32_t KA = (BM / 32) * 512;                   // stride between K-atoms
          da_hi = smem_desc(a_hi + kk * 2 * KA, 512, KA, kLayoutSW128Base32);
          da_lo = smem_desc(a_lo + kk * 2 * KA, 512, KA, kLayoutSW128Base32);
        }
        if constexpr (B_K) {
          db_hi = smem_desc(b_hi + kk * 32, 0, 1024, kLayoutSW128);
          db_lo = smem_desc(b_lo + kk * 32, 0, 1024, kLayoutSW128);
        } else {
          constexpr uint32_t KB = (BN / 32) * 512;
          db_hi = smem_desc(b_hi + kk * 2 * KB, 512, KB, kLayoutSW128Base32);
          db_lo = smem_desc(b_lo + kk * 2 * KB, 512, KB, kLayoutSW128Base32);
        }
        const uint32_t acc = (i > 0 || kk > 0) ? 1u : 0u;
        umma_tf32(tmem_d, da_lo, db_hi, idesc, acc);                 // small terms first
        umma_tf32(tmem_d, da_hi, db_lo, idesc, 1u);
        umma_tf32(tmem_d, da_hi, db_hi, idesc, 1u);
      }
      umma_commit(&empty[s]);                                        // stage free once these MMAs retire
    }
    umma_commit(accum_full);
  }

  // ------------------------------ epilogue (warps 0-3) ------------------------------
  uint32_t r[BN];
  if (warp < 4) {
    mb_wait(accum_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]),
          "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),
          "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]),
          "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]),
          "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");

    const int row = warp * 32 + lane, gm = m0 + row;
    const int tile_id = blockIdx.y * gridDim.x + blockIdx.x;
    if (g.splits > 1) {                                              // partial tile -> workspace; reduced by splitk_reduce_kernel
      float* part = g.ws + ((size_t)split * gridDim.x * gridDim.y + tile_id) * (BM * BN) + (size_t)row * BN;
#pragma unroll
      for (int j = 0; j < BN; j += 4)
        *reinterpret_cast<float4*>(part + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                           __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
    }
    if (g.splits == 1) {
      if (gm < g.M) {
        if (!g.c_transposed) {
          // handled below (staged through shared memory so that the stores are coalesced)
        } else {
#pragma unroll
          for (int j = 0; j < BN; ++j) {
            if (n0 + j < g.N) {
              float* c = g.C + (size_t)(n0 + j) * g.ldc + gm;
              const float v = __uint_as_float(r[j]);
              *c = g.accumulate ? *c + v : v;
            }
          }
        }
      }
    }
  }
  if (warp < 4 && !g.c_transposed && g.splits == 1) {
    // every warp transposes its 32 x 64 block through its own shared-memory patch (the pipeline stages are
    // idle by now): thread-per-row registers -> lane-per-column stores, 128 B per instruction
    float* patch = reinterpret_cast<float*>(smem) + warp * (32 * (BN + 1));
#pragma unroll
    for (int j = 0; j < BN; ++j) patch[lane * (BN + 1) + j] = __uint_as_float(r[j]);
    __syncwarp();
    const bool pair_ok = (g.ldc % 2 == 0) && (n0 + BN <= g.N) && ((reinterpret_cast<uintptr_t>(g.C) & 7u) == 0);
    for (int rr = 0; rr < 32; ++rr) {
      const int gmr = m0 + warp * 32 + rr;
      if (gmr >= g.M) break;
      float* crow = g.C + (size_t)gmr * g.ldc + n0;
      if (pair_ok) {                                                   // lane writes columns 2*lane, 2*lane+1: 256 B per warp
        float2 v = make_float2(patch[rr * (BN + 1) + 2 * lane], patch[rr * (BN + 1) + 2 * lane + 1]);
        float2* dst = reinterpret_cast<float2*>(crow + 2 * lane);
        if (g.accumulate) { const float2 o = *dst; v.x += o.x; v.y += o.y; }
        *dst = v;
      } else {
#pragma unroll
        for (int h = 0; h < BN / 32; ++h) {
          const int j = lane + 32 * h;
          if (n0 + j < g.N) {
            const float v = patch[rr * (BN + 1) + j];
            crow[j] = g.accumulate ? crow[j] + v : v;
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(BN) : "memory");
  }
}

// Sums the split-K partial tiles in split order (deterministic) and writes / accumulates C.
// One thread per (row, 4-column chunk) of the M x N result.
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const GemmArgs g, int m_tiles, int n_tiles) {
  pdl_prologue();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int n4 = (g.N + 3) / 4;
  if (idx >= (long long)g.M * n4) return;
  const int m = (int)(idx / n4), n = (int)(idx - (long long)m * n4) * 4;
  const int tm = m / BM, tn = n / BN, tile_id = tn * m_tiles + tm;
  const size_t in_tile = (size_t)(m - tm * BM) * BN + (n - tn * BN);
  const size_t split_stride = (size_t)m_tiles * n_tiles * (BM * BN);
  const float* p = g.ws + (size_t)tile_id * (BM * BN) + in_tile;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 v[16];                                                      // splits <= 16: every load in flight at once
#pragma unroll
  for (int sp = 0; sp < 16; ++sp)
    v[sp] = (sp < g.splits) ? __ldcs(reinterpret_cast<const float4*>(p + (size_t)sp * split_stride))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int sp = 0; sp < 16; ++sp) { acc.x += v[sp].x; acc.y += v[sp].y; acc.z += v[sp].z; acc.w += v[sp].w; }
  const float a[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (n + j < g.N) {
      float* c = g.c_transposed ? g.C + (size_t)(n + j) * g.ldc + m : g.C + (size_t)m * g.ldc + n + j;
      *c = g.accumulate ? *c + a[j] : a[j];
    }
  }
}

}  // namespace dgn

using namespace dgn;

extern thread_local cudaError_t g_dgn_last_cuda;

// C[M,N] (+)= op(A) * op(B), fp32 accuracy on the tensor cores.
//   a_kmajor = 1: A is [M][K] (row stride lda);  0: A is [K][M]
//   b_kmajor = 1: B is [N][K] (row stride ldb);  0: B is [K][N]
//   c_transposed = 1: the result is stored as C[N][M] (row stride ldc)
// ws: workspace of dgn_gemm_ws_floats() floats, ZERO before the first call (the kernel re-zeroes its counters).
// Split-K is only used when the output has fewer than 148 tiles, and splits * tiles stays below 448: the
// workspace is 256 counters followed by 448 partial tiles, independent of the problem shape.
constexpr int kWsCounters = 256, kWsTiles = 448;
extern "C" int64_t dgn_gemm_ws_floats(void) { return kWsCounters + (int64_t)kWsTiles * (BM * BN); }

static int pick_splits(int mt, int nt, int kb) {
  const int tiles = mt * nt;
  if (tiles >= 148 || kb <= 2) return 1;
  int s = (2 * 148) / tiles;                                          // one wave at 2 CTAs per SM
  if (s > 16) s = 16;
  if (s > kb / 2) s = kb / 2;                                         // at least 2 k-blocks per split
  return s < 1 ? 1 : s;
}

extern "C" int dgn_gemm_tf32x3(int32_t M, int32_t N, int32_t K, const float* A, int32_t lda, int32_t a_kmajor,
                               const float* B, int32_t ldb, int32_t b_kmajor, float* C, int32_t ldc,
                               int32_t accumulate, int32_t c_transposed, float* ws, void* stream) {
  if (M < 0 || N < 0 || K < 0 || !A || !B || !C || !ws) return DGN_ERR_INVALID;
  if (M == 0 || N == 0) return DGN_OK;
  if (K == 0) return DGN_ERR_UNSUPPORTED;
  // 128-bit loads: every row start must be 16 B aligned and the contiguous extent a multiple of 4
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  if (!al(A) || !al(B) || lda % 4 || ldb % 4) return DGN_ERR_UNSUPPORTED;
  if ((a_kmajor ? K : M) % 4 || (b_kmajor ? K : N) % 4) return DGN_ERR_UNSUPPORTED;
  GemmArgs g;
  g.M = M; g.N = N; g.K = K; g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc;
  g.accumulate = accumulate; g.c_transposed = c_transposed;
  const int mt = (M + BM - 1) / BM, nt = (N + BN - 1) / BN, kb = (K + BK - 1) / BK;
  int splits = pick_splits(mt, nt, kb);
  g.kb_per_split = (kb + splits - 1) / splits;
  splits = (kb + g.kb_per_split - 1) / g.kb_per_split;                // no empty split
  g.splits = splits;
  if (splits > 1 && (int64_t)splits * mt * nt > kWsTiles) return DGN_ERR_UNSUPPORTED;
  g.ws = ws + kWsCounters;
  g.counters = reinterpret_cast<unsigned*>(ws);
  const dim3 grid(mt, nt, splits);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaSuccess;
#define LAUNCH(AK, BK_)                                                                                          \
  do {                                                                                                           \
    static bool attr = false;                                                                                    \
    if (!attr) { e = cudaFuncSetAttribute(gemm_tf32x3_kernel<AK, BK_>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM); attr = true; } \
    if (e == cudaSuccess) launch_pdl(gemm_tf32x3_kernel<AK, BK_>, grid, dim3(kGemmThreads), GEMM_SMEM, st, g);                  \
  } while (0)
  if (a_kmajor && b_kmajor) LAUNCH(true, true);
  else if (a_kmajor && !b_kmajor) LAUNCH(true, false);
  else if (!a_kmajor && b_kmajor) LAUNCH(false, true);
  else LAUNCH(false, false);
#undef LAUNCH
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e == cudaSuccess && splits > 1) {
    const long long threads = (long long)M * ((N + 3) / 4);
    launch_pdl(splitk_reduce_kernel, dim3((unsigned)((threads + 255) / 256)), dim3(256), 0, st, g, mt, nt);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) { g_dgn_last_cuda = e; return DGN_ERR_CUDA; }
  return DGN_OK;
}
