// Internal: building blocks shared by the tcgen05 kernels (dgn_gemm.cu, dgn_post.cu), sm_100a.
//   * mbarrier / tcgen05 / TMEM / cluster wrappers (inline PTX)
//   * shared-memory matrix descriptors and the kind::tf32 instruction descriptor
//   * operand tile loaders: global fp32 -> registers -> 3xTF32 {hi, lo} split -> shared memory in the UMMA canonical
//     SWIZZLE_128B (K-major) or SWIZZLE_128B_BASE32B (MN-major) layout
// Not part of the C ABI.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace dgn {
namespace umma {

constexpr int BK = 32;                    // floats of K per shared-memory tile = one 128 B swizzle row

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mb_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n)); }
__device__ __forceinline__ void mb_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mb_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(s32(b)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, leading / stride byte offsets
// (all >> 4), version 1 (Blackwell) at bit 46, layout type at bits 61..63:
//   2 = SWIZZLE_128B          K-major tf32 operands   (8 rows x 128 B atoms, 16 B chunks XOR row)
//   1 = SWIZZLE_128B_BASE32B  MN-major tf32 operands  (4 rows x 128 B atoms, 32 B chunks XOR row) - the only
//                             MN-major layout the tensor core accepts for 32-bit inputs
constexpr uint32_t kLayoutSW128 = 2, kLayoutSW128Base32 = 1;
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

// descriptor of k-step kk (8 floats of K) of an operand tile with ROWS rows (M or N extent)
template <int ROWS, bool KMAJOR>
__device__ __forceinline__ uint64_t tile_desc(uint32_t tile_addr, int kk) {
  if constexpr (KMAJOR) {
    return smem_desc(tile_addr + kk * 32, 0, 1024, kLayoutSW128);               // advance 32 B inside the swizzled row
  } else {
    constexpr uint32_t KA = (ROWS / 32) * 512;                                   // stride between 4-row K-atoms
    return smem_desc(tile_addr + kk * 2 * KA, 512, KA, kLayoutSW128Base32);      // two K-atoms per k-step
  }
}

// instruction descriptor (cute::UMMA::InstrDescriptor) for kind::tf32 with fp32 accumulation
__host__ __device__ constexpr uint32_t instr_desc_tf32(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}

// the three products of the 3xTF32 split of one k-step, small terms first
__device__ __forceinline__ void umma_tf32x3(uint32_t d_tmem, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                            uint32_t idesc, uint32_t accumulate) {
  umma_tf32(d_tmem, a_lo, b_hi, idesc, accumulate);
  umma_tf32(d_tmem, a_hi, b_lo, idesc, 1u);
  umma_tf32(d_tmem, a_hi, b_hi, idesc, 1u);
}

template <int COLS> __device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(slot)), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS> __device__ __forceinline__ void tmem_dealloc(uint32_t addr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tmem_alloc_n(uint32_t* slot, int cols) {      // cols in {64, 128, 256, 512}
  if (cols <= 64) tmem_alloc<64>(slot);
  else if (cols <= 128) tmem_alloc<128>(slot);
  else if (cols <= 256) tmem_alloc<256>(slot);
  else tmem_alloc<512>(slot);
}
__device__ __forceinline__ void tmem_dealloc_n(uint32_t addr, int cols) {
  if (cols <= 64) tmem_dealloc<64>(addr);
  else if (cols <= 128) tmem_dealloc<128>(addr);
  else if (cols <= 256) tmem_dealloc<256>(addr);
  else tmem_dealloc<512>(addr);
}

// 32 accumulator columns of this thread's TMEM lane (lane = 32 * (warp % 4) + lane id): tcgen05.ld 32x32b.x32 + wait
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}

// ---- thread-block cluster / distributed shared memory -------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
// every thread of every CTA of the cluster calls this at the same point (release / acquire: shared-memory writes made
// before it are visible to the peers' ld.shared::cluster after it)
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ uint32_t dsmem_addr(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ float dsmem_ld(uint32_t addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
// Cluster reduction of 32 / NR elements per call: element e lives at the same shared-memory offset addr_of(e) in every
// CTA of the cluster; out(e, sum over ranks 0 .. NR-1 in rank order) - deterministic.  All 32 remote loads are issued
// before the first add: a DSMEM load takes ~200 cycles, a load -> add chain per rank would serialise them.
template <int NR, typename AddrFn, typename OutFn>
__device__ __forceinline__ void dsmem_reduce32(AddrFn&& addr_of, OutFn&& out) {
  constexpr int PER = 32 / NR;
  float v[32];
#pragma unroll
  for (int e = 0; e < PER; ++e) {
    const uint32_t a = addr_of(e);
#pragma unroll
    for (int p = 0; p < NR; ++p) v[e * NR + p] = dsmem_ld(dsmem_addr(a, (uint32_t)p));
  }
#pragma unroll
  for (int e = 0; e < PER; ++e) {
    float acc = 0.f;
#pragma unroll
    for (int p = 0; p < NR; ++p) acc += v[e * NR + p];
    out(e, acc);
  }
}

// ---- 3xTF32 split + swizzled stores ------------------------------------------------------------------
// byte offset of element chunk (row r, 16-byte chunk ch of the 128-byte row) inside a swizzled atom stack
__device__ __forceinline__ uint32_t sw128(uint32_t row_in_atom, uint32_t ch) { return row_in_atom * 128u + ((ch ^ row_in_atom) << 4); }

// round-to-nearest TF32 (the tensor core itself just ignores the low 13 mantissa bits, so pre-rounded values are
// consumed exactly): |a - hi| <= 2^-12 |a|, and the residual is rounded once more, leaving ~2^-23 |a| unaccounted
// (integer add + mask = round-half-away in magnitude; the cvt.rna.tf32.f32 instruction does the same but runs on a
//  low-throughput conversion pipe and made the loader warps the bottleneck of the whole kernel)
__device__ __forceinline__ float rn_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }

__device__ __forceinline__ void split_store(unsigned char* hi_tile, unsigned char* lo_tile, uint32_t off, float4 v) {
  float4 h, l;
  h.x = rn_tf32(v.x); l.x = rn_tf32(v.x - h.x);
  h.y = rn_tf32(v.y); l.y = rn_tf32(v.y - h.y);
  h.z = rn_tf32(v.z); l.z = rn_tf32(v.z - h.z);
  h.w = rn_tf32(v.w); l.w = rn_tf32(v.w - h.w);
  *reinterpret_cast<float4*>(hi_tile + off) = h;
  *reinterpret_cast<float4*>(lo_tile + off) = l;
}

// ---- operand tiles: ROWS (M or N extent) x BK floats, NT loader threads -------------------------------------------
// K-major source: element (r, k) at src[(r0 + r) * ld + k0 + k]; rows >= r_max and columns >= k_max read as zero.
// k0, k_max multiples of 4 and 16 B aligned rows (float4 granularity).
template <int ROWS, int NT>
__device__ __forceinline__ void fetch_k(const float* __restrict__ src, int ld, int r0, int r_max, int k0, int k_max,
                                        float4 (&v)[ROWS * 8 / NT], int t) {
#pragma unroll
  for (int i = 0; i < ROWS * 8 / NT; ++i) {
    const int id = t + i * NT;
    const int row = id >> 3, ch = id & 7;
    const int gr = r0 + row, gk = k0 + ch * 4;
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gr < r_max && gk < k_max) v[i] = __ldg(reinterpret_cast<const float4*>(src + (size_t)gr * ld + gk));
  }
}
template <int ROWS, int NT>
__device__ __forceinline__ void store_k(const float4 (&v)[ROWS * 8 / NT], unsigned char* hi, unsigned char* lo, int t) {
#pragma unroll
  for (int i = 0; i < ROWS * 8 / NT; ++i) {
    const int id = t + i * NT;
    const int row = id >> 3, ch = id & 7;                        // 8 chunks (32 floats of K) per row
    split_store(hi, lo, (uint32_t)(row >> 3) * 1024u + sw128(row & 7, ch), v[i]);
  }
}
// row-scaled variant: v *= scale (per 16 B chunk's row) before the split
__device__ __forceinline__ float4 scale4(float4 v, float s) { return make_float4(v.x * s, v.y * s, v.z * s, v.w * s); }

struct NoScale { __device__ __forceinline__ float operator()(int) const { return 1.f; } static constexpr bool on = false; };

// MN-major source: element (r, k) at src[(k0 + k) * ld + r0 + r]; k-rows >= k_max and columns >= r_max read as zero;
// optional per-k scale functor (kscale(k0 + k)); ones_col >= 0: column r == ones_col reads as 1 on valid k-rows.
// r0, r_max multiples of 4, 16 B aligned rows.
template <int ROWS, int NT, typename Scale = NoScale>
__device__ __forceinline__ void fetch_mn(const float* __restrict__ src, int ld, int r0, int r_max, int k0, int k_max,
                                         int ones_col, float4 (&v)[ROWS * 8 / NT], int t, Scale kscale = Scale()) {
  constexpr int CPR = ROWS / 4;                                  // 16 B chunks per K-row
#pragma unroll
  for (int i = 0; i < ROWS * 8 / NT; ++i) {
    const int id = t + i * NT;
    const int krow = id / CPR, ch = id % CPR;
    const int gk = k0 + krow, gr = r0 + ch * 4;
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gk < k_max) {
      if (gr < r_max) {
        v[i] = __ldg(reinterpret_cast<const float4*>(src + (size_t)gk * ld + gr));
        if constexpr (Scale::on) v[i] = scale4(v[i], kscale(gk));
      } else if (gr == ones_col) {
        v[i].x = 1.f;
      }
    }
  }
}
template <int ROWS, int NT>
__device__ __forceinline__ void store_mn(const float4 (&v)[ROWS * 8 / NT], unsigned char* hi, unsigned char* lo, int t) {
  constexpr int CPR = ROWS / 4;
#pragma unroll
  for (int i = 0; i < ROWS * 8 / NT; ++i) {
    const int id = t + i * NT;
    const int krow = id / CPR, ch = id % CPR;
    // SW128_32B atoms: 4 K-rows x 32 MN-floats (512 B); atoms contiguous along MN, then along K
    const uint32_t kl = krow & 3, c16 = ch & 7;
    const uint32_t off = (uint32_t)(krow >> 2) * (uint32_t)(ROWS / 32) * 512u + (uint32_t)(ch >> 3) * 512u + kl * 128u +
                         ((((c16 >> 1) ^ kl)) << 5) + ((c16 & 1u) << 4);
    split_store(hi, lo, off, v[i]);
  }
}

// MN-major tile that is part of a WIDE operand of rows_total (a multiple of 64) MN-floats: the 64-column tile `term`
// lands at MN-columns [64 term, 64 term + 64) of the wide tile, so that one tcgen05.mma with N = rows_total consumes all
// terms at once (a narrow N = 64 MMA is bound by re-reading the A operand from shared memory: measured 69 instead of 32
// cycles per instruction).
template <int NT>
__device__ __forceinline__ void store_mn_wide(const float4 (&v)[64 * 8 / NT], unsigned char* hi, unsigned char* lo, int t,
                                              int rows_total, int term) {
  constexpr int CPR = 64 / 4;
#pragma unroll
  for (int i = 0; i < 64 * 8 / NT; ++i) {
    const int id = t + i * NT;
    const int krow = id / CPR, ch = term * CPR + id % CPR;
    const uint32_t kl = krow & 3, c16 = ch & 7;
    const uint32_t off = (uint32_t)(krow >> 2) * (uint32_t)(rows_total / 32) * 512u + (uint32_t)(ch >> 3) * 512u + kl * 128u +
                         ((((c16 >> 1) ^ kl)) << 5) + ((c16 & 1u) << 4);
    split_store(hi, lo, off, v[i]);
  }
}
__device__ __forceinline__ uint64_t tile_desc_mn(uint32_t tile_addr, int kk, int rows_total) {
  const uint32_t KA = (uint32_t)(rows_total / 32) * 512u;
  return smem_desc(tile_addr + kk * 2 * KA, 512, KA, kLayoutSW128Base32);
}

}  // namespace umma
}  // namespace dgn
