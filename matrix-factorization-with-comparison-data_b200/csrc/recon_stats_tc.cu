// K5 (tensor-core variant): reconstruction statistics with tcgen05 + TMEM + bulk async copies.
//
// Same contract as k_recon_stats (recon_stats.cu; reference structure.py:939-955, :980-1064): one pass over
// W = U V^T and X that leaves six fp64 sums per row.  Here the 128 x 64 tiles of W are produced by the
// 5th-generation tensor cores:
//   * A = 128 rows of U, B = 64 rows of V, both K-major in shared memory with the 128-byte swizzle the
//     UMMA descriptors expect; operands are split  x = hi + lo  (hi = top 10 mantissa bits, exactly a
//     TF32 value) and three MMAs  hi.hi + hi.lo + lo.hi  accumulate in TMEM, which restores fp32-grade
//     products (the reference multiplies in fp32);
//   * tcgen05.mma (kind::tf32, M=128, N=64, K=8 per instruction) is issued by one elected thread;
//     accumulators live in TMEM (2 stages x 64 columns) and come back with tcgen05.ld for the epilogue;
//   * the matching 128 x 64 tile of X -- the only large HBM stream, 4 n m bytes -- is staged by the bulk
//     async-copy engine (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP) two tiles ahead of use;
//   * warp roles: warps 0-3 epilogue (TMEM lane quarter = warp id), warps 4-5 operand producers,
//     warp 6 TMEM owner + MMA issuer; four mbarriers per stage connect them.
// Every mbarrier wait is bounded: a pipeline bug sets an error flag and lets the kernel drain instead
// of hanging the GPU.
#include "internal.h"

namespace mfcd {
namespace tc {

constexpr int TM = 128;                 // tile rows  (UMMA M, TMEM lanes)
constexpr int TN = 64;                  // tile cols  (UMMA N, TMEM columns per stage)
constexpr int KMAX = 64;                // largest (padded) K handled by this kernel
constexpr int SLAB_K = 32;              // tf32 elements per 128-byte swizzle row
constexpr int A_SLAB_BYTES = TM * 128;  // one K-slab of the A tile
constexpr int B_SLAB_BYTES = TN * 128;
constexpr int X_PITCH = TN + 4;         // floats; 272-byte rows keep LDS.128 conflict-free and 16-byte aligned
constexpr int NSTAGE = 2;
constexpr int NTHREADS = 7 * 32;
constexpr int EPI_THREADS = 128;
constexpr int PROD_THREADS = 64;

struct Smem {
  // 1024-byte aligned slabs first
  float a_hi[2][TM * SLAB_K];
  float a_lo[2][TM * SLAB_K];
  float b_hi[NSTAGE][2][TN * SLAB_K];
  float b_lo[NSTAGE][2][TN * SLAB_K];
  float xs[NSTAGE][TM * X_PITCH];
  float b_col[NSTAGE][TN];
  float vbar[KMAX], ubar[KMAX];
  unsigned long long bar_full_b[NSTAGE], bar_full_x[NSTAGE], bar_mma_done[NSTAGE], bar_epi_done[NSTAGE];
  uint32_t tmem_base;
  int error;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// bounded wait: returns false (and raises the CTA's error flag) instead of spinning forever
__device__ __forceinline__ bool mbar_wait(unsigned long long* bar, uint32_t parity, volatile int* err) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (*err) return false;
    if (clock64() - t0 > (1ll << 28)) { *err = 1; return false; }
  }
  return true;
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* result_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(result_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = __uint_as_float(r[k]);
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start address >> 4 in [0,14), LBO >> 4 in [16,30) (unused for swizzled K-major: 1), SBO >> 4 in [32,46)
// (1024 B between 8-row groups), version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10),
// both K-major (bits 15, 16 = 0), N >> 3 in [17,23), M >> 4 in [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// split 4 floats into tf32-exact hi and the remainder lo, store both at the swizzled position of (row, k4)
__device__ __forceinline__ void store_split(float* hi_slabs, float* lo_slabs, int slab_floats, int row, int k, float4 v) {
  const int slab = k / SLAB_K;
  const int chunk = (k % SLAB_K) >> 2;                         // 16-byte chunk inside the 128-byte row
  const int off = slab * slab_floats + row * SLAB_K + ((chunk ^ (row & 7)) << 2);
  float4 h, l;
  h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
  h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
  h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
  h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
  *reinterpret_cast<float4*>(hi_slabs + off) = h;
  *reinterpret_cast<float4*>(lo_slabs + off) = l;
}

__device__ __forceinline__ float4 load_row4(const float* __restrict__ T, int64_t row, int64_t rows, int k, int d, bool vec) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row < rows && k < d) {
    const float* p = T + row * d + k;
    if (vec) {
      v = __ldg(reinterpret_cast<const float4*>(p));
    } else {
      v.x = __ldg(p);
      if (k + 1 < d) v.y = __ldg(p + 1);
      if (k + 2 < d) v.z = __ldg(p + 2);
      if (k + 3 < d) v.w = __ldg(p + 3);
    }
  }
  return v;
}

__global__ void __launch_bounds__(NTHREADS, 1)
k_recon_stats_tc(const float* __restrict__ U, const float* __restrict__ V, int64_t n, int64_t m, int d, int kp,
                 const float* __restrict__ X, int64_t ldx, float s, const float* __restrict__ ubar_g,
                 const float* __restrict__ vbar_g, int col_splits, double* __restrict__ row_stats,
                 int* __restrict__ error_flag) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  volatile int* err = &sm.error;
  const bool vec_uv = (d & 3) == 0;
  const int nslab = (kp + SLAB_K - 1) / SLAB_K;
  const int64_t row_tiles = (n + TM - 1) / TM;
  const int64_t col_tiles = (m + TN - 1) / TN;

  if (threadIdx.x == 0) {
    sm.error = 0;
    for (int st = 0; st < NSTAGE; ++st) {
      mbar_init(&sm.bar_full_b[st], PROD_THREADS);
      mbar_init(&sm.bar_full_x[st], 1);
      mbar_init(&sm.bar_mma_done[st], 1);
      mbar_init(&sm.bar_epi_done[st], EPI_THREADS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int k = threadIdx.x; k < KMAX; k += NTHREADS) {
    sm.vbar[k] = k < d ? vbar_g[k] : 0.f;
    sm.ubar[k] = k < d ? ubar_g[k] : 0.f;
  }
  if (warp == 6) tmem_alloc(&sm.tmem_base, NSTAGE * TN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;
  const uint32_t idesc = make_idesc(TM, TN);

  uint32_t use = 0;       // tiles processed so far by this CTA: stage = use % 2, phase = (use / 2) & 1

  for (int64_t work = blockIdx.x; work < row_tiles * col_splits; work += gridDim.x) {
    const int64_t rt = work / col_splits;
    const int split = (int)(work % col_splits);
    const int64_t row0 = rt * TM;
    const int64_t ct_begin = col_tiles * split / col_splits;
    const int64_t ct_end = col_tiles * (split + 1) / col_splits;
    const int ntiles = (int)(ct_end - ct_begin);

    // A tile: all threads cooperate; the previous work item's MMAs have completed (its epilogues waited on them)
    for (int idx = threadIdx.x; idx < TM * (kp >> 2); idx += NTHREADS) {
      const int row = idx / (kp >> 2), k = (idx % (kp >> 2)) << 2;
      store_split(&sm.a_hi[0][0], &sm.a_lo[0][0], TM * SLAB_K, row, k, load_row4(U, row0 + row, n, k, d, vec_uv));
    }
    fence_proxy_async();
    __syncthreads();

    if (warp < 4) {
      // ================= epilogue: thread t owns tile row t (TMEM lane t) =================
      const int t = threadIdx.x;
      const int64_t gr = row0 + t;
      float a_row = 0.f;
      if (gr < n)
        for (int k = 0; k < d; ++k) a_row = fmaf(__ldg(U + gr * d + k), sm.vbar[k], a_row);
      double acc[6] = {0, 0, 0, 0, 0, 0};
      for (int it = 0; it < ntiles; ++it) {
        const uint32_t u = use + it, st = u & 1, ph = (u >> 1) & 1;
        const int64_t col0 = (ct_begin + it) * TN;
        if (!mbar_wait(&sm.bar_mma_done[st], ph, err)) break;
        if (!mbar_wait(&sm.bar_full_x[st], ph, err)) break;
        tc_fence_after();
        const float* xrow = &sm.xs[st][t * X_PITCH];
#pragma unroll 1
        for (int c0 = 0; c0 < TN; c0 += 16) {
          float w[16];
          tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + st * TN + c0, w);
          float sx = 0.f, sxx = 0.f, sw = 0.f, sww = 0.f, sxw = 0.f, see = 0.f;
#pragma unroll
          for (int q = 0; q < 16; q += 4) {
            const float4 xv = *reinterpret_cast<const float4*>(xrow + c0 + q);
            const float xs4[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (col0 + c0 + q + e < m) {
                const float x = xs4[e];
                const float wa = w[q + e] - a_row;
                const float ee = (w[q + e] - sm.b_col[st][c0 + q + e]) - s * x;
                sx += x; sxx = fmaf(x, x, sxx);
                sw += wa; sww = fmaf(wa, wa, sww); sxw = fmaf(x, wa, sxw);
                see = fmaf(ee, ee, see);
              }
            }
          }
          acc[0] += (double)sx; acc[1] += (double)sxx; acc[2] += (double)sw;
          acc[3] += (double)sww; acc[4] += (double)sxw; acc[5] += (double)see;
        }
        tc_fence_before();
        mbar_arrive(&sm.bar_epi_done[st]);
      }
      if (gr < n) {
#pragma unroll
        for (int q = 0; q < 6; ++q) atomicAdd(row_stats + gr * 8 + q, acc[q]);
        if (split == 0) row_stats[gr * 8 + 6] = (double)a_row;
      }
    } else if (warp < 6) {
      // ================= producers: X tile via bulk copies, B tile split + swizzled =================
      const int pt = threadIdx.x - 128;      // 0..63
      for (int it = 0; it < ntiles; ++it) {
        const uint32_t u = use + it, st = u & 1, ph = (u >> 1) & 1;
        const int64_t col0 = (ct_begin + it) * TN;
        if (u >= NSTAGE) {                    // stage reuse: previous MMAs done with B, previous epilogue done with X
          if (!mbar_wait(&sm.bar_mma_done[st], ph ^ 1, err)) break;
          if (!mbar_wait(&sm.bar_epi_done[st], ph ^ 1, err)) break;
        }
        const int ncols = (int)((m - col0) < TN ? (m - col0) : TN);
        const int nrows = (int)((n - row0) < TM ? (n - row0) : TM);
        if (pt == 0) mbar_arrive_expect_tx(&sm.bar_full_x[st], (uint32_t)(nrows * ncols * 4));
        __syncwarp();
        if (warp == 4) {
          for (int r = lane; r < nrows; r += 32)
            bulk_g2s(&sm.xs[st][r * X_PITCH], X + (row0 + r) * ldx + col0, (uint32_t)(ncols * 4), &sm.bar_full_x[st]);
        }
        // B tile (64 rows of V) and its column means <ubar, V_c>
        for (int idx = pt; idx < TN * (kp >> 2); idx += PROD_THREADS) {
          const int row = idx / (kp >> 2), k = (idx % (kp >> 2)) << 2;
          store_split(&sm.b_hi[st][0][0], &sm.b_lo[st][0][0], TN * SLAB_K, row, k,
                      load_row4(V, col0 + row, m, k, d, vec_uv));
        }
        {
          const int64_t gc = col0 + pt;
          float b = 0.f;
          if (gc < m)
            for (int k = 0; k < d; ++k) b = fmaf(sm.ubar[k], __ldg(V + gc * d + k), b);
          sm.b_col[st][pt] = b;
        }
        fence_proxy_async();                 // generic-proxy smem writes -> visible to the tensor core's async proxy
        mbar_arrive(&sm.bar_full_b[st]);
      }
    } else {
      // ================= MMA issuer (one elected thread of warp 6) =================
      if (lane == 0) {
        for (int it = 0; it < ntiles; ++it) {
          const uint32_t u = use + it, st = u & 1, ph = (u >> 1) & 1;
          if (!mbar_wait(&sm.bar_full_b[st], ph, err)) break;
          if (u >= NSTAGE && !mbar_wait(&sm.bar_epi_done[st], ph ^ 1, err)) break;   // TMEM stage drained
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + st * TN;
          uint32_t accumulate = 0;
          for (int ks = 0; ks < (kp >> 3); ++ks) {
            const int slab = ks >> 2, within = (ks & 3) * 32;          // 8 tf32 = 32 bytes per K step
            const uint64_t a_hi = make_desc(smem_u32(&sm.a_hi[slab][0]) + within);
            const uint64_t a_lo = make_desc(smem_u32(&sm.a_lo[slab][0]) + within);
            const uint64_t b_hi = make_desc(smem_u32(&sm.b_hi[st][slab][0]) + within);
            const uint64_t b_lo = make_desc(smem_u32(&sm.b_lo[st][slab][0]) + within);
            umma_tf32(d_tmem, a_hi, b_hi, idesc, accumulate);
            umma_tf32(d_tmem, a_hi, b_lo, idesc, 1u);
            umma_tf32(d_tmem, a_lo, b_hi, idesc, 1u);
            accumulate = 1u;
          }
          umma_commit(&sm.bar_mma_done[st]);   // arrives when the MMAs above have finished (implies fence::before_thread_sync)
        }
      }
      __syncwarp();
    }
    use += (uint32_t)ntiles;
    (void)nslab;
    tc_fence_before();
    __syncthreads();                          // work-item boundary: all roles done with this row block
    tc_fence_after();
    if (sm.error) break;
  }

  __syncthreads();
  if (warp == 6) tmem_dealloc(tmem_base, NSTAGE * TN);
  if (threadIdx.x == 0 && sm.error) atomicExch(error_flag, 1);
}

}  // namespace tc
}  // namespace mfcd

using namespace mfcd;

extern "C" int mfcd_recon_stats_tc(const float* U, const float* V, int64_t n, int64_t m, int32_t d,
                                   const mfcd_xview* X, float s, const float* ubar, const float* vbar,
                                   double* row_stats, int32_t* error_flag, void* stream) {
  MFCD_REQUIRE(U && V && X && ubar && vbar && row_stats && error_flag, "mfcd_recon_stats_tc: NULL pointer");
  MFCD_REQUIRE(n >= 1 && m >= 1 && d >= 1, "mfcd_recon_stats_tc: bad sizes");
  // eligibility: dense X with 16-byte aligned rows (bulk copies), K = d padded to a multiple of 8 up to 64
  if (X->X == nullptr || d > tc::KMAX || (m & 3) != 0 || (X->ldx & 3) != 0 ||
      (reinterpret_cast<uintptr_t>(X->X) & 15u) != 0) {
    set_error("mfcd_recon_stats_tc: shape not eligible (needs dense X, d <= 64, m %% 4 == 0, 16-byte aligned rows)");
    return MFCD_ERR_UNSUPPORTED;
  }
  cudaStream_t st = as_stream(stream);
  const int kp = (d + 7) & ~7;
  MFCD_CUDA(cudaMemsetAsync(row_stats, 0, sizeof(double) * 8 * n, st));
  MFCD_CUDA(cudaMemsetAsync(error_flag, 0, sizeof(int32_t), st));
  const int64_t row_tiles = (n + tc::TM - 1) / tc::TM;
  const int64_t col_tiles = (m + tc::TN - 1) / tc::TN;
  const int64_t sms = sm_count();
  int64_t splits = (2 * sms + row_tiles - 1) / row_tiles;       // >= 2 work items per SM when possible
  if (splits < 1) splits = 1;
  if (splits > col_tiles) splits = col_tiles;
  int64_t blocks = row_tiles * splits;
  if (blocks > sms) blocks = sms;                                // persistent: one CTA per SM
  const size_t smem = sizeof(tc::Smem) + 1024;
  MFCD_CUDA(cudaFuncSetAttribute(tc::k_recon_stats_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc::k_recon_stats_tc<<<(int)blocks, tc::NTHREADS, smem, st>>>(U, V, n, m, d, kp, X->X, X->ldx, s, ubar, vbar,
                                                               (int)splits, row_stats, error_flag);
  MFCD_CHECK_LAUNCH();
  return MFCD_OK;
}
