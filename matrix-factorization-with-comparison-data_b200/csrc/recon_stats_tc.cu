// K5 (tensor-core variant): reconstruction statistics with tcgen05 + TMEM + TMA.
//
// Same contract as k_recon_stats (recon_stats.cu; reference structure.py:939-955, :980-1064): one pass over
// W = U V^T and X that leaves six fp64 sums per row.  X (4 n m bytes) is the only large stream, so the
// kernel is organised around keeping that stream moving at HBM speed:
//   * a tiny pre-pass splits V into  hi + lo  (hi = top 10 mantissa bits = an exact TF32 value) K-padded
//     staging tables, and computes the row / column means of W (a_r = <U_r, vbar>, b_c = <ubar, V_c>);
//   * the A operand (128 rows of U, split hi / lo on the fly) lives in TENSOR MEMORY, not shared memory: at the
//     start of a row block the epilogue warps read their rows of U from global memory and write them with
//     tcgen05.st next to the accumulators (TMEM columns [256, 256 + K) = hi, [384, 384 + K) = lo).  That
//     frees 64 KB of shared memory at K = 64 for a deeper X ring (5 tiles instead of 3: the round-1 profile
//     showed the X stream starved by a 3-deep ring) and is what makes K = 128 fit at all;
//   * ONE producer thread per CTA feeds the rest with TMA tensor copies (cp.async.bulk.tensor.2d, 128-byte
//     swizzle, zero fill outside the matrices): per 128 x 64 tile the B operand (64 rows of V_hi / V_lo) and,
//     through a separate ring that runs 2-5 tiles ahead (3-6 slots of 32 KB), the matching X tile;
//   * ONE thread issues tcgen05.mma (kind::tf32, A from TMEM, B from shared memory, M = 128, N = 64, K = 8 per
//     instruction), three MMAs per K step (hi.hi + hi.lo + lo.hi: fp32-grade products like the reference's
//     sgemm), accumulating in TMEM (4 stages x 64 columns);
//   * sixteen epilogue warps in two groups of eight that take alternate tiles (TMEM lane quarter = warp id % 4,
//     32-column half = (warp id / 4) % 2, group = warp id / 8) read the accumulators with tcgen05.ld, the X tile
//     from swizzled shared memory, and fold both into per-row fp64 sums with packed fp32x2 arithmetic.
// mbarriers connect the roles; every wait is bounded, so a pipeline bug raises an error flag and lets
// the kernel drain instead of hanging the GPU.
#include <cuda.h>
#include <stdlib.h>
#include "internal.h"

namespace mfcd {
namespace tc {

constexpr int TM = 128;                 // tile rows  (UMMA M, TMEM lanes)
constexpr int TN = 64;                  // tile cols  (UMMA N, TMEM columns per stage)
constexpr int KMAX = 128;               // largest K handled (4 swizzle slabs of B; A = 2 x 128 TMEM columns)
constexpr int SLAB_K = 32;              // tf32 elements per 128-byte swizzle row
constexpr int NSTAGE = 4;                // TMEM accumulator stages (see the note at the MMA issuer)
constexpr int MAX_BSTAGE = 4;            // shared-memory stages of the B operand (2 - 4, by K)
constexpr int EPI_WARPS = 16;           // warp w: TMEM lane quarter w % 4, 32-column half (w / 4) % 2, tile group w / 8
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int PRODUCER_WARP = EPI_WARPS, MMA_WARP = EPI_WARPS + 1, BPRODUCER_WARP = EPI_WARPS + 2;
constexpr int NTHREADS = (EPI_WARPS + 3) * 32;
constexpr uint32_t B_SLAB_BYTES = TN * 128;
constexpr uint32_t TMEM_A_HI = NSTAGE * TN;            // TMEM column of the A operand's hi part
constexpr uint32_t TMEM_A_LO = TMEM_A_HI + KMAX;       // ... and of its lo part
constexpr uint32_t X_BOX_BYTES = TM * 128;   // 128 rows x 32 columns of X

constexpr int MAX_RING = 6;              // X-tile ring depth (tiles): 6 next to K <= 32, 5 at K = 64, 3 at K = 128
constexpr uint32_t XSLOT_BYTES = 2 * X_BOX_BYTES;      // one ring slot = the two 32-column boxes of a tile

// Shared-memory plan (all regions 1024-byte aligned, sizes depend on the number of K slabs):
//   B  : 2-4 stages x (b_hi[nslab], b_lo[nslab])   8 KB each.  The B tile of tile t+nbst is requested when the MMAs of
//        tile t complete, so with 2 stages the L2 latency of a 32 KB operand tile sat on the critical path (round-2
//        profile at K = 64: DRAM 39 %, tensor pipe 29 %, issue 45 % -- nothing saturated, tile period 1.6 us)
//   X  : ring of `ring` tile slots                 32 KB each
//   tail: b_col[ring][64], mbarriers, TMEM base, error flag
struct Tail {
  float b_col[MAX_RING][TN];
  unsigned long long x_full[MAX_RING], x_free[MAX_RING];
  unsigned long long b_full[MAX_BSTAGE], b_free[MAX_BSTAGE], mma_done[NSTAGE], tmem_free[NSTAGE], a_full;
  uint32_t tmem_base;
  int error;
};
__host__ __device__ constexpr uint32_t b_stage_bytes(int nslab) { return 2u * nslab * B_SLAB_BYTES; }
__host__ __device__ constexpr uint32_t smem_bytes(int nslab, int nbst, int ring) {
  return nbst * b_stage_bytes(nslab) + ring * XSLOT_BYTES + (uint32_t)sizeof(Tail) + 1024u;
}
// TMEM columns to allocate (power of two >= 32): 2 accumulator stages + hi and lo of the A operand
__host__ __device__ constexpr uint32_t tmem_cols(int kpad) { return (NSTAGE * TN + 2 * kpad) <= 256 ? 256u : 512u; }

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
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (*err) return false;
    if (clock64() - t0 > (1ll << 28)) { *err = 1; return false; }
  }
  return true;
}
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* map, int c_inner, int c_outer,
                                            unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c_inner), "r"(c_outer), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// one lane of a converged warp (elect.sync): unlike `lane == 0`, ptxas knows the region that follows runs in a
// single thread, so the uniform-datapath operands of UTCHMMA / UTMALDG need no per-instruction election loop
// (the round-2 profile showed ~40 scalar instructions and ~130 cycles around every tcgen05.mma: the ISSUER, not the
// tensor pipe -- 26 % busy -- set the 3 165-cycle tile period at K = 64, four times the MMA floor)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
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
// A operand from tensor memory (lane = tile row, one 32-bit column per K element), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// 8 consecutive TMEM columns of this thread's lane (the warp covers its 32-lane quarter)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                 "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// issue the load of 16 consecutive TMEM columns of this thread's lane; the registers are valid after tmem_ld16_wait
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
// the "+r" operands tie every later use of the registers to this wait (the compiler cannot hoist a use above it)
__device__ __forceinline__ void tmem_ld16_wait(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :: "memory");
}

__device__ __forceinline__ f32x2 pk2u(uint32_t lo, uint32_t hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
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

// ---- pre-pass (ONE launch): hi/lo split of both tables into K-padded staging arrays, their row-wise dots with
// the other table's column mean (a_r = <U_r, vbar>, b_c = <ubar, V_c>), and the zeroing of the outputs ----
__device__ __forceinline__ void split_table(const float* __restrict__ T, int64_t rows, int d, int kpad,
                                            const float* __restrict__ other_mean, float* __restrict__ hi,
                                            float* __restrict__ lo, float* __restrict__ dots, int64_t dots_len,
                                            int64_t tid, int64_t nth) {
  const int64_t total = hi ? rows * kpad : 0;          // hi == NULL: only the dots (A is split in the main kernel)
  for (int64_t idx = tid; idx < total; idx += nth) {
    const int64_t r = idx / kpad;
    const int k = (int)(idx - r * kpad);
    const float v = k < d ? __ldg(T + r * d + k) : 0.f;
    const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    hi[idx] = h;
    lo[idx] = v - h;
  }
  for (int64_t r = tid; r < dots_len; r += nth) {
    float acc = 0.f;
    if (r < rows)
      for (int k = 0; k < d; ++k) acc = fmaf(__ldg(T + r * d + k), __ldg(other_mean + k), acc);
    dots[r] = acc;                                   // padding entries (r >= rows) are zero
  }
}

__global__ void __launch_bounds__(256)
k_tc_prepass(const float* __restrict__ U, const float* __restrict__ V, int64_t n, int64_t m, int d, int kpad,
             const float* __restrict__ ubar, const float* __restrict__ vbar, float* __restrict__ v_hi,
             float* __restrict__ v_lo, float* __restrict__ avec,
             float* __restrict__ bvec, int64_t bvec_len, double* __restrict__ row_stats, int* __restrict__ error_flag) {
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;
  split_table(U, n, d, kpad, vbar, nullptr, nullptr, avec, n, tid, nth);
  split_table(V, m, d, kpad, ubar, v_hi, v_lo, bvec, bvec_len, tid, nth);
  for (int64_t k = tid; k < 8 * n; k += nth) row_stats[k] = 0.0;
  if (tid == 0) *error_flag = 0;
}

struct Maps {
  CUtensorMap x, v_hi, v_lo;
};

__global__ void __launch_bounds__(NTHREADS, 1)
k_recon_stats_tc(const __grid_constant__ Maps maps, const float* __restrict__ U, int64_t n, int64_t m, int d, int kp,
                 int nbst, int ring, float s, const float* __restrict__ avec, const float* __restrict__ bvec, int probe,
                 double* __restrict__ row_stats, int* __restrict__ error_flag) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // align inside the shared window with shared-space arithmetic so the compiler keeps LDS/STS addressing
  unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nslab = (kp + SLAB_K - 1) / SLAB_K;
  unsigned char* b_base = base;                                   // stage st: b_hi[nslab] then b_lo[nslab]
  unsigned char* x_base = b_base + nbst * b_stage_bytes(nslab);   // slot sl: box 0, box 1
  Tail& tl = *reinterpret_cast<Tail*>(x_base + ring * XSLOT_BYTES);
  volatile int* err = &tl.error;
  const int64_t row_tiles = (n + TM - 1) / TM;
  const int64_t col_tiles = (m + TN - 1) / TN;

  if (threadIdx.x == 0) {
    tl.error = 0;
    for (int k = 0; k < MAX_RING; ++k) {
      mbar_init(&tl.x_full[k], 1);
      mbar_init(&tl.x_free[k], EPI_THREADS / 2);        // one group of eight warps per tile
    }
    for (int st = 0; st < MAX_BSTAGE; ++st) {
      mbar_init(&tl.b_full[st], 1);
      mbar_init(&tl.b_free[st], 1);
    }
    for (int st = 0; st < NSTAGE; ++st) {
      mbar_init(&tl.mma_done[st], 1);
      mbar_init(&tl.tmem_free[st], EPI_THREADS / 2);
    }
    mbar_init(&tl.a_full, EPI_THREADS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) tmem_alloc(&tl.tmem_base, tmem_cols(nslab * SLAB_K));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tl.tmem_base;
  const uint32_t idesc = make_idesc(TM, TN);

  uint32_t use = 0;       // tiles processed so far by this CTA: B/TMEM stage = use % 2, X slot = use % ring
  uint32_t item = 0;      // work items processed so far: phase of a_full

  // Balanced persistent schedule: the row-major tile sequence is cut into gridDim.x equal ranges (+-1 tile);
  // a CTA walks its range one row block at a time (row blocks x fixed column splits left 24 of 148 CTAs with
  // a third work item while the others idled: 72 % wave efficiency).
  const int64_t total_tiles = row_tiles * col_tiles;
  const int64_t t_begin = total_tiles * blockIdx.x / gridDim.x;
  const int64_t t_end = total_tiles * (blockIdx.x + 1) / gridDim.x;
  for (int64_t tpos = t_begin; tpos < t_end; ++item) {
    const int64_t rt = tpos / col_tiles;
    const int row0 = (int)(rt * TM);
    const int64_t ct_begin = tpos - rt * col_tiles;
    const int64_t ct_end = (t_end - rt * col_tiles) < col_tiles ? (t_end - rt * col_tiles) : col_tiles;
    const int ntiles = (int)(ct_end - ct_begin);
    tpos += ntiles;

    if (warp < EPI_WARPS) {
      // ===== epilogue: thread owns tile row t (TMEM lane t) and one 16-column quarter of the tile =====
      const int t = (warp & 3) * 32 + lane;
      const int quarter = warp >> 2;                              // columns quarter*16 .. +15 of the tile
      const int64_t gr = (int64_t)row0 + t;
      const float a_row = gr < n ? __ldg(avec + gr) : 0.f;
      const int sw = t & 7;
      {
        // A operand of this row block -> tensor memory: the thread's row of U, this warp's quarter of the K range,
        // split into hi (exact TF32) and lo = v - hi, eight columns per tcgen05.st.  The previous row block's MMAs
        // have completed (work-item barrier below), so the columns are free to overwrite.
        const int kc = (nslab * SLAB_K) >> 2;                       // K columns per warp quarter: 8, 16, 24 or 32
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const float* urow = U + gr * (int64_t)d;
        const bool vec = (d & 3) == 0;
        for (int k0 = quarter * kc; k0 < (quarter + 1) * kc; k0 += 8) {
          float v[8], hi[8], lo[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = 0.f;
          if (gr < n && k0 < d) {
            if (vec && k0 + 8 <= d) {
              const float4 p = __ldg(reinterpret_cast<const float4*>(urow + k0));
              const float4 q = __ldg(reinterpret_cast<const float4*>(urow + k0 + 4));
              v[0] = p.x; v[1] = p.y; v[2] = p.z; v[3] = p.w; v[4] = q.x; v[5] = q.y; v[6] = q.z; v[7] = q.w;
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) if (k0 + e < d) v[e] = __ldg(urow + k0 + e);
            }
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            hi[e] = __uint_as_float(__float_as_uint(v[e]) & 0xFFFFE000u);
            lo[e] = v[e] - hi[e];
          }
          tmem_st8(tmem_base + lane_base + TMEM_A_HI + (uint32_t)k0, hi);
          tmem_st8(tmem_base + lane_base + TMEM_A_LO + (uint32_t)k0, lo);
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&tl.a_full);
      }
      // Two groups of eight warps take ALTERNATE tiles (group = warp / 8; inside a group: TMEM lane quarter =
      // warp % 4, 32-column half of the tile = (warp / 4) % 2).  The round-2 profiles showed the sixteen-warp,
      // one-tile-at-a-time epilogue as the pipeline's slow stage: ~170 instructions per thread and tile taking
      // ~1 950 cycles (4 warps per scheduler, all in the same phase of the same tile, long-scoreboard and FP64-pipe
      // stalls exposed) against 768 cycles of MMAs.  With 32 columns per thread the per-tile overhead (barrier
      // waits, address arithmetic, the six fp32 -> fp64 folds) is paid once per 32 elements instead of once per 16,
      // and while one group waits for its accumulators the other one computes.
      double acc[6] = {0, 0, 0, 0, 0, 0};
      const int half = (warp >> 2) & 1;                         // columns half*32 .. +31 of the tile = X box `half`
      const int egrp = warp >> 3;                               // this group's tiles: it = egrp, egrp + 2, ...
      const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + half * 32;
      const uint32_t xoff = half * X_BOX_BYTES + (uint32_t)t * (SLAB_K * 4);
      const uint32_t u0 = use + egrp;
      uint32_t slot = u0 % (uint32_t)ring, xph = (u0 / (uint32_t)ring) & 1;     // ring position, advanced by 2 per tile
      const f32x2 a2 = pk2(a_row, a_row), ns2 = pk2(-s, -s);
      for (int it = egrp; it < ntiles; it += 2) {
        const uint32_t u = use + it, st = u % NSTAGE, ph = (u / NSTAGE) & 1;
        const int64_t col0 = (ct_begin + it) * TN + half * 32;
        if (!mbar_wait(&tl.x_full[slot], xph, err)) break;      // X tile + column means landed
        if (probe == 1) {                                       // microbenchmark: the X stream alone
          mbar_arrive(&tl.x_free[slot]);
          slot += 2;
          if (slot >= (uint32_t)ring) { slot -= (uint32_t)ring; xph ^= 1; }
          continue;
        }
        if (!mbar_wait(&tl.mma_done[st], ph, err)) break;       // accumulators complete
        tc_fence_after();
        if (probe == 2) {                                       // microbenchmark: X stream + MMAs, no epilogue math
          tc_fence_before();
          mbar_arrive(&tl.tmem_free[st]);
          mbar_arrive(&tl.x_free[slot]);
          slot += 2;
          if (slot >= (uint32_t)ring) { slot -= (uint32_t)ring; xph ^= 1; }
          continue;
        }
        uint32_t w[32];
        {
          uint32_t (&wa)[16] = *reinterpret_cast<uint32_t (*)[16]>(&w[0]);
          uint32_t (&wb)[16] = *reinterpret_cast<uint32_t (*)[16]>(&w[16]);
          tmem_ld16_issue(lane_addr + st * TN, wa);
          tmem_ld16_issue(lane_addr + st * TN + 16, wb);
          tmem_ld16_wait(wa);
          tmem_ld16_wait(wb);
        }
        tc_fence_before();
        mbar_arrive(&tl.tmem_free[st]);                         // the accumulator stage is in registers
        const unsigned char* xrow = x_base + slot * XSLOT_BYTES + xoff;
        const float* bcol = &tl.b_col[slot][half * 32];
        const int64_t left = m - col0;                          // columns of this half inside X
        float sx = 0.f, sxx = 0.f, swm = 0.f, sww = 0.f, sxw = 0.f, see = 0.f;
        f32x2 px = 0, pxx = 0, pwm = 0, pww = 0, pxw = 0, pee = 0;           // (+0.f, +0.f)
#pragma unroll
        for (int hq = 0; hq < 2; ++hq) {                        // two passes of 16 columns (register budget)
          float4 xv[4], bv[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            xv[q] = *reinterpret_cast<const float4*>(xrow + (((hq * 4 + q) ^ sw) << 4));   // un-swizzle the 16-byte chunk
            bv[q] = *reinterpret_cast<const float4*>(bcol + hq * 16 + 4 * q);
          }
          if (hq == 1) mbar_arrive(&tl.x_free[slot]);           // the whole X half-row has been read
          if (left >= 32) {                                     // interior tile: packed pairs, no bounds checks
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const f32x2 x01 = pk2(xv[q].x, xv[q].y), x23 = pk2(xv[q].z, xv[q].w);
              const f32x2 b01 = pk2(bv[q].x, bv[q].y), b23 = pk2(bv[q].z, bv[q].w);
              const int c = hq * 16 + 4 * q;
              const f32x2 w01 = pk2u(w[c], w[c + 1]), w23 = pk2u(w[c + 2], w[c + 3]);
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const f32x2 x = h ? x23 : x01, b = h ? b23 : b01, wv = h ? w23 : w01;
                const f32x2 wa = sub2(wv, a2);
                const f32x2 ee = fma2(ns2, x, sub2(wv, b));                  // (w - b) - s x
                px = add2(px, x); pxx = fma2(x, x, pxx);
                pwm = add2(pwm, wa); pww = fma2(wa, wa, pww); pxw = fma2(x, wa, pxw);
                pee = fma2(ee, ee, pee);
              }
            }
          } else {
            const int valid = (int)(left < 0 ? 0 : left);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float xs4[4] = {xv[q].x, xv[q].y, xv[q].z, xv[q].w};
              const float bs4[4] = {bv[q].x, bv[q].y, bv[q].z, bv[q].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int c = hq * 16 + 4 * q + e;
                if (c < valid) {
                  const float x = xs4[e], wf = __uint_as_float(w[c]);
                  const float wa = wf - a_row;
                  const float ee = (wf - bs4[e]) - s * x;
                  sx += x; sxx = fmaf(x, x, sxx);
                  swm += wa; sww = fmaf(wa, wa, sww); sxw = fmaf(x, wa, sxw);
                  see = fmaf(ee, ee, see);
                }
              }
            }
          }
        }
        if (left >= 32) {
          sx = sum2(px); sxx = sum2(pxx); swm = sum2(pwm); sww = sum2(pww); sxw = sum2(pxw); see = sum2(pee);
        }
        // 32-element fp32 partials, fp64 across tiles
        acc[0] += (double)sx; acc[1] += (double)sxx; acc[2] += (double)swm;
        acc[3] += (double)sww; acc[4] += (double)sxw; acc[5] += (double)see;
        slot += 2;
        if (slot >= (uint32_t)ring) { slot -= (uint32_t)ring; xph ^= 1; }
      }
      if (gr < n) {
#pragma unroll
        for (int q = 0; q < 6; ++q) atomicAdd(row_stats + gr * 8 + q, acc[q]);
        if (ct_begin == 0 && warp < 4) row_stats[gr * 8 + 6] = (double)a_row;
      }
    } else if (warp == PRODUCER_WARP) {
      // ================= TMA producer of the X stream (one elected thread) =================
      // X and the B operand have their own producer warps: with one thread issuing both, the request for the B
      // tile of tile t sat behind the wait for a free X ring slot (the ring is full whenever HBM is the limit), so
      // B ran only ~2 tiles ahead of its MMAs whatever the number of B stages.  (An L2 prefetch of the X tiles ahead
      // of the ring, cp.async.bulk.prefetch.tensor, was measured and lost: 0.167 -> 0.196 / 0.242 ms at 8 / 16 tiles.)
      if (elect_one()) {
        for (int it = 0; it < ntiles; ++it) {
          const uint32_t u = use + it, slot = u % (uint32_t)ring, xph = (u / (uint32_t)ring) & 1;
          const int col0 = (int)((ct_begin + it) * TN);
          if (u >= (uint32_t)ring && !mbar_wait(&tl.x_free[slot], xph ^ 1, err)) break;
          mbar_arrive_expect_tx(&tl.x_full[slot], XSLOT_BYTES + TN * (uint32_t)sizeof(float));
          unsigned char* dst = x_base + slot * XSLOT_BYTES;
          tma_load_2d(dst, &maps.x, col0, row0, &tl.x_full[slot]);
          tma_load_2d(dst + X_BOX_BYTES, &maps.x, col0 + SLAB_K, row0, &tl.x_full[slot]);
          bulk_g2s(&tl.b_col[slot][0], bvec + col0, TN * (uint32_t)sizeof(float), &tl.x_full[slot]);   // bvec padded to 64
        }
      }
      __syncwarp();
    } else if (warp == BPRODUCER_WARP) {
      // ================= TMA producer of the B operand (one elected thread) =================
      if (probe != 1 && elect_one()) {
        for (int it = 0; it < ntiles; ++it) {
          const uint32_t u = use + it, bs = u % (uint32_t)nbst, bph = (u / (uint32_t)nbst) & 1;
          const int col0 = (int)((ct_begin + it) * TN);
          if (u >= (uint32_t)nbst && !mbar_wait(&tl.b_free[bs], bph ^ 1, err)) break;   // B stage free again
          mbar_arrive_expect_tx(&tl.b_full[bs], b_stage_bytes(nslab));
          unsigned char* bh = b_base + bs * b_stage_bytes(nslab);
          unsigned char* bl = bh + nslab * B_SLAB_BYTES;
          for (int sl = 0; sl < nslab; ++sl) {
            tma_load_2d(bh + sl * B_SLAB_BYTES, &maps.v_hi, sl * SLAB_K, col0, &tl.b_full[bs]);
            tma_load_2d(bl + sl * B_SLAB_BYTES, &maps.v_lo, sl * SLAB_K, col0, &tl.b_full[bs]);
          }
        }
      }
      __syncwarp();
    } else {
      // ================= MMA issuer (one elected thread) =================
      // FOUR accumulator stages: with two, the issuer could only start tile t+2 after the epilogue had drained tile
      // t, and the epilogue could only start a tile after its MMAs had completed -- the round-2 profile at K = 64
      // showed 28 % of all warp samples on the epilogue's mma_done wait while nothing was saturated (DRAM 39 %, tensor
      // pipe 29 %, issue 45 %) and neither more B stages nor a deeper X ring moved the time.  With four the MMAs run
      // up to three tiles ahead of the epilogue.
      if (probe != 1 && elect_one()) {
        bool ok = mbar_wait(&tl.a_full, item & 1, err);
        for (int it = 0; ok && it < ntiles; ++it) {
          const uint32_t u = use + it, st = u % NSTAGE, ph = (u / NSTAGE) & 1;
          const uint32_t bs = u % (uint32_t)nbst, bph = (u / (uint32_t)nbst) & 1;
          if (!mbar_wait(&tl.b_full[bs], bph, err)) break;
          if (u >= NSTAGE && !mbar_wait(&tl.tmem_free[st], ph ^ 1, err)) break;      // TMEM stage drained
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + st * TN;
          const uint32_t bh = smem_u32(b_base + bs * b_stage_bytes(nslab));
          // descriptors of the stage's first K step; a K step advances 32 bytes inside a 128-byte swizzle row
          // (2 units of 16 bytes in the address field), a slab advances B_SLAB_BYTES
          const uint64_t desc_hi = make_desc(bh);
          const uint64_t desc_lo = make_desc(bh + nslab * B_SLAB_BYTES);
          uint32_t accumulate = 0;
#pragma unroll 4
          for (int ks = 0; ks < (kp >> 3); ++ks) {
            const uint64_t off = (uint64_t)((ks >> 2) * (B_SLAB_BYTES >> 4) + (ks & 3) * 2);
            const uint32_t t_a_hi = tmem_base + TMEM_A_HI + ks * 8;      // 8 tf32 = 8 TMEM columns of A per K step
            const uint32_t t_a_lo = tmem_base + TMEM_A_LO + ks * 8;
            umma_tf32_ts(d_tmem, t_a_hi, desc_hi + off, idesc, accumulate);
            umma_tf32_ts(d_tmem, t_a_hi, desc_lo + off, idesc, 1u);
            umma_tf32_ts(d_tmem, t_a_lo, desc_hi + off, idesc, 1u);
            accumulate = 1u;
          }
          umma_commit(&tl.b_free[bs]);     // both arrive when the MMAs above have finished (implies fence::before_thread_sync):
          umma_commit(&tl.mma_done[st]);   // the B stage can be refilled, the accumulators can be read
        }
      }
      __syncwarp();
    }
    use += (uint32_t)ntiles;
    tc_fence_before();
    __syncthreads();                          // work-item boundary: all roles done with this row block
    tc_fence_after();
    if (tl.error) break;
  }

  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem_base, tmem_cols(nslab * SLAB_K));
  if (threadIdx.x == 0 && tl.error) atomicExch(error_flag, 1);
}

// ---- host: tensor maps through the driver entry point (no link-time dependency on libcuda) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// fp32 row-major [rows][cols] with leading dimension ld (elements); box = box_rows x 32 columns, 128-byte swizzle
static int make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return MFCD_ERR_UNSUPPORTED; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)SLAB_K, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return MFCD_ERR_ARG; }
  return MFCD_OK;
}

static inline size_t align_up(size_t x) { return (x + 255) & ~size_t(255); }
static int kpad_for(int d) { return (d + SLAB_K - 1) / SLAB_K * SLAB_K; }      // 32, 64, 96 or 128

struct TcLayout { size_t v_hi, v_lo, avec, bvec, total; };
static TcLayout tc_layout(int64_t n, int64_t m, int d) {
  TcLayout L; size_t off = 0; const int kpad = kpad_for(d);
  L.v_hi = off; off += align_up(sizeof(float) * m * kpad);
  L.v_lo = off; off += align_up(sizeof(float) * m * kpad);
  L.avec = off; off += align_up(sizeof(float) * n);
  L.bvec = off; off += align_up(sizeof(float) * ((m + TN - 1) / TN * TN));   // padded: bulk copies read whole tiles
  L.total = off;
  return L;
}

}  // namespace tc
}  // namespace mfcd

using namespace mfcd;

static bool tc_eligible(int64_t n, int64_t m, int d, const mfcd_xview* X) {
  return X && X->X != nullptr && d >= 1 && d <= tc::KMAX && (X->ldx & 3) == 0 &&
         (reinterpret_cast<uintptr_t>(X->X) & 15u) == 0 && n < (int64_t(1) << 31) && m < (int64_t(1) << 31);
}

extern "C" int mfcd_recon_stats_tc_workspace_bytes(int64_t n, int64_t m, int32_t d, size_t* bytes) {
  MFCD_REQUIRE(bytes && n >= 1 && m >= 1 && d >= 1, "mfcd_recon_stats_tc_workspace_bytes: bad argument");
  *bytes = d <= tc::KMAX ? tc::tc_layout(n, m, d).total : 0;
  return MFCD_OK;
}

extern "C" int mfcd_recon_stats_tc(const float* U, const float* V, int64_t n, int64_t m, int32_t d,
                                   const mfcd_xview* X, float s, const float* ubar, const float* vbar,
                                   double* row_stats, int32_t* error_flag, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  MFCD_REQUIRE(U && V && X && ubar && vbar && row_stats && error_flag, "mfcd_recon_stats_tc: NULL pointer");
  MFCD_REQUIRE(n >= 1 && m >= 1 && d >= 1, "mfcd_recon_stats_tc: bad sizes");
  if (!tc_eligible(n, m, d, X)) {
    set_error("mfcd_recon_stats_tc: shape not eligible (needs dense X with 16-byte aligned rows and d <= 128)");
    return MFCD_ERR_UNSUPPORTED;
  }
  const tc::TcLayout L = tc::tc_layout(n, m, d);
  if (workspace == nullptr || workspace_bytes < L.total) {
    set_error("mfcd_recon_stats_tc: workspace too small (%zu < %zu bytes)", workspace_bytes, L.total);
    return MFCD_ERR_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  char* base = static_cast<char*>(workspace);
  float* v_hi = reinterpret_cast<float*>(base + L.v_hi);
  float* v_lo = reinterpret_cast<float*>(base + L.v_lo);
  float* avec = reinterpret_cast<float*>(base + L.avec);
  float* bvec = reinterpret_cast<float*>(base + L.bvec);
  const int kpad = tc::kpad_for(d);
  const int kp = (d + 7) & ~7;

  tc::Maps maps;
  int rc;
  if ((rc = tc::make_map(&maps.x, X->X, n, m, X->ldx, tc::TM)) != MFCD_OK) return rc;
  if ((rc = tc::make_map(&maps.v_hi, v_hi, m, kpad, kpad, tc::TN)) != MFCD_OK) return rc;
  if ((rc = tc::make_map(&maps.v_lo, v_lo, m, kpad, kpad, tc::TN)) != MFCD_OK) return rc;

  const int64_t bvec_len = (m + tc::TN - 1) / tc::TN * tc::TN;
  tc::k_tc_prepass<<<grid_for((n + m) * kpad, 256, 8), 256, 0, st>>>(U, V, n, m, d, kpad, ubar, vbar, v_hi,
                                                                     v_lo, avec, bvec, bvec_len, row_stats, error_flag);
  MFCD_CHECK_LAUNCH();

  const int64_t row_tiles = (n + tc::TM - 1) / tc::TM;
  const int64_t col_tiles = (m + tc::TN - 1) / tc::TN;
  const int64_t sms = sm_count();
  int64_t splits = (2 * sms + row_tiles - 1) / row_tiles;       // >= 2 work items per SM when possible
  if (splits < 1) splits = 1;
  if (splits > col_tiles) splits = col_tiles;
  int64_t blocks = row_tiles * splits;
  if (blocks > sms) blocks = sms;                                // persistent: one CTA per SM
  const int nslab = kpad / tc::SLAB_K;
  // shared-memory split between B stages and X ring slots (both want ~3 tiles in flight), by operand size
  static const int kBst[5] = {0, 3, 3, 2, 2};     // with the even ring: 4 / 4 / 4 / 2 slots (measured best, r02 notes)
  int nbst = kBst[nslab];
  if (getenv("MFCD_K5_BSTAGES")) nbst = atoi(getenv("MFCD_K5_BSTAGES"));
  if (nbst < 2) nbst = 2;
  if (nbst > tc::MAX_BSTAGE) nbst = tc::MAX_BSTAGE;
  while (nbst > 2 && tc::smem_bytes(nslab, nbst, 2) > 232448u) --nbst;
  int ring = tc::MAX_RING;                                       // deepest X ring that fits in 227 KB
  if (getenv("MFCD_K5_RING")) ring = atoi(getenv("MFCD_K5_RING"));
  if (ring > tc::MAX_RING) ring = tc::MAX_RING;
  while (ring > 2 && tc::smem_bytes(nslab, nbst, ring) > 232448u) --ring;
  // The ring depth must be EVEN: the two epilogue groups take alternate tiles, so with an even ring every slot is
  // always served by the same group, in order.  With an odd ring the groups share slots, and a parity wait only
  // looks at the barrier's current phase bit: a group that runs ahead asks for "tile u + ring has landed" while
  // tile u (the other group's, same slot) is still in flight and passes at once -- stale data, double arrivals on
  // x_free and soon a trap (seen as intermittent launch failures at ring = 3 and 5).
  ring &= ~1;
  if (ring < 2) ring = 2;
  const size_t smem = tc::smem_bytes(nslab, nbst, ring);
  // MFCD_K5_PROBE (tools/bench_k5.py --probe): pipeline microbenchmarks whose RESULTS ARE INVALID -- 1: the X stream
  // alone (TMA ring + barriers), 2: X stream + MMAs without the epilogue arithmetic.  0 / unset: the real kernel.
  const int probe = getenv("MFCD_K5_PROBE") ? atoi(getenv("MFCD_K5_PROBE")) : 0;
  MFCD_CUDA(cudaFuncSetAttribute(tc::k_recon_stats_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc::k_recon_stats_tc<<<(int)blocks, tc::NTHREADS, smem, st>>>(maps, U, n, m, d, kp, nbst, ring, s, avec, bvec, probe,
                                                               row_stats, error_flag);
  MFCD_CHECK_LAUNCH();
  return MFCD_OK;
}
