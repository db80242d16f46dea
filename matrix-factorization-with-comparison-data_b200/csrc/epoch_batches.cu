// Epoch batching: the reshuffle of an epoch and the per-batch user grouping in ONE streaming pass.
//
// The reference draws a new random order of the training set every epoch (DataLoader(shuffle=True) ->
// RandomSampler -> randperm, structure.py:738) and cuts it into batches of B (structure.py:845).  Only the
// batch MEMBERSHIP matters to the step (a batch-mean gradient is a sum); the order inside a batch does not.
// So instead of materialising the permutation, gathering 16-byte records at random and then sorting every
// batch by user (what K1's user-run path wants), this file does a stable multisplit:
//
//   * record r of the store has epoch position pos(r): pos[r] from memory (the inverse of a given
//     permutation -- reference RNG mode), or a keyed bijection of [0, N) evaluated on the fly (see ShuffleKey;
//     device RNG mode);
//   * batch(r) = pos(r) / B -- exactly the batches a loader walking the permutation in chunks of B forms,
//     every batch has B members (the last one N - (nb-1) B);
//   * out[] receives batch after batch, each batch's records in STORE order.  When the store is sorted by
//     user (done once per dataset) every batch comes out user-grouped for free (MFCD_FLAG_USER_GROUPED).
//
// Two kernels over the store + a scan of the (batch x warp-segment) histogram:
//   k_epoch_count    warp w derives the batch of every record of its segment, stores it (1 - 2 bytes per record)
//                    and counts the records per batch
//   scan             exclusive prefix sum of hist[batch][segment] = final output offsets  (3 small kernels)
//   k_epoch_scatter  warp w ranks each record among the equal-batch lanes of its 32-record tile (match.any) and
//                    stores it at offset[batch] + rank
// HBM traffic: 16 N read + 16 N written + 2 - 4 N for the batch ids (+ 4 N for pos[] when given) -- close to the
// floor of any out-of-place shuffle.
#include "internal.h"

namespace mfcd {

constexpr int kEpochMaxBatches = 1024;     // counters per warp in shared memory (4 KB per warp, 32 KB per CTA)
constexpr int kEpochBlock = 256;
constexpr int kEpochWarps = kEpochBlock / 32;

// Keyed bijection of [0, M), M = c * 2^k >= N with c <= 64 (so M - N < M / 32: a pass leaves [0, N) with
// probability < 3 %, and the cycle walk below almost never iterates).  A value is split into hi in [0, c) and
// lo in [0, 2^k); three rounds of
//     lo <- ((lo ^ f_r(hi)) * odd_r  mod 2^k) ^ (itself >> k/2), + key_r  mod 2^k     (a bijection of lo for fixed hi)
//     hi <- hi + g_r(lo)  mod c,  g_r(lo) = mulhi(hash32(lo + key_r), c)               (a bijection of hi for fixed lo)
// -- every step is invertible, so the composition is a permutation of [0, M).
// Why this shape: the first version of this file walked an 8-round Feistel network over 2^ceil(log2 N) values
// (~100 instructions per pass); the second a 4-round multiply / xor-shift chain (~25).  Both reject up to half of
// their outputs, and a WARP repeats the pass until its slowest lane is inside [0, N): 5.6 - 9 passes per 32 records,
// which left k_epoch_count and k_epoch_scatter ALU-bound (profiles/r02_notes.md: 380 - 410 warp instructions per
// tile, 72 % issue-active, 0 - 33 % DRAM).  With a domain that hugs N the walk disappears, and the batch ids are
// computed ONCE (k_epoch_count stores them for k_epoch_scatter).
struct ShuffleKey {
  uint32_t k, c;                      // lo bits (>= 1), hi range (1..64)
  uint32_t fmul[3], lmul[3], ladd[3], hadd[3];
};

__host__ __device__ __forceinline__ uint32_t shuffle_once(uint32_t x, const ShuffleKey& K) {
  const uint32_t mask = (1u << K.k) - 1u;
  const uint32_t sh = (K.k >> 1) ? (K.k >> 1) : 1u;
  uint32_t hi = x >> K.k, lo = x & mask;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const uint32_t f = ((hi + 1u) * K.fmul[r]) >> (32u - K.k);
    lo = ((lo ^ f) * K.lmul[r]) & mask;
    lo ^= lo >> sh;
    lo = (lo + K.ladd[r]) & mask;
    uint32_t h = (lo + K.hadd[r]) * 0x9E3779B1u;
    h ^= h >> 15;
    h *= 0x85EBCA77u;
#ifdef __CUDA_ARCH__
    hi += __umulhi(h, K.c);
#else
    hi += (uint32_t)(((uint64_t)h * K.c) >> 32);
#endif
    if (hi >= K.c) hi -= K.c;
  }
  return (hi << K.k) | lo;
}

// bijection of [0, N): walk the cycle of the permutation of [0, M) until it re-enters [0, N)
__host__ __device__ __forceinline__ uint32_t epoch_position(uint32_t r, uint32_t N, const ShuffleKey& K) {
  uint32_t x = shuffle_once(r, K);
  while (x >= N) x = shuffle_once(x, K);
  return x;
}

static ShuffleKey make_key(int64_t N, uint64_t seed) {
  ShuffleKey K;
  int bits = 2;
  while ((int64_t(1) << bits) < N) ++bits;
  const int k = (bits / 2) > (bits - 6) ? (bits / 2) : (bits - 6);
  K.k = (uint32_t)k;
  K.c = (uint32_t)((N + (int64_t(1) << k) - 1) >> k);
  if (K.c < 1) K.c = 1;
  uint32_t w[12];
  uint64_t s = seed;
  for (int r = 0; r < 12; ++r) {         // splitmix64 stream -> 12 key words (bits 16..47 of each output)
    s += 0x9E3779B97F4A7C15ull;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    w[r] = (uint32_t)(z >> 16);
  }
  for (int r = 0; r < 3; ++r) {
    K.fmul[r] = w[3 * r] | 1u;
    K.lmul[r] = (w[3 * r + 1] << 1) | 1u;            // odd multipliers
    K.ladd[r] = w[3 * r + 2];
    K.hadd[r] = w[9 + r];
  }
  return K;
}

struct EpochPlan {
  int64_t N, B;
  float inv_B;       // 1 / B: pos / B by a float multiply + one correction step (the quotient is < 2^10)
  int nb;            // batches
  int seg;           // records per warp segment
  int64_t n_seg;     // warp segments
  int64_t hist_len;  // nb * n_seg
  int64_t n_tiles;   // scan tiles of kScanTile counters
  int wide;          // batch ids stored as uint16 (more than 256 batches) instead of uint8
};
constexpr int kScanTile = 4096;

static bool make_plan(int64_t N, int64_t B, EpochPlan* P) {
  if (N < 0 || B < 1 || N >= (int64_t(1) << 31)) return false;
  const int64_t nb = N == 0 ? 0 : (N + B - 1) / B;
  if (nb > kEpochMaxBatches) return false;
  P->N = N; P->B = B; P->nb = (int)nb;
  P->inv_B = 1.0f / (float)B;
  int seg = 2048;                                   // 64 tiles of 32 records per warp
  while (seg < 8 * nb) seg <<= 1;                   // keep the histogram below N/8 counters (<= 8192 records)
  P->seg = seg;
  P->n_seg = (N + seg - 1) / seg;
  P->hist_len = nb * P->n_seg;
  P->n_tiles = (P->hist_len + kScanTile - 1) / kScanTile;
  P->wide = nb > 256;
  return true;
}

// pos / B for pos < 2^31 and a quotient below kEpochMaxBatches: float estimate (off by at most one), one correction
__device__ __forceinline__ uint32_t div_by_batch(uint32_t p, uint32_t B, float inv_B) {
  uint32_t q = (uint32_t)(__uint2float_rz(p) * inv_B);
  const int32_t r = (int32_t)(p - q * B);
  if (r < 0) --q;
  else if ((uint32_t)r >= B) ++q;
  return q;
}

template <bool HAVE_POS>
__device__ __forceinline__ uint32_t batch_of(int64_t r, const int32_t* __restrict__ pos, const EpochPlan& P,
                                             const ShuffleKey& K) {
  const uint32_t p = HAVE_POS ? (uint32_t)__ldg(pos + r) : epoch_position((uint32_t)r, (uint32_t)P.N, K);
  return div_by_batch(p, (uint32_t)P.B, P.inv_B);
}

// Counting pass, up to kEpochLaneCounters batches: every LANE keeps its own 16-bit counter per batch in shared
// memory ([batch][lane]: conflict-free), so a record costs its hash plus one shared-memory increment -- no
// match.any, no leader election, no warp barrier per tile (those made the first version MIO-bound: 7 cycles of
// short-scoreboard stall per issue).  The counters are summed over the lanes once per segment.
constexpr int kEpochLaneCounters = 64;

template <bool HAVE_POS, typename IdT>
__global__ void __launch_bounds__(kEpochBlock)
k_epoch_count_lanes(const int32_t* __restrict__ pos, EpochPlan P, ShuffleKey K, uint32_t* __restrict__ hist,
                    IdT* __restrict__ ids) {
  extern __shared__ uint32_t s_dyn[];
  unsigned short* s_cnt = reinterpret_cast<unsigned short*>(s_dyn);      // [kEpochWarps][nb][32]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned short* cnt = s_cnt + (size_t)warp * P.nb * 32 + lane;
  for (int64_t sgm = blockIdx.x * (int64_t)kEpochWarps + warp; sgm < P.n_seg; sgm += (int64_t)gridDim.x * kEpochWarps) {
    for (int b = 0; b < P.nb; ++b) cnt[b * 32] = 0;
    const int64_t r0 = sgm * P.seg;
    const int64_t r1 = (r0 + P.seg) < P.N ? (r0 + P.seg) : P.N;
#pragma unroll 2
    for (int64_t r = r0 + lane; r < r1; r += 32) {
      const uint32_t b = batch_of<HAVE_POS>(r, pos, P, K);
      ids[r] = (IdT)b;                                      // computed once: k_epoch_scatter reads it back
      cnt[b * 32] += 1;                                     // segments hold at most 2^15 records (make_plan)
    }
    for (int b = 0; b < P.nb; ++b) {
      uint32_t v = cnt[b * 32];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if (lane == 0) hist[(int64_t)b * P.n_seg + sgm] = v;
    }
  }
}

template <bool HAVE_POS, typename IdT>
__global__ void __launch_bounds__(kEpochBlock)
k_epoch_count(const int32_t* __restrict__ pos, EpochPlan P, ShuffleKey K, uint32_t* __restrict__ hist,
              IdT* __restrict__ ids) {
  extern __shared__ uint32_t s_cnt[];                       // [kEpochWarps][nb]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* cnt = s_cnt + warp * P.nb;
  for (int64_t sgm = blockIdx.x * (int64_t)kEpochWarps + warp; sgm < P.n_seg; sgm += (int64_t)gridDim.x * kEpochWarps) {
    for (int b = lane; b < P.nb; b += 32) cnt[b] = 0;
    __syncwarp();
    const int64_t r0 = sgm * P.seg;
    const int64_t r1 = (r0 + P.seg) < P.N ? (r0 + P.seg) : P.N;
    for (int64_t t = r0; t < r1; t += 32) {
      const int64_t r = t + lane;
      uint32_t b = 0xffffffffu;
      if (r < r1) {
        b = batch_of<HAVE_POS>(r, pos, P, K);
        ids[r] = (IdT)b;                                    // computed once: k_epoch_scatter reads it back
      }
      const uint32_t same = __match_any_sync(0xffffffffu, b);
      if (b != 0xffffffffu && lane == __ffs(same) - 1) cnt[b] += __popc(same);   // one leader per distinct batch
      __syncwarp();
    }
    for (int b = lane; b < P.nb; b += 32) hist[(int64_t)b * P.n_seg + sgm] = cnt[b];
    __syncwarp();
  }
}

template <typename IdT>
__global__ void __launch_bounds__(kEpochBlock)
k_epoch_scatter(const mfcd_triplet* __restrict__ rec, const IdT* __restrict__ ids, EpochPlan P,
                const uint32_t* __restrict__ offs, mfcd_triplet* __restrict__ out) {
  extern __shared__ uint32_t s_base[];                      // [kEpochWarps][nb]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t* base = s_base + warp * P.nb;
  for (int64_t sgm = blockIdx.x * (int64_t)kEpochWarps + warp; sgm < P.n_seg; sgm += (int64_t)gridDim.x * kEpochWarps) {
    for (int b = lane; b < P.nb; b += 32) base[b] = offs[(int64_t)b * P.n_seg + sgm];
    __syncwarp();
    const int64_t r0 = sgm * P.seg;
    const int64_t r1 = (r0 + P.seg) < P.N ? (r0 + P.seg) : P.N;
    // kTiles tiles in flight per warp: all their loads are issued before the first one is ranked and stored (with one
    // tile ahead the kernel sat at 15 % issue-active and 31 % of DRAM peak, every warp waiting on a single load)
    constexpr int kTiles = 4;
    for (int64_t t0 = r0; t0 < r1; t0 += 32 * kTiles) {
      int4 v[kTiles];
      uint32_t b[kTiles];
#pragma unroll
      for (int q = 0; q < kTiles; ++q) {
        const int64_t r = t0 + q * 32 + lane;
        v[q] = make_int4(0, 0, 0, 0);
        b[q] = 0xffffffffu;
        if (r < r1) {
          v[q] = __ldg(reinterpret_cast<const int4*>(rec) + r);
          b[q] = (uint32_t)__ldg(ids + r);
        }
      }
#pragma unroll
      for (int q = 0; q < kTiles; ++q) {
        const bool ok = b[q] != 0xffffffffu;
        const uint32_t same = __match_any_sync(0xffffffffu, b[q]);
        uint32_t dst = 0;
        if (ok) dst = base[b[q]] + __popc(same & lt);       // stable: lower lanes = earlier records
        __syncwarp();
        if (ok && lane == __ffs(same) - 1) base[b[q]] += __popc(same);
        if (ok) reinterpret_cast<int4*>(out)[dst] = v[q];
        __syncwarp();
      }
    }
  }
}

// ---- exclusive prefix sum of the histogram (in place), three small kernels -----------------------------
__global__ void __launch_bounds__(256) k_scan_tile_sums(const uint32_t* __restrict__ a, int64_t n, uint32_t* __restrict__ sums) {
  __shared__ uint32_t s_w[8];
  const int64_t t0 = (int64_t)blockIdx.x * kScanTile;
  uint32_t acc = 0;
  for (int e = threadIdx.x; e < kScanTile; e += 256) {
    const int64_t k = t0 + e;
    if (k < n) acc += a[k];
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int q = 0; q < 8; ++q) t += s_w[q];
    sums[blockIdx.x] = t;
  }
}

// block-wide exclusive scan of one value per thread (1024 threads); returns the exclusive prefix, *total = block sum
__device__ __forceinline__ uint32_t block_excl_scan_1024(uint32_t v, uint32_t* s_w, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += t;
  }
  if (lane == 31) s_w[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = s_w[lane], wi = w;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, wi, off);
      if (lane >= off) wi += t;
    }
    s_w[lane] = wi - w;                 // exclusive prefix of the warp sums
    if (lane == 31) s_w[32] = wi;
  }
  __syncthreads();
  const uint32_t out = s_w[warp] + incl - v;
  *total = s_w[32];
  __syncthreads();
  return out;
}

__global__ void __launch_bounds__(1024) k_scan_sums(uint32_t* __restrict__ sums, int64_t n) {
  __shared__ uint32_t s_w[33];
  uint32_t carry = 0;
  for (int64_t c0 = 0; c0 < n; c0 += 1024) {
    const int64_t k = c0 + threadIdx.x;
    const uint32_t v = k < n ? sums[k] : 0u;
    uint32_t total;
    const uint32_t ex = block_excl_scan_1024(v, s_w, &total);
    if (k < n) sums[k] = carry + ex;
    carry += total;
  }
}

__global__ void __launch_bounds__(1024) k_scan_tiles(uint32_t* __restrict__ a, int64_t n, const uint32_t* __restrict__ sums) {
  __shared__ uint32_t s_w[33];
  const int64_t t0 = (int64_t)blockIdx.x * kScanTile;
  uint32_t carry = sums[blockIdx.x];
  // each thread owns 4 consecutive counters of the tile
  const int64_t k = t0 + threadIdx.x * 4;
  uint32_t v[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) v[q] = (k + q) < n ? a[k + q] : 0u;
  uint32_t total;
  const uint32_t ex = block_excl_scan_1024(v[0] + v[1] + v[2] + v[3], s_w, &total);
  uint32_t run = carry + ex;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if ((k + q) < n) a[k + q] = run;
    run += v[q];
  }
}

__global__ void k_epoch_positions(int64_t N, ShuffleKey K, int32_t* __restrict__ pos) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < N; r += (int64_t)gridDim.x * blockDim.x)
    pos[r] = (int32_t)epoch_position((uint32_t)r, (uint32_t)N, K);
}

__global__ void k_invert_perm(const int32_t* __restrict__ perm, int64_t N, int32_t* __restrict__ pos) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < N; k += (int64_t)gridDim.x * blockDim.x)
    pos[perm[k]] = (int32_t)k;
}

static size_t epoch_hist_bytes(const EpochPlan& P) {
  return (sizeof(uint32_t) * (size_t)(P.hist_len + P.n_tiles + 8) + 255) & ~size_t(255);
}
static size_t epoch_ws_bytes(const EpochPlan& P) {
  return epoch_hist_bytes(P) + (size_t)P.N * (P.wide ? 2 : 1) + 256;
}

}  // namespace mfcd

using namespace mfcd;

extern "C" int mfcd_epoch_max_batches(int32_t* out) {
  MFCD_REQUIRE(out != nullptr, "mfcd_epoch_max_batches: out is NULL");
  *out = kEpochMaxBatches;
  return MFCD_OK;
}

extern "C" int mfcd_epoch_positions(int64_t N, uint64_t seed, int32_t* pos, void* stream) {
  MFCD_REQUIRE(N >= 0 && N < (int64_t(1) << 31), "mfcd_epoch_positions: N out of range");
  if (N == 0) return MFCD_OK;
  MFCD_REQUIRE(pos != nullptr, "mfcd_epoch_positions: pos is NULL");
  k_epoch_positions<<<grid_for(N, 256, 8), 256, 0, as_stream(stream)>>>(N, make_key(N, seed), pos);
  MFCD_CHECK_LAUNCH();
  return MFCD_OK;
}

extern "C" int mfcd_invert_perm(const int32_t* perm, int64_t N, int32_t* pos, void* stream) {
  MFCD_REQUIRE(N >= 0, "mfcd_invert_perm: N < 0");
  if (N == 0) return MFCD_OK;
  MFCD_REQUIRE(perm && pos, "mfcd_invert_perm: NULL pointer");
  k_invert_perm<<<grid_for(N, 256, 8), 256, 0, as_stream(stream)>>>(perm, N, pos);
  MFCD_CHECK_LAUNCH();
  return MFCD_OK;
}

extern "C" int mfcd_epoch_batches_workspace(int64_t N, int64_t B, size_t* bytes) {
  MFCD_REQUIRE(bytes != nullptr, "mfcd_epoch_batches_workspace: bytes is NULL");
  EpochPlan P;
  if (!make_plan(N, B, &P)) {
    set_error("mfcd_epoch_batches: needs 0 <= N < 2^31, B >= 1 and at most %d batches per epoch (N=%lld, B=%lld)",
              kEpochMaxBatches, (long long)N, (long long)B);
    return MFCD_ERR_UNSUPPORTED;
  }
  *bytes = epoch_ws_bytes(P);
  return MFCD_OK;
}

extern "C" int mfcd_epoch_batches(const mfcd_triplet* rec, int64_t N, int64_t B, const int32_t* pos, uint64_t seed,
                                  mfcd_triplet* out, void* workspace, size_t workspace_bytes, void* stream) {
  EpochPlan P;
  if (!make_plan(N, B, &P)) {
    set_error("mfcd_epoch_batches: needs 0 <= N < 2^31, B >= 1 and at most %d batches per epoch (N=%lld, B=%lld)",
              kEpochMaxBatches, (long long)N, (long long)B);
    return MFCD_ERR_UNSUPPORTED;
  }
  if (N == 0) return MFCD_OK;
  MFCD_REQUIRE(rec && out && rec != out, "mfcd_epoch_batches: NULL pointer, or in-place call (out must differ from rec)");
  if (workspace == nullptr || workspace_bytes < epoch_ws_bytes(P)) {
    set_error("mfcd_epoch_batches: workspace too small (%zu < %zu bytes)", workspace_bytes, epoch_ws_bytes(P));
    return MFCD_ERR_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  uint32_t* hist = static_cast<uint32_t*>(workspace);
  uint32_t* sums = hist + P.hist_len;
  const ShuffleKey K = make_key(N, seed);
  const size_t smem = sizeof(uint32_t) * (size_t)kEpochWarps * P.nb;
  const int grid = grid_for(P.n_seg, kEpochWarps, 8);
  void* ids = static_cast<char*>(workspace) + epoch_hist_bytes(P);
  if (P.nb <= kEpochLaneCounters) {                      // nb <= 64 implies 8-bit ids
    const size_t smem_l = sizeof(unsigned short) * (size_t)kEpochWarps * P.nb * 32;
    if (pos) k_epoch_count_lanes<true, uint8_t><<<grid, kEpochBlock, smem_l, st>>>(pos, P, K, hist, static_cast<uint8_t*>(ids));
    else k_epoch_count_lanes<false, uint8_t><<<grid, kEpochBlock, smem_l, st>>>(pos, P, K, hist, static_cast<uint8_t*>(ids));
  } else if (P.wide) {
    if (pos) k_epoch_count<true, uint16_t><<<grid, kEpochBlock, smem, st>>>(pos, P, K, hist, static_cast<uint16_t*>(ids));
    else k_epoch_count<false, uint16_t><<<grid, kEpochBlock, smem, st>>>(pos, P, K, hist, static_cast<uint16_t*>(ids));
  } else {
    if (pos) k_epoch_count<true, uint8_t><<<grid, kEpochBlock, smem, st>>>(pos, P, K, hist, static_cast<uint8_t*>(ids));
    else k_epoch_count<false, uint8_t><<<grid, kEpochBlock, smem, st>>>(pos, P, K, hist, static_cast<uint8_t*>(ids));
  }
  MFCD_CHECK_LAUNCH();
  k_scan_tile_sums<<<(unsigned)P.n_tiles, 256, 0, st>>>(hist, P.hist_len, sums);
  MFCD_CHECK_LAUNCH();
  k_scan_sums<<<1, 1024, 0, st>>>(sums, P.n_tiles);
  MFCD_CHECK_LAUNCH();
  k_scan_tiles<<<(unsigned)P.n_tiles, 1024, 0, st>>>(hist, P.hist_len, sums);
  MFCD_CHECK_LAUNCH();
  if (P.wide) k_epoch_scatter<uint16_t><<<grid, kEpochBlock, smem, st>>>(rec, static_cast<const uint16_t*>(ids), P, hist, out);
  else k_epoch_scatter<uint8_t><<<grid, kEpochBlock, smem, st>>>(rec, static_cast<const uint8_t*>(ids), P, hist, out);
  MFCD_CHECK_LAUNCH();
  return MFCD_OK;
}
