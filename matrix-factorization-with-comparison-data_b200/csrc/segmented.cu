// K2: deterministic gradient scatter for large batches --
// "sort by destination row, then segmented reduction".
//
// For each of the three destination index streams of a batch (u -> gU,
// i -> gV with +g, j -> gV with -g; reference: the three index_put_(accumulate)
// of the autograd backward of structure.py:787-792):
//   1. (row, b) pairs are sorted by row with a STABLE LSD radix sort
//      (cub::DeviceRadixSort -- a library call, like cuBLAS for a plain GEMM),
//      so inside a row the contributions stay in batch order;
//   2. k_gather_meta writes one 16-byte record per sorted position
//      {row, src row(s), g_b};
//   3. k_seg_reduce: a lane group walks a chunk of kChunk sorted entries,
//      summing runs of equal row in registers (sequentially, in batch order);
//      runs that live entirely inside the chunk are added to the gradient
//      table by their single owner, runs that cross a chunk edge leave a
//      partial row;
//   4. k_seg_fixup: the chunk where a straddling run starts adds up its
//      partials in chunk order and writes the row.
// No atomics anywhere, fixed summation order => bit-reproducible.
#include <cub/device/device_radix_sort.cuh>
#include "internal.h"
#include "shape_dispatch.cuh"

namespace mfcd {

constexpr int kChunk = 32;
constexpr int kLongSpan = 48;     // straddling runs longer than this many chunks get a whole CTA
constexpr int kSegBlock = 256;
constexpr int kMaxLossPartials = 4096;

int launch_det_forward(const float* U, const float* V, const mfcd_triplet* rec, const int32_t* perm, int64_t start,
                       int64_t B, int d, float inv_batch, float* gbuf, float* partials, int grid, cudaStream_t st);

static inline size_t align_up(size_t x) { return (x + 255) & ~size_t(255); }

struct DetLayout {
  size_t gbuf, partials, keys_a, keys_b, vals_a, vals_b, meta, part_first, part_last, longlist, cub_temp, total;
  size_t cub_bytes;
};

static DetLayout det_layout(int64_t B, int d) {
  DetLayout L;
  size_t off = 0;
  const int64_t nchunks = (B + kChunk - 1) / kChunk;
  L.gbuf = off;       off += align_up(sizeof(float) * B);
  L.partials = off;   off += align_up(sizeof(float) * kMaxLossPartials);
  L.keys_a = off;     off += align_up(sizeof(int32_t) * B);
  L.keys_b = off;     off += align_up(sizeof(int32_t) * B);
  L.vals_a = off;     off += align_up(sizeof(int32_t) * B);
  L.vals_b = off;     off += align_up(sizeof(int32_t) * B);
  L.meta = off;       off += align_up(sizeof(int4) * (B + 1));
  L.part_first = off; off += align_up(sizeof(float) * nchunks * d);
  L.part_last = off;  off += align_up(sizeof(float) * nchunks * d);
  // long straddling runs (> kLongSpan chunks, each owning >= kLongSpan-1 chunks exclusively); [0] is the counter
  L.longlist = off;   off += align_up(sizeof(int64_t) * (2 * (nchunks / (kLongSpan / 2) + 4) + 2));
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const int32_t*)nullptr, (int32_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, B, 0, 32, (cudaStream_t)0);
  L.cub_bytes = cub_bytes;
  L.cub_temp = off;   off += align_up(cub_bytes);
  L.total = off;
  return L;
}

size_t det_large_workspace_bytes(int64_t B, int d) { return det_layout(B, d).total; }

__global__ void k_loss_finish(const float* __restrict__ partials, int n, float inv_batch, float* loss) {
  float t = 0.f;
  for (int k = threadIdx.x; k < n; k += 32) t += partials[k];
  t = warp_sum(t);
  if (threadIdx.x == 0) *loss += t * inv_batch;
}

// SIDE 0: key = u ; 1: key = i ; 2: key = j
template <int SIDE>
__global__ void k_extract_keys(const mfcd_triplet* __restrict__ rec, const int32_t* __restrict__ perm,
                               int64_t start, int64_t B, int32_t* __restrict__ keys, int32_t* __restrict__ vals) {
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
    const int64_t idx = perm ? (int64_t)__ldg(perm + start + b) : (start + b);
    const int4 r = __ldg(reinterpret_cast<const int4*>(rec) + idx);
    keys[b] = SIDE == 0 ? r.x : (SIDE == 1 ? r.y : r.z);
    vals[b] = (int32_t)b;
  }
}

template <int SIDE>
__global__ void k_gather_meta(const mfcd_triplet* __restrict__ rec, const int32_t* __restrict__ perm,
                              int64_t start, int64_t B, const int32_t* __restrict__ sorted_b,
                              const float* __restrict__ gbuf, int4* __restrict__ meta) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < B; t += (int64_t)gridDim.x * blockDim.x) {
    const int b = sorted_b[t];
    const int64_t idx = perm ? (int64_t)__ldg(perm + start + b) : (start + b);
    const int4 r = __ldg(reinterpret_cast<const int4*>(rec) + idx);
    const float g = gbuf[b];
    int4 mt;
    if (SIDE == 0) { mt.x = r.x; mt.y = r.y; mt.z = r.z; mt.w = __float_as_int(g); }
    else if (SIDE == 1) { mt.x = r.y; mt.y = r.x; mt.z = 0; mt.w = __float_as_int(g); }
    else { mt.x = r.z; mt.y = r.x; mt.z = 0; mt.w = __float_as_int(-g); }
    meta[t] = mt;
  }
}

// SIDE_U: contributions g*(T[y]-T[z]) from table T = V into gU; else g*T[y] with T = U into gV.
template <int VEC, int LPT, int NITER, bool SIDE_U>
__global__ void __launch_bounds__(kSegBlock)
k_seg_reduce(const float* __restrict__ T, const int4* __restrict__ meta, int64_t B, int d,
             float* __restrict__ out, float* __restrict__ part_first, float* __restrict__ part_last) {
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPT;
  const int64_t group0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / LPT;
  const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) / LPT;
  const int64_t nchunks = (B + kChunk - 1) / kChunk;

  for (int64_t ch = group0; ch < nchunks; ch += ngroups) {
    const int64_t s = ch * kChunk;
    const int64_t e = (s + kChunk < B) ? (s + kChunk) : B;
    const bool head_cont = (s > 0) && (meta[s - 1].x == meta[s].x);
    const bool tail_cont = (e < B) && (meta[e].x == meta[e - 1].x);
    Frag<VEC> acc[NITER];
#pragma unroll
    for (int it = 0; it < NITER; ++it) acc[it] = frag_zero<VEC>();
    int cur = meta[s].x;
    bool first_run = true;

    auto flush = [&](bool last_run) {
      const bool from_prev = first_run && head_cont;
      const bool to_next = last_run && tail_cont;
#pragma unroll
      for (int it = 0; it < NITER; ++it) {
        const int c = (it * LPT + sub) * VEC;
        if (c >= d) continue;
        if (!from_prev && !to_next) {
          float* dst = out + (int64_t)cur * d + c;
          Frag<VEC> v = ld_frag<VEC>(dst);
#pragma unroll
          for (int kk = 0; kk < VEC; ++kk) v.v[kk] += acc[it].v[kk];
          st_frag<VEC>(dst, v);
        } else if (from_prev) {
          st_frag<VEC>(part_first + ch * d + c, acc[it]);
        } else {
          st_frag<VEC>(part_last + ch * d + c, acc[it]);
        }
        acc[it] = frag_zero<VEC>();
      }
    };

#pragma unroll 4
    for (int64_t t = s; t < e; ++t) {
      const int4 mt = meta[t];
      if (mt.x != cur) {
        flush(false);
        cur = mt.x;
        first_run = false;
      }
      const float g = __int_as_float(mt.w);
#pragma unroll
      for (int it = 0; it < NITER; ++it) {
        const int c = (it * LPT + sub) * VEC;
        if (c >= d) continue;
        if (SIDE_U) {
          Frag<VEC> a = ldg_frag<VEC>(T + (int64_t)mt.y * d + c);
          Frag<VEC> b = ldg_frag<VEC>(T + (int64_t)mt.z * d + c);
#pragma unroll
          for (int kk = 0; kk < VEC; ++kk) acc[it].v[kk] += g * (a.v[kk] - b.v[kk]);
        } else {
          Frag<VEC> a = ldg_frag<VEC>(T + (int64_t)mt.y * d + c);
#pragma unroll
          for (int kk = 0; kk < VEC; ++kk) acc[it].v[kk] += g * a.v[kk];
        }
      }
    }
    flush(true);
  }
}

// last chunk that still belongs to the run of `key` which continues past chunk `ch`
__device__ __forceinline__ int64_t run_last_chunk(const int4* __restrict__ meta, int64_t B, int64_t from, int key) {
  int64_t lo = from, hi = B;                 // first position in [from, B) whose key differs (keys are sorted)
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (meta[mid].x == key) lo = mid + 1; else hi = mid;
  }
  return (lo - 1) / kChunk;
}

template <int VEC, int LPT, int NITER>
__global__ void __launch_bounds__(kSegBlock)
k_seg_fixup(const int4* __restrict__ meta, int64_t B, int d, float* __restrict__ out,
            const float* __restrict__ part_first, const float* __restrict__ part_last,
            unsigned long long* __restrict__ longlist) {
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPT;
  const int64_t group0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / LPT;
  const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) / LPT;
  const int64_t nchunks = (B + kChunk - 1) / kChunk;
  for (int64_t ch = group0; ch < nchunks; ch += ngroups) {
    const int64_t s = ch * kChunk;
    const int64_t e = (s + kChunk < B) ? (s + kChunk) : B;
    const int key = meta[e - 1].x;
    const bool tail_cont = (e < B) && (meta[e].x == key);
    if (!tail_cont) continue;
    const bool single = meta[s].x == key;
    const bool head_cont = (s > 0) && (meta[s - 1].x == meta[s].x);
    if (single && head_cont) continue;              // the run started in an earlier chunk
    const int64_t last = run_last_chunk(meta, B, e, key);     // chunks ch+1 .. last hold a part_first row of this run
    if (last - ch > kLongSpan) {                    // hot row: hand it to a whole CTA (k_seg_fixup_long)
      if (sub == 0) {
        const unsigned long long slot = atomicAdd(longlist, 1ull);
        longlist[1 + 2 * slot] = (unsigned long long)ch;
        longlist[2 + 2 * slot] = (unsigned long long)last;
      }
      continue;
    }
#pragma unroll
    for (int it = 0; it < NITER; ++it) {
      const int c = (it * LPT + sub) * VEC;
      if (c >= d) continue;
      Frag<VEC> tot = ld_frag<VEC>(part_last + ch * d + c);
      for (int64_t c2 = ch + 1; c2 <= last; ++c2) {
        Frag<VEC> pf = ld_frag<VEC>(part_first + c2 * d + c);
#pragma unroll
        for (int kk = 0; kk < VEC; ++kk) tot.v[kk] += pf.v[kk];
      }
      float* dst = out + (int64_t)key * d + c;
      Frag<VEC> v = ld_frag<VEC>(dst);
#pragma unroll
      for (int kk = 0; kk < VEC; ++kk) v.v[kk] += tot.v[kk];
      st_frag<VEC>(dst, v);
    }
  }
}

// One CTA per long run: column sums of the contiguous block part_first[ch+1 .. last] in a fixed
// order (row r is summed by row-lane r % RL sequentially, lanes combined in index order), plus part_last[ch].
__global__ void __launch_bounds__(256)
k_seg_fixup_long(const int4* __restrict__ meta, int d, float* __restrict__ out, const float* __restrict__ part_first,
                 const float* __restrict__ part_last, const unsigned long long* __restrict__ longlist) {
  extern __shared__ float s_part[];                 // [RL][d]
  const unsigned long long count = longlist[0];
  const int cols = d < 256 ? d : 256;               // columns handled per pass
  const int RL = 256 / cols;                        // row lanes
  for (unsigned long long w = blockIdx.x; w < count; w += gridDim.x) {
    const int64_t ch = (int64_t)longlist[1 + 2 * w];
    const int64_t last = (int64_t)longlist[2 + 2 * w];
    const int key = meta[(ch + 1) * kChunk].x;
    for (int c0 = 0; c0 < d; c0 += cols) {
      const int c = c0 + (threadIdx.x % cols);
      const int rl = threadIdx.x / cols;
      float acc = 0.f;
      if (rl < RL && c < d)
        for (int64_t r = ch + 1 + rl; r <= last; r += RL) acc += part_first[r * d + c];
      __syncthreads();
      if (rl < RL && c < d) s_part[rl * cols + (c - c0)] = acc;
      __syncthreads();
      if (rl == 0 && c < d) {
        float tot = part_last[ch * d + c];
        for (int q = 0; q < RL; ++q) tot += s_part[q * cols + (c - c0)];
        out[(int64_t)key * d + c] += tot;
      }
    }
  }
}

template <int VEC, int LPT, int NITER>
struct SegLauncher {
  static int run(bool side_u, const float* T, const int4* meta, int64_t B, int d, float* out, float* pf, float* pl,
                 unsigned long long* longlist, cudaStream_t st) {
    const int64_t nchunks = (B + kChunk - 1) / kChunk;
    const int grid = grid_for(nchunks, kSegBlock / LPT, 8);
    if (side_u) k_seg_reduce<VEC, LPT, NITER, true><<<grid, kSegBlock, 0, st>>>(T, meta, B, d, out, pf, pl);
    else k_seg_reduce<VEC, LPT, NITER, false><<<grid, kSegBlock, 0, st>>>(T, meta, B, d, out, pf, pl);
    MFCD_CHECK_LAUNCH();
    MFCD_CUDA(cudaMemsetAsync(longlist, 0, sizeof(unsigned long long), st));
    k_seg_fixup<VEC, LPT, NITER><<<grid, kSegBlock, 0, st>>>(meta, B, d, out, pf, pl, longlist);
    MFCD_CHECK_LAUNCH();
    const int long_grid = (int)((nchunks / kLongSpan + 1) < 512 ? (nchunks / kLongSpan + 1) : 512);
    const int cols = d < 256 ? d : 256;
    k_seg_fixup_long<<<long_grid, 256, sizeof(float) * (256 / cols) * cols, st>>>(meta, d, out, pf, pl, longlist);
    MFCD_CHECK_LAUNCH();
    return MFCD_OK;
  }
};

static int launch_seg(bool side_u, const float* T, const int4* meta, int64_t B, int d, float* out, float* pf,
                      float* pl, unsigned long long* longlist, cudaStream_t st) {
  MFCD_DISPATCH_ROW_SHAPE(SegLauncher, d, side_u, T, meta, B, d, out, pf, pl, longlist, st);
}

static int bits_for(int64_t n) {
  int b = 1;
  while (b < 32 && (int64_t(1) << b) < n) ++b;
  return b;
}

int launch_det_large(const float* U, const float* V, const mfcd_triplet* rec, const int32_t* perm, int64_t start,
                     int64_t B, int d, float inv_batch, int64_t n_users, int64_t n_items, float* gU, float* gV,
                     float* loss, void* ws, size_t ws_bytes, cudaStream_t st) {
  (void)ws_bytes;
  const DetLayout L = det_layout(B, d);
  char* base = static_cast<char*>(ws);
  float* gbuf = reinterpret_cast<float*>(base + L.gbuf);
  float* partials = reinterpret_cast<float*>(base + L.partials);
  int32_t* keys_a = reinterpret_cast<int32_t*>(base + L.keys_a);
  int32_t* keys_b = reinterpret_cast<int32_t*>(base + L.keys_b);
  int32_t* vals_a = reinterpret_cast<int32_t*>(base + L.vals_a);
  int32_t* vals_b = reinterpret_cast<int32_t*>(base + L.vals_b);
  int4* meta = reinterpret_cast<int4*>(base + L.meta);
  float* pf = reinterpret_cast<float*>(base + L.part_first);
  float* pl = reinterpret_cast<float*>(base + L.part_last);
  void* cub_temp = base + L.cub_temp;
  unsigned long long* longlist = reinterpret_cast<unsigned long long*>(base + L.longlist);

  // forward with a FIXED grid so the loss partials reduce in a fixed order
  int grid = (int)((B + 255) / 256);
  if (grid > 1024) grid = 1024;
  int rc = launch_det_forward(U, V, rec, perm, start, B, d, inv_batch, gbuf, partials, grid, st);
  if (rc != MFCD_OK) return rc;
  k_loss_finish<<<1, 32, 0, st>>>(partials, grid, inv_batch, loss);
  MFCD_CHECK_LAUNCH();

  const int eg = grid_for(B, 256, 8);
  for (int side = 0; side < 3; ++side) {
    if (side == 0) k_extract_keys<0><<<eg, 256, 0, st>>>(rec, perm, start, B, keys_a, vals_a);
    else if (side == 1) k_extract_keys<1><<<eg, 256, 0, st>>>(rec, perm, start, B, keys_a, vals_a);
    else k_extract_keys<2><<<eg, 256, 0, st>>>(rec, perm, start, B, keys_a, vals_a);
    MFCD_CHECK_LAUNCH();
    size_t cub_bytes = L.cub_bytes;
    const int end_bit = bits_for(side == 0 ? n_users : n_items);
    MFCD_CUDA(cub::DeviceRadixSort::SortPairs(cub_temp, cub_bytes, (const int32_t*)keys_a, keys_b,
                                              (const int32_t*)vals_a, vals_b, B, 0, end_bit, st));
    if (side == 0) k_gather_meta<0><<<eg, 256, 0, st>>>(rec, perm, start, B, vals_b, gbuf, meta);
    else if (side == 1) k_gather_meta<1><<<eg, 256, 0, st>>>(rec, perm, start, B, vals_b, gbuf, meta);
    else k_gather_meta<2><<<eg, 256, 0, st>>>(rec, perm, start, B, vals_b, gbuf, meta);
    MFCD_CHECK_LAUNCH();
    rc = launch_seg(side == 0, side == 0 ? V : U, meta, B, d, side == 0 ? gU : gV, pf, pl, longlist, st);
    if (rc != MFCD_OK) return rc;
  }
  return MFCD_OK;
}

}  // namespace mfcd
