// K6: per-row ranks (average ties) and per-row Pearson, for Spearman.
//
// Reference: the python loop of scipy.stats.spearmanr over the n rows of X and
// UV^T (structure.py:1024-1031).  spearmanr(a, b) is the Pearson correlation of
// rankdata(a, 'average') and rankdata(b, 'average').
//
// Rows are sorted as segments of one flat array (cub::DeviceSegmentedSort --
// library sort, like cuBLAS for a plain GEMM); the rank of an element is then
// read off the sorted row: with lb/ub the lower/upper bound of its value,
// rank = (lb + ub + 1) / 2  (1-based, ties share the mean of their positions).
#include <cub/device/device_segmented_sort.cuh>
#include "internal.h"

namespace mfcd {

static inline size_t align_up(size_t x) { return (x + 255) & ~size_t(255); }

struct RankLayout {
  size_t keys_out, idx_in, idx_out, offsets, cub_temp, total, cub_bytes;
};

static RankLayout rank_layout(int64_t rows, int64_t m) {
  RankLayout L;
  size_t off = 0;
  const int64_t N = rows * m;
  L.keys_out = off; off += align_up(sizeof(float) * N);
  L.idx_in = off;   off += align_up(sizeof(int32_t) * N);
  L.idx_out = off;  off += align_up(sizeof(int32_t) * N);
  L.offsets = off;  off += align_up(sizeof(int64_t) * (rows + 1));
  size_t cb = 0;
  cub::DeviceSegmentedSort::SortPairs(nullptr, cb, (const float*)nullptr, (float*)nullptr, (const int32_t*)nullptr,
                                      (int32_t*)nullptr, N, rows, (const int64_t*)nullptr, (const int64_t*)nullptr,
                                      (cudaStream_t)0);
  L.cub_bytes = cb;
  L.cub_temp = off; off += align_up(cb);
  L.total = off;
  return L;
}

__global__ void k_rank_prepare(int64_t rows, int64_t m, int32_t* __restrict__ idx, int64_t* __restrict__ offsets) {
  const int64_t N = rows * m;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < N; k += (int64_t)gridDim.x * blockDim.x)
    idx[k] = (int32_t)(k % m);
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r <= rows; r += (int64_t)gridDim.x * blockDim.x)
    offsets[r] = r * m;
}

__global__ void __launch_bounds__(256)
k_rank_scatter(const float* __restrict__ sorted, const int32_t* __restrict__ sidx, int64_t rows, int64_t m,
               float* __restrict__ ranks) {
  const int64_t N = rows * m;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < N; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = k / m;
    const int64_t t = k - r * m;
    const float* row = sorted + r * m;
    const float v = row[t];
    int64_t lb = t, ub = t + 1;
    if (t > 0 && row[t - 1] == v) {          // tie to the left: lower bound by bisection
      int64_t lo = 0, hi = t;
      while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (row[mid] < v) lo = mid + 1; else hi = mid; }
      lb = lo;
    }
    if (t + 1 < m && row[t + 1] == v) {      // tie to the right: upper bound by bisection
      int64_t lo = t + 1, hi = m;
      while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (row[mid] <= v) lo = mid + 1; else hi = mid; }
      ub = lo;
    }
    ranks[r * m + sidx[k]] = 0.5f * (float)(lb + ub + 1);
  }
}

// one block per row (grid-stride over rows), fp64 sums
__global__ void __launch_bounds__(256)
k_row_pearson(const float* __restrict__ a, const float* __restrict__ b, int64_t rows, int64_t m,
              double* __restrict__ rho) {
  __shared__ double s[5][8];
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const float* pa = a + r * m;
    const float* pb = b + r * m;
    double sa = 0, sb = 0, saa = 0, sbb = 0, sab = 0;
    for (int64_t c = threadIdx.x; c < m; c += blockDim.x) {
      const double x = (double)pa[c], y = (double)pb[c];
      sa += x; sb += y; saa += x * x; sbb += y * y; sab += x * y;
    }
    sa = warp_sum(sa); sb = warp_sum(sb); saa = warp_sum(saa); sbb = warp_sum(sbb); sab = warp_sum(sab);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) { s[0][w] = sa; s[1][w] = sb; s[2][w] = saa; s[3][w] = sbb; s[4][w] = sab; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double t[5] = {0, 0, 0, 0, 0};
      for (int q = 0; q < 5; ++q)
        for (int k = 0; k < 8; ++k) t[q] += s[q][k];
      const double M = (double)m;
      const double cov = t[4] - t[0] * t[1] / M;
      const double va = t[2] - t[0] * t[0] / M;
      const double vb = t[3] - t[1] * t[1] / M;
      rho[r] = cov / sqrt(va * vb);          // NaN when a row is constant, like scipy
    }
  }
}

}  // namespace mfcd

using namespace mfcd;

extern "C" int mfcd_rank_workspace_bytes(int64_t rows, int64_t m, size_t* bytes) {
  MFCD_REQUIRE(bytes && rows >= 0 && m >= 1, "mfcd_rank_workspace_bytes: bad argument");
  MFCD_REQUIRE(m < (int64_t(1) << 31), "mfcd_rank_workspace_bytes: m must fit int32");
  *bytes = rank_layout(rows, m).total;
  return MFCD_OK;
}

extern "C" int mfcd_row_ranks(const float* vals, int64_t rows, int64_t m, float* ranks, void* workspace,
                              size_t workspace_bytes, void* stream) {
  MFCD_REQUIRE(rows >= 0 && m >= 1 && m < (int64_t(1) << 31), "mfcd_row_ranks: bad sizes");
  if (rows == 0) return MFCD_OK;
  MFCD_REQUIRE(vals && ranks, "mfcd_row_ranks: NULL pointer");
  const RankLayout L = rank_layout(rows, m);
  if (workspace == nullptr || workspace_bytes < L.total) {
    set_error("mfcd_row_ranks: workspace too small (%zu < %zu bytes)", workspace_bytes, L.total);
    return MFCD_ERR_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  char* base = static_cast<char*>(workspace);
  float* keys_out = reinterpret_cast<float*>(base + L.keys_out);
  int32_t* idx_in = reinterpret_cast<int32_t*>(base + L.idx_in);
  int32_t* idx_out = reinterpret_cast<int32_t*>(base + L.idx_out);
  int64_t* offsets = reinterpret_cast<int64_t*>(base + L.offsets);
  const int64_t N = rows * m;
  k_rank_prepare<<<grid_for(N, 256, 8), 256, 0, st>>>(rows, m, idx_in, offsets);
  MFCD_CHECK_LAUNCH();
  size_t cb = L.cub_bytes;
  MFCD_CUDA(cub::DeviceSegmentedSort::SortPairs(base + L.cub_temp, cb, vals, keys_out, (const int32_t*)idx_in,
                                                idx_out, N, rows, (const int64_t*)offsets,
                                                (const int64_t*)(offsets + 1), st));
  k_rank_scatter<<<grid_for(N, 256, 8), 256, 0, st>>>(keys_out, idx_out, rows, m, ranks);
  MFCD_CHECK_LAUNCH();
  return MFCD_OK;
}

extern "C" int mfcd_row_pearson(const float* a, const float* b, int64_t rows, int64_t m, double* rho,
                                void* stream) {
  MFCD_REQUIRE(rows >= 0 && m >= 1, "mfcd_row_pearson: bad sizes");
  if (rows == 0) return MFCD_OK;
  MFCD_REQUIRE(a && b && rho, "mfcd_row_pearson: NULL pointer");
  const int grid = (int)(rows < (int64_t)sm_count() * 8 ? rows : (int64_t)sm_count() * 8);
  k_row_pearson<<<grid, 256, 0, as_stream(stream)>>>(a, b, rows, m, rho);
  MFCD_CHECK_LAUNCH();
  return MFCD_OK;
}
