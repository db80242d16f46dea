// K5: reconstruction statistics -- one pass over W = U V^T and X.
//
// Reference: compute_reconstruction_error (structure.py:939-955) and the dense
// part of compute_alpha_and_norm_ratios (structure.py:980-1064) materialise
// UV^T with torch.mm and then run python loops over its rows.  Here W is
// produced tile by tile in registers (K = d is small) and consumed on the spot
// together with the matching X tile, so the only large stream is X itself
// (4 n m bytes, HBM-bound; FLOPs 2 n m d).  Per-row sums are accumulated in
// fp64 (B200 has a full-rate-enough fp64 pipe) so that the derived quantities
// (alpha, ||alpha W - X||, Pearson, slopes) keep 1e-3 parity despite the
// cancellation in expressions like  alpha^2 Sww - 2 alpha Sxw + Sxx.
//
// This file holds the fp32 SIMT tile engine (exact fp32 FMA like the
// reference's sgemm); recon_stats_tc.cu holds the tcgen05/TMEM variant.
#include "internal.h"

namespace mfcd {

constexpr int TM = 64, TN = 64, TK = 32;
constexpr int PAD = 68;             // smem row pitch in floats (multiple of 4 for 128-bit reads)
constexpr int kThreads = 256;       // 16 x 16 threads, 4 x 4 outputs each

__global__ void k_col_means(const float* __restrict__ T, int64_t rows, int d, float* __restrict__ mean) {
  // one block per column chunk of 32; fp64 accumulation, fixed order
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int wy = threadIdx.x >> 5;          // 8 row lanes
  __shared__ double s[8][33];
  double acc = 0.0;
  if (c < d)
    for (int64_t r = wy; r < rows; r += 8) acc += (double)T[r * d + c];
  s[wy][threadIdx.x & 31] = acc;
  __syncthreads();
  if (wy == 0 && c < d) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += s[k][threadIdx.x & 31];
    mean[c] = (float)(t / (double)rows);     // torch.mean of an fp32 column
  }
}

// loads a TM x TK tile of a row-major [rows x ld] table, transposed, into smem[k][row]
__device__ __forceinline__ void load_tile_T(float (*dst)[PAD], const float* __restrict__ src, int64_t row0,
                                            int64_t rows, int k0, int kmax, int ld) {
  // 256 threads: lane = k (32 wide, coalesced along the row), 8 rows per pass
  const int kk = threadIdx.x & 31;
  const int r8 = threadIdx.x >> 5;
#pragma unroll
  for (int p = 0; p < TM / 8; ++p) {
    const int r = p * 8 + r8;
    const int64_t gr = row0 + r;
    float v = 0.f;
    if (gr < rows && (k0 + kk) < kmax) v = __ldg(src + gr * ld + k0 + kk);
    dst[kk][r] = v;
  }
}

// XMODE 0: dense X ; 1: low-rank X = scale * A B^T
template <int XMODE>
__global__ void __launch_bounds__(kThreads)
k_recon_stats(const float* __restrict__ U, const float* __restrict__ V, int64_t n, int64_t m, int d, mfcd_xview X,
              float s, const float* __restrict__ ubar, const float* __restrict__ vbar, int col_splits,
              double* __restrict__ row_stats) {
  __shared__ __align__(16) float Us[TK][PAD];
  __shared__ __align__(16) float Vs[TK][PAD];
  __shared__ __align__(16) float As[XMODE ? TK : 1][PAD];
  __shared__ __align__(16) float Bs[XMODE ? TK : 1][PAD];
  __shared__ float a_row[TM];       // <U_r, vbar>  = row mean of W
  __shared__ float b_col[TN];       // <ubar, V_c>  = column mean of W

  const int tx = threadIdx.x & 15;  // column quad
  const int ty = threadIdx.x >> 4;  // row quad
  const int64_t row_tiles = (n + TM - 1) / TM;
  const int64_t col_tiles = (m + TN - 1) / TN;

  for (int64_t work = blockIdx.x; work < row_tiles * col_splits; work += gridDim.x) {
    const int64_t rt = work / col_splits;
    const int split = (int)(work % col_splits);
    const int64_t row0 = rt * TM;
    const int64_t ct_begin = col_tiles * split / col_splits;
    const int64_t ct_end = col_tiles * (split + 1) / col_splits;

    __syncthreads();
    if (threadIdx.x < TM) {
      const int64_t gr = row0 + threadIdx.x;
      float acc = 0.f;
      if (gr < n)
        for (int k = 0; k < d; ++k) acc = fmaf(__ldg(U + gr * d + k), __ldg(vbar + k), acc);
      a_row[threadIdx.x] = acc;
    }

    double st[4][6];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int q = 0; q < 6; ++q) st[a][q] = 0.0;

    for (int64_t ct = ct_begin; ct < ct_end; ++ct) {
      const int64_t col0 = ct * TN;
      __syncthreads();
      if (threadIdx.x < TN) {
        const int64_t gc = col0 + threadIdx.x;
        float acc = 0.f;
        if (gc < m)
          for (int k = 0; k < d; ++k) acc = fmaf(__ldg(ubar + k), __ldg(V + gc * d + k), acc);
        b_col[threadIdx.x] = acc;
      }
      float w[4][4], xv[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) { w[a][b] = 0.f; xv[a][b] = 0.f; }

      for (int k0 = 0; k0 < d; k0 += TK) {
        __syncthreads();
        load_tile_T(Us, U, row0, n, k0, d, d);
        load_tile_T(Vs, V, col0, m, k0, d, d);
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < TK; ++kk) {
          const float4 ua = *reinterpret_cast<const float4*>(&Us[kk][ty * 4]);
          const float4 vb = *reinterpret_cast<const float4*>(&Vs[kk][tx * 4]);
          const float uu[4] = {ua.x, ua.y, ua.z, ua.w};
          const float vv[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) w[a][b] = fmaf(uu[a], vv[b], w[a][b]);
        }
      }
      if (XMODE == 1) {
        for (int k0 = 0; k0 < X.dx; k0 += TK) {
          __syncthreads();
          load_tile_T(As, X.A, row0, n, k0, X.dx, X.dx);
          load_tile_T(Bs, X.B, col0, m, k0, X.dx, X.dx);
          __syncthreads();
#pragma unroll 8
          for (int kk = 0; kk < TK; ++kk) {
            const float4 ua = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 vb = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float uu[4] = {ua.x, ua.y, ua.z, ua.w};
            const float vv[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
              for (int b = 0; b < 4; ++b) xv[a][b] = fmaf(uu[a], vv[b], xv[a][b]);
          }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) xv[a][b] *= X.scale;
      } else {
        const bool vec_ok = ((X.ldx & 3) == 0) && ((reinterpret_cast<uintptr_t>(X.X) & 15u) == 0) &&
                            (col0 + tx * 4 + 3 < m);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int64_t gr = row0 + ty * 4 + a;
          if (vec_ok && gr < n) {          // the X stream: one 128-bit load per row quad
            const float4 t = __ldg(reinterpret_cast<const float4*>(X.X + gr * X.ldx + col0 + tx * 4));
            xv[a][0] = t.x; xv[a][1] = t.y; xv[a][2] = t.z; xv[a][3] = t.w;
          } else {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              const int64_t gc = col0 + tx * 4 + b;
              xv[a][b] = (gr < n && gc < m) ? __ldg(X.X + gr * X.ldx + gc) : 0.f;
            }
          }
        }
      }
      // consume the tile
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int64_t gr = row0 + ty * 4 + a;
        const float ar = a_row[ty * 4 + a];
        float sx = 0.f, sxx = 0.f, sw = 0.f, sww = 0.f, sxw = 0.f, see = 0.f;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int64_t gc = col0 + tx * 4 + b;
          if (gr < n && gc < m) {
            const float x = xv[a][b];
            const float wa = w[a][b] - ar;                    // row-centred UV^T  (structure.py:985)
            const float e = (w[a][b] - b_col[tx * 4 + b]) - s * x;   // column-centred minus sX (:943-949)
            sx += x; sxx = fmaf(x, x, sxx);
            sw += wa; sww = fmaf(wa, wa, sww); sxw = fmaf(x, wa, sxw);
            see = fmaf(e, e, see);
          }
        }
        st[a][0] += (double)sx; st[a][1] += (double)sxx; st[a][2] += (double)sw;
        st[a][3] += (double)sww; st[a][4] += (double)sxw; st[a][5] += (double)see;
      }
    }

    // reduce the 16 column-quad threads of each row (lanes tx = 0..15 are contiguous in a half-warp)
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        double v = st[a][q];
#pragma unroll
        for (int off = 8; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        st[a][q] = v;
      }
      const int64_t gr = row0 + ty * 4 + a;
      if (tx == 0 && gr < n) {
#pragma unroll
        for (int q = 0; q < 6; ++q) atomicAdd(row_stats + gr * 8 + q, st[a][q]);
        if (split == 0) row_stats[gr * 8 + 6] = (double)a_row[ty * 4 + a];
      }
    }
  }
}

__global__ void __launch_bounds__(256)
k_reconstruct_rows(const float* __restrict__ U, const float* __restrict__ V, int64_t r0, int64_t nr, int64_t m,
                   int d, float* __restrict__ out) {
  // small helper (a handful of rows at a time): out[r][c] = <U[r0+r], V[c]>
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < nr * m; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / m, c = idx % m;
    const float* pu = U + (r0 + r) * d;
    const float* pv = V + c * d;
    float acc = 0.f;
    for (int k = 0; k < d; ++k) acc = fmaf(__ldg(pu + k), __ldg(pv + k), acc);
    out[idx] = acc;
  }
}

__global__ void __launch_bounds__(256)
k_xview_rows(mfcd_xview X, int64_t r0, int64_t nr, int64_t m, float* __restrict__ out) {
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < nr * m; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / m, c = idx % m;
    out[idx] = xview_at(X, r0 + r, c);
  }
}

}  // namespace mfcd

using namespace mfcd;

extern "C" int mfcd_table_col_means(const float* T, int64_t rows, int32_t d, float* mean, void* stream) {
  MFCD_REQUIRE(T && mean && rows >= 1 && d >= 1, "mfcd_table_col_means: bad argument");
  k_col_means<<<(d + 31) / 32, 256, 0, as_stream(stream)>>>(T, rows, d, mean);
  MFCD_CHECK_LAUNCH();
  return MFCD_OK;
}

extern "C" int mfcd_recon_stats(const float* U, const float* V, int64_t n, int64_t m, int32_t d,
                                const mfcd_xview* X, float s, const float* ubar, const float* vbar,
                                double* row_stats, void* stream) {
  MFCD_REQUIRE(U && V && X && ubar && vbar && row_stats, "mfcd_recon_stats: NULL pointer");
  MFCD_REQUIRE(n >= 1 && m >= 1 && d >= 1, "mfcd_recon_stats: bad sizes");
  MFCD_REQUIRE(X->X || (X->A && X->B && X->dx >= 1), "mfcd_recon_stats: bad xview");
  cudaStream_t st = as_stream(stream);
  MFCD_CUDA(cudaMemsetAsync(row_stats, 0, sizeof(double) * 8 * n, st));
  const int64_t row_tiles = (n + TM - 1) / TM;
  const int64_t col_tiles = (m + TN - 1) / TN;
  const int64_t target = (int64_t)sm_count() * 8;                // ~4 waves at 2 CTAs/SM
  int64_t splits = (target + row_tiles - 1) / row_tiles;
  if (splits < 1) splits = 1;
  if (splits > col_tiles) splits = col_tiles;
  int64_t blocks = row_tiles * splits;
  if (blocks > target * 4) blocks = target * 4;
  if (X->X)
    k_recon_stats<0><<<(int)blocks, kThreads, 0, st>>>(U, V, n, m, d, *X, s, ubar, vbar, (int)splits, row_stats);
  else
    k_recon_stats<1><<<(int)blocks, kThreads, 0, st>>>(U, V, n, m, d, *X, s, ubar, vbar, (int)splits, row_stats);
  MFCD_CHECK_LAUNCH();
  return MFCD_OK;
}

extern "C" int mfcd_reconstruct_rows(const float* U, const float* V, int64_t r0, int64_t nr, int64_t m, int32_t d,
                                     float* out, void* stream) {
  MFCD_REQUIRE(U && V && out && r0 >= 0 && nr >= 0 && m >= 1 && d >= 1, "mfcd_reconstruct_rows: bad argument");
  if (nr == 0) return MFCD_OK;
  k_reconstruct_rows<<<grid_for(nr * m, 256, 8), 256, 0, as_stream(stream)>>>(U, V, r0, nr, m, d, out);
  MFCD_CHECK_LAUNCH();
  return MFCD_OK;
}

extern "C" int mfcd_xview_rows(const mfcd_xview* X, int64_t r0, int64_t nr, int64_t m, float* out, void* stream) {
  MFCD_REQUIRE(X && out && r0 >= 0 && nr >= 0 && m >= 1, "mfcd_xview_rows: bad argument");
  MFCD_REQUIRE(X->X || (X->A && X->B && X->dx >= 1), "mfcd_xview_rows: bad xview");
  if (nr == 0) return MFCD_OK;
  k_xview_rows<<<grid_for(nr * m, 256, 8), 256, 0, as_stream(stream)>>>(*X, r0, nr, m, out);
  MFCD_CHECK_LAUNCH();
  return MFCD_OK;
}
