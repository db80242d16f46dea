// K1-fix: the deterministic fused forward + BCE + backward for LARGE batches, without a sort.
//
// The reference's CPU index_put_(accumulate) adds a row's contributions in batch order.  For batches of up to 256
// triplets k_det_small reproduces that order exactly.  For large batches the first implementation sorted (row, b)
// pairs three times per step with cub::DeviceRadixSort and ran a chunked segmented reduction (segmented.cu:
// 17 launches, 0.6 - 1.7 ms at 2^20 triplets).  This kernel gets bit-reproducible gradients another way: every
// contribution g * x is rounded ONCE to 64-bit fixed point (2^-shift units) and added with integer atomics --
// integer addition is associative, so the sum does not depend on the order in which the atomics land, on the grid
// or on what else runs on the GPU.  k_fix_finish converts the sums to fp32 (one rounding per element; the result
// is at least as close to the exact sum as any fp32 summation order, the reference's included) and adds them to
// the gradient tables.  Two launches per step instead of seventeen, no library call on the step path.
//
// Arithmetic per triplet is the reference's (structure.py:787-795, :849-850; exact sigmoid, the 1e-12 clamp of
// binary_cross_entropy_backward); the batch loss is accumulated the same way (2^-32 units, exact).
#include "internal.h"
#include "shape_dispatch.cuh"

namespace mfcd {

constexpr int kFixBlock = 256;

// 2^shift: |sum per element| <= 2 B inv_batch max|x|; with max|x| < 2^10 the sums stay below 2^62
static inline int fix_shift(int64_t B, float inv_batch) {
  double bound = 2.0 * (double)B * (double)inv_batch * 1024.0;      // bound on |sum|
  int shift = 62;
  while (shift > 0 && bound * (double)(1ull << shift) >= 4.6e18) --shift;
  return shift > 52 ? 52 : shift;
}

__device__ __forceinline__ void fix_add(long long* dst, float c, float scale) {
  const long long f = __float2ll_rn(c * scale);
  if (f != 0) atomicAdd(reinterpret_cast<unsigned long long*>(dst), (unsigned long long)f);
}

template <int VEC, int LPT, int NITER>
__global__ void __launch_bounds__(kFixBlock)
k_fwd_bwd_fix(const float* __restrict__ U, const float* __restrict__ V, const mfcd_triplet* __restrict__ rec,
              const int32_t* __restrict__ perm, int64_t start, int64_t B, int d, float inv_batch, float scale,
              long long* __restrict__ fixU, long long* __restrict__ fixV, long long* __restrict__ loss_fix) {
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPT;
  const int grp = lane / LPT;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  long long loss_acc = 0;
  for (int64_t base = warp0 * 32; base < B; base += nwarps * 32) {
    const int64_t k = base + lane;
    int4 r = make_int4(0, 0, 0, 0);
    if (k < B) {
      const int64_t idx = perm ? (int64_t)__ldg(perm + start + k) : (start + k);
      r = __ldg(reinterpret_cast<const int4*>(rec) + idx);
    }
    const int nvalid = (B - base) < 32 ? (int)(B - base) : 32;
    float x_home = 0.f;
#pragma unroll 2
    for (int rr = 0; rr < LPT; ++rr) {
      const int e = grp * LPT + rr;
      const int tu = __shfl_sync(0xffffffffu, r.x, e);
      const int ti = __shfl_sync(0xffffffffu, r.y, e);
      const int tj = __shfl_sync(0xffffffffu, r.z, e);
      const float tz = __int_as_float(__shfl_sync(0xffffffffu, r.w, e));
      const bool ok = e < nvalid;
      TripletRows<VEC, LPT, NITER> rows;
      load_rows<VEC, LPT, NITER>(rows, U, V, tu, ti, tj, d, sub, ok);
      const float x = group_sum<LPT>(partial_dot<VEC, LPT, NITER>(rows), 0xffffffffu);
      x_home = (sub == rr) ? x : x_home;
      const float g = ok ? bce_grad_score_ref(sigmoidf_ref(x), tz, inv_batch) : 0.f;
      if (g != 0.f) {
        long long* du = fixU + (int64_t)tu * d;
        long long* di = fixV + (int64_t)ti * d;
        long long* dj = fixV + (int64_t)tj * d;
#pragma unroll
        for (int it = 0; it < NITER; ++it) {
          const int c0 = (it * LPT + sub) * VEC;
#pragma unroll
          for (int kk = 0; kk < VEC; ++kk) {
            const int c = c0 + kk;
            if (c < d) {
              const float b = g * rows.uu[it].v[kk];
              fix_add(du + c, g * rows.dv[it].v[kk], scale);
              fix_add(di + c, b, scale);
              fix_add(dj + c, -b, scale);
            }
          }
        }
      }
    }
    if (lane < nvalid) loss_acc += __float2ll_rn(bce_ref(sigmoidf_ref(x_home), __int_as_float(r.w)) * 4294967296.f);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, off);
  if (lane == 0 && loss_acc != 0) atomicAdd(reinterpret_cast<unsigned long long*>(loss_fix), (unsigned long long)loss_acc);
}

// grads += fixed-point sums (one rounding per element); *loss += batch loss * inv_batch
__global__ void __launch_bounds__(256)
k_fix_finish(const long long* __restrict__ fixU, const long long* __restrict__ fixV, int64_t nU, int64_t nV,
             double inv_scale, float* __restrict__ gU, float* __restrict__ gV, const long long* __restrict__ loss_fix,
             float inv_batch, float* __restrict__ loss) {
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = tid; e < nU; e += nth) {
    const long long f = fixU[e];
    if (f != 0) gU[e] += (float)((double)f * inv_scale);
  }
  for (int64_t e = tid; e < nV; e += nth) {
    const long long f = fixV[e];
    if (f != 0) gV[e] += (float)((double)f * inv_scale);
  }
  if (tid == 0) *loss += (float)((double)(*loss_fix) * (1.0 / 4294967296.0)) * inv_batch;
}

template <int VEC, int LPT, int NITER>
struct FixLauncher {
  static int run(const float* U, const float* V, const mfcd_triplet* rec, const int32_t* perm, int64_t start, int64_t B,
                 int d, float inv_batch, float scale, long long* fixU, long long* fixV, long long* loss_fix,
                 cudaStream_t st) {
    const int grid = grid_for(B, kFixBlock, 6);
    k_fwd_bwd_fix<VEC, LPT, NITER><<<grid, kFixBlock, 0, st>>>(U, V, rec, perm, start, B, d, inv_batch, scale, fixU, fixV,
                                                               loss_fix);
    MFCD_CHECK_LAUNCH();
    return MFCD_OK;
  }
};

static int dispatch_fix(const float* U, const float* V, const mfcd_triplet* rec, const int32_t* perm, int64_t start,
                        int64_t B, int d, float inv_batch, float scale, long long* fixU, long long* fixV,
                        long long* loss_fix, cudaStream_t st) {
  MFCD_DISPATCH_ROW_SHAPE(FixLauncher, d, U, V, rec, perm, start, B, d, inv_batch, scale, fixU, fixV, loss_fix, st);
}

size_t det_fixed_workspace_bytes(int64_t n_users, int64_t n_items, int d) {
  return sizeof(long long) * (size_t)((n_users + n_items) * (int64_t)d + 32);
}

int launch_det_fixed(const float* U, const float* V, const mfcd_triplet* rec, const int32_t* perm, int64_t start,
                     int64_t B, int d, float inv_batch, int64_t n_users, int64_t n_items, float* gU, float* gV,
                     float* loss, void* ws, cudaStream_t st) {
  const int64_t nU = n_users * d, nV = n_items * d;
  long long* fixU = static_cast<long long*>(ws);
  long long* fixV = fixU + nU;
  long long* loss_fix = fixV + nV;
  MFCD_CUDA(cudaMemsetAsync(ws, 0, det_fixed_workspace_bytes(n_users, n_items, d), st));
  const int shift = fix_shift(B, inv_batch);
  const float scale = (float)(double)(1ull << shift);
  const int rc = dispatch_fix(U, V, rec, perm, start, B, d, inv_batch, scale, fixU, fixV, loss_fix, st);
  if (rc != MFCD_OK) return rc;
  k_fix_finish<<<grid_for(nU + nV, 256, 8), 256, 0, st>>>(fixU, fixV, nU, nV, 1.0 / (double)(1ull << shift), gU, gV,
                                                         loss_fix, inv_batch, loss);
  MFCD_CHECK_LAUNCH();
  return MFCD_OK;
}

}  // namespace mfcd
