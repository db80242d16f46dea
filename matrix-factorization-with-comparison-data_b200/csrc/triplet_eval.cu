// K4: evaluation kernels (no gradients).
//   mfcd_triplet_eval       evaluate_model (structure.py:896-921) and the validation
//                           pass of train_model (structure.py:858-868)
//   mfcd_ground_truth_eval  compute_ground_truth_metrics (structure.py:1100-1127)
//   mfcd_triplet_scores     MatrixFactorization.forward (structure.py:773-795)
// The reference reports "mean over batches of the batch-mean loss", so the
// kernels produce one mean per batch of `batch_size` consecutive records.
#include "internal.h"
#include "shape_dispatch.cuh"

namespace mfcd {

constexpr int kEvalBlock = 256;

// Per-batch loss accumulation that is EXACT and order-independent for any batch size: every sample's loss is
// converted to 64-bit fixed point (2^-32 units: exact for fp32 values >= 2^-9, absolute error <= 2^-33 below that) and integer sums are associative, so the batch sums do not depend on how the grid, concurrent
// kernels or the atomics interleave -- a sweep that runs experiments concurrently reports the same bits as the
// sequential one.  A warp keeps a running sum for the batch it is currently in (tiles are 32 consecutive
// positions) and issues ONE 64-bit atomic when the batch index changes: 64-sample batches cost two atomics, a
// 4M-sample batch one per warp.  k_eval_finish turns the sums into fp32 batch means.
// fixed-point scale: 2^32 for batches of up to 2^24 samples (100 -- the BCE clamp -- * 2^24 * 2^32 < 2^63), one bit
// less for every doubling of the batch beyond that
static inline int fix_shift_for(int64_t batch_size) {
  int shift = 32;
  for (int64_t cap = int64_t(1) << 24; cap < batch_size && shift > 0; cap <<= 1) --shift;
  return shift;
}
struct BatchAcc {
  int64_t batch;
  long long sum;
};
__device__ __forceinline__ long long to_fix(float v, float scale) { return __float2ll_rn(v * scale); }
__device__ __forceinline__ long long warp_sum_ll(long long x) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
  return x;
}
__device__ __forceinline__ void batch_acc_flush(BatchAcc& a, long long* __restrict__ batch_fix, int lane) {
  if (a.batch >= 0) {
    const long long t = warp_sum_ll(a.sum);
    if (lane == 0 && t != 0)
      atomicAdd(reinterpret_cast<unsigned long long*>(batch_fix + a.batch), (unsigned long long)t);
  }
  a.batch = -1;
  a.sum = 0;
}
// adds this lane's value (for position pos, valid or not) to the warp's running batch sum
__device__ __forceinline__ void batch_acc_add(BatchAcc& a, float v, int64_t pos, bool valid,
                                              long long* __restrict__ batch_fix, int64_t batch_size, float scale,
                                              int lane) {
  const int64_t b = valid ? pos / batch_size : -1;
  const int64_t b0 = __shfl_sync(0xffffffffu, b, 0);
  const bool uniform = __all_sync(0xffffffffu, b == b0 || !valid) && b0 >= 0;
  if (uniform) {                      // the common case: the whole tile lies in one batch
    if (b0 != a.batch) batch_acc_flush(a, batch_fix, lane), a.batch = b0;
    a.sum += valid ? to_fix(v, scale) : 0;
  } else {                            // tile straddles batches (or is the ragged end): per-lane atomics
    batch_acc_flush(a, batch_fix, lane);
    if (valid) atomicAdd(reinterpret_cast<unsigned long long*>(batch_fix + b), (unsigned long long)to_fix(v, scale));
  }
}
// batch_out[b] = (sum of batch b) / (its sample count), rounded once to fp32
__global__ void k_eval_finish(const long long* __restrict__ batch_fix, int64_t N, int64_t batch_size, int64_t nb,
                              double inv_scale, float* __restrict__ batch_out) {
  const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const int64_t cnt = (b == nb - 1) ? (N - b * batch_size) : batch_size;
  batch_out[b] = (float)((double)batch_fix[b] * inv_scale / (double)cnt);
}
static inline int launch_eval_finish(const long long* batch_fix, int64_t N, int64_t batch_size, float* batch_out,
                                     cudaStream_t st) {
  const int64_t nb = (N + batch_size - 1) / batch_size;
  k_eval_finish<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(batch_fix, N, batch_size, nb,
                                                               1.0 / (double)(int64_t(1) << fix_shift_for(batch_size)), batch_out);
  MFCD_CHECK_LAUNCH();
  return MFCD_OK;
}

template <int VEC, int LPT, int NITER>
__global__ void __launch_bounds__(kEvalBlock)
k_eval(const float* __restrict__ U, const float* __restrict__ V, const mfcd_triplet* __restrict__ rec, int64_t N,
       int d, int64_t batch_size, float scale, long long* __restrict__ batch_fix,
       unsigned long long* __restrict__ correct) {
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPT;
  const int grp = lane / LPT;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  unsigned int hits = 0;
  BatchAcc acc{-1, 0};

  for (int64_t base = warp0 * 32; base < N; base += nwarps * 32) {
    const int64_t k = base + lane;
    int4 r = make_int4(0, 0, 0, 0);
    if (k < N) r = __ldg(reinterpret_cast<const int4*>(rec) + k);
    const int nvalid = (N - base) < 32 ? (int)(N - base) : 32;
    float x_home = 0.f;               // score of the slot this lane loaded (see k1_rounds in triplet_fwd_bwd.cu)
#pragma unroll 2
    for (int rr = 0; rr < LPT; ++rr) {
      const int e = grp * LPT + rr;
      const int tu = __shfl_sync(0xffffffffu, r.x, e);
      const int ti = __shfl_sync(0xffffffffu, r.y, e);
      const int tj = __shfl_sync(0xffffffffu, r.z, e);
      const bool ok = e < nvalid;
      TripletRows<VEC, LPT, NITER> rows;
      load_rows<VEC, LPT, NITER>(rows, U, V, tu, ti, tj, d, sub, ok);
      const float x = group_sum<LPT>(partial_dot<VEC, LPT, NITER>(rows), 0xffffffffu);
      x_home = (sub == rr) ? x : x_home;
    }
    const bool valid = lane < nvalid;
    const float z = __int_as_float(r.w);
    const float p = sigmoidf_ref(x_home);
    batch_acc_add(acc, bce_ref(p, z), k, valid, batch_fix, batch_size, scale, lane);
    const float hard = (p > 0.5f) ? 1.f : 0.f;               // (pred > 0.5).float() == z
    hits += (valid && hard == z) ? 1u : 0u;
  }
  batch_acc_flush(acc, batch_fix, lane);
  hits = __reduce_add_sync(0xffffffffu, hits);
  if (lane == 0 && hits) atomicAdd(correct, (unsigned long long)hits);
}

template <int VEC, int LPT, int NITER>
struct EvalLauncher {
  static int run(const float* U, const float* V, const mfcd_triplet* rec, int64_t N, int d, int64_t batch_size,
                 long long* batch_fix, unsigned long long* correct, cudaStream_t st) {
    const int grid = grid_for(N, kEvalBlock, 4);
    k_eval<VEC, LPT, NITER><<<grid, kEvalBlock, 0, st>>>(
        U, V, rec, N, d, batch_size, (float)(int64_t(1) << fix_shift_for(batch_size)), batch_fix, correct);
    MFCD_CHECK_LAUNCH();
    return MFCD_OK;
  }
};

static int launch_eval(const float* U, const float* V, const mfcd_triplet* rec, int64_t N, int d,
                       int64_t batch_size, long long* batch_fix, unsigned long long* correct, cudaStream_t st) {
  MFCD_DISPATCH_ROW_SHAPE(EvalLauncher, d, U, V, rec, N, d, batch_size, batch_fix, correct, st);
}

__global__ void __launch_bounds__(256)
k_gt_eval(mfcd_xview X, const mfcd_triplet* __restrict__ rec, int64_t N, int64_t batch_size, float scale,
          long long* __restrict__ batch_fix, unsigned long long* __restrict__ correct) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  unsigned int hits = 0;
  BatchAcc acc{-1, 0};
  for (int64_t base = warp0 * 32; base < N; base += nwarps * 32) {
    const int64_t k = base + lane;
    const bool valid = k < N;
    float e2 = 0.f;
    if (valid) {
      const int4 r = __ldg(reinterpret_cast<const int4*>(rec) + k);
      const float z = __int_as_float(r.w);
      const float diff = xview_at(X, r.x, r.y) - xview_at(X, r.x, r.z);   // no scale s (structure.py:1108)
      const float e = sigmoidf_ref(diff) - z;
      e2 = e * e;
      const float hard = (diff > 0.f) ? 1.f : 0.f;
      hits += (hard == z) ? 1u : 0u;
    }
    batch_acc_add(acc, e2, k, valid, batch_fix, batch_size, scale, lane);
  }
  batch_acc_flush(acc, batch_fix, lane);
  hits = __reduce_add_sync(0xffffffffu, hits);
  if (lane == 0 && hits) atomicAdd(correct, (unsigned long long)hits);
}

template <int VEC, int LPT, int NITER>
__global__ void __launch_bounds__(kEvalBlock)
k_scores(const float* __restrict__ U, const float* __restrict__ V, const int64_t* __restrict__ u,
         const int64_t* __restrict__ i, const int64_t* __restrict__ j, int64_t N, int d, float* __restrict__ p) {
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPT;
  const int grp = lane / LPT;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t base = warp0 * 32; base < N; base += nwarps * 32) {
    const int64_t k = base + lane;
    int ru = 0, ri = 0, rj = 0;
    if (k < N) { ru = (int)u[k]; ri = (int)i[k]; rj = (int)j[k]; }
    const int nvalid = (N - base) < 32 ? (int)(N - base) : 32;
    float x_home = 0.f;
    for (int rr = 0; rr < LPT; ++rr) {
      const int e = grp * LPT + rr;
      const int tu = __shfl_sync(0xffffffffu, ru, e);
      const int ti = __shfl_sync(0xffffffffu, ri, e);
      const int tj = __shfl_sync(0xffffffffu, rj, e);
      const bool ok = e < nvalid;
      TripletRows<VEC, LPT, NITER> rows;
      load_rows<VEC, LPT, NITER>(rows, U, V, tu, ti, tj, d, sub, ok);
      const float x = group_sum<LPT>(partial_dot<VEC, LPT, NITER>(rows), 0xffffffffu);
      x_home = (sub == rr) ? x : x_home;
    }
    if (lane < nvalid) p[k] = sigmoidf_ref(x_home);          // coalesced: one probability per lane
  }
}

template <int VEC, int LPT, int NITER>
struct ScoresLauncher {
  static int run(const float* U, const float* V, const int64_t* u, const int64_t* i, const int64_t* j, int64_t N,
                 int d, float* p, cudaStream_t st) {
    const int grid = grid_for(N, kEvalBlock, 4);
    k_scores<VEC, LPT, NITER><<<grid, kEvalBlock, 0, st>>>(U, V, u, i, j, N, d, p);
    MFCD_CHECK_LAUNCH();
    return MFCD_OK;
  }
};

static int launch_scores(const float* U, const float* V, const int64_t* u, const int64_t* i, const int64_t* j,
                         int64_t N, int d, float* p, cudaStream_t st) {
  MFCD_DISPATCH_ROW_SHAPE(ScoresLauncher, d, U, V, u, i, j, N, d, p, st);
}

}  // namespace mfcd

using namespace mfcd;

static int check_xview(const char* fn, const mfcd_xview* X) {
  MFCD_REQUIRE(X != nullptr, "%s: xview is NULL", fn);
  MFCD_REQUIRE(X->X != nullptr || (X->A != nullptr && X->B != nullptr && X->dx >= 1),
               "%s: xview has neither a dense matrix nor factors", fn);
  return MFCD_OK;
}

extern "C" int mfcd_triplet_eval(const float* U, const float* V, const mfcd_triplet* rec, int64_t N, int32_t d,
                                 int64_t batch_size, float* batch_loss, unsigned long long* correct,
                                 int64_t* batch_acc, void* stream) {
  MFCD_REQUIRE(N >= 0 && d >= 1 && batch_size >= 1, "mfcd_triplet_eval: bad sizes");
  if (N == 0) return MFCD_OK;
  MFCD_REQUIRE(U && V && rec && batch_loss && correct && batch_acc, "mfcd_triplet_eval: NULL pointer");
  long long* fix = reinterpret_cast<long long*>(batch_acc);
  const int rc = launch_eval(U, V, rec, N, d, batch_size, fix, correct, as_stream(stream));
  if (rc != MFCD_OK) return rc;
  return launch_eval_finish(fix, N, batch_size, batch_loss, as_stream(stream));
}

extern "C" int mfcd_ground_truth_eval(const mfcd_xview* X, const mfcd_triplet* rec, int64_t N, int64_t batch_size,
                                      float* batch_mse, unsigned long long* correct, int64_t* batch_acc,
                                      void* stream) {
  int rc = check_xview("mfcd_ground_truth_eval", X);
  if (rc != MFCD_OK) return rc;
  MFCD_REQUIRE(N >= 0 && batch_size >= 1, "mfcd_ground_truth_eval: bad sizes");
  if (N == 0) return MFCD_OK;
  MFCD_REQUIRE(rec && batch_mse && correct && batch_acc, "mfcd_ground_truth_eval: NULL pointer");
  long long* fix = reinterpret_cast<long long*>(batch_acc);
  k_gt_eval<<<grid_for(N, 256, 8), 256, 0, as_stream(stream)>>>(
      *X, rec, N, batch_size, (float)(int64_t(1) << fix_shift_for(batch_size)), fix, correct);
  MFCD_CHECK_LAUNCH();
  return launch_eval_finish(fix, N, batch_size, batch_mse, as_stream(stream));
}

extern "C" int mfcd_triplet_scores(const float* U, const float* V, const int64_t* u, const int64_t* i,
                                   const int64_t* j, int64_t N, int32_t d, float* p, void* stream) {
  MFCD_REQUIRE(N >= 0 && d >= 1, "mfcd_triplet_scores: bad sizes");
  if (N == 0) return MFCD_OK;
  MFCD_REQUIRE(U && V && u && i && j && p, "mfcd_triplet_scores: NULL pointer");
  return launch_scores(U, V, u, i, j, N, d, p, as_stream(stream));
}
