// Internal (non-ABI) launchers shared between translation units.
#pragma once
#include <math.h>
#include "common.cuh"

namespace mfcd {
// records a CUDA event pair around a K1 launch on its stream while mfcd_profile_k1(1) is in effect (bench.py's
// live roofline measurement of the dominant kernel inside the timed region); free otherwise
struct K1Timer {
  explicit K1Timer(cudaStream_t st);
  ~K1Timer();
  cudaStream_t st_;
  bool on_;
};

int launch_adam(float* p, float* g, float* m, float* v, int64_t numel, float lr, float beta1, float beta2,
                float eps, float wd, int64_t step, int zero_grad, cudaStream_t st);
int launch_sgd(float* p, float* g, float* buf, int64_t numel, float lr, float momentum, float wd,
               int64_t step, int zero_grad, cudaStream_t st);
int launch_fwd_bwd_atomic(const float* U, const float* V, const mfcd_triplet* rec, const int32_t* perm,
                          int64_t start, int64_t B, int d, float inv_batch, float* gU, float* gV,
                          float* loss, const int8_t* item_slot, const int32_t* hot_items, int n_hot,
                          int flags, cudaStream_t st);
int max_hot_rows(int d);
int launch_fwd_bwd_det(const float* U, const float* V, const mfcd_triplet* rec, const int32_t* perm,
                       int64_t start, int64_t B, int d, float inv_batch, int64_t n_users, int64_t n_items,
                       float* gU, float* gV, float* loss, void* ws, size_t ws_bytes, cudaStream_t st);
size_t det_workspace_bytes(int64_t B, int d);                      // sort + segmented reduction engine (segmented.cu)
// fixed-point engine (k1_fixed.cu): integer atomics into a 64-bit image of the gradient tables, no sort
size_t det_fixed_workspace_bytes(int64_t n_users, int64_t n_items, int d);
size_t det_workspace_bytes_nm(int64_t B, int d, int64_t n_users, int64_t n_items);   // what the default engine wants
int launch_det_fixed(const float* U, const float* V, const mfcd_triplet* rec, const int32_t* perm, int64_t start,
                     int64_t B, int d, float inv_batch, int64_t n_users, int64_t n_items, float* gU, float* gV,
                     float* loss, void* ws, cudaStream_t st);

__device__ __forceinline__ float xview_at(const mfcd_xview& X, int64_t r, int64_t c) {
  if (X.X != nullptr) return __ldg(X.X + r * X.ldx + c);
  const float* a = X.A + r * X.dx;
  const float* b = X.B + c * X.dx;
  float acc = 0.f;
  for (int k = 0; k < X.dx; ++k) acc = fmaf(__ldg(a + k), __ldg(b + k), acc);
  return X.scale * acc;
}

// ---- Adam, one element (torch.optim.Adam single-tensor path, coupled L2; structure.py:364, :851) ----
struct AdamScalars {
  float lr_over_bc1;   // lr / (1 - beta1^t)
  float bc2_sqrt;      // sqrt(1 - beta2^t)
  float one_minus_b1, b2, one_minus_b2, eps, wd;
};

// bias corrections in double like the python floats of torch's _single_tensor_adam
__host__ __device__ inline AdamScalars adam_scalars(float lr, float beta1, float beta2, float eps, float wd,
                                                    int64_t step) {
  AdamScalars s;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  s.lr_over_bc1 = (float)((double)lr / bc1);
  s.bc2_sqrt = (float)sqrt(bc2);
  s.one_minus_b1 = (float)(1.0 - (double)beta1);
  s.b2 = beta2;
  s.one_minus_b2 = (float)(1.0 - (double)beta2);
  s.eps = eps;
  s.wd = wd;
  return s;
}

__device__ __forceinline__ void adam_elem(float& p, float& g, float& m, float& v, const AdamScalars& s) {
  float gg = (s.wd != 0.f) ? fmaf(s.wd, p, g) : g;          // grad.add(param, alpha=wd)
  m = fmaf(gg - m, s.one_minus_b1, m);                      // exp_avg.lerp_(grad, 1-beta1)
  v = fmaf(s.one_minus_b2 * gg, gg, v * s.b2);              // mul_(beta2).addcmul_(g, g, 1-beta2)
  float denom = sqrtf(v) / s.bc2_sqrt + s.eps;              // (sqrt(v)/bc2_sqrt).add_(eps)
  p = fmaf(-s.lr_over_bc1, m / denom, p);                   // addcdiv_(m, denom, value=-step_size)
}

// persistent small-batch epoch (epoch_small.cu): one cooperative kernel runs every optimiser step of the epoch
size_t epoch_small_workspace_bytes(const mfcd_epoch_args* a);      // 0 = this epoch is not eligible
int launch_epoch_small(const mfcd_epoch_args* a, cudaStream_t st); // MFCD_ERR_UNSUPPORTED = not eligible
}  // namespace mfcd
