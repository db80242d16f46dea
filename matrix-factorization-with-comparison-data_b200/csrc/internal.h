// Internal (non-ABI) launchers shared between translation units.
#pragma once
#include "common.cuh"

namespace mfcd {
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
size_t det_workspace_bytes(int64_t B, int d);

__device__ __forceinline__ float xview_at(const mfcd_xview& X, int64_t r, int64_t c) {
  if (X.X != nullptr) return __ldg(X.X + r * X.ldx + c);
  const float* a = X.A + r * X.dx;
  const float* b = X.B + c * X.dx;
  float acc = 0.f;
  for (int k = 0; k < X.dx; ++k) acc = fmaf(__ldg(a + k), __ldg(b + k), acc);
  return X.scale * acc;
}
}  // namespace mfcd
