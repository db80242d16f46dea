// K3: fused dense optimiser update (Adam with coupled L2, SGD with momentum).
// Reference: torch.optim.Adam(model.parameters(), lr, weight_decay) at
// structure.py:364 stepped at :851, gradients cleared at :847.  HBM-bound
// elementwise pass: reads p,g,m,v and writes p,m,v,(g=0) = 32 bytes/element.
#include <math.h>
#include "internal.h"

namespace mfcd {

template <bool ZERO>
__global__ void __launch_bounds__(256) k_adam(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                              float* __restrict__ v, int64_t numel, AdamScalars s) {
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = numel >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  float4* g4 = reinterpret_cast<float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (int64_t k = tid; k < n4; k += nth) {
    float4 pp = p4[k], gg = g4[k], mm = m4[k], vv = v4[k];
    adam_elem(pp.x, gg.x, mm.x, vv.x, s);
    adam_elem(pp.y, gg.y, mm.y, vv.y, s);
    adam_elem(pp.z, gg.z, mm.z, vv.z, s);
    adam_elem(pp.w, gg.w, mm.w, vv.w, s);
    p4[k] = pp; m4[k] = mm; v4[k] = vv;
    if (ZERO) g4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int64_t k = (n4 << 2) + tid; k < numel; k += nth) {
    float pp = p[k], gg = g[k], mm = m[k], vv = v[k];
    adam_elem(pp, gg, mm, vv, s);
    p[k] = pp; m[k] = mm; v[k] = vv;
    if (ZERO) g[k] = 0.f;
  }
}

template <bool ZERO, bool MOM>
__global__ void __launch_bounds__(256) k_sgd(float* __restrict__ p, float* __restrict__ g, float* __restrict__ buf,
                                             int64_t numel, float lr, float mu, float wd, int first) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < numel; k += (int64_t)gridDim.x * blockDim.x) {
    float pp = p[k];
    float gg = g[k];
    if (wd != 0.f) gg = fmaf(wd, pp, gg);
    if (MOM) {
      float b = first ? gg : fmaf(mu, buf[k], gg);
      buf[k] = b;
      gg = b;
    }
    p[k] = fmaf(-lr, gg, pp);
    if (ZERO) g[k] = 0.f;
  }
}

static inline bool aligned16(const void* a) { return (reinterpret_cast<uintptr_t>(a) & 15u) == 0; }

int launch_adam(float* p, float* g, float* m, float* v, int64_t numel, float lr, float beta1, float beta2,
                float eps, float wd, int64_t step, int zero_grad, cudaStream_t st) {
  if (numel == 0) return MFCD_OK;
  MFCD_REQUIRE(p && g && m && v, "adam: NULL pointer");
  MFCD_REQUIRE(step >= 1, "adam: step must be >= 1 (1-based)");
  MFCD_REQUIRE(aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v), "adam: buffers must be 16-byte aligned");
  const AdamScalars s = adam_scalars(lr, beta1, beta2, eps, wd, step);
  const int grid = grid_for((numel + 3) / 4, 256, 8);
  if (zero_grad) k_adam<true><<<grid, 256, 0, st>>>(p, g, m, v, numel, s);
  else k_adam<false><<<grid, 256, 0, st>>>(p, g, m, v, numel, s);
  MFCD_CHECK_LAUNCH();
  return MFCD_OK;
}

int launch_sgd(float* p, float* g, float* buf, int64_t numel, float lr, float momentum, float wd, int64_t step,
               int zero_grad, cudaStream_t st) {
  if (numel == 0) return MFCD_OK;
  MFCD_REQUIRE(p && g, "sgd: NULL pointer");
  MFCD_REQUIRE(step >= 1, "sgd: step must be >= 1 (1-based)");
  MFCD_REQUIRE(momentum == 0.f || buf != nullptr, "sgd: momentum buffer is NULL");
  const int grid = grid_for(numel, 256, 8);
  const int first = (step == 1);
  if (momentum != 0.f) {
    if (zero_grad) k_sgd<true, true><<<grid, 256, 0, st>>>(p, g, buf, numel, lr, momentum, wd, first);
    else k_sgd<false, true><<<grid, 256, 0, st>>>(p, g, buf, numel, lr, momentum, wd, first);
  } else {
    if (zero_grad) k_sgd<true, false><<<grid, 256, 0, st>>>(p, g, buf, numel, lr, momentum, wd, first);
    else k_sgd<false, false><<<grid, 256, 0, st>>>(p, g, buf, numel, lr, momentum, wd, first);
  }
  MFCD_CHECK_LAUNCH();
  return MFCD_OK;
}

}  // namespace mfcd

extern "C" int mfcd_adam_update(float* p, float* g, float* m, float* v, int64_t numel, float lr, float beta1,
                                float beta2, float eps, float weight_decay, int64_t step, int32_t zero_grad,
                                void* stream) {
  MFCD_REQUIRE(numel >= 0, "mfcd_adam_update: numel < 0");
  return mfcd::launch_adam(p, g, m, v, numel, lr, beta1, beta2, eps, weight_decay, step, zero_grad,
                           mfcd::as_stream(stream));
}

extern "C" int mfcd_sgd_update(float* p, float* g, float* buf, int64_t numel, float lr, float momentum,
                               float weight_decay, int64_t step, int32_t zero_grad, void* stream) {
  MFCD_REQUIRE(numel >= 0, "mfcd_sgd_update: numel < 0");
  return mfcd::launch_sgd(p, g, buf, numel, lr, momentum, weight_decay, step, zero_grad, mfcd::as_stream(stream));
}
