// Shared device/host helpers for the mfcd_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/mfcd_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "mfcd_b200 is written for sm_100a (B200) only"
#endif

namespace mfcd {

// ---------------------------------------------------------------------------
// host-side error plumbing: no exceptions cross the C ABI
// ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define MFCD_CUDA(call)                                                        \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) return ::mfcd::cuda_fail(e__, #call, __FILE__, __LINE__); \
  } while (0)

#define MFCD_CHECK_LAUNCH() MFCD_CUDA(cudaGetLastError())

#define MFCD_REQUIRE(cond, ...)                                                \
  do {                                                                         \
    if (!(cond)) { ::mfcd::set_error(__VA_ARGS__); return MFCD_ERR_ARG; }      \
  } while (0)

int sm_count();                 // cached cudaDevAttrMultiProcessorCount of the current device
static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---------------------------------------------------------------------------
// vector row fragments: VEC floats per lane, moved with one 32/64/128-bit access
// ---------------------------------------------------------------------------
template <int VEC> struct Frag { float v[VEC]; };

template <int VEC>
__device__ __forceinline__ Frag<VEC> frag_zero() {
  Frag<VEC> f;
#pragma unroll
  for (int k = 0; k < VEC; ++k) f.v[k] = 0.f;
  return f;
}

// read-only (non-coherent) path: the tables are never written by the kernels that gather them
template <int VEC>
__device__ __forceinline__ Frag<VEC> ldg_frag(const float* __restrict__ p) {
  Frag<VEC> f;
  if constexpr (VEC == 4) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    f.v[0] = t.x; f.v[1] = t.y; f.v[2] = t.z; f.v[3] = t.w;
  } else if constexpr (VEC == 2) {
    float2 t = __ldg(reinterpret_cast<const float2*>(p));
    f.v[0] = t.x; f.v[1] = t.y;
  } else {
    f.v[0] = __ldg(p);
  }
  return f;
}

// plain (coherent) load, for buffers written earlier in the same kernel sequence
template <int VEC>
__device__ __forceinline__ Frag<VEC> ld_frag(const float* p) {
  Frag<VEC> f;
  if constexpr (VEC == 4) {
    float4 t = *reinterpret_cast<const float4*>(p);
    f.v[0] = t.x; f.v[1] = t.y; f.v[2] = t.z; f.v[3] = t.w;
  } else if constexpr (VEC == 2) {
    float2 t = *reinterpret_cast<const float2*>(p);
    f.v[0] = t.x; f.v[1] = t.y;
  } else {
    f.v[0] = *p;
  }
  return f;
}

template <int VEC>
__device__ __forceinline__ void st_frag(float* p, const Frag<VEC>& f) {
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(f.v[0], f.v[1], f.v[2], f.v[3]);
  } else if constexpr (VEC == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(f.v[0], f.v[1]);
  } else {
    *p = f.v[0];
  }
}

// fire-and-forget reduction into global memory: compiles to REDG.E.ADD.F32{,x2,x4}
template <int VEC>
__device__ __forceinline__ void red_frag(float* p, const Frag<VEC>& f) {
  if constexpr (VEC == 4) {
    atomicAdd(reinterpret_cast<float4*>(p), make_float4(f.v[0], f.v[1], f.v[2], f.v[3]));
  } else if constexpr (VEC == 2) {
    atomicAdd(reinterpret_cast<float2*>(p), make_float2(f.v[0], f.v[1]));
  } else {
    atomicAdd(p, f.v[0]);
  }
}

// packed fp32 pairs (FADD2 / FMUL2 / FFMA2 on sm_100): two elements per issued instruction for issue-bound kernels
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpk2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ float sum2(f32x2 v) {
  float lo, hi;
  unpk2(v, lo, hi);
  return lo + hi;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// sum over the LPT lanes of a sub-warp group (LPT power of two, groups are lane-aligned)
template <int LPT>
__device__ __forceinline__ float group_sum(float x, unsigned mask) {
#pragma unroll
  for (int off = LPT / 2; off > 0; off >>= 1) x += __shfl_xor_sync(mask, x, off);
  return x;
}

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
  return x;
}
__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
  return x;
}

// ---------------------------------------------------------------------------
// the model's scalar math (reference: structure.py:787-795, :849; ATen
// binary_cross_entropy / binary_cross_entropy_backward / sigmoid_backward)
// ---------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_ref(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float bce_ref(float p, float z) {
  float lp = fmaxf(logf(p), -100.f);
  float lq = fmaxf(log1pf(-p), -100.f);
  return (z - 1.f) * lq - z * lp;
}

// d(mean BCE)/dx for x the pre-sigmoid score; inv_batch = 1/B of the GLOBAL batch
__device__ __forceinline__ float bce_grad_score_ref(float p, float z, float inv_batch) {
  float q = (1.f - p) * p;
  float g_in = inv_batch * (p - z) / fmaxf(q, 1e-12f);
  return g_in * (1.f - p) * p;
}

// ---------------------------------------------------------------------------
// Philox4x32-10 counter-based generator (Salmon et al. 2011), one call = 4 u32
// ---------------------------------------------------------------------------
struct Philox {
  static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  static constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  __host__ __device__ static inline void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
    uint64_t p = static_cast<uint64_t>(a) * b;
    hi = static_cast<uint32_t>(p >> 32);
    lo = static_cast<uint32_t>(p);
  }
  __host__ __device__ static inline uint4 run(uint64_t seed, uint64_t counter, uint32_t stream) {
    uint32_t c0 = static_cast<uint32_t>(counter), c1 = static_cast<uint32_t>(counter >> 32);
    uint32_t c2 = stream, c3 = 0;
    uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
      uint32_t h0, l0, h1, l1;
      mulhilo(M0, c0, h0, l0);
      mulhilo(M1, c2, h1, l1);
      uint32_t n0 = h1 ^ c1 ^ k0, n1 = l1, n2 = h0 ^ c3 ^ k1, n3 = l0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += W0; k1 += W1;
    }
    uint4 out; out.x = c0; out.y = c1; out.z = c2; out.w = c3;
    return out;
  }
};

// unbiased-enough integer in [0, range): high word of a 32x32 multiply
__host__ __device__ static inline uint32_t bounded(uint32_t r, uint32_t range) {
  return static_cast<uint32_t>((static_cast<uint64_t>(r) * range) >> 32);
}
// 24-bit uniform in [0,1), the same mapping torch's CPU float uniform uses
__host__ __device__ static inline float u01_24(uint32_t r) { return (r & 0xFFFFFFu) * (1.0f / 16777216.0f); }

// ---------------------------------------------------------------------------
// grid sizing: persistent-style grids in multiples of the SM count
// ---------------------------------------------------------------------------
static inline int grid_for(int64_t work_items, int items_per_block, int blocks_per_sm) {
  int64_t need = (work_items + items_per_block - 1) / items_per_block;
  int64_t cap = static_cast<int64_t>(sm_count()) * blocks_per_sm;
  if (need < 1) need = 1;
  return static_cast<int>(need < cap ? need : cap);
}

}  // namespace mfcd
