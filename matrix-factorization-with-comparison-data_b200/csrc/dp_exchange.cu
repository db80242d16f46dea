// K9: fused gradient exchange + optimiser update over NVLink peer memory.
//
// Data-parallel step, after every rank's K1 has produced its local dense gradient
// ("NCCL allreduce of gU||gV ... fuse/overlap with K3"):
// instead of  ncclAllReduce(grads)  followed by a full-size Adam on every rank,
// rank r owns the r-th slice of the flat parameter vector and, in ONE kernel,
//   1. reduce-scatter: sums that slice of the gradient over all ranks, reading the
//      peers' buffers directly -- either W peer-to-peer loads (fixed rank order, so
//      the sum is deterministic) or a single in-switch reduction
//      (multimem.ld_reduce on the NVSwitch multicast address),
//   2. applies Adam to its slice (moments exist only for the owned slice's elements),
//   3. all-gather: writes the updated parameters straight into every rank's replica
//      (W peer stores, or one multimem.st that the switch fans out).
// NVLink traffic per rank is what a ring all-reduce moves (about 2 x (W-1)/W x 4 numel
// bytes), but there is one launch, no separate Adam pass over the other W-1 slices, and
// the transfer overlaps the arithmetic element by element.
// The caller brackets the kernel with two symmetric-memory barriers (all gradients
// final before; all replicas written and all gradient reads done after) and clears its
// own gradient buffer afterwards.
#include <math.h>
#include "internal.h"

namespace mfcd {

constexpr int kMaxPeers = 8;
struct PeerTable { float* p[kMaxPeers]; };

struct AdamK {
  float lr_over_bc1, bc2_sqrt, one_minus_b1, b2, one_minus_b2, eps, wd;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamK& s) {
  float gg = (s.wd != 0.f) ? fmaf(s.wd, p, g) : g;
  m = fmaf(gg - m, s.one_minus_b1, m);
  v = fmaf(s.one_minus_b2 * gg, gg, v * s.b2);
  const float denom = sqrtf(v) / s.bc2_sqrt + s.eps;
  p = fmaf(-s.lr_over_bc1, m / denom, p);
}

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(mc) : "memory");
  return r;
}
__device__ __forceinline__ void multimem_st(float* mc, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <bool MULTIMEM>
__global__ void __launch_bounds__(256)
k_dp_fused_adam(PeerTable grads, PeerTable params, const float* __restrict__ mc_grads, float* __restrict__ mc_params,
                int rank, int world, int64_t begin, int64_t end /* element range owned, begin % 4 == 0 */,
                float* __restrict__ m, float* __restrict__ v, AdamK s) {
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = (end - begin) >> 2;
  float* my_params = params.p[rank];
  for (int64_t k = tid; k < n4; k += nth) {
    const int64_t e = begin + (k << 2);
    float4 g;
    if (MULTIMEM) {
      g = multimem_ld_reduce_add(mc_grads + e);
    } else {
      g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int q = 0; q < kMaxPeers; ++q) {
        if (q < world) {
          const float4 t = __ldcv(reinterpret_cast<const float4*>(grads.p[q] + e));   // peer data: never from a stale L1 line
          g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
        }
      }
    }
    float4 pp = *reinterpret_cast<const float4*>(my_params + e);
    float4 mm = *reinterpret_cast<const float4*>(m + e);
    float4 vv = *reinterpret_cast<const float4*>(v + e);
    adam_one(pp.x, g.x, mm.x, vv.x, s);
    adam_one(pp.y, g.y, mm.y, vv.y, s);
    adam_one(pp.z, g.z, mm.z, vv.z, s);
    adam_one(pp.w, g.w, mm.w, vv.w, s);
    *reinterpret_cast<float4*>(m + e) = mm;
    *reinterpret_cast<float4*>(v + e) = vv;
    if (MULTIMEM) {
      multimem_st(mc_params + e, pp);
    } else {
#pragma unroll
      for (int q = 0; q < kMaxPeers; ++q)
        if (q < world) *reinterpret_cast<float4*>(params.p[q] + e) = pp;
    }
  }
  // ragged tail (numel % 4) of the last owner: scalar peer loads / stores
  for (int64_t e = begin + (n4 << 2) + tid; e < end; e += nth) {
    float g = 0.f;
    for (int q = 0; q < world; ++q) g += __ldcv(grads.p[q] + e);
    float pp = my_params[e], mm = m[e], vv = v[e];
    adam_one(pp, g, mm, vv, s);
    m[e] = mm; v[e] = vv;
    for (int q = 0; q < world; ++q) params.p[q][e] = pp;
  }
}

}  // namespace mfcd

using namespace mfcd;

extern "C" int mfcd_dp_shard_range(int64_t numel, int32_t rank, int32_t world, int64_t* begin, int64_t* end) {
  MFCD_REQUIRE(begin && end && numel >= 0 && world >= 1 && rank >= 0 && rank < world, "mfcd_dp_shard_range: bad argument");
  const int64_t per = ((numel + world - 1) / world + 3) & ~int64_t(3);      // 16-byte aligned slice starts
  int64_t b = per * rank, e = per * (rank + 1);
  if (b > numel) b = numel;
  if (e > numel || rank == world - 1) e = numel;
  *begin = b; *end = e;
  return MFCD_OK;
}

extern "C" int mfcd_dp_fused_adam(const uint64_t* peer_grads, const uint64_t* peer_params, uint64_t mc_grads,
                                  uint64_t mc_params, int32_t rank, int32_t world, int64_t numel, float* m, float* v,
                                  float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step,
                                  void* stream) {
  MFCD_REQUIRE(peer_grads && peer_params && m && v, "mfcd_dp_fused_adam: NULL pointer");
  MFCD_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "mfcd_dp_fused_adam: world must be 1..8");
  MFCD_REQUIRE(step >= 1 && numel >= 0, "mfcd_dp_fused_adam: bad step / numel");
  PeerTable g, p;
  for (int q = 0; q < kMaxPeers; ++q) {
    g.p[q] = q < world ? reinterpret_cast<float*>(peer_grads[q]) : nullptr;
    p.p[q] = q < world ? reinterpret_cast<float*>(peer_params[q]) : nullptr;
    if (q < world) {
      MFCD_REQUIRE(g.p[q] && p.p[q], "mfcd_dp_fused_adam: NULL peer pointer");
      MFCD_REQUIRE(((peer_grads[q] | peer_params[q]) & 15u) == 0, "mfcd_dp_fused_adam: peer buffers must be 16-byte aligned");
    }
  }
  int64_t begin = 0, end = 0;
  int rc = mfcd_dp_shard_range(numel, rank, world, &begin, &end);
  if (rc != MFCD_OK) return rc;
  if (end <= begin) return MFCD_OK;
  AdamK s;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  s.lr_over_bc1 = (float)((double)lr / bc1);
  s.bc2_sqrt = (float)sqrt(bc2);
  s.one_minus_b1 = (float)(1.0 - (double)beta1);
  s.b2 = beta2;
  s.one_minus_b2 = (float)(1.0 - (double)beta2);
  s.eps = eps;
  s.wd = weight_decay;
  const int grid = grid_for((end - begin + 3) / 4, 256, 8);
  cudaStream_t st = as_stream(stream);
  if (mc_grads != 0 && mc_params != 0)
    k_dp_fused_adam<true><<<grid, 256, 0, st>>>(g, p, reinterpret_cast<const float*>(mc_grads),
                                                reinterpret_cast<float*>(mc_params), rank, world, begin, end, m, v, s);
  else
    k_dp_fused_adam<false><<<grid, 256, 0, st>>>(g, p, nullptr, nullptr, rank, world, begin, end, m, v, s);
  MFCD_CHECK_LAUNCH();
  return MFCD_OK;
}
