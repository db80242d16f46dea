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


// ---------------------------------------------------------------------------------------------------------
// K9 with in-kernel synchronisation (the default): no host-visible barrier and no separate gradient memset.
//
//   flags (peer-mapped, one array per rank):  ready[q] = sequence number of the last step for which rank q's
//   gradients are final;  done[q] = last step for which rank q has finished reading MY gradients, writing MY
//   parameter replica and clearing its slice of MY gradient buffer.  Sequence numbers only grow, so nothing is
//   ever reset.
//
//   start:  CTA 0 publishes ready[rank] = seq into every rank's flag array (K1 of this step has completed in
//           stream order, its reductions are in this GPU's L2 where peer loads are served);
//           every CTA waits until all ranks' ready[] >= seq (acquire, system scope);
//   body:   reduce-scatter of the owned slice -> Adam -> all-gather.  The next K1 needs a clean gradient buffer:
//           with `zero_local` the gradients are DOUBLE-BUFFERED -- step k accumulates into buffer k & 1, and this
//           kernel clears the rank's OTHER buffer locally (every peer finished reading it before the previous K9
//           completed, see `end`), so no zeros cross the fabric; without it the owner writes zeros over the slice
//           it has just consumed in every rank's buffer (as much NVLink traffic again as the all-gather);
//   end:    the last CTA to finish publishes done[rank] = seq everywhere and waits for every rank's done[] >= seq
//           before it exits: the kernel -- and with it the next K1 in the stream -- cannot complete / start
//           before all replicas are written and all gradient slices are cleared.
//   Spins are bounded (~4 s of SM clocks): on time-out *error is set and the kernel leaves instead of hanging.
// ---------------------------------------------------------------------------------------------------------
struct FlagTable { uint32_t* p[kMaxPeers]; };      // p[q][0..8) = ready[], p[q][8..16) = done[]  (rank q's array)

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// wait until word >= seq (wrap-safe compare); false on time-out
__device__ __forceinline__ bool spin_until(const uint32_t* word, uint32_t seq) {
  const long long t0 = clock64();
  while ((int32_t)(ld_acquire_sys(word) - seq) < 0) {
    if (clock64() - t0 > 8000000000ll) return false;
    __nanosleep(64);
  }
  return true;
}

// W = number of peer loads per element on the p2p path (compile-time, so the in-flight registers are exactly W x UNR
// float4), UNR = elements whose gradient loads are ALL issued before the first one is consumed.  A remote load
// takes ~4-5 us through the switch; with one element in flight per thread the 8 dependent iterations of a 2-rank
// slice alone cost ~40 us of the 85 us the kernel took (tools/k9_diag.py), so the loads of UNR elements overlap.
template <bool MULTIMEM, int W, int UNR>
__global__ void __launch_bounds__(256, 3)
k_dp_fused_adam_sync(PeerTable grads, PeerTable params, FlagTable flags, float* __restrict__ mc_grads,
                     float* __restrict__ mc_params, int rank, int world, int64_t begin, int64_t end,
                     float* __restrict__ m, float* __restrict__ v, AdamK s, uint32_t seq,
                     unsigned int* __restrict__ cta_counter, int* __restrict__ error, float* __restrict__ zero_local,
                     int64_t zero_n4) {
  uint32_t* mine = flags.p[rank];
  if (blockIdx.x == 0 && threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(flags.p[threadIdx.x] + rank, seq);                     // ready[rank] in rank threadIdx.x's array
  }
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = (end - begin) >> 2;
  float* my_params = params.p[rank];
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  // the other gradient buffer, whole, local stores only -- needs nothing from the peers, so it is cleared while
  // the slower ranks are still finishing their K1
  if (zero_local != nullptr)
    for (int64_t k = tid; k < zero_n4; k += nth) reinterpret_cast<float4*>(zero_local)[k] = zero4;
  if (threadIdx.x < world) {
    if (!spin_until(mine + threadIdx.x, seq)) atomicExch(error, 1);
  }
  __syncthreads();

  for (int64_t k0 = tid; k0 < n4; k0 += nth * UNR) {
    float4 g[UNR], t[UNR][MULTIMEM ? 1 : W];
    // 1. every gradient load of the UNR elements goes out before anything waits for one
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int64_t k = k0 + u * nth;
      if (k < n4) {
        const int64_t e = begin + (k << 2);
        if (MULTIMEM) {
          t[u][0] = multimem_ld_reduce_add(mc_grads + e);
        } else {
#pragma unroll
          for (int q = 0; q < W; ++q)
            t[u][q] = __ldcv(reinterpret_cast<const float4*>(grads.p[q] + e));      // peer data: never a stale L1 line
        }
      }
    }
    // 2. sums in rank order (deterministic), Adam on the owned slice, all-gather
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int64_t k = k0 + u * nth;
      if (k >= n4) continue;
      const int64_t e = begin + (k << 2);
      if (MULTIMEM) {
        g[u] = t[u][0];
      } else {
        g[u] = zero4;
#pragma unroll
        for (int q = 0; q < W; ++q) { g[u].x += t[u][q].x; g[u].y += t[u][q].y; g[u].z += t[u][q].z; g[u].w += t[u][q].w; }
      }
      float4 pp = *reinterpret_cast<const float4*>(my_params + e);
      float4 mm = *reinterpret_cast<const float4*>(m + e);
      float4 vv = *reinterpret_cast<const float4*>(v + e);
      adam_one(pp.x, g[u].x, mm.x, vv.x, s);
      adam_one(pp.y, g[u].y, mm.y, vv.y, s);
      adam_one(pp.z, g[u].z, mm.z, vv.z, s);
      adam_one(pp.w, g[u].w, mm.w, vv.w, s);
      *reinterpret_cast<float4*>(m + e) = mm;
      *reinterpret_cast<float4*>(v + e) = vv;
      // The parameter stores depend on g, so a warp (in-order issue) reaches the zero stores below only after the
      // gradient loads have RETURNED: clearing a slice can never overtake the read of that slice.
      if (MULTIMEM) {
        multimem_st(mc_params + e, pp);
        if (zero_local == nullptr) multimem_st(mc_grads + e, zero4);
      } else {
#pragma unroll
        for (int q = 0; q < W; ++q) *reinterpret_cast<float4*>(params.p[q] + e) = pp;
        if (zero_local == nullptr) {
#pragma unroll
          for (int q = 0; q < W; ++q) *reinterpret_cast<float4*>(grads.p[q] + e) = zero4;
        }
      }
    }
  }
  for (int64_t e = begin + (n4 << 2) + tid; e < end; e += nth) {        // ragged tail of the last owner
    float g = 0.f;
    for (int q = 0; q < world; ++q) g += __ldcv(grads.p[q] + e);
    float pp = my_params[e], mm = m[e], vv = v[e];
    adam_one(pp, g, mm, vv, s);
    m[e] = mm; v[e] = vv;
    for (int q = 0; q < world; ++q) params.p[q][e] = pp;
    if (zero_local == nullptr)
      for (int q = 0; q < world; ++q) grads.p[q][e] = 0.f;
  }

  // all of this CTA's peer stores are ordered before its arrival at the counter
  __threadfence_system();
  __syncthreads();
  __shared__ int s_last;
  if (threadIdx.x == 0) s_last = (atomicAdd(cta_counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    if (threadIdx.x == 0) { *cta_counter = 0; __threadfence_system(); }
    __syncthreads();
    if (threadIdx.x < world) {
      st_release_sys(flags.p[threadIdx.x] + kMaxPeers + rank, seq);       // done[rank] in every rank's array
      if (!spin_until(mine + kMaxPeers + threadIdx.x, seq)) atomicExch(error, 1);
    }
  }
}

}  // namespace mfcd

using namespace mfcd;

extern "C" int mfcd_dp_fused_adam_sync(const uint64_t* peer_grads, const uint64_t* peer_params,
                                       const uint64_t* peer_flags, uint64_t mc_grads, uint64_t mc_params,
                                       int32_t rank, int32_t world, int64_t numel, float* m, float* v, float lr,
                                       float beta1, float beta2, float eps, float weight_decay, int64_t step,
                                       uint32_t seq, uint32_t* cta_counter, int32_t* error, float* zero_local,
                                       void* stream) {
  MFCD_REQUIRE(peer_grads && peer_params && peer_flags && m && v && cta_counter && error,
               "mfcd_dp_fused_adam_sync: NULL pointer");
  MFCD_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "mfcd_dp_fused_adam_sync: world must be 1..8");
  MFCD_REQUIRE(step >= 1 && numel >= 0 && seq >= 1, "mfcd_dp_fused_adam_sync: bad step / numel / seq");
  PeerTable g, p;
  FlagTable f;
  for (int q = 0; q < kMaxPeers; ++q) {
    g.p[q] = q < world ? reinterpret_cast<float*>(peer_grads[q]) : nullptr;
    p.p[q] = q < world ? reinterpret_cast<float*>(peer_params[q]) : nullptr;
    f.p[q] = q < world ? reinterpret_cast<uint32_t*>(peer_flags[q]) : nullptr;
    if (q < world) {
      MFCD_REQUIRE(g.p[q] && p.p[q] && f.p[q], "mfcd_dp_fused_adam_sync: NULL peer pointer");
      MFCD_REQUIRE(((peer_grads[q] | peer_params[q]) & 15u) == 0, "mfcd_dp_fused_adam_sync: peer buffers must be 16-byte aligned");
    }
  }
  int64_t begin = 0, end = 0;
  int rc = mfcd_dp_shard_range(numel, rank, world, &begin, &end);
  if (rc != MFCD_OK) return rc;
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
  // every CTA spins at the start, so the whole grid must be co-resident: at most 3 CTAs of 256 threads per SM
  // (the launch bound keeps the kernel at <= 85 registers).  A rank with an empty slice still takes part in the
  // flag exchange (one CTA).
  // zero_local: a buffer of numel floats padded to a multiple of 4 (the exchange allocates it that way)
  const int64_t zero_n4 = zero_local ? (numel + 3) / 4 : 0;
  MFCD_REQUIRE((reinterpret_cast<uintptr_t>(zero_local) & 15u) == 0, "mfcd_dp_fused_adam_sync: zero_local must be 16-byte aligned");
  int grid = grid_for((end - begin + 3) / 4 + zero_n4, 256, 3);
  if (end <= begin && zero_n4 == 0) grid = 1;
  cudaStream_t st = as_stream(stream);
#define MFCD_K9_GO(MM, W, UNR)                                                                                   \
  k_dp_fused_adam_sync<MM, W, UNR><<<grid, 256, 0, st>>>(g, p, f, reinterpret_cast<float*>(mc_grads),                \
                                                         reinterpret_cast<float*>(mc_params), rank, world, begin, end, \
                                                         m, v, s, seq, cta_counter, error, zero_local, zero_n4)
  if (mc_grads != 0 && mc_params != 0) MFCD_K9_GO(true, 1, 4);
  else if (world == 1) MFCD_K9_GO(false, 1, 4);
  else if (world == 2) MFCD_K9_GO(false, 2, 4);
  else if (world == 3) MFCD_K9_GO(false, 3, 2);
  else if (world == 4) MFCD_K9_GO(false, 4, 2);
  else if (world == 5) MFCD_K9_GO(false, 5, 1);
  else if (world == 6) MFCD_K9_GO(false, 6, 1);
  else if (world == 7) MFCD_K9_GO(false, 7, 1);
  else MFCD_K9_GO(false, 8, 1);
#undef MFCD_K9_GO
  MFCD_CHECK_LAUNCH();
  return MFCD_OK;
}


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
