// K1 lean: the atomic-mode fused forward + BCE + backward kernel for rows that exactly fill their lane
// group with float4 fragments (d == 4*LPT*NITER: d = 4, 8, 16, 32, 64, 128, 256, 384, 512).
// Same arithmetic as k_fwd_bwd<..., MODE 0> (structure.py:787-795, :849-850), restructured after the ncu
// profiles of that kernel (profiles/r01b_*):
//
//   * issue slots (103 warp instructions per triplet, 68 % issue-active on zipf items): D is a compile-time
//     constant, the score gradient is g = (p - z)/B with the reference's saturation branch (p(1-p) < 1e-12)
//     kept as a rare slow path and p = rcp.approx(1 + expf(-x)), so no IEEE division sequences run per lane
//     group (the REPORTED loss still uses the exact sigmoid / BCE, once per tile, SIMD over the home lanes);
//   * hot item rows: one shared-memory image per warp as before, but the lane groups of a warp simply take
//     turns in program order (SHARE) -- the clash detection, votes and __syncwarp of k_fwd_bwd are gone; an
//     image per lane group (MFCD_K1_SHARE=0) costs twice the shared memory per row and measured slower;
//   * RUNS: on uniform items the kernel is bound by the L1 -> crossbar request port (82 % busy: every row
//     read that misses L1 and every RED is a request).  When the batch keeps each user's triplets adjacent
//     (mfcd_group_by_user), the U row is read once and its gradient leaves as ONE reduction per run instead
//     of once per triplet: a third of the requests disappear.  Correct for any order, faster when grouped.
#pragma once
#include "shape_dispatch.cuh"

namespace mfcd {

// read-modify-write of a lane's fragment in shared memory, addressed by a 32-bit shared-window address
__device__ __forceinline__ void smem_add4_at(uint32_t addr, const Frag<4>& f) {
  float4 t;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "r"(addr));
  t.x += f.v[0]; t.y += f.v[1]; t.z += f.v[2]; t.w += f.v[3];
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(t.x), "f"(t.y), "f"(t.z), "f"(t.w) : "memory");
}

__device__ __forceinline__ float sigmoid_fast(float x) {
  float p;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(1.f + expf(-x)));
  return p;
}

// (p - z)/B; the reference's value (0, or its clamped form) where p(1-p) underflows 1e-12
__device__ __forceinline__ float bce_grad_score_fast(float p, float z, float inv_batch) {
  const float omp = 1.f - p;
  const float q = omp * p;
  float g = inv_batch * (p - z);
  if (q < 1e-12f) g = g * 1e12f * omp * p;
  return g;
}

constexpr int kLeanBlock = 256;

// CTAs per SM the register budget is sized for
template <int NITER, bool HOT, bool SHARE>
constexpr int lean_min_blocks() {
  return NITER > 2 ? 1 : (HOT ? (SHARE ? 3 : 2) : (NITER == 2 ? 3 : 4));
}

// SHARE: the lane groups of a warp share ONE hot image (half the shared memory per row at d = 64) and take
// turns updating it, group after group, in program order -- no detection, no barrier.
// WIRE: `rec` is one user-grouped batch in the run-length staging format of wire.cu (start = 0, perm = NULL):
// the step then needs no unpack pass between the host-to-device copy and K1.
template <int LPT, int NITER, bool HOT, bool RUNS, bool SHARE, bool WIRE = false>
__global__ void __launch_bounds__(kLeanBlock, lean_min_blocks<NITER, HOT, SHARE>())
k_fwd_bwd_lean(const float* __restrict__ U, const float* __restrict__ V, const mfcd_triplet* __restrict__ rec,
               const int32_t* __restrict__ perm, int64_t start, int64_t B, float inv_batch,
               float* __restrict__ gU, float* __restrict__ gV, float* __restrict__ loss_out,
               const int8_t* __restrict__ item_slot, const int32_t* __restrict__ hot_items, int n_hot) {
  constexpr int VEC = 4;
  constexpr int D = VEC * LPT * NITER;                 // floats per row
  constexpr uint32_t ROWB = D * 4;                     // bytes per row
  constexpr int STEP = LPT * VEC * 4;                  // bytes between a lane's NITER fragments
  constexpr int UNR = LPT >= 2 ? 2 : 1;                // triplets in flight per lane group (4 was measured: no gain)
  constexpr int GPW = 32 / LPT;                        // lane groups per warp
  constexpr int IMAGES = SHARE ? kLeanBlock / 32 : kLeanBlock / LPT;   // hot images per CTA
  constexpr uint32_t NO_USER = 0xffffffffu;
  __shared__ float s_red[kLeanBlock / 32];
  extern __shared__ __align__(16) float s_hot[];       // HOT: [IMAGES][n_hot][D]
  if constexpr (HOT) {
    for (int e = threadIdx.x; e < IMAGES * n_hot * D; e += kLeanBlock) s_hot[e] = 0.f;
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPT;
  const int grp = lane / LPT;
  // row r of a table starts at base + r * ROWB; the lane's fragment sits lane_off bytes into the row
  const uint64_t lane_off = (uint64_t)(sub * (VEC * 4));
  const char* Ub = reinterpret_cast<const char*>(U);
  const char* Vb = reinterpret_cast<const char*>(V);
  char* gUb = reinterpret_cast<char*>(gU);
  char* gVb = reinterpret_cast<char*>(gV);
  // this lane's column of its image, as a shared-window address
  const uint32_t my_hot = (uint32_t)__cvta_generic_to_shared(s_hot) +
                          (uint32_t)(threadIdx.x / (SHARE ? 32 : LPT)) * (uint32_t)n_hot * ROWB + sub * (VEC * 4);

  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float loss_acc = 0.f;
  for (int64_t base = warp0 * 32; base < B; base += nwarps * 32) {
    const int64_t k = base + lane;
    int4 r = make_int4(0, 0, 0, 0);
    int slots = 0xffff;                                // (slot_i & 0xff) | (slot_j & 0xff) << 8, 0xff = cold
    if (k < B) {
      if constexpr (WIRE) {
        const uint32_t* wire = reinterpret_cast<const uint32_t*>(rec);
        const int64_t nw = (B + 31) >> 5, w = base >> 5;           // [hdr 4 | run0 nw | zbits nw | nbits nw | ij B | users]
        const uint32_t bits = __ldg(wire + 4 + 2 * nw + w);
        const uint32_t zw = __ldg(wire + 4 + nw + w);
        const uint32_t ij = __ldg(wire + 4 + 3 * nw + k);
        const uint32_t run = __ldg(wire + 4 + w) + __popc(bits & (0xffffffffu >> (31 - lane))) - 1u;
        r.x = (int)__ldg(wire + 4 + 3 * nw + B + run);
        r.y = (int)(ij & 0xffffu);
        r.z = (int)(ij >> 16);
        r.w = __float_as_int(((zw >> lane) & 1u) ? 1.f : 0.f);
      } else {
        const int64_t idx = perm ? (int64_t)__ldg(perm + start + k) : (start + k);
        r = __ldg(reinterpret_cast<const int4*>(rec) + idx);
      }
      if constexpr (HOT)
        slots = ((int)__ldg(item_slot + r.y) & 0xff) | (((int)__ldg(item_slot + r.z) & 0xff) << 8);
    }
    const int nvalid = (B - base) < 32 ? (int)(B - base) : 32;
    float x_home = 0.f;
    // RUNS: the user row in flight and its gradient, carried across this group's LPT slots of the tile
    uint32_t cur_u = NO_USER;
    Frag<VEC> cu[NITER], accU[NITER];
#pragma unroll
    for (int it = 0; it < NITER; ++it) { cu[it] = frag_zero<VEC>(); accU[it] = frag_zero<VEC>(); }

#pragma unroll 1
    for (int r0 = 0; r0 < LPT; r0 += UNR) {
      Frag<VEC> uu[UNR][NITER], dv[UNR][NITER];
      uint32_t tu[UNR], ti[UNR], tj[UNR];
      int ts[UNR];
      float tz[UNR];
      bool fresh[UNR];                                 // RUNS: this triplet starts a new user run
#pragma unroll
      for (int q = 0; q < UNR; ++q) {
        const int e = grp * LPT + r0 + q;              // tile slot this group handles in this round
        tu[q] = (uint32_t)__shfl_sync(0xffffffffu, r.x, e);
        ti[q] = (uint32_t)__shfl_sync(0xffffffffu, r.y, e);
        tj[q] = (uint32_t)__shfl_sync(0xffffffffu, r.z, e);
        tz[q] = __int_as_float(__shfl_sync(0xffffffffu, r.w, e));
        ts[q] = HOT ? __shfl_sync(0xffffffffu, slots, e) : 0xffff;
        if constexpr (RUNS) {
          if (e >= nvalid) tu[q] = (q == 0) ? cur_u : tu[q > 0 ? q - 1 : 0];   // padding never opens a run
          fresh[q] = tu[q] != ((q == 0) ? cur_u : tu[q > 0 ? q - 1 : 0]);
        } else {
          fresh[q] = true;
        }
        const char* pu = Ub + ((uint64_t)tu[q] * ROWB + lane_off);
        const char* pi = Vb + ((uint64_t)ti[q] * ROWB + lane_off);
        const char* pj = Vb + ((uint64_t)tj[q] * ROWB + lane_off);
#pragma unroll
        for (int it = 0; it < NITER; ++it) {
          if (fresh[q] && (!RUNS || tu[q] != NO_USER))
            uu[q][it] = ldg_frag<VEC>(reinterpret_cast<const float*>(pu + it * STEP));
          const Frag<VEC> a = ldg_frag<VEC>(reinterpret_cast<const float*>(pi + it * STEP));
          const Frag<VEC> b = ldg_frag<VEC>(reinterpret_cast<const float*>(pj + it * STEP));
#pragma unroll
          for (int kk = 0; kk < VEC; ++kk) dv[q][it].v[kk] = a.v[kk] - b.v[kk];
        }
      }
#pragma unroll
      for (int q = 0; q < UNR; ++q) {
        const bool ok = grp * LPT + r0 + q < nvalid;   // false only in the last, partial tile
        if constexpr (RUNS) {
          if (fresh[q]) {
            if (cur_u != NO_USER) {                    // close the previous run: one reduction per row
              char* du = gUb + ((uint64_t)cur_u * ROWB + lane_off);
#pragma unroll
              for (int it = 0; it < NITER; ++it) red_frag<VEC>(reinterpret_cast<float*>(du + it * STEP), accU[it]);
            }
#pragma unroll
            for (int it = 0; it < NITER; ++it) { accU[it] = frag_zero<VEC>(); cu[it] = uu[q][it]; }
            cur_u = tu[q];
          }
        } else {
#pragma unroll
          for (int it = 0; it < NITER; ++it) cu[it] = uu[q][it];
        }
        float part = 0.f;
#pragma unroll
        for (int it = 0; it < NITER; ++it)
#pragma unroll
          for (int kk = 0; kk < VEC; ++kk) part = fmaf(cu[it].v[kk], dv[q][it].v[kk], part);
        const float x = group_sum<LPT>(part, 0xffffffffu);
        x_home = (sub == r0 + q) ? x : x_home;
        float g = bce_grad_score_fast(sigmoid_fast(x), tz[q], inv_batch);
        if (!ok) g = 0.f;
        const float ng = -g;
        const int si = ts[q] & 0xff, sj = ts[q] >> 8;
        const bool cold_i = !HOT || si == 0xff;
        const bool cold_j = !HOT || sj == 0xff;
        if constexpr (RUNS) {
#pragma unroll
          for (int it = 0; it < NITER; ++it)
#pragma unroll
            for (int kk = 0; kk < VEC; ++kk) accU[it].v[kk] = fmaf(g, dv[q][it].v[kk], accU[it].v[kk]);
        }
        if (ok) {
          char* di = gVb + ((uint64_t)ti[q] * ROWB + lane_off);
          char* dj = gVb + ((uint64_t)tj[q] * ROWB + lane_off);
#pragma unroll
          for (int it = 0; it < NITER; ++it) {
            Frag<VEC> b, nb;
#pragma unroll
            for (int kk = 0; kk < VEC; ++kk) {
              b.v[kk] = g * cu[it].v[kk];
              nb.v[kk] = ng * cu[it].v[kk];
            }
            if constexpr (!RUNS) {
              Frag<VEC> a;
#pragma unroll
              for (int kk = 0; kk < VEC; ++kk) a.v[kk] = g * dv[q][it].v[kk];
              char* du = gUb + ((uint64_t)tu[q] * ROWB + lane_off);
              red_frag<VEC>(reinterpret_cast<float*>(du + it * STEP), a);
            }
            if (cold_i) red_frag<VEC>(reinterpret_cast<float*>(di + it * STEP), b);
            if (cold_j) red_frag<VEC>(reinterpret_cast<float*>(dj + it * STEP), nb);
            if constexpr (HOT && !SHARE) {
              // this group's own image: no other lane touches these 16 bytes, program order is enough
              if (!cold_i) smem_add4_at(my_hot + si * ROWB + it * STEP, b);
              if (!cold_j) smem_add4_at(my_hot + sj * ROWB + it * STEP, nb);
            }
          }
        }
        if constexpr (HOT && SHARE) {
          // one image per warp: its lane groups update it one after the other (same-warp shared-memory
          // accesses complete in program order, so group 1's read sees group 0's write)
#pragma unroll
          for (int ph = 0; ph < GPW; ++ph) {
            if (grp == ph && ok) {
#pragma unroll
              for (int it = 0; it < NITER; ++it) {
                Frag<VEC> b, nb;
#pragma unroll
                for (int kk = 0; kk < VEC; ++kk) {
                  b.v[kk] = g * cu[it].v[kk];
                  nb.v[kk] = ng * cu[it].v[kk];
                }
                if (!cold_i) smem_add4_at(my_hot + si * ROWB + it * STEP, b);
                if (!cold_j) smem_add4_at(my_hot + sj * ROWB + it * STEP, nb);
              }
            }
            // independent thread scheduling gives no ordering between the divergent phases of a warp:
            // the barrier (and its memory ordering among the participating lanes) makes group ph+1's
            // read of the image see group ph's write whatever code the compiler emits
            __syncwarp();
          }
        }
      }
    }
    if constexpr (RUNS) {
      if (cur_u != NO_USER) {                          // the run still open at the end of the tile
        char* du = gUb + ((uint64_t)cur_u * ROWB + lane_off);
#pragma unroll
        for (int it = 0; it < NITER; ++it) red_frag<VEC>(reinterpret_cast<float*>(du + it * STEP), accU[it]);
      }
    }
    // the tile's losses, one triplet per lane (exact sigmoid / BCE: this is the reported number)
    if (lane < nvalid) loss_acc += bce_ref(sigmoidf_ref(x_home), __int_as_float(r.w));
  }

  if constexpr (HOT) {
    // one reduction per privatised row element and CTA
    __syncthreads();
    const int per_img = n_hot * D;
    for (int e = threadIdx.x; e < per_img; e += kLeanBlock) {
      float t = 0.f;
#pragma unroll 4
      for (int w = 0; w < IMAGES; ++w) t += s_hot[w * per_img + e];
      if (t != 0.f) atomicAdd(gV + (int64_t)__ldg(hot_items + e / D) * D + (e % D), t);
    }
  }
  // block loss: warp sums -> warp 0
  loss_acc = warp_sum(loss_acc);
  if (lane == 0) s_red[threadIdx.x >> 5] = loss_acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = (lane < kLeanBlock / 32) ? s_red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) atomicAdd(loss_out, t * inv_batch);
  }
}

// ===========================================================================================================
// K1 span: the user-grouped atomic kernel for d >= 64 (LPT 16 or 32), third profile pass (profiles/r02*).
//
// k_fwd_bwd_lean spends a third of its 81 warp instructions per triplet on the hot item rows: every triplet does
// LDS.128 + 4 FADD + STS.128 per hot row, and the lane groups of a warp take turns (one phase per group).  But
// the gradient a RUN of one user's triplets sends to hot row h is  (sum over the run of +-g) * U_u : a scalar per
// (run, hot row) times the row the group already holds.  So here
//   * lane `sub` of a group keeps the scalar weights of hot slots sub, sub + LPT, ... in registers; a triplet costs
//     two compares + two predicated adds per slot register, no shared-memory traffic at all;
//   * when the run closes, the group walks the slots with a non-zero weight (one ballot, then ffs) and adds
//     weight * U_u to its warp's shared image: one LDS/4 FFMA/STS per DISTINCT hot row of the run instead of one per
//     reference.  The two groups of a warp walk their lists from opposite ends, so they meet on the same row at
//     most once per flush (then, and only then, they take turns);
//   * every lane group owns a CONTIGUOUS span of the batch instead of LPT slots of every warp tile, so a run is cut
//     only at span ends (a few hundred triplets apart): with ~42 triplets per user per batch at config 4 the user
//     row is read once and its gradient leaves as one reduction per ~42 triplets (it was ~16), and a flush covers
//     a whole run.
// Same arithmetic per triplet as the lean kernel (fast sigmoid for g, exact sigmoid/BCE for the reported loss).
// ===========================================================================================================
struct Frag2;
template <int LPT, int NITER>
__device__ __forceinline__ void span_image_add(uint32_t addr_row, float wv, const Frag2 (&cu)[NITER]);

// rows of a warp's hot image in the span kernel: all 32 slots at LPT = 16 (the flush walks them unconditionally),
// the privatised rows only at LPT = 32
template <int LPT>
__host__ __device__ constexpr int span_image_rows(int n_hot) { return LPT == 16 ? 32 : n_hot; }

// Called by ALL lanes of the warp; `need` is uniform within a lane group.  Adds weight * cu to the warp's image
// for every slot, then clears the weights.  No search for the non-zero weights: a first version that walked them
// (ballot + ffs, the two groups from opposite ends, a clash check per step) cost 45 instructions per visited row;
// this one costs 7 (SHFL, LDS.128, 4 FFMA, STS.128) for each of the 32 slots, once per run of ~42 triplets.
// LPT = 16: while group 0 is on slot h, group 1 is on slot h ^ 16, so the two never touch the same row at once.
template <int LPT, int NITER>
__device__ __forceinline__ void span_flush_weights(bool need, float (&w)[32 / LPT], const Frag2 (&cu)[NITER],
                                                   uint32_t my_hot, int n_hot, int lane) {
  constexpr uint32_t ROWB = 16u * LPT * NITER;
  if constexpr (LPT == 32) {
    (void)need;                                          // one group per warp: need is warp-uniform and true
#pragma unroll 4
    for (int h = 0; h < n_hot; ++h) {
      const float wv = __shfl_sync(0xffffffffu, w[0], h);
      span_image_add<LPT, NITER>(my_hot + h * ROWB, wv, cu);
    }
    w[0] = 0.f;
  } else {
    static_assert(LPT == 16, "span kernel: LPT 16 or 32");
    const bool g1 = (lane & 16) != 0;
    // group 0 walks slots 0..31, group 1 walks 16..31, 0..15: first its `wa` half, then its `wb` half
    const float wa = need ? (g1 ? w[1] : w[0]) : 0.f;
    const float wb = need ? (g1 ? w[0] : w[1]) : 0.f;
    const uint32_t base_a = my_hot + (g1 ? 16u * ROWB : 0u);
    const uint32_t base_b = my_hot + (g1 ? 0u : 16u * ROWB);
    const int src = lane & 16;
#pragma unroll 4
    for (int h = 0; h < 16; ++h) {
      const float wv = __shfl_sync(0xffffffffu, wa, src | h);
      span_image_add<LPT, NITER>(base_a + h * ROWB, wv, cu);
    }
    __syncwarp();                                        // the groups swap halves
#pragma unroll 4
    for (int h = 0; h < 16; ++h) {
      const float wv = __shfl_sync(0xffffffffu, wb, src | h);
      span_image_add<LPT, NITER>(base_b + h * ROWB, wv, cu);
    }
    __syncwarp();
    if (need) { w[0] = 0.f; w[1] = 0.f; }
  }
}

// sigmoid for the gradient: ex2.approx on -x log2(e), rcp.approx (relative error ~1e-6 for |x| < 30; the reported
// loss uses the exact functions)
__device__ __forceinline__ float sigmoid_fast2(float x) {
  float e, p;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(1.f + e));
  return p;
}

// a row fragment as two packed pairs: (x, y) and (z, w) of the lane's float4
struct Frag2 {
  f32x2 lo, hi;
};
__device__ __forceinline__ Frag2 ldg_frag2(const float4* __restrict__ p) {
  const float4 t = __ldg(p);
  Frag2 f;
  f.lo = pk2(t.x, t.y);
  f.hi = pk2(t.z, t.w);
  return f;
}
__device__ __forceinline__ void red_frag2(float4* p, const Frag2& f) {
  float4 t;
  unpk2(f.lo, t.x, t.y);
  unpk2(f.hi, t.z, t.w);
  atomicAdd(p, t);
}

template <int LPT, int NITER>
__device__ __forceinline__ void span_image_add(uint32_t addr_row, float wv, const Frag2 (&cu)[NITER]) {
  constexpr int STEP = LPT * 16;
  const f32x2 w2 = pk2(wv, wv);
#pragma unroll
  for (int it = 0; it < NITER; ++it) {
    f32x2 lo, hi;
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "r"(addr_row + it * STEP));
    lo = fma2(w2, cu[it].lo, lo);
    hi = fma2(w2, cu[it].hi, hi);
    asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(addr_row + it * STEP), "l"(lo), "l"(hi) : "memory");
  }
}

template <int LPT, int NITER, bool HOT>
__global__ void __launch_bounds__(kLeanBlock, NITER > 2 ? 1 : (NITER == 2 ? 2 : 3))
k_fwd_bwd_span(const float* __restrict__ U, const float* __restrict__ V, const mfcd_triplet* __restrict__ rec,
               int64_t start, int64_t B, int64_t span, float inv_batch, float* __restrict__ gU,
               float* __restrict__ gV, float* __restrict__ loss_out, const int8_t* __restrict__ item_slot,
               const int32_t* __restrict__ hot_items, int n_hot) {
  constexpr int VEC = 4;
  constexpr int D = VEC * LPT * NITER;
  constexpr uint32_t ROWB = D * 4;
  constexpr uint32_t ROW4 = LPT * NITER;               // float4 fragments per row
  constexpr int UNR = 2;
  constexpr int NW = 32 / LPT;                         // hot slots per lane (up to 32 hot rows)
  constexpr int IMAGES = kLeanBlock / 32;
  constexpr uint32_t NO_USER = 0xffffffffu;
  __shared__ float s_red[kLeanBlock / 32];
  extern __shared__ __align__(16) float s_hot[];       // HOT: [IMAGES][img_rows][D], one image per warp
  const int img_rows = span_image_rows<LPT>(n_hot);
  if constexpr (HOT) {
    for (int e = threadIdx.x; e < IMAGES * img_rows * D; e += kLeanBlock) s_hot[e] = 0.f;
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const uint32_t sub = lane % LPT;
  // tables as arrays of float4 fragments: fragment (row * ROW4 + sub [+ it * LPT]) -- a 32-bit index (the launcher
  // checks the tables have fewer than 2^32 fragments), so an address is one IMAD + one IMAD.WIDE off the
  // kernel-parameter base
  const float4* U4 = reinterpret_cast<const float4*>(U);
  const float4* V4 = reinterpret_cast<const float4*>(V);
  float4* gU4 = reinterpret_cast<float4*>(gU);
  float4* gV4 = reinterpret_cast<float4*>(gV);
  const uint32_t my_hot = (uint32_t)__cvta_generic_to_shared(s_hot) +
                          (uint32_t)(threadIdx.x >> 5) * (uint32_t)img_rows * ROWB + sub * (VEC * 4);

  const int64_t gid = (blockIdx.x * (int64_t)kLeanBlock + threadIdx.x) / LPT;
  const int64_t s0 = gid * span;                       // this group's records: [s0, s0 + span) of the batch
  const mfcd_triplet* my_rec = rec + start;
  float loss_acc = 0.f;
  uint32_t cur_u = NO_USER;                            // the user run in flight, carried over the whole span
  Frag2 cu[NITER], accU[NITER];
  float w[NW];
#pragma unroll
  for (int it = 0; it < NITER; ++it) { cu[it].lo = cu[it].hi = 0; accU[it].lo = accU[it].hi = 0; }
#pragma unroll
  for (int k = 0; k < NW; ++k) w[k] = 0.f;

  for (int64_t tb = s0; tb < s0 + span; tb += LPT) {
    const int nvalid = tb >= B ? 0 : ((B - tb) < LPT ? (int)(B - tb) : LPT);
    if (__all_sync(0xffffffffu, nvalid == 0)) break;
    int4 r = make_int4(0, 0, 0, 0);
    int slots = 0xffff;                                // (slot_i & 0xff) | (slot_j & 0xff) << 8, 0xff = cold
    if ((int)sub < nvalid) {
      r = __ldg(reinterpret_cast<const int4*>(my_rec) + (tb + sub));
      if constexpr (HOT)
        slots = ((int)__ldg(item_slot + r.y) & 0xff) | (((int)__ldg(item_slot + r.z) & 0xff) << 8);
    } else {
      r.x = (int)NO_USER;                              // padding: continues whatever run is open (see below)
    }
    float x_home = 0.f;
#pragma unroll 1
    for (int r0 = 0; r0 < LPT; r0 += UNR) {
      Frag2 uu[UNR][NITER], dv[UNR][NITER];
      uint32_t tu[UNR], ti[UNR], tj[UNR];
      int ts[UNR];
      float tz[UNR];
      bool fresh[UNR];
#pragma unroll
      for (int q = 0; q < UNR; ++q) {
        const int e = (lane & ~(LPT - 1)) + r0 + q;    // source lane: slot r0 + q of this group's tile
        tu[q] = (uint32_t)__shfl_sync(0xffffffffu, r.x, e);
        ti[q] = (uint32_t)__shfl_sync(0xffffffffu, r.y, e);
        tj[q] = (uint32_t)__shfl_sync(0xffffffffu, r.z, e);
        tz[q] = __int_as_float(__shfl_sync(0xffffffffu, r.w, e));
        ts[q] = HOT ? __shfl_sync(0xffffffffu, slots, e) : 0xffff;
        const uint32_t prev = (q == 0) ? cur_u : tu[q > 0 ? q - 1 : 0];
        if (tu[q] == NO_USER) tu[q] = prev;            // padding never opens a run
        fresh[q] = tu[q] != prev;
        const uint32_t fi = ti[q] * ROW4 + sub, fj = tj[q] * ROW4 + sub;
        if (fresh[q]) {
          const uint32_t fu = tu[q] * ROW4 + sub;
#pragma unroll
          for (int it = 0; it < NITER; ++it) uu[q][it] = ldg_frag2(U4 + (fu + it * LPT));
        }
#pragma unroll
        for (int it = 0; it < NITER; ++it) {
          const Frag2 a = ldg_frag2(V4 + (fi + it * LPT));
          const Frag2 b = ldg_frag2(V4 + (fj + it * LPT));
          dv[q][it].lo = sub2(a.lo, b.lo);
          dv[q][it].hi = sub2(a.hi, b.hi);
        }
      }
#pragma unroll
      for (int q = 0; q < UNR; ++q) {
        const bool ok = r0 + q < nvalid;
        // a run closes (rare: ~1 in 42 triplets per group at config 4): its hot-row weights go to the image, its
        // user gradient leaves as one reduction per row, the new user's row becomes the current one
        if (__any_sync(0xffffffffu, fresh[q])) {
          if constexpr (HOT)
            span_flush_weights<LPT, NITER>(fresh[q] && cur_u != NO_USER, w, cu, my_hot, n_hot, lane);
          if (fresh[q]) {
            if (cur_u != NO_USER) {
              const uint32_t fu = cur_u * ROW4 + sub;
#pragma unroll
              for (int it = 0; it < NITER; ++it) red_frag2(gU4 + (fu + it * LPT), accU[it]);
            }
#pragma unroll
            for (int it = 0; it < NITER; ++it) { accU[it].lo = accU[it].hi = 0; cu[it] = uu[q][it]; }
            cur_u = tu[q];
          }
        }
        f32x2 part2 = mul2(cu[0].lo, dv[q][0].lo);
        part2 = fma2(cu[0].hi, dv[q][0].hi, part2);
#pragma unroll
        for (int it = 1; it < NITER; ++it) {
          part2 = fma2(cu[it].lo, dv[q][it].lo, part2);
          part2 = fma2(cu[it].hi, dv[q][it].hi, part2);
        }
        const float x = group_sum<LPT>(sum2(part2), 0xffffffffu);
        x_home = ((int)sub == r0 + q) ? x : x_home;
        const float p = sigmoid_fast2(x);
        float g = inv_batch * (p - tz[q]);
        if (fabsf(x) > 27.f) {                         // p(1-p) may be below the reference's 1e-12 clamp: rare
          const float omp = 1.f - p, qq = omp * p;
          if (qq < 1e-12f) g = g * 1e12f * omp * p;
          asm volatile("" ::: "memory");               // keep this a branch, not five predicated instructions
        }
        if (!ok) g = 0.f;
        const int si = ts[q] & 0xff, sj = ts[q] >> 8;
        const bool cold_i = !HOT || si == 0xff;
        const bool cold_j = !HOT || sj == 0xff;
        const f32x2 g2 = pk2(g, g);
#pragma unroll
        for (int it = 0; it < NITER; ++it) {
          accU[it].lo = fma2(g2, dv[q][it].lo, accU[it].lo);
          accU[it].hi = fma2(g2, dv[q][it].hi, accU[it].hi);
        }
        if constexpr (HOT) {
#pragma unroll
          for (int k = 0; k < NW; ++k) {
            const int s = (int)sub + LPT * k;
            if (si == s) w[k] += g;
            if (sj == s) w[k] -= g;
          }
        }
        if (ok && (cold_i || cold_j)) {
          const f32x2 ng2 = pk2(-g, -g);
          const uint32_t fi = ti[q] * ROW4 + sub, fj = tj[q] * ROW4 + sub;
#pragma unroll
          for (int it = 0; it < NITER; ++it) {
            Frag2 b, nb;
            b.lo = mul2(g2, cu[it].lo); b.hi = mul2(g2, cu[it].hi);
            nb.lo = mul2(ng2, cu[it].lo); nb.hi = mul2(ng2, cu[it].hi);
            if (cold_i) red_frag2(gV4 + (fi + it * LPT), b);
            if (cold_j) red_frag2(gV4 + (fj + it * LPT), nb);
          }
        }
      }
    }
    // the tile's losses, one triplet per lane (exact sigmoid / BCE: this is the reported number)
    if ((int)sub < nvalid) loss_acc += bce_ref(sigmoidf_ref(x_home), __int_as_float(r.w));
  }
  // the run still open at the end of the span
  if constexpr (HOT) {
    const bool close = cur_u != NO_USER;
    if (__any_sync(0xffffffffu, close)) span_flush_weights<LPT, NITER>(close, w, cu, my_hot, n_hot, lane);
  }
  if (cur_u != NO_USER) {
    const uint32_t fu = cur_u * ROW4 + sub;
#pragma unroll
    for (int it = 0; it < NITER; ++it) red_frag2(gU4 + (fu + it * LPT), accU[it]);
  }

  if constexpr (HOT) {
    __syncthreads();
    const int per_img = img_rows * D;
    for (int e = threadIdx.x; e < n_hot * D; e += kLeanBlock) {
      float t = 0.f;
#pragma unroll 4
      for (int wi = 0; wi < IMAGES; ++wi) t += s_hot[wi * per_img + e];
      if (t != 0.f) atomicAdd(gV + (int64_t)__ldg(hot_items + e / D) * D + (e % D), t);
    }
  }
  loss_acc = warp_sum(loss_acc);
  if (lane == 0) s_red[threadIdx.x >> 5] = loss_acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = (lane < kLeanBlock / 32) ? s_red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) atomicAdd(loss_out, t * inv_batch);
  }
}

static inline bool span_enabled() {
  static const bool on = !(getenv("MFCD_K1_SPAN") && atoi(getenv("MFCD_K1_SPAN")) == 0);
  return on;
}

template <int LPT, int NITER, bool HOT>
static int launch_span_kernel(const float* U, const float* V, const mfcd_triplet* rec, int64_t start, int64_t B,
                              float inv_batch, float* gU, float* gV, float* loss, const int8_t* item_slot,
                              const int32_t* hot_items, int n_hot, cudaStream_t st) {
  auto kern = k_fwd_bwd_span<LPT, NITER, HOT>;
  constexpr int want = NITER > 2 ? 1 : (NITER == 2 ? 2 : 3);
  const size_t smem = HOT ? (size_t)(kLeanBlock / 32) * span_image_rows<LPT>(n_hot) * (4 * LPT * NITER) * sizeof(float) : 0;
  int per_sm = want;
  if (HOT) {
    if (smem > 40 * 1024)
      MFCD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int fit = (int)((227 * 1024) / (smem + 1024));
    per_sm = fit < want ? (fit < 1 ? 1 : fit) : want;
  }
  static const int min_tiles = getenv("MFCD_K1_SPAN_MIN_TILES") ? atoi(getenv("MFCD_K1_SPAN_MIN_TILES")) : 4;
  constexpr int groups_per_cta = kLeanBlock / LPT;
  const int grid = grid_for(B, groups_per_cta * LPT * (min_tiles < 1 ? 1 : min_tiles), per_sm);
  const int64_t groups = (int64_t)grid * groups_per_cta;
  int64_t span = (B + groups - 1) / groups;
  span = (span + LPT - 1) / LPT * LPT;                 // whole tiles of LPT records
  kern<<<grid, kLeanBlock, smem, st>>>(U, V, rec, start, B, span, inv_batch, gU, gV, loss, item_slot, hot_items, n_hot);
  MFCD_CHECK_LAUNCH();
  return MFCD_OK;
}

// shared memory one hot row costs in the lean kernel: an image per lane group, or per warp (SHARE)
static inline bool lean_share() {
  static const bool on = !(getenv("MFCD_K1_SHARE") && atoi(getenv("MFCD_K1_SHARE")) == 0);
  return on;
}
static inline size_t lean_hot_row_bytes(int lpt, int niter) {
  const int images = lean_share() ? kLeanBlock / 32 : kLeanBlock / lpt;
  return (size_t)images * 4 * lpt * niter * sizeof(float);
}
static inline int hot_smem_budget() {
  static const int kb = getenv("MFCD_HOT_SMEM_KB") ? atoi(getenv("MFCD_HOT_SMEM_KB")) : 64;
  return kb * 1024;
}
static inline bool lean_enabled() {
  static const bool on = !(getenv("MFCD_K1_LEAN") && atoi(getenv("MFCD_K1_LEAN")) == 0);
  return on;
}
// d that the lean kernel covers: float4 fragments that exactly fill the lane group
static inline bool lean_shape(int d, const RowShape& s) { return s.vec == 4 && d == 4 * s.lpt * s.niter; }

template <int LPT, int NITER, bool HOT, bool RUNS, bool SHARE, bool WIRE = false>
static int launch_lean_kernel(const float* U, const float* V, const mfcd_triplet* rec, const int32_t* perm,
                              int64_t start, int64_t B, float inv_batch, float* gU, float* gV, float* loss,
                              const int8_t* item_slot, const int32_t* hot_items, int n_hot, size_t smem,
                              cudaStream_t st) {
  auto kern = k_fwd_bwd_lean<LPT, NITER, HOT, RUNS, SHARE, WIRE>;
  constexpr int want = lean_min_blocks<NITER, HOT, SHARE>();
  int per_sm = want;
  if (HOT) {
    if (smem > 40 * 1024)   // opt-in above the default 48 KB (static + dynamic) limit; per device, so not cached
      MFCD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int fit = (int)((227 * 1024) / (smem + 1024));
    per_sm = fit < want ? (fit < 1 ? 1 : fit) : want;
  }
  // a block pass covers 8 warp tiles of 32 triplets; HOT CTAs take at least two passes (each CTA flushes its hot
  // images once).  The batch-size sweep showed the earlier 8 passes per CTA starving mid-size batches: 32 CTAs
  // for 65536 triplets, K1 0.12 ms whatever the batch below 2^20 (profiles/r01_notes.md).
  const int grid = grid_for(B, HOT ? kLeanBlock * 2 : kLeanBlock, per_sm);
  kern<<<grid, kLeanBlock, smem, st>>>(U, V, rec, perm, start, B, inv_batch, gU, gV, loss, item_slot, hot_items, n_hot);
  MFCD_CHECK_LAUNCH();
  return MFCD_OK;
}

template <int LPT, int NITER>
static int launch_lean(const float* U, const float* V, const mfcd_triplet* rec, const int32_t* perm, int64_t start,
                       int64_t B, float inv_batch, float* gU, float* gV, float* loss, const int8_t* item_slot,
                       const int32_t* hot_items, int n_hot, bool runs, bool wire, cudaStream_t st) {
  if constexpr (LPT >= 16) {
    if (runs && !wire && perm == nullptr && n_hot <= 32 && span_enabled()) {
      if (n_hot > 0)
        return launch_span_kernel<LPT, NITER, true>(U, V, rec, start, B, inv_batch, gU, gV, loss, item_slot,
                                                    hot_items, n_hot, st);
      return launch_span_kernel<LPT, NITER, false>(U, V, rec, start, B, inv_batch, gU, gV, loss, nullptr, nullptr, 0, st);
    }
  }
#define MFCD_LEAN_GO(HOT, RUNS, SHARE, WIRE)                                                              \
  return launch_lean_kernel<LPT, NITER, HOT, RUNS, SHARE, WIRE>(U, V, rec, perm, start, B, inv_batch, gU, gV, \
                                                                loss, item_slot, hot_items, n_hot, smem, st)
  if (n_hot > 0) {
    const size_t smem = lean_hot_row_bytes(LPT, NITER) * n_hot;
    if constexpr (LPT < 32) {
      if (lean_share()) {
        if (wire) MFCD_LEAN_GO(true, true, true, true);
        if (runs) MFCD_LEAN_GO(true, true, true, false); else MFCD_LEAN_GO(true, false, true, false);
      }
    }
    if (wire) MFCD_LEAN_GO(true, true, false, true);
    if (runs) MFCD_LEAN_GO(true, true, false, false); else MFCD_LEAN_GO(true, false, false, false);
  } else {
    const size_t smem = 0;
    if (wire) MFCD_LEAN_GO(false, true, false, true);
    if (runs) MFCD_LEAN_GO(false, true, false, false); else MFCD_LEAN_GO(false, false, false, false);
  }
#undef MFCD_LEAN_GO
}

static int dispatch_lean(const RowShape& s, const float* U, const float* V, const mfcd_triplet* rec,
                         const int32_t* perm, int64_t start, int64_t B, float inv_batch, float* gU, float* gV,
                         float* loss, const int8_t* item_slot, const int32_t* hot_items, int n_hot, bool runs,
                         bool wire, cudaStream_t st) {
#define MFCD_LEAN_CASE(L, N) \
  return launch_lean<L, N>(U, V, rec, perm, start, B, inv_batch, gU, gV, loss, item_slot, hot_items, n_hot, runs, wire, st)
  switch (s.lpt) {
    case 1: MFCD_LEAN_CASE(1, 1);
    case 2: MFCD_LEAN_CASE(2, 1);
    case 4: MFCD_LEAN_CASE(4, 1);
    case 8: MFCD_LEAN_CASE(8, 1);
    case 16: MFCD_LEAN_CASE(16, 1);
    default:
      switch (s.niter) {
        case 1: MFCD_LEAN_CASE(32, 1);
        case 2: MFCD_LEAN_CASE(32, 2);
        case 3: MFCD_LEAN_CASE(32, 3);
        default: MFCD_LEAN_CASE(32, 4);
      }
  }
#undef MFCD_LEAN_CASE
}

}  // namespace mfcd
