// K1: fused forward + BCE + backward for one batch of comparison triplets.
//
// Replaces, per optimiser step of the reference (structure.py:848-850):
//   pred = sigmoid(sum(U[u] * (V[i] - V[j]), dim=1))        structure.py:787-795
//   loss = F.binary_cross_entropy(pred, z.float())           structure.py:849
//   loss.backward()   -> dense U.grad / V.grad               structure.py:850
//
// Layout: a lane-aligned sub-warp group of LPT lanes owns one triplet; each
// lane holds VEC consecutive floats of the three rows (128-bit loads when
// d % 4 == 0).  A warp loads 32 consecutive 16-byte records with one coalesced
// LDG.128 per lane and broadcasts the indices with shuffles; UNR triplets per
// group are kept in flight so that 3*UNR row loads overlap.  Gradients leave as
// REDG.E.ADD.F32x4 (fast mode) -- HBM/L2-bound byte traffic, no tensor cores.
//
// Algorithmic bytes per triplet: 16 (record) + 3*4*d (row reads) + 3*4*d
// (gradient row updates) = 16 + 24 d.
#include <stdlib.h>
#include "internal.h"
#include "shape_dispatch.cuh"
#include <string.h>
#include "k1_lean.cuh"

namespace mfcd {

constexpr int kBlock = 256;

// sums a per-thread fp32 value over the block (all threads must call)
__device__ __forceinline__ float block_sum_256(float x, float* smem8) {
  x = warp_sum(x);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) smem8[w] = x;
  __syncthreads();
  float t = 0.f;
  if (w == 0) {
    t = (lane < (kBlock / 32)) ? smem8[lane] : 0.f;
    t = warp_sum(t);
  }
  return t;  // valid in warp 0
}

// Hot-row privatisation (HOT = true, atomic mode only).  Under popularity-biased
// sampling a handful of item rows receive a large share of all gradient updates
// (zipf 1.5: item 0 is in 38% of the pairs) and the L2 atomic unit serialises
// same-line REDs at ~1 op/cycle, which caps the whole kernel.  The caller names
// up to n_hot such rows (item_slot[row] = slot or -1); every warp keeps a private
// fp32 accumulator image of those rows in shared memory, updates it with plain
// LDS/STS (lane groups of one warp take turns, so there are no shared-memory
// atomics), and the CTA flushes the 8 images as ONE reduction per row at the end.
struct HotRows {
  const int8_t* item_slot;   // [n_items] slot id or -1
  const int32_t* hot_items;  // [n_hot] row id of each slot
  int n_hot;
};

template <int VEC>
__device__ __forceinline__ void smem_add_frag(float* p, const Frag<VEC>& f) {
  Frag<VEC> cur;
  if constexpr (VEC == 4) {
    float4 t = *reinterpret_cast<float4*>(p);
    t.x += f.v[0]; t.y += f.v[1]; t.z += f.v[2]; t.w += f.v[3];
    *reinterpret_cast<float4*>(p) = t;
  } else if constexpr (VEC == 2) {
    float2 t = *reinterpret_cast<float2*>(p);
    t.x += f.v[0]; t.y += f.v[1];
    *reinterpret_cast<float2*>(p) = t;
  } else {
    *p += f.v[0];
  }
  (void)cur;
}

// MODE 0: atomic scatter into gU/gV.  MODE 1: write g_b to gbuf (deterministic path).
// VAR (tuning variant, env MFCD_K1_VARIANT): 2 (default) = 2 triplets in flight per lane group, 4 CTAs/SM
// (56-62 registers); 0 = 4 in flight, 3 CTAs/SM (80 registers).  Measured on B200 at d=64, B=2^21:
// var 2 is 8% faster with hot rows, equal on uniform items, 10% faster at d=32 (profiles/r01_notes.md).
template <int NITER, int VAR, bool HOT>
constexpr int k1_min_blocks() { return NITER > 2 ? 2 : (NITER == 2 ? 3 : (VAR == 0 ? 3 : 4)); }

// One batch of UNR rounds of a warp tile.  Round rr of lane group `grp` handles tile slot grp*LPT + rr, so
// the slot's "home lane" (the lane that loaded its record) sits in the same group with sub == rr: the score x
// is parked there with a select and the tile's 32 losses are evaluated once, fully SIMD, after the rounds.
// CHECK = false for full tiles and rows that exactly fill the lane group (no bounds tests in the hot loop).
template <int VEC, int LPT, int NITER, int MODE, bool HOT, int UNR, bool CHECK>
__device__ __forceinline__ void k1_rounds(const float* __restrict__ U, const float* __restrict__ V, const int4& r,
                                          int slots, int nvalid, int d, float inv_batch, float* __restrict__ gU,
                                          float* __restrict__ gV, float* __restrict__ gbuf, int64_t base,
                                          float* my_hot, int lane, float& x_home) {
  constexpr int GPW = 32 / LPT;
  const int sub = lane % LPT;
  const int grp = lane / LPT;
#pragma unroll 1
  for (int r0 = 0; r0 < LPT; r0 += UNR) {
    TripletRows<VEC, LPT, NITER> rows[UNR];
    int tu[UNR], ti[UNR], tj[UNR], ts[UNR];
    float tz[UNR];
    bool ok[UNR];
#pragma unroll
    for (int q = 0; q < UNR; ++q) {
      const int e = grp * LPT + r0 + q;                        // tile slot this group handles in this round
      tu[q] = __shfl_sync(0xffffffffu, r.x, e);
      ti[q] = __shfl_sync(0xffffffffu, r.y, e);
      tj[q] = __shfl_sync(0xffffffffu, r.z, e);
      tz[q] = __int_as_float(__shfl_sync(0xffffffffu, r.w, e));
      ts[q] = HOT ? __shfl_sync(0xffffffffu, slots, e) : 0xffff;
      ok[q] = CHECK ? (e < nvalid) : true;
      if constexpr (CHECK) {
        load_rows<VEC, LPT, NITER>(rows[q], U, V, tu[q], ti[q], tj[q], d, sub, ok[q]);
      } else {
        const float* pu = U + (int64_t)tu[q] * d;
        const float* pi = V + (int64_t)ti[q] * d;
        const float* pj = V + (int64_t)tj[q] * d;
#pragma unroll
        for (int it = 0; it < NITER; ++it) {
          const int c = (it * LPT + sub) * VEC;
          rows[q].uu[it] = ldg_frag<VEC>(pu + c);
          Frag<VEC> a = ldg_frag<VEC>(pi + c);
          Frag<VEC> b = ldg_frag<VEC>(pj + c);
#pragma unroll
          for (int k = 0; k < VEC; ++k) rows[q].dv[it].v[k] = a.v[k] - b.v[k];
        }
      }
    }
#pragma unroll
    for (int q = 0; q < UNR; ++q) {
      const float x = group_sum<LPT>(partial_dot<VEC, LPT, NITER>(rows[q]), 0xffffffffu);
      x_home = (sub == r0 + q) ? x : x_home;
      const float p = sigmoidf_ref(x);
      const float g = bce_grad_score_ref(p, tz[q], inv_batch);
      if (MODE == 1) {
        if (ok[q] && sub == 0) gbuf[base + grp * LPT + r0 + q] = g;
      } else {
        if (ok[q]) {
          float* du = gU + (int64_t)tu[q] * d;
          float* di = gV + (int64_t)ti[q] * d;
          float* dj = gV + (int64_t)tj[q] * d;
#pragma unroll
          for (int it = 0; it < NITER; ++it) {
            const int c = (it * LPT + sub) * VEC;
            if (!CHECK || c < d) {
              Frag<VEC> a, b, nb;
#pragma unroll
              for (int kk = 0; kk < VEC; ++kk) {
                a.v[kk] = g * rows[q].dv[it].v[kk];
                b.v[kk] = g * rows[q].uu[it].v[kk];
                nb.v[kk] = -b.v[kk];
              }
              red_frag<VEC>(du + c, a);
              if (!HOT || (ts[q] & 0xff) == 0xff) red_frag<VEC>(di + c, b);
              if (!HOT || (ts[q] >> 8) == 0xff) red_frag<VEC>(dj + c, nb);
            }
          }
        }
        if constexpr (HOT) {
          // updates to privatised rows go to the warp's own shared-memory image.  Lane groups may run
          // concurrently unless two of them name the same hot row in this round; then they take turns.
          const bool mine = ok[q] && ts[q] != 0xffff;
          if (__any_sync(0xffffffffu, mine)) {
            const int si = ts[q] & 0xff, sj = ts[q] >> 8;
            bool clash = false;
#pragma unroll
            for (int off = LPT; off < 32; off += LPT) {
              const int o = __shfl_sync(0xffffffffu, ts[q], (lane + off) & 31);
              const int oi = o & 0xff, oj = o >> 8;
              clash = clash || (si != 0xff && (si == oi || si == oj)) || (sj != 0xff && (sj == oi || sj == oj));
            }
            const bool serial = GPW > 1 && __any_sync(0xffffffffu, clash && mine);
            auto add_hot = [&]() {
#pragma unroll
              for (int it = 0; it < NITER; ++it) {
                const int c = (it * LPT + sub) * VEC;
                if (!CHECK || c < d) {
                  Frag<VEC> b, nb;
#pragma unroll
                  for (int kk = 0; kk < VEC; ++kk) {
                    b.v[kk] = g * rows[q].uu[it].v[kk];
                    nb.v[kk] = -b.v[kk];
                  }
                  if (si != 0xff) smem_add_frag<VEC>(my_hot + si * d + c, b);
                  if (sj != 0xff) smem_add_frag<VEC>(my_hot + sj * d + c, nb);
                }
              }
            };
            if (!serial) {
              if (mine) add_hot();
              __syncwarp();
            } else {
#pragma unroll 1
              for (int ph = 0; ph < GPW; ++ph) {
                if (mine && grp == ph) add_hot();
                __syncwarp();
              }
            }
          }
        }
      }
    }
  }
}

template <int VEC, int LPT, int NITER, int MODE, bool HOT, int VAR = 0>
__global__ void __launch_bounds__(kBlock, k1_min_blocks<NITER, VAR, HOT>())
k_fwd_bwd(const float* __restrict__ U, const float* __restrict__ V, const mfcd_triplet* __restrict__ rec,
          const int32_t* __restrict__ perm, int64_t start, int64_t B, int d, float inv_batch,
          float* __restrict__ gU, float* __restrict__ gV, float* __restrict__ gbuf,
          float* __restrict__ loss_out /* MODE 0: scalar accumulator; MODE 1: per-block partials */,
          HotRows hot) {
  constexpr int UNR0 = (NITER > 1 || VAR == 2) ? 2 : 4;
  constexpr int UNR = UNR0 < LPT ? UNR0 : LPT;                    // triplets in flight per group
  __shared__ float s_red[kBlock / 32];
  extern __shared__ __align__(16) float s_hot[];                  // HOT: [8 warps][n_hot][d]
  float* my_hot = nullptr;
  if constexpr (HOT) {
    const int per_warp = hot.n_hot * d;
    for (int e = threadIdx.x; e < (kBlock / 32) * per_warp; e += kBlock) s_hot[e] = 0.f;
    __syncthreads();
    my_hot = s_hot + (threadIdx.x >> 5) * per_warp;
  }

  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const bool exact = (d == LPT * VEC * NITER);                    // every lane of a group owns valid columns

  float loss_acc = 0.f;
  for (int64_t base = warp0 * 32; base < B; base += nwarps * 32) {
    const int64_t k = base + lane;
    int4 r = make_int4(0, 0, 0, 0);
    int slots = 0xffff;                                            // (slot_i & 0xff) | (slot_j & 0xff) << 8, 0xff = cold
    if (k < B) {
      const int64_t idx = perm ? (int64_t)__ldg(perm + start + k) : (start + k);
      r = __ldg(reinterpret_cast<const int4*>(rec) + idx);
      if constexpr (HOT)
        slots = ((int)__ldg(hot.item_slot + r.y) & 0xff) | (((int)__ldg(hot.item_slot + r.z) & 0xff) << 8);
    }
    const int nvalid = (B - base) < 32 ? (int)(B - base) : 32;
    float x_home = 0.f;
    if (exact && nvalid == 32)
      k1_rounds<VEC, LPT, NITER, MODE, HOT, UNR, false>(U, V, r, slots, nvalid, d, inv_batch, gU, gV, gbuf, base,
                                                       my_hot, lane, x_home);
    else
      k1_rounds<VEC, LPT, NITER, MODE, HOT, UNR, true>(U, V, r, slots, nvalid, d, inv_batch, gU, gV, gbuf, base,
                                                      my_hot, lane, x_home);
    // the tile's losses, one triplet per lane
    if (lane < nvalid) loss_acc += bce_ref(sigmoidf_ref(x_home), __int_as_float(r.w));
  }

  if constexpr (HOT && MODE == 0) {
    // one reduction per privatised row and CTA
    __syncthreads();
    const int per_warp = hot.n_hot * d;
    for (int e = threadIdx.x; e < per_warp; e += kBlock) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kBlock / 32; ++w) t += s_hot[w * per_warp + e];
      if (t != 0.f) atomicAdd(gV + (int64_t)__ldg(hot.hot_items + e / d) * d + (e % d), t);
    }
  }

  const float tot = block_sum_256(loss_acc, s_red);
  if (threadIdx.x == 0) {
    if (MODE == 0) atomicAdd(loss_out, tot * inv_batch);
    else loss_out[blockIdx.x] = tot;
  }
}

template <int VEC, int LPT, int NITER>
struct AtomicLauncher {
  static int run(const float* U, const float* V, const mfcd_triplet* rec, const int32_t* perm, int64_t start,
                 int64_t B, int d, float inv_batch, float* gU, float* gV, float* loss, HotRows hot,
                 cudaStream_t st) {
    static const int var = getenv("MFCD_K1_VARIANT") ? atoi(getenv("MFCD_K1_VARIANT")) : 2;
    if (hot.n_hot > 0) {
      // privatised hot rows: dynamic smem = 8 warp images; taking turns needs <= 4 groups per warp
      if constexpr (LPT >= 8) {
        const size_t smem = sizeof(float) * (kBlock / 32) * (size_t)hot.n_hot * d;
        auto kern = var == 2 ? k_fwd_bwd<VEC, LPT, NITER, 0, true, 2> : k_fwd_bwd<VEC, LPT, NITER, 0, true, 0>;
        if (smem > 40 * 1024)   // opt-in above the default 48 KB (static + dynamic) limit; per device, so not cached
          MFCD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int per_sm = smem > 52 * 1024 ? 3 : 4;
        const int grid = grid_for(B, kBlock * 8, per_sm);     // long-lived CTAs: few flushes per hot row
        kern<<<grid, kBlock, smem, st>>>(U, V, rec, perm, start, B, d, inv_batch, gU, gV, nullptr, loss, hot);
        MFCD_CHECK_LAUNCH();
        return MFCD_OK;
      }
    }
    auto kern = var == 2 ? k_fwd_bwd<VEC, LPT, NITER, 0, false, 2> : k_fwd_bwd<VEC, LPT, NITER, 0, false, 0>;
    const int grid = grid_for(B, kBlock, var == 0 ? 3 : 4);   // a block pass covers 8 warp tiles of 32 triplets
    kern<<<grid, kBlock, 0, st>>>(U, V, rec, perm, start, B, d, inv_batch, gU, gV, nullptr, loss,
                                  HotRows{nullptr, nullptr, 0});
    MFCD_CHECK_LAUNCH();
    return MFCD_OK;
  }
};

int max_hot_rows(int d) {
  // 64 KB of shared memory for the 8 per-warp images; slots are int8 (<= 127)
  RowShape s;
  if (!row_shape_for(d, &s)) return 0;
  int h;
  if (lean_shape(d, s) && lean_enabled())     // lean kernel: an image per lane group
    h = hot_smem_budget() / (int)lean_hot_row_bytes(s.lpt, s.niter);
  else if (s.lpt < 8) return 0;
  else {
    static const int budget_kb = getenv("MFCD_HOT_SMEM_KB") ? atoi(getenv("MFCD_HOT_SMEM_KB")) : 64;
    h = (budget_kb * 1024) / ((kBlock / 32) * d * (int)sizeof(float));
  }
  return h > 127 ? 127 : h;
}

int launch_fwd_bwd_atomic(const float* U, const float* V, const mfcd_triplet* rec, const int32_t* perm,
                          int64_t start, int64_t B, int d, float inv_batch, float* gU, float* gV, float* loss,
                          const int8_t* item_slot, const int32_t* hot_items, int n_hot, int flags,
                          cudaStream_t st) {
  if (B == 0) return MFCD_OK;
  K1Timer timer(st);
  HotRows hot{item_slot, hot_items, (item_slot && hot_items) ? n_hot : 0};
  if (hot.n_hot > max_hot_rows(d)) {
    set_error("too many hot rows for d=%d: %d > %d", d, hot.n_hot, max_hot_rows(d));
    return MFCD_ERR_ARG;
  }
  RowShape shape_l;
  const bool wire = (flags & MFCD_FLAG_WIRE_RLE) != 0;
  if (wire && (perm != nullptr || start != 0)) {
    set_error("MFCD_FLAG_WIRE_RLE: the wire batch is read whole (start = 0, perm = NULL)");
    return MFCD_ERR_ARG;
  }
  if (row_shape_for(d, &shape_l) && lean_shape(d, shape_l) && (lean_enabled() || wire))
    return dispatch_lean(shape_l, U, V, rec, perm, start, B, inv_batch, gU, gV, loss, hot.item_slot, hot.hot_items,
                         hot.n_hot, (flags & MFCD_FLAG_USER_GROUPED) != 0 && perm == nullptr, wire, st);
  if (wire) {
    set_error("MFCD_FLAG_WIRE_RLE needs d in {4, 8, 16, 32, 64, 128, 256, 384, 512}; unpack the batch first (d=%d)", d);
    return MFCD_ERR_UNSUPPORTED;
  }
  MFCD_DISPATCH_ROW_SHAPE(AtomicLauncher, d, U, V, rec, perm, start, B, d, inv_batch, gU, gV, loss, hot, st);
}

// ===========================================================================
// Deterministic mode
// ===========================================================================
// (a) B <= kSmallB: ONE CTA does the whole batch.  Indices, g_b and losses sit
//     in shared memory; the owner of a destination row is the first batch
//     entry that names it, and it sums that row's contributions in batch order
//     -- exactly the order of the reference's sequential index_put_.
// (b) larger B: g_b from the MODE 1 forward, three stable radix sorts of
//     (row, b) pairs, and a chunked segmented reduction staged through shared
//     memory with a fix-up pass for rows that straddle chunks.
constexpr int kSmallB = 256;
constexpr int kSmallThreads = 1024;     // one CTA; enough lane groups for a group per batch entry

struct SmallBatchSmem {
  int su[kSmallB], si[kSmallB], sj[kSmallB];
  float sg[kSmallB], sl[kSmallB];
};

// One deterministic forward + backward of a batch of B <= kSmallB triplets by ONE CTA of kSmallThreads threads.
// NC selects the read-only load path for the tables.  Must be called by all threads of the CTA.
// (A persistent one-CTA-per-epoch variant that also ran the optimiser in-kernel was measured and dropped:
// 20.9 us/step against 14.0 us/step for this kernel + K3 as two launches at config 2, profiles/r01_notes.md.)
template <int VEC, int LPT, int NITER, bool NC>
__device__ __forceinline__ void det_small_step(SmallBatchSmem& sm, const float* U, const float* V,
                                               const mfcd_triplet* __restrict__ rec,
                                               const int32_t* __restrict__ perm, int64_t start, int B, int d,
                                               float inv_batch, float* gU, float* gV, float* loss_out) {
  int* su = sm.su; int* si = sm.si; int* sj = sm.sj;
  float* sg = sm.sg; float* sl = sm.sl;
  constexpr int NG = kSmallThreads / LPT;     // groups in the CTA
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPT;
  const int gid = threadIdx.x / LPT;
  const unsigned gmask = group_mask<LPT>(lane);

  for (int b = threadIdx.x; b < B; b += kSmallThreads) {
    const int64_t idx = perm ? (int64_t)perm[start + b] : (start + b);
    const int4 r = __ldg(reinterpret_cast<const int4*>(rec) + idx);
    su[b] = r.x; si[b] = r.y; sj[b] = r.z; sg[b] = __int_as_float(r.w);   // sg holds z until phase 1 ends
  }
  __syncthreads();

  // phase 1: forward, g_b, per-sample loss   (uniform trip count per group)
  const int rounds = (B + NG - 1) / NG;
  for (int rr = 0; rr < rounds; ++rr) {
    const int b = rr * NG + gid;
    const bool ok = b < B;
    TripletRows<VEC, LPT, NITER> rows;
    const int tu = ok ? su[b] : 0, ti = ok ? si[b] : 0, tj = ok ? sj[b] : 0;
    const float z = ok ? sg[b] : 0.f;
    load_rows<VEC, LPT, NITER, NC>(rows, U, V, tu, ti, tj, d, sub, ok);
    const float x = group_sum<LPT>(partial_dot<VEC, LPT, NITER>(rows), gmask);
    const float p = sigmoidf_ref(x);
    __syncwarp(gmask);
    if (ok && sub == 0) {
      sl[b] = bce_ref(p, z);
      sg[b] = bce_grad_score_ref(p, z, inv_batch);
    }
  }
  __syncthreads();

  // loss: fixed-order reduction by warp 0
  if (threadIdx.x < 32) {
    float t = 0.f;
    for (int b = lane; b < B; b += 32) t += sl[b];
    t = warp_sum(t);
    if (lane == 0) *loss_out += t * inv_batch;
  }

  // phases 2 and 3: a lane group owns a destination row iff its entry is the first of the batch to name it,
  // and then sums that row's contributions in batch order.  The scans over the batch are done LPT entries
  // at a time: every lane of the group tests one entry, a ballot turns the tests into a bit mask, and only
  // the set bits (the actual matches, ascending) are visited.
  constexpr unsigned GBITS = (LPT == 32) ? 0xffffffffu : ((1u << LPT) - 1u);
  const int gshift = (lane / LPT) * LPT;
  const int entry_rounds = (B + NG - 1) / NG;

  // ---- phase 2: gU rows -------------------------------------------------------------------------------
  for (int er = 0; er < entry_rounds; ++er) {
    const int b = er * NG + gid;
    const bool valid = b < B;
    const int row = valid ? su[b] : -1;
    bool owner = valid;
    Frag<VEC> acc[NITER];
#pragma unroll
    for (int it = 0; it < NITER; ++it) acc[it] = frag_zero<VEC>();
    for (int t0 = 0; t0 < B; t0 += LPT) {
      const int t = t0 + sub;
      const bool hit = valid && t < B && su[t] == row;
      unsigned m = (__ballot_sync(0xffffffffu, hit) >> gshift) & GBITS;
      if (!valid) continue;
      // bits for entries before b: somebody earlier names the row -> not the owner
      const int rel = b - t0;                                  // entry b sits at bit `rel` of this chunk (if 0 <= rel < LPT)
      const unsigned before = rel <= 0 ? 0u : (rel >= LPT ? GBITS : ((1u << rel) - 1u));
      if (m & before) owner = false;
      m &= ~before;
      if (!owner) continue;
      while (m) {                                              // matches at or after b, ascending = batch order
        const int tt = t0 + __ffs(m) - 1;
        m &= m - 1;
        const float g = sg[tt];
        const float* pi = V + (int64_t)si[tt] * d;
        const float* pj = V + (int64_t)sj[tt] * d;
#pragma unroll
        for (int it = 0; it < NITER; ++it) {
          const int c = (it * LPT + sub) * VEC;
          if (c < d) {
            Frag<VEC> a = rd_frag<VEC, NC>(pi + c);
            Frag<VEC> bb = rd_frag<VEC, NC>(pj + c);
#pragma unroll
            for (int kk = 0; kk < VEC; ++kk) acc[it].v[kk] += g * (a.v[kk] - bb.v[kk]);
          }
        }
      }
    }
    if (owner) {
#pragma unroll
      for (int it = 0; it < NITER; ++it) {
        const int c = (it * LPT + sub) * VEC;
        if (c < d) {
          float* dst = gU + (int64_t)row * d + c;
          Frag<VEC> cur = ld_frag<VEC>(dst);
#pragma unroll
          for (int kk = 0; kk < VEC; ++kk) cur.v[kk] += acc[it].v[kk];
          st_frag<VEC>(dst, cur);
        }
      }
    }
  }

  // ---- phase 3: gV rows.  A row can be named as i or as j; its owner is the first entry naming it either
  // way.  sum_i and sum_j are formed separately in batch order and then added (autograd adds the two
  // index_put_ results).
  const int task_rounds = (2 * B + NG - 1) / NG;          // task = (entry, side): both sides run side by side
  for (int er = 0; er < task_rounds; ++er) {
    const int task = er * NG + gid;
    const int b = task >> 1;
    {
      const int side = task & 1;
      const bool valid = b < B && !(side == 1 && sj[b] == si[b]);
      const int row = valid ? (side == 0 ? si[b] : sj[b]) : -1;
      bool owner = valid;
      Frag<VEC> acc_i[NITER], acc_j[NITER];
#pragma unroll
      for (int it = 0; it < NITER; ++it) { acc_i[it] = frag_zero<VEC>(); acc_j[it] = frag_zero<VEC>(); }
      for (int t0 = 0; t0 < B; t0 += LPT) {
        const int t = t0 + sub;
        const bool in = valid && t < B;
        unsigned mi = (__ballot_sync(0xffffffffu, in && si[t] == row) >> gshift) & GBITS;
        unsigned mj = (__ballot_sync(0xffffffffu, in && sj[t] == row) >> gshift) & GBITS;
        if (!valid) continue;
        const int rel = b - t0;
        const unsigned before = rel <= 0 ? 0u : (rel >= LPT ? GBITS : ((1u << rel) - 1u));
        if ((mi | mj) & before) owner = false;
        // side 1 (row = sj[b]): entry b itself names the row only as j; as i it names a different row
        mi &= ~before;
        mj &= ~before;
        if (!owner) continue;
        unsigned m = mi | mj;
        while (m) {
          const int bit = __ffs(m) - 1;
          m &= m - 1;
          const int tt = t0 + bit;
          const float g = sg[tt];
          const float* pu = U + (int64_t)su[tt] * d;
          const bool hi = (mi >> bit) & 1u, hj = (mj >> bit) & 1u;
#pragma unroll
          for (int it = 0; it < NITER; ++it) {
            const int c = (it * LPT + sub) * VEC;
            if (c < d) {
              Frag<VEC> uu = rd_frag<VEC, NC>(pu + c);
#pragma unroll
              for (int kk = 0; kk < VEC; ++kk) {
                const float gu = g * uu.v[kk];
                if (hi) acc_i[it].v[kk] += gu;
                if (hj) acc_j[it].v[kk] -= gu;
              }
            }
          }
        }
      }
      if (owner) {
#pragma unroll
        for (int it = 0; it < NITER; ++it) {
          const int c = (it * LPT + sub) * VEC;
          if (c < d) {
            float* dst = gV + (int64_t)row * d + c;
            Frag<VEC> cur = ld_frag<VEC>(dst);
#pragma unroll
            for (int kk = 0; kk < VEC; ++kk) cur.v[kk] += acc_i[it].v[kk] + acc_j[it].v[kk];
            st_frag<VEC>(dst, cur);
          }
        }
      }
    }
  }
}

template <int VEC, int LPT, int NITER>
__global__ void __launch_bounds__(kSmallThreads)
k_det_small(const float* __restrict__ U, const float* __restrict__ V, const mfcd_triplet* __restrict__ rec,
            const int32_t* __restrict__ perm, int64_t start, int B, int d, float inv_batch,
            float* __restrict__ gU, float* __restrict__ gV, float* __restrict__ loss_out) {
  __shared__ SmallBatchSmem sm;
  det_small_step<VEC, LPT, NITER, true>(sm, U, V, rec, perm, start, B, d, inv_batch, gU, gV, loss_out);
}

template <int VEC, int LPT, int NITER>
struct DetSmallLauncher {
  static int run(const float* U, const float* V, const mfcd_triplet* rec, const int32_t* perm, int64_t start,
                 int B, int d, float inv_batch, float* gU, float* gV, float* loss, cudaStream_t st) {
    k_det_small<VEC, LPT, NITER><<<1, kSmallThreads, 0, st>>>(U, V, rec, perm, start, B, d, inv_batch, gU, gV, loss);
    MFCD_CHECK_LAUNCH();
    return MFCD_OK;
  }
};

static int launch_det_small(const float* U, const float* V, const mfcd_triplet* rec, const int32_t* perm,
                            int64_t start, int B, int d, float inv_batch, float* gU, float* gV, float* loss,
                            cudaStream_t st) {
  // at least 8 lanes per group: the groups also scan the batch, 8 entries per ballot (narrow rows idle the rest)
  MFCD_DISPATCH_ROW_SHAPE_MIN(DetSmallLauncher, d, 8, U, V, rec, perm, start, B, d, inv_batch, gU, gV, loss, st);
}

}  // namespace mfcd

// ---------------------------------------------------------------------------
// large-batch deterministic path (sort + segmented reduction): segmented.cu
// ---------------------------------------------------------------------------
namespace mfcd {
int launch_det_large(const float* U, const float* V, const mfcd_triplet* rec, const int32_t* perm, int64_t start,
                     int64_t B, int d, float inv_batch, int64_t n_users, int64_t n_items, float* gU, float* gV,
                     float* loss, void* ws, size_t ws_bytes, cudaStream_t st);
size_t det_large_workspace_bytes(int64_t B, int d);

// MODE 1 forward used by the large path: g_b -> gbuf, per-block loss partials
template <int VEC, int LPT, int NITER>
struct DetForwardLauncher {
  static int run(const float* U, const float* V, const mfcd_triplet* rec, const int32_t* perm, int64_t start,
                 int64_t B, int d, float inv_batch, float* gbuf, float* partials, int grid, cudaStream_t st) {
    k_fwd_bwd<VEC, LPT, NITER, 1, false><<<grid, kBlock, 0, st>>>(U, V, rec, perm, start, B, d, inv_batch, nullptr,
                                                                   nullptr, gbuf, partials, HotRows{nullptr, nullptr, 0});
    MFCD_CHECK_LAUNCH();
    return MFCD_OK;
  }
};

int launch_det_forward(const float* U, const float* V, const mfcd_triplet* rec, const int32_t* perm, int64_t start,
                       int64_t B, int d, float inv_batch, float* gbuf, float* partials, int grid, cudaStream_t st) {
  MFCD_DISPATCH_ROW_SHAPE(DetForwardLauncher, d, U, V, rec, perm, start, B, d, inv_batch, gbuf, partials, grid, st);
}

size_t det_workspace_bytes(int64_t B, int d) {
  if (B <= kSmallB) return 0;
  return det_large_workspace_bytes(B, d);
}

// Engine of the large-batch deterministic mode.
//   "fixed" (k1_fixed.cu): integer atomics into a 64-bit fixed-point image of the gradient tables; 2 launches, no
//           library call, and the result does not even depend on the order of the batch.  64-bit atomics have a
//           quarter of the vector-fp32 reduction rate and serialise on hot rows, so it wins where the step is
//           launch-bound and loses on big batches (measured at 2^20 triplets, config-4 tables: 0.98 ms against
//           0.62 ms uniform, 19 ms against 1.7 ms under zipf(1.5) items; profiles/r02_notes.md);
//   "sort"  (segmented.cu): three stable radix sorts by destination row (cub::DeviceRadixSort) + a chunked segmented
//           reduction that keeps batch order inside a row, like the reference's sequential index_put_; 17 launches.
// Default: fixed up to kFixedMaxB triplets per batch, sort above (tools/det_sweep.py, profiles/r02_det_engines_us.json:
// uniform items cross over near 5e5 triplets, zipf(1.5) items near 1e4); MFCD_DET_ENGINE=fixed|sort forces one.
constexpr int64_t kFixedMaxB = 16384;
static int det_engine_env() {          // 0 = auto, 1 = fixed, 2 = sort
  static const int v = !getenv("MFCD_DET_ENGINE") ? 0 : (strcmp(getenv("MFCD_DET_ENGINE"), "fixed") == 0 ? 1 :
                                                         (strcmp(getenv("MFCD_DET_ENGINE"), "sort") == 0 ? 2 : 0));
  return v;
}
static bool det_use_fixed(int64_t B) {
  const int e = det_engine_env();
  return e == 1 || (e == 0 && B <= kFixedMaxB);
}

size_t det_workspace_bytes_nm(int64_t B, int d, int64_t n_users, int64_t n_items) {
  if (B <= kSmallB) return 0;
  return det_use_fixed(B) ? det_fixed_workspace_bytes(n_users, n_items, d) : det_large_workspace_bytes(B, d);
}

int launch_fwd_bwd_det(const float* U, const float* V, const mfcd_triplet* rec, const int32_t* perm, int64_t start,
                       int64_t B, int d, float inv_batch, int64_t n_users, int64_t n_items, float* gU, float* gV,
                       float* loss, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (B == 0) return MFCD_OK;
  K1Timer timer(st);
  if (B <= kSmallB) return launch_det_small(U, V, rec, perm, start, (int)B, d, inv_batch, gU, gV, loss, st);
  const size_t need_fix = det_fixed_workspace_bytes(n_users, n_items, d);
  const size_t need_sort = det_large_workspace_bytes(B, d);
  const bool fits_fix = ws != nullptr && ws_bytes >= need_fix, fits_sort = ws != nullptr && ws_bytes >= need_sort;
  // the preferred engine when its workspace is there, else the other one (a caller that sized the workspace with
  // mfcd_det_workspace_bytes gets the sort engine as before)
  const bool fixed = det_use_fixed(B) ? (fits_fix || !fits_sort) : !fits_sort && fits_fix;
  if (fixed && fits_fix)
    return launch_det_fixed(U, V, rec, perm, start, B, d, inv_batch, n_users, n_items, gU, gV, loss, ws, st);
  if (!fits_sort) {
    set_error("deterministic mode: workspace too small (%zu bytes; the fixed-point engine wants %zu, the sort "
              "engine %zu; see mfcd_det_workspace_bytes_nm)", ws_bytes, need_fix, need_sort);
    return MFCD_ERR_WORKSPACE;
  }
  return launch_det_large(U, V, rec, perm, start, B, d, inv_batch, n_users, n_items, gU, gV, loss, ws, ws_bytes, st);
}
}  // namespace mfcd

using namespace mfcd;

static int check_common(const char* fn, const float* U, const float* V, const mfcd_triplet* rec, int64_t start,
                        int64_t B, int d, float* gU, float* gV, float* loss) {
  MFCD_REQUIRE(B >= 0 && start >= 0, "%s: negative size", fn);
  MFCD_REQUIRE(d >= 1, "%s: d must be >= 1", fn);
  if (B == 0) return MFCD_OK;
  MFCD_REQUIRE(U && V && rec && gU && gV && loss, "%s: NULL pointer", fn);
  MFCD_REQUIRE(B < (int64_t(1) << 31), "%s: batch too large", fn);
  return MFCD_OK;
}

extern "C" int mfcd_triplet_fwd_bwd(const float* U, const float* V, const mfcd_triplet* rec, const int32_t* perm,
                                    int64_t start, int64_t B, int32_t d, float inv_batch, float* gU, float* gV,
                                    float* loss, void* stream) {
  int rc = check_common("mfcd_triplet_fwd_bwd", U, V, rec, start, B, d, gU, gV, loss);
  if (rc != MFCD_OK) return rc;
  return launch_fwd_bwd_atomic(U, V, rec, perm, start, B, d, inv_batch, gU, gV, loss, nullptr, nullptr, 0, 0,
                               as_stream(stream));
}

extern "C" int mfcd_max_hot_items(int32_t d, int32_t* out) {
  MFCD_REQUIRE(out != nullptr && d >= 1, "mfcd_max_hot_items: bad argument");
  *out = max_hot_rows(d);
  return MFCD_OK;
}

extern "C" int mfcd_triplet_fwd_bwd_hot(const float* U, const float* V, const mfcd_triplet* rec,
                                        const int32_t* perm, int64_t start, int64_t B, int32_t d, float inv_batch,
                                        float* gU, float* gV, float* loss, const int8_t* item_slot,
                                        const int32_t* hot_items, int32_t n_hot, void* stream) {
  int rc = check_common("mfcd_triplet_fwd_bwd_hot", U, V, rec, start, B, d, gU, gV, loss);
  if (rc != MFCD_OK) return rc;
  MFCD_REQUIRE(n_hot >= 0, "mfcd_triplet_fwd_bwd_hot: n_hot < 0");
  return launch_fwd_bwd_atomic(U, V, rec, perm, start, B, d, inv_batch, gU, gV, loss, item_slot, hot_items, n_hot, 0,
                               as_stream(stream));
}

extern "C" int mfcd_triplet_fwd_bwd_ex(const float* U, const float* V, const mfcd_triplet* rec, const int32_t* perm,
                                       int64_t start, int64_t B, int32_t d, float inv_batch, float* gU, float* gV,
                                       float* loss, const int8_t* item_slot, const int32_t* hot_items,
                                       int32_t n_hot, int32_t flags, void* stream) {
  int rc = check_common("mfcd_triplet_fwd_bwd_ex", U, V, rec, start, B, d, gU, gV, loss);
  if (rc != MFCD_OK) return rc;
  MFCD_REQUIRE(n_hot >= 0, "mfcd_triplet_fwd_bwd_ex: n_hot < 0");
  return launch_fwd_bwd_atomic(U, V, rec, perm, start, B, d, inv_batch, gU, gV, loss, item_slot, hot_items, n_hot,
                               flags, as_stream(stream));
}

extern "C" int mfcd_det_workspace_bytes(int64_t B, int32_t d, size_t* bytes) {
  MFCD_REQUIRE(bytes != nullptr && B >= 0 && d >= 1, "mfcd_det_workspace_bytes: bad argument");
  *bytes = det_workspace_bytes(B, d);
  return MFCD_OK;
}

extern "C" int mfcd_det_workspace_bytes_nm(int64_t B, int32_t d, int64_t n_users, int64_t n_items, size_t* bytes) {
  MFCD_REQUIRE(bytes != nullptr && B >= 0 && d >= 1 && n_users >= 1 && n_items >= 1,
               "mfcd_det_workspace_bytes_nm: bad argument");
  *bytes = det_workspace_bytes_nm(B, d, n_users, n_items);
  return MFCD_OK;
}

extern "C" int mfcd_triplet_fwd_bwd_det(const float* U, const float* V, const mfcd_triplet* rec,
                                        const int32_t* perm, int64_t start, int64_t B, int32_t d, float inv_batch,
                                        int64_t n_users, int64_t n_items, float* gU, float* gV, float* loss,
                                        void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_common("mfcd_triplet_fwd_bwd_det", U, V, rec, start, B, d, gU, gV, loss);
  if (rc != MFCD_OK) return rc;
  MFCD_REQUIRE(n_users > 0 && n_items > 0, "mfcd_triplet_fwd_bwd_det: table sizes must be positive");
  return launch_fwd_bwd_det(U, V, rec, perm, start, B, d, inv_batch, n_users, n_items, gU, gV, loss, workspace,
                            workspace_bytes, as_stream(stream));
}

extern "C" int mfcd_triplet_fwd_bwd_det_fixed(const float* U, const float* V, const mfcd_triplet* rec,
                                              const int32_t* perm, int64_t start, int64_t B, int32_t d,
                                              float inv_batch, int64_t n_users, int64_t n_items, float* gU, float* gV,
                                              float* loss, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_common("mfcd_triplet_fwd_bwd_det_fixed", U, V, rec, start, B, d, gU, gV, loss);
  if (rc != MFCD_OK) return rc;
  MFCD_REQUIRE(n_users > 0 && n_items > 0, "mfcd_triplet_fwd_bwd_det_fixed: table sizes must be positive");
  if (B == 0) return MFCD_OK;
  const size_t need = det_fixed_workspace_bytes(n_users, n_items, d);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("mfcd_triplet_fwd_bwd_det_fixed: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
    return MFCD_ERR_WORKSPACE;
  }
  return launch_det_fixed(U, V, rec, perm, start, B, d, inv_batch, n_users, n_items, gU, gV, loss, workspace,
                          as_stream(stream));
}

extern "C" int mfcd_det_fixed_workspace_bytes(int32_t d, int64_t n_users, int64_t n_items, size_t* bytes) {
  MFCD_REQUIRE(bytes != nullptr && d >= 1 && n_users >= 1 && n_items >= 1, "mfcd_det_fixed_workspace_bytes: bad argument");
  *bytes = det_fixed_workspace_bytes(n_users, n_items, d);
  return MFCD_OK;
}

extern "C" int mfcd_triplet_fwd_bwd_det_sort(const float* U, const float* V, const mfcd_triplet* rec,
                                             const int32_t* perm, int64_t start, int64_t B, int32_t d,
                                             float inv_batch, int64_t n_users, int64_t n_items, float* gU, float* gV,
                                             float* loss, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_common("mfcd_triplet_fwd_bwd_det_sort", U, V, rec, start, B, d, gU, gV, loss);
  if (rc != MFCD_OK) return rc;
  MFCD_REQUIRE(n_users > 0 && n_items > 0, "mfcd_triplet_fwd_bwd_det_sort: table sizes must be positive");
  if (B == 0) return MFCD_OK;
  if (B <= kSmallB)
    return launch_det_small(U, V, rec, perm, start, (int)B, d, inv_batch, gU, gV, loss, as_stream(stream));
  if (workspace == nullptr || workspace_bytes < det_large_workspace_bytes(B, d)) {
    set_error("mfcd_triplet_fwd_bwd_det_sort: workspace too small (%zu < %zu bytes)", workspace_bytes,
              det_large_workspace_bytes(B, d));
    return MFCD_ERR_WORKSPACE;
  }
  return launch_det_large(U, V, rec, perm, start, B, d, inv_batch, n_users, n_items, gU, gV, loss, workspace,
                          workspace_bytes, as_stream(stream));
}

// one epoch of structure.py:845-852 launched from C (no per-step python)
extern "C" int mfcd_train_epoch(const mfcd_epoch_args* a) {
  MFCD_REQUIRE(a != nullptr, "mfcd_train_epoch: args is NULL");
  MFCD_REQUIRE(a->batch_size > 0 && a->n_samples >= 0 && a->d >= 1, "mfcd_train_epoch: bad sizes");
  MFCD_REQUIRE(a->params && a->grads && a->rec && a->step_losses, "mfcd_train_epoch: NULL pointer");
  MFCD_REQUIRE(a->optimizer == MFCD_OPT_ADAM || a->optimizer == MFCD_OPT_SGD, "mfcd_train_epoch: bad optimizer");
  MFCD_REQUIRE(a->optimizer != MFCD_OPT_ADAM || (a->state1 && a->state2), "mfcd_train_epoch: Adam state is NULL");
  cudaStream_t st = as_stream(a->stream);
  const int64_t numel = (a->n_users + a->n_items) * a->d;
  float* U = a->params;
  float* V = a->params + a->n_users * a->d;
  float* gU = a->grads;
  float* gV = a->grads + a->n_users * a->d;
  const int64_t n_steps = (a->n_samples + a->batch_size - 1) / a->batch_size;
  if (n_steps == 0) return MFCD_OK;
  {   // reference regime (small batch, deterministic, Adam, small tables): one persistent kernel for the epoch
    const int rc = launch_epoch_small(a, st);
    if (rc != MFCD_ERR_UNSUPPORTED) return rc;
  }
  MFCD_CUDA(cudaMemsetAsync(a->step_losses, 0, sizeof(float) * n_steps, st));
  for (int64_t k = 0; k < n_steps; ++k) {
    const int64_t start = k * a->batch_size;
    const int64_t B = (a->n_samples - start) < a->batch_size ? (a->n_samples - start) : a->batch_size;
    const float inv_b = 1.0f / (float)B;
    int rc;
    if (a->mode == MFCD_MODE_DETERMINISTIC)
      rc = launch_fwd_bwd_det(U, V, a->rec, a->perm, start, B, a->d, inv_b, a->n_users, a->n_items, gU, gV,
                              a->step_losses + k, a->workspace, a->workspace_bytes, st);
    else
      rc = launch_fwd_bwd_atomic(U, V, a->rec, a->perm, start, B, a->d, inv_b, gU, gV, a->step_losses + k,
                                 a->item_slot, a->hot_items, a->n_hot, a->flags, st);
    if (rc != MFCD_OK) return rc;
    const int64_t step = a->step0 + k + 1;
    if (a->optimizer == MFCD_OPT_ADAM)
      rc = launch_adam(a->params, a->grads, a->state1, a->state2, numel, a->lr, a->beta1, a->beta2, a->eps,
                       a->weight_decay, step, 1, st);
    else
      rc = launch_sgd(a->params, a->grads, a->state1, numel, a->lr, a->momentum, a->weight_decay, step, 1, st);
    if (rc != MFCD_OK) return rc;
  }
  return MFCD_OK;
}
