// Persistent small-batch epoch: ONE cooperative kernel runs every optimiser step of train_model's inner
// loop (structure.py:845-852) for the reference's own regime -- batch 64, tables of a few thousand rows,
// thousands of steps per epoch, where two launches per step (K1 deterministic + K3) cost 12.3 us/step and
// the GPU is idle most of that time.
//
// G CTAs of 1024 threads (one per SM, co-resident: cooperative launch).  CTA c OWNS the rows
// [c*R, (c+1)*R) of the flat table [U | V]: their Adam moments and parameters live in its registers for
// the whole epoch.  Per step every CTA
//   1. runs the batch's forward redundantly (64 triplets: latency, not throughput, is what matters) from
//      the parameter buffer p[k & 1] through L2 (ld.global.cg: other CTAs wrote it);
//   2. forms the gradient rows that fall into ITS slice, summed in batch order exactly like the
//      reference's sequential index_put_ (same owner-scans as k_det_small), in shared memory;
//   3. applies Adam to all of its elements (dense: untouched rows decay too) and writes them to the other
//      parameter buffer p[(k + 1) & 1] -- ping-pong, so nobody reads what is being written;
//   4. meets the others at one grid barrier.
// The next batch's records are fetched while the current step computes.  Bias corrections come from a
// per-step table computed in double precision (torch's python floats).  No gradient buffer is touched.
#include "internal.h"
#include "shape_dispatch.cuh"

namespace mfcd {

constexpr int kEpThreads = 1024;
constexpr int kEpMaxB = 256;
constexpr int kEpElems = 4;                       // table elements per thread, kept in registers

struct EpochSmallArgs {
  float* p0; float* p1;                           // step k reads p[k & 1], writes p[(k + 1) & 1]
  float* m; float* v;
  const mfcd_triplet* rec; const int32_t* perm;
  int64_t n_samples; int64_t n_steps; int64_t n_users; int64_t rows_total;
  int batch; int d; int rows_per_cta;
  int stage;                                      // 1: the batch's U rows and V_i - V_j rows are staged in shared memory
  int sgrad_floats;                               // floats reserved for the slice gradient (multiple of 4)
  float* step_losses;
  const float2* bias;                             // [n_steps] (lr / (1 - b1^t), sqrt(1 - b2^t))
  float one_minus_b1, b2, one_minus_b2, eps, wd;
  unsigned int* barrier;                          // zeroed by the launcher
  int cluster;                                    // 1: the grid is ONE thread-block cluster -> hardware cluster barrier
};

template <int VEC>
__device__ __forceinline__ Frag<VEC> ldcg_frag(const float* p) {
  Frag<VEC> f;
  if constexpr (VEC == 4) {
    const float4 t = __ldcg(reinterpret_cast<const float4*>(p));
    f.v[0] = t.x; f.v[1] = t.y; f.v[2] = t.z; f.v[3] = t.w;
  } else if constexpr (VEC == 2) {
    const float2 t = __ldcg(reinterpret_cast<const float2*>(p));
    f.v[0] = t.x; f.v[1] = t.y;
  } else {
    f.v[0] = __ldcg(p);
  }
  return f;
}

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <int VEC, int LPT, int NITER>
__global__ void __launch_bounds__(kEpThreads, 1) k_epoch_small(const EpochSmallArgs a) {
  __shared__ int su[2][kEpMaxB], si[2][kEpMaxB], sj[2][kEpMaxB];
  __shared__ float sz[2][kEpMaxB];
  __shared__ float sg[kEpMaxB], sl[kEpMaxB];
  extern __shared__ __align__(16) float sgrad[];                 // [rows of this CTA][d] | s_dv [batch][d] | s_uu [batch][d]
  float* s_dv = sgrad + a.sgrad_floats;                          // V_i - V_j of every batch entry (stage)
  float* s_uu = s_dv + (size_t)a.batch * a.d;                    // U_u of every batch entry (stage)
  const bool stage = a.stage != 0;
  int* s_cnt = reinterpret_cast<int*>(sgrad + a.sgrad_floats + (stage ? 2 * (size_t)a.batch * a.d : 0));
  int* s_first = s_cnt + a.rows_per_cta;                          // per owned row: #references, first referencing entry

  constexpr int NG = kEpThreads / LPT;                            // lane groups in the CTA
  constexpr unsigned GBITS = (LPT == 32) ? 0xffffffffu : ((1u << LPT) - 1u);
  const int tid = threadIdx.x, lane = tid & 31, sub = lane % LPT, gid = tid / LPT;
  const int gshift = (lane / LPT) * LPT;
  const int d = a.d;
  const int64_t row0 = (int64_t)blockIdx.x * a.rows_per_cta;
  const int64_t row1 = (row0 + a.rows_per_cta < a.rows_total) ? row0 + a.rows_per_cta : a.rows_total;
  const int64_t e0 = row0 * d;
  const int n_el = row1 > row0 ? (int)((row1 - row0) * d) : 0;   // trailing CTAs may own nothing

  // this CTA's parameters and Adam moments: registers for the whole epoch
  float rp[kEpElems], rm[kEpElems], rv[kEpElems];
#pragma unroll
  for (int q = 0; q < kEpElems; ++q) {
    const int e = tid + q * kEpThreads;
    rp[q] = e < n_el ? a.p0[e0 + e] : 0.f;
    rm[q] = e < n_el ? a.m[e0 + e] : 0.f;
    rv[q] = e < n_el ? a.v[e0 + e] : 0.f;
  }
  {   // records of step 0
    const int B0 = a.n_samples < a.batch ? (int)a.n_samples : a.batch;
    if (tid < B0) {
      const int64_t idx = a.perm ? (int64_t)a.perm[tid] : tid;
      const int4 r = __ldg(reinterpret_cast<const int4*>(a.rec) + idx);
      su[0][tid] = r.x; si[0][tid] = r.y; sj[0][tid] = r.z; sz[0][tid] = __int_as_float(r.w);
    }
  }
  __syncthreads();

  for (int64_t k = 0; k < a.n_steps; ++k) {
    const int cb = (int)(k & 1);
    const float* pold = cb ? a.p1 : a.p0;
    float* pnew = cb ? a.p0 : a.p1;
    const float* U = pold;
    const float* V = pold + a.n_users * d;
    const int64_t start = k * a.batch;
    const int B = (a.n_samples - start) < a.batch ? (int)(a.n_samples - start) : a.batch;
    const float inv_batch = 1.0f / (float)B;
    const int* cu = su[cb]; const int* ci = si[cb]; const int* cj = sj[cb];
    const float* cz = sz[cb];

    const float2 bc = __ldg(a.bias + k);                          // needed by the Adam phase: fetch it now
    // next step's records: in flight while this step computes
    int4 nxt = make_int4(0, 0, 0, 0);
    int Bn = 0;
    if (k + 1 < a.n_steps) {
      const int64_t s1 = start + a.batch;
      Bn = (a.n_samples - s1) < a.batch ? (int)(a.n_samples - s1) : a.batch;
      if (tid < Bn) {
        const int64_t idx = a.perm ? (int64_t)a.perm[s1 + tid] : (s1 + tid);
        nxt = __ldg(reinterpret_cast<const int4*>(a.rec) + idx);
      }
    }
    for (int e = tid; e < n_el; e += kEpThreads) sgrad[e] = 0.f;
    for (int r = tid; r < a.rows_per_cta; r += kEpThreads) { s_cnt[r] = 0; s_first[r] = 0x7fffffff; }
    __syncthreads();
    // who touches which of MY rows: most rows are named once per batch and need no scan at all
    for (int t = tid; t < 3 * B; t += kEpThreads) {
      const int b = t / 3, kind = t - 3 * b;
      const int64_t frow = kind == 0 ? (int64_t)cu[b] : a.n_users + (kind == 1 ? ci[b] : cj[b]);
      if (frow >= row0 && frow < row1) {       // an entry with i == j counts twice: its two updates cancel in the scan
        atomicAdd(&s_cnt[frow - row0], 1);
        atomicMin(&s_first[frow - row0], b);
      }
    }

    // ---- forward, g_b, per-sample loss (every CTA, all entries) ----------------------------------------
    const int rounds = (B + NG - 1) / NG;
    for (int rr = 0; rr < rounds; ++rr) {
      const int b = rr * NG + gid;
      const bool ok = b < B;
      const int tu = ok ? cu[b] : 0, ti = ok ? ci[b] : 0, tj = ok ? cj[b] : 0;
      const float z = ok ? cz[b] : 0.f;
      const float* pu = U + (int64_t)tu * d;
      const float* pi = V + (int64_t)ti * d;
      const float* pj = V + (int64_t)tj * d;
      float part = 0.f;
#pragma unroll
      for (int it = 0; it < NITER; ++it) {
        const int c = (it * LPT + sub) * VEC;
        if (ok && c < d) {
          const Frag<VEC> uu = ldcg_frag<VEC>(pu + c);
          const Frag<VEC> fa = ldcg_frag<VEC>(pi + c);
          const Frag<VEC> fb = ldcg_frag<VEC>(pj + c);
          Frag<VEC> dv;
#pragma unroll
          for (int kk = 0; kk < VEC; ++kk) {
            dv.v[kk] = fa.v[kk] - fb.v[kk];
            part = fmaf(uu.v[kk], dv.v[kk], part);
          }
          if (stage) {                                            // the scatter phases read these, not L2
            st_frag<VEC>(s_dv + b * d + c, dv);
            st_frag<VEC>(s_uu + b * d + c, uu);
          }
        }
      }
      const float x = group_sum<LPT>(part, 0xffffffffu);
      const float p = sigmoidf_ref(x);
      if (ok && sub == 0) {
        if (blockIdx.x == 0) sl[b] = bce_ref(p, z);               // only CTA 0 reports the loss
        sg[b] = bce_grad_score_ref(p, z, inv_batch);
      }
    }
    __syncthreads();

    if (blockIdx.x == 0 && tid < 32) {                            // loss: fixed-order reduction by warp 0
      float t = 0.f;
      for (int b = lane; b < B; b += 32) t += sl[b];
      t = warp_sum(t);
      if (lane == 0) a.step_losses[k] = t * inv_batch;
    }

    // ---- gradient rows of THIS CTA's slice, batch-order sums (see k_det_small) ---------------------------
    const int entry_rounds = (B + NG - 1) / NG;
    for (int er = 0; er < entry_rounds; ++er) {                   // gU rows: flat row id = u
      const int b = er * NG + gid;
      const int row = b < B ? cu[b] : -1;
      const bool mine = b < B && row >= row0 && row < row1;
      const int refs = mine ? s_cnt[row - row0] : 0;
      if (refs == 1) {                                            // the only reference: its contribution IS the row
        const float g = sg[b];
#pragma unroll
        for (int it = 0; it < NITER; ++it) {
          const int c = (it * LPT + sub) * VEC;
          if (c < d) {
            Frag<VEC> t;
            if (stage) {
              const Frag<VEC> dv = ld_frag<VEC>(s_dv + b * d + c);
#pragma unroll
              for (int kk = 0; kk < VEC; ++kk) t.v[kk] = g * dv.v[kk];
            } else {
              const Frag<VEC> fa = ldcg_frag<VEC>(V + (int64_t)ci[b] * d + c);
              const Frag<VEC> fb = ldcg_frag<VEC>(V + (int64_t)cj[b] * d + c);
#pragma unroll
              for (int kk = 0; kk < VEC; ++kk) t.v[kk] = g * (fa.v[kk] - fb.v[kk]);
            }
            st_frag<VEC>(sgrad + (int64_t)(row - row0) * d + c, t);
          }
        }
      }
      // rows named several times: the first entry that names one sums its contributions in batch order
      const bool valid = refs > 1 && s_first[row - row0] == b;
      if (!__any_sync(0xffffffffu, valid)) continue;
      bool owner = valid;
      Frag<VEC> acc[NITER];
#pragma unroll
      for (int it = 0; it < NITER; ++it) acc[it] = frag_zero<VEC>();
      for (int t0 = 0; t0 < B; t0 += LPT) {
        const int t = t0 + sub;
        const bool hit = valid && t < B && cu[t] == row;
        unsigned mm = (__ballot_sync(0xffffffffu, hit) >> gshift) & GBITS;
        if (!valid) continue;
        const int rel = b - t0;
        const unsigned before = rel <= 0 ? 0u : (rel >= LPT ? GBITS : ((1u << rel) - 1u));
        if (mm & before) owner = false;
        mm &= ~before;
        if (!owner) continue;
        while (mm) {                                              // matches at or after b, ascending = batch order
          const int tt = t0 + __ffs(mm) - 1;
          mm &= mm - 1;
          const float g = sg[tt];
          const float* pi = V + (int64_t)ci[tt] * d;
          const float* pj = V + (int64_t)cj[tt] * d;
#pragma unroll
          for (int it = 0; it < NITER; ++it) {
            const int c = (it * LPT + sub) * VEC;
            if (c < d) {
              if (stage) {
                const Frag<VEC> dv = ld_frag<VEC>(s_dv + tt * d + c);
#pragma unroll
                for (int kk = 0; kk < VEC; ++kk) acc[it].v[kk] += g * dv.v[kk];
              } else {
                const Frag<VEC> fa = ldcg_frag<VEC>(pi + c);
                const Frag<VEC> fb = ldcg_frag<VEC>(pj + c);
#pragma unroll
                for (int kk = 0; kk < VEC; ++kk) acc[it].v[kk] += g * (fa.v[kk] - fb.v[kk]);
              }
            }
          }
        }
      }
      if (owner) {
#pragma unroll
        for (int it = 0; it < NITER; ++it) {
          const int c = (it * LPT + sub) * VEC;
          if (c < d) st_frag<VEC>(sgrad + (int64_t)(row - row0) * d + c, acc[it]);
        }
      }
    }
    const int task_rounds = (2 * B + NG - 1) / NG;                // gV rows: flat row id = n_users + item
    for (int er = 0; er < task_rounds; ++er) {
      const int task = er * NG + gid;
      const int b = task >> 1;
      const int side = task & 1;
      const int item = b < B ? (side == 0 ? ci[b] : cj[b]) : -1;
      const int64_t frow = a.n_users + item;
      const bool mine = b < B && !(side == 1 && cj[b] == ci[b]) && frow >= row0 && frow < row1;
      const int refs = mine ? s_cnt[frow - row0] : 0;
      if (refs == 1) {
        const float g = side == 0 ? sg[b] : -sg[b];
#pragma unroll
        for (int it = 0; it < NITER; ++it) {
          const int c = (it * LPT + sub) * VEC;
          if (c < d) {
            const Frag<VEC> uu = stage ? ld_frag<VEC>(s_uu + b * d + c) : ldcg_frag<VEC>(U + (int64_t)cu[b] * d + c);
            Frag<VEC> t;
#pragma unroll
            for (int kk = 0; kk < VEC; ++kk) t.v[kk] = g * uu.v[kk];
            st_frag<VEC>(sgrad + (int64_t)(frow - row0) * d + c, t);
          }
        }
      }
      const bool valid = refs > 1 && s_first[frow - row0] == b;
      if (!__any_sync(0xffffffffu, valid)) continue;
      bool owner = valid;
      Frag<VEC> acc_i[NITER], acc_j[NITER];
#pragma unroll
      for (int it = 0; it < NITER; ++it) { acc_i[it] = frag_zero<VEC>(); acc_j[it] = frag_zero<VEC>(); }
      for (int t0 = 0; t0 < B; t0 += LPT) {
        const int t = t0 + sub;
        const bool in = valid && t < B;
        unsigned mi = (__ballot_sync(0xffffffffu, in && ci[t] == item) >> gshift) & GBITS;
        unsigned mj = (__ballot_sync(0xffffffffu, in && cj[t] == item) >> gshift) & GBITS;
        if (!valid) continue;
        const int rel = b - t0;
        const unsigned before = rel <= 0 ? 0u : (rel >= LPT ? GBITS : ((1u << rel) - 1u));
        if ((mi | mj) & before) owner = false;
        mi &= ~before;
        mj &= ~before;
        if (!owner) continue;
        unsigned mm = mi | mj;
        while (mm) {
          const int bit = __ffs(mm) - 1;
          mm &= mm - 1;
          const int tt = t0 + bit;
          const float g = sg[tt];
          const float* pu = U + (int64_t)cu[tt] * d;
          const bool hi = (mi >> bit) & 1u, hj = (mj >> bit) & 1u;
#pragma unroll
          for (int it = 0; it < NITER; ++it) {
            const int c = (it * LPT + sub) * VEC;
            if (c < d) {
              const Frag<VEC> uu = stage ? ld_frag<VEC>(s_uu + tt * d + c) : ldcg_frag<VEC>(pu + c);
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
            Frag<VEC> t;
#pragma unroll
            for (int kk = 0; kk < VEC; ++kk) t.v[kk] = acc_i[it].v[kk] + acc_j[it].v[kk];
            st_frag<VEC>(sgrad + (int64_t)(frow - row0) * d + c, t);
          }
        }
      }
    }
    __syncthreads();

    // ---- Adam on every element of the slice; new parameters go to the other buffer ------------------------
    {
      AdamScalars s;
      s.lr_over_bc1 = bc.x; s.bc2_sqrt = bc.y;
      s.one_minus_b1 = a.one_minus_b1; s.b2 = a.b2; s.one_minus_b2 = a.one_minus_b2; s.eps = a.eps; s.wd = a.wd;
#pragma unroll
      for (int q = 0; q < kEpElems; ++q) {
        const int e = tid + q * kEpThreads;
        if (e < n_el) {
          float g = sgrad[e];
          adam_elem(rp[q], g, rm[q], rv[q], s);
          __stcg(pnew + e0 + e, rp[q]);
        }
      }
    }
    if (tid < Bn) { su[cb ^ 1][tid] = nxt.x; si[cb ^ 1][tid] = nxt.y; sj[cb ^ 1][tid] = nxt.z; sz[cb ^ 1][tid] = __int_as_float(nxt.w); }

    // ---- grid barrier: every slice of p[(k + 1) & 1] is written before anybody reads it --------------------
    if (a.cluster) {
      // the whole grid is one cluster: the hardware barrier (release / acquire at cluster scope) orders the
      // st.global.cg stores above before the other CTAs' ld.global.cg -- ~10x cheaper than a trip through L2
      asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else {
      // (the pattern of cooperative groups' grid sync: the CTA barrier orders every thread's stores before
      // thread 0's fence, whose release is cumulative -- ONE gpu-scope fence per CTA instead of 1024)
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        atomicAdd(a.barrier, 1u);
        const unsigned int want = (unsigned int)gridDim.x * (unsigned int)(k + 1);
        while (ld_acquire_u32(a.barrier) < want) { }
      }
      __syncthreads();
    }
  }

#pragma unroll
  for (int q = 0; q < kEpElems; ++q) {
    const int e = tid + q * kEpThreads;
    if (e < n_el) { a.m[e0 + e] = rm[q]; a.v[e0 + e] = rv[q]; }
  }
}

__global__ void k_bias_table(float2* __restrict__ out, int64_t step0, int64_t n_steps, float lr, float beta1,
                             float beta2) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= n_steps) return;
  const AdamScalars s = adam_scalars(lr, beta1, beta2, 0.f, 0.f, step0 + k + 1);
  out[k] = make_float2(s.lr_over_bc1, s.bc2_sqrt);
}

constexpr int kClusterMax = 8;                    // portable cluster size

struct EpochSmallPlan {
  bool ok;
  int grid, rows_per_cta, stage, sgrad_floats, cluster;
  size_t smem, off_p1, off_bias, off_barrier, total;
  int64_t n_steps, numel;
};

static bool epoch_small_enabled() {
  static const bool on = !(getenv("MFCD_EPOCH_KERNEL") && atoi(getenv("MFCD_EPOCH_KERNEL")) == 0);
  return on;
}

static EpochSmallPlan epoch_small_plan(const mfcd_epoch_args* a) {
  EpochSmallPlan P{};
  P.ok = false;
  if (!epoch_small_enabled()) return P;
  if (a->mode != MFCD_MODE_DETERMINISTIC || a->optimizer != MFCD_OPT_ADAM) return P;
  if (a->batch_size < 1 || a->batch_size > kEpMaxB || a->n_samples < 1 || a->d < 1) return P;
  RowShape shape;
  if (!row_shape_for(a->d, &shape)) return P;
  const int64_t rows = a->n_users + a->n_items;
  P.numel = rows * a->d;
  P.n_steps = (a->n_samples + a->batch_size - 1) / a->batch_size;
  if (P.n_steps >= (int64_t(1) << 31) / 148 || (a->n_users * a->d) % shape.vec != 0) return P;   // barrier counter; V alignment
  const int per_cta_max = (kEpElems * kEpThreads) / a->d;                               // rows a CTA can own
  if (per_cta_max < 1) return P;
  int64_t G = (P.numel + 2047) / 2048;                 // ~2 elements per thread
  if (G < 1) G = 1;
  if (G > sm_count()) G = sm_count();
  int64_t rpc = (rows + G - 1) / G;
  if (rpc > per_cta_max) { rpc = per_cta_max; G = (rows + rpc - 1) / rpc; }
  if (G > sm_count()) return P;
  G = (rows + rpc - 1) / rpc;
  // a grid that fits one thread-block cluster synchronises with the hardware cluster barrier: shrink G to the
  // cluster limit when the per-thread register budget allows it
  static const bool use_cluster = !(getenv("MFCD_EPOCH_CLUSTER") && atoi(getenv("MFCD_EPOCH_CLUSTER")) == 0);
  P.cluster = 0;
  if (use_cluster) {
    if (G > kClusterMax && (rows + kClusterMax - 1) / kClusterMax <= per_cta_max) {
      rpc = (rows + kClusterMax - 1) / kClusterMax;
      G = (rows + rpc - 1) / rpc;
    }
    if (G <= kClusterMax) P.cluster = 1;
  }
  P.grid = (int)G;
  P.rows_per_cta = (int)rpc;
  P.sgrad_floats = (int)(((rpc * a->d) + 3) & ~int64_t(3));
  const size_t stage_bytes = sizeof(float) * 2 * (size_t)a->batch_size * a->d;
  P.stage = (sizeof(float) * P.sgrad_floats + stage_bytes <= 160 * 1024) ? 1 : 0;
  P.smem = sizeof(float) * (size_t)P.sgrad_floats + (P.stage ? stage_bytes : 0) + 2 * sizeof(int) * (size_t)rpc;
  auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
  size_t off = 0;
  P.off_p1 = off; off += up(sizeof(float) * (size_t)P.numel);
  P.off_bias = off; off += up(sizeof(float2) * (size_t)P.n_steps);
  P.off_barrier = off; off += 256;
  P.total = off;
  P.ok = true;
  return P;
}

size_t epoch_small_workspace_bytes(const mfcd_epoch_args* a) {
  const EpochSmallPlan P = epoch_small_plan(a);
  return P.ok ? P.total : 0;
}

template <int VEC, int LPT, int NITER>
struct EpochSmallLauncher {
  static int run(const EpochSmallArgs& ea, const EpochSmallPlan& P, cudaStream_t st) {
    auto kern = k_epoch_small<VEC, LPT, NITER>;
    int per_sm = 0;
    if (P.smem > 32 * 1024)
      MFCD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem));
    MFCD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kEpThreads, P.smem));
    if (per_sm < 1 || (int64_t)per_sm * sm_count() < P.grid) return MFCD_ERR_UNSUPPORTED;   // not co-resident
    EpochSmallArgs args = ea;
    if (P.cluster) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(P.grid); cfg.blockDim = dim3(kEpThreads); cfg.dynamicSmemBytes = P.smem; cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = P.grid; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      int clusters = 0;
      if (cudaOccupancyMaxActiveClusters(&clusters, kern, &cfg) == cudaSuccess && clusters >= 1) {
        args.cluster = 1;
        MFCD_CUDA(cudaLaunchKernelEx(&cfg, kern, args));
        return MFCD_OK;
      }
      (void)cudaGetLastError();                      // the cluster does not fit: grid barrier through L2 instead
    }
    args.cluster = 0;
    void* params[] = {&args};
    MFCD_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(P.grid), dim3(kEpThreads), params, P.smem, st));
    return MFCD_OK;
  }
};

static int dispatch_epoch_small(int d, const EpochSmallArgs& ea, const EpochSmallPlan& P, cudaStream_t st) {
  MFCD_DISPATCH_ROW_SHAPE_MIN(EpochSmallLauncher, d, 8, ea, P, st);
}

int launch_epoch_small(const mfcd_epoch_args* a, cudaStream_t st) {
  const EpochSmallPlan P = epoch_small_plan(a);
  if (!P.ok || a->workspace == nullptr || a->workspace_bytes < P.total) return MFCD_ERR_UNSUPPORTED;
  char* base = static_cast<char*>(a->workspace);
  EpochSmallArgs ea;
  ea.p0 = a->params;
  ea.p1 = reinterpret_cast<float*>(base + P.off_p1);
  ea.m = a->state1; ea.v = a->state2;
  ea.rec = a->rec; ea.perm = a->perm;
  ea.n_samples = a->n_samples; ea.n_steps = P.n_steps; ea.n_users = a->n_users;
  ea.rows_total = a->n_users + a->n_items;
  ea.batch = (int)a->batch_size; ea.d = a->d; ea.rows_per_cta = P.rows_per_cta;
  ea.stage = P.stage; ea.sgrad_floats = P.sgrad_floats; ea.cluster = 0;
  ea.step_losses = a->step_losses;
  float2* bias = reinterpret_cast<float2*>(base + P.off_bias);
  ea.bias = bias;
  ea.one_minus_b1 = (float)(1.0 - (double)a->beta1);
  ea.b2 = a->beta2;
  ea.one_minus_b2 = (float)(1.0 - (double)a->beta2);
  ea.eps = a->eps; ea.wd = a->weight_decay;
  ea.barrier = reinterpret_cast<unsigned int*>(base + P.off_barrier);
  MFCD_CUDA(cudaMemsetAsync(ea.barrier, 0, sizeof(unsigned int), st));
  k_bias_table<<<(unsigned)((P.n_steps + 255) / 256), 256, 0, st>>>(bias, a->step0, P.n_steps, a->lr, a->beta1, a->beta2);
  MFCD_CHECK_LAUNCH();
  const int rc = dispatch_epoch_small(a->d, ea, P, st);
  if (rc != MFCD_OK) return rc;
  if (P.n_steps & 1)                                   // the last step wrote the workspace copy
    MFCD_CUDA(cudaMemcpyAsync(a->params, ea.p1, sizeof(float) * (size_t)P.numel, cudaMemcpyDeviceToDevice, st));
  return MFCD_OK;
}

}  // namespace mfcd

extern "C" int mfcd_train_epoch_workspace(const mfcd_epoch_args* a, size_t* bytes) {
  using namespace mfcd;
  MFCD_REQUIRE(a != nullptr && bytes != nullptr, "mfcd_train_epoch_workspace: NULL pointer");
  size_t need = epoch_small_workspace_bytes(a);
  if (a->mode == MFCD_MODE_DETERMINISTIC && a->batch_size > 0 && a->d >= 1) {
    const int64_t b = a->n_samples < a->batch_size ? a->n_samples : a->batch_size;
    const size_t det = (a->n_users >= 1 && a->n_items >= 1) ? det_workspace_bytes_nm(b, a->d, a->n_users, a->n_items)
                                                            : det_workspace_bytes(b, a->d);
    if (det > need) need = det;
  }
  *bytes = need;
  return MFCD_OK;
}
