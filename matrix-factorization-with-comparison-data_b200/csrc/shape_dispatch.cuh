// Row-shape dispatch shared by every kernel that walks embedding rows.
//
// A row of d floats is covered by LPT lanes (a lane-aligned sub-warp group,
// LPT a power of two), each lane moving VEC floats per access (VEC = 4, 2 or 1
// depending on the alignment d allows), NITER accesses per lane:
//     d = 64  -> VEC 4, LPT 16, NITER 1   (two triplets per warp side by side)
//     d = 10  -> VEC 2, LPT 8 (5 active), NITER 1
//     d = 2   -> VEC 2, LPT 1
//     d = 128 -> VEC 4, LPT 32, NITER 1 ;  d = 256 -> NITER 2
#pragma once
#include "common.cuh"

namespace mfcd {

struct RowShape {
  int vec, lpt, niter;
};

static inline bool row_shape_for(int d, RowShape* s) {
  if (d <= 0) return false;
  int vec = (d % 4 == 0) ? 4 : (d % 2 == 0) ? 2 : 1;
  int chunks = d / vec;
  int lpt = 1;
  while (lpt < chunks && lpt < 32) lpt <<= 1;
  int niter = (chunks + lpt - 1) / lpt;
  if (niter > 4) return false;
  s->vec = vec; s->lpt = lpt; s->niter = niter;
  return true;
}

// L is a class template L<VEC, LPT, NITER> with a static `int run(Args...)`.
#define MFCD_DISPATCH_LPT(L, VEC, ...)                                      \
  switch (shape.lpt) {                                                      \
    case 1:  return L<VEC, 1, 1>::run(__VA_ARGS__);                         \
    case 2:  return L<VEC, 2, 1>::run(__VA_ARGS__);                         \
    case 4:  return L<VEC, 4, 1>::run(__VA_ARGS__);                         \
    case 8:  return L<VEC, 8, 1>::run(__VA_ARGS__);                         \
    case 16: return L<VEC, 16, 1>::run(__VA_ARGS__);                        \
    default:                                                                \
      switch (shape.niter) {                                                \
        case 1:  return L<VEC, 32, 1>::run(__VA_ARGS__);                    \
        case 2:  return L<VEC, 32, 2>::run(__VA_ARGS__);                    \
        case 3:  return L<VEC, 32, 3>::run(__VA_ARGS__);                    \
        default: return L<VEC, 32, 4>::run(__VA_ARGS__);                    \
      }                                                                     \
  }

// MINLPT widens the lane group beyond what the row needs (extra lanes hold zeros): the small-batch
// deterministic kernel uses its groups for batch scans too and wants at least 8 lanes per group.
#define MFCD_DISPATCH_ROW_SHAPE(L, d, ...) MFCD_DISPATCH_ROW_SHAPE_MIN(L, d, 1, __VA_ARGS__)

#define MFCD_DISPATCH_ROW_SHAPE_MIN(L, d, MINLPT, ...)                      \
  do {                                                                      \
    ::mfcd::RowShape shape;                                                 \
    if (!::mfcd::row_shape_for((d), &shape)) {                              \
      ::mfcd::set_error("unsupported embedding width d=%d (need d <= 512 for d%%4==0, <= 256 for even d, <= 128 otherwise)", (int)(d)); \
      return MFCD_ERR_UNSUPPORTED;                                          \
    }                                                                       \
    if (shape.lpt < (MINLPT)) shape.lpt = (MINLPT);                         \
    if (shape.vec == 4) { MFCD_DISPATCH_LPT(L, 4, __VA_ARGS__) }            \
    else if (shape.vec == 2) { MFCD_DISPATCH_LPT(L, 2, __VA_ARGS__) }       \
    else { MFCD_DISPATCH_LPT(L, 1, __VA_ARGS__) }                           \
  } while (0)

// Per-lane register image of one triplet's rows: u-row fragment(s) and the
// fragment(s) of V[i]-V[j].
template <int VEC, int LPT, int NITER>
struct TripletRows {
  Frag<VEC> uu[NITER];
  Frag<VEC> dv[NITER];
};

// NC = true: read-only (ld.global.nc) path, for kernels that never write the tables;
// NC = false: coherent loads, for the persistent epoch kernel that updates the tables between steps.
template <int VEC, bool NC>
__device__ __forceinline__ Frag<VEC> rd_frag(const float* p) {
  if constexpr (NC) return ldg_frag<VEC>(p);
  else return ld_frag<VEC>(p);
}

template <int VEC, int LPT, int NITER, bool NC = true>
__device__ __forceinline__ void load_rows(TripletRows<VEC, LPT, NITER>& t, const float* __restrict__ U,
                                          const float* __restrict__ V, int tu, int ti, int tj, int d,
                                          int sub, bool ok) {
  const float* pu = U + static_cast<int64_t>(tu) * d;
  const float* pi = V + static_cast<int64_t>(ti) * d;
  const float* pj = V + static_cast<int64_t>(tj) * d;
#pragma unroll
  for (int it = 0; it < NITER; ++it) {
    const int c = (it * LPT + sub) * VEC;
    if (ok && c < d) {
      t.uu[it] = rd_frag<VEC, NC>(pu + c);
      Frag<VEC> a = rd_frag<VEC, NC>(pi + c);
      Frag<VEC> b = rd_frag<VEC, NC>(pj + c);
#pragma unroll
      for (int k = 0; k < VEC; ++k) t.dv[it].v[k] = a.v[k] - b.v[k];
    } else {
      t.uu[it] = frag_zero<VEC>();
      t.dv[it] = frag_zero<VEC>();
    }
  }
}

// lane-partial of <U_u, V_i - V_j>
template <int VEC, int LPT, int NITER>
__device__ __forceinline__ float partial_dot(const TripletRows<VEC, LPT, NITER>& t) {
  float acc = 0.f;
#pragma unroll
  for (int it = 0; it < NITER; ++it)
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc = fmaf(t.uu[it].v[k], t.dv[it].v[k], acc);
  return acc;
}

// Butterfly shuffles with offsets < LPT never leave a lane-aligned group, and every
// caller runs them warp-convergently, so the constant full mask is correct -- and it
// spares the MATCH.ANY / REDUX / VOTE mask validation ptxas emits for a variable mask.
template <int LPT>
__device__ __forceinline__ unsigned group_mask(int lane) {
  (void)lane;
  return 0xffffffffu;
}

}  // namespace mfcd
