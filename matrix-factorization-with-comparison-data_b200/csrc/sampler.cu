// K7 / K8: GPU triplet samplers and the BTL label sampler.
//
// Reference (python, one triplet at a time, hash-set dedup):
//   choose_items_random              generation_data.py:16-26
//   choose_items_by_margin           generation_data.py:46-84
//   choose_items_by_popularity       generation_data.py:103-128
//   choose_items_by_svd_projection   generation_data.py:131-179 (sampling loop :164-174)
//   BTLPreferenceDataset._generate_labels   structure.py:493-519
//
// Here a "round" draws `count` candidates in parallel from a counter-based
// generator (candidate c = Philox(seed, counter0 + c)), applies the strategy's
// own acceptance test, and encodes survivors as 64-bit keys (u*m + i)*m + j.
// mfcd_unique_accept then reproduces the reference's sequential
// "if t not in exclude and t not in triplets" rule over the candidate stream
// with one stable radix sort (cub::DeviceRadixSort, library) + a scan.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include "internal.h"

namespace mfcd {

__device__ __forceinline__ uint64_t make_key(uint32_t u, uint32_t i, uint32_t j, int64_t m) {
  return ((uint64_t)u * (uint64_t)m + i) * (uint64_t)m + j;
}

__device__ __forceinline__ double u01_53(uint32_t hi, uint32_t lo) {
  return (double)(((uint64_t)hi << 21) | (uint64_t)(lo >> 11)) * (1.0 / 9007199254740992.0);
}

__global__ void __launch_bounds__(256)
k_sample_random(int64_t n, int64_t m, int64_t count, uint64_t seed, uint64_t counter0, uint64_t* __restrict__ keys) {
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < count; c += (int64_t)gridDim.x * blockDim.x) {
    const uint4 r = Philox::run(seed, counter0 + (uint64_t)c, 0u);
    const uint32_t u = bounded(r.x, (uint32_t)n), i = bounded(r.y, (uint32_t)m), j = bounded(r.z, (uint32_t)m);
    keys[c] = (i != j) ? make_key(u, i, j, m) : MFCD_KEY_NONE;
  }
}

__global__ void __launch_bounds__(256)
k_sample_margin(int64_t n, int64_t m, int64_t count, uint64_t seed, uint64_t counter0, mfcd_xview X, float margin,
                uint64_t* __restrict__ keys) {
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < count; c += (int64_t)gridDim.x * blockDim.x) {
    const uint4 r = Philox::run(seed, counter0 + (uint64_t)c, 0u);
    const uint32_t u = bounded(r.x, (uint32_t)n), i = bounded(r.y, (uint32_t)m), j = bounded(r.z, (uint32_t)m);
    bool ok = (i != j);
    if (ok) ok = fabsf(xview_at(X, u, i) - xview_at(X, u, j)) <= margin;     // generation_data.py:72-73
    keys[c] = ok ? make_key(u, i, j, m) : MFCD_KEY_NONE;
  }
}

// smallest idx with cdf[idx] > v   (cdf inclusive prefix sums, increasing)
__device__ __forceinline__ uint32_t cdf_search(const double* __restrict__ cdf, int64_t m, double v) {
  int64_t lo = 0, hi = m - 1;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(cdf + mid) > v) hi = mid; else lo = mid + 1;
  }
  return (uint32_t)lo;
}

__global__ void __launch_bounds__(256)
k_sample_popularity(int64_t n, int64_t m, int64_t count, uint64_t seed, uint64_t counter0,
                    const double* __restrict__ cdf, uint64_t* __restrict__ keys) {
  const double total = __ldg(cdf + m - 1);
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < count; c += (int64_t)gridDim.x * blockDim.x) {
    const uint4 r = Philox::run(seed, counter0 + (uint64_t)c, 0u);
    const uint32_t u = bounded(r.x, (uint32_t)n);
    const uint32_t i = cdf_search(cdf, m, u01_53(r.y, r.z) * total);
    // second item ~ probs restricted to != i and renormalised == redraw until different
    uint32_t j = i;
    for (uint32_t attempt = 1; attempt <= 64 && j == i; ++attempt) {
      const uint4 r2 = Philox::run(seed, counter0 + (uint64_t)c, attempt);
      j = cdf_search(cdf, m, u01_53(r2.x, r2.y) * total);
    }
    keys[c] = (i != j) ? make_key(u, i, j, m) : MFCD_KEY_NONE;
  }
}

__global__ void __launch_bounds__(256)
k_sample_block(int64_t m, int64_t count, uint64_t seed, uint64_t counter0, const int32_t* __restrict__ top_users,
               int64_t nu, const int32_t* __restrict__ top_items, int64_t ni, uint64_t* __restrict__ keys) {
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < count; c += (int64_t)gridDim.x * blockDim.x) {
    const uint4 r = Philox::run(seed, counter0 + (uint64_t)c, 0u);
    const uint32_t u = (uint32_t)__ldg(top_users + bounded(r.x, (uint32_t)nu));
    const uint32_t a = bounded(r.y, (uint32_t)ni);
    uint32_t b = bounded(r.z, (uint32_t)(ni - 1));
    if (b >= a) ++b;                                           // ordered pair without replacement
    const uint32_t i = (uint32_t)__ldg(top_items + a), j = (uint32_t)__ldg(top_items + b);
    keys[c] = (i != j) ? make_key(u, i, j, m) : MFCD_KEY_NONE;
  }
}

// proximity (generation_data.py:29-43) and top_k (generation_data.py:189-224): per-user candidate lists
// (the user's top / bottom items by X[u], built once by the caller); i ~ U(list_i[u]), j ~ U(list_j[u]).
// same_list != 0: both items come from one list and must differ (top_k redraws j until it does: drawing
// j uniformly among the other k-1 positions is the same law).
__global__ void __launch_bounds__(256)
k_sample_lists(int64_t n, int64_t m, int64_t count, uint64_t seed, uint64_t counter0,
               const int32_t* __restrict__ list_i, int ki, const int32_t* __restrict__ list_j, int kj, int same_list,
               uint64_t* __restrict__ keys) {
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < count; c += (int64_t)gridDim.x * blockDim.x) {
    const uint4 r = Philox::run(seed, counter0 + (uint64_t)c, 0u);
    const uint32_t u = bounded(r.x, (uint32_t)n);
    const uint32_t a = bounded(r.y, (uint32_t)ki);
    uint32_t b;
    if (same_list) {
      b = bounded(r.z, (uint32_t)(kj - 1));
      if (b >= a) ++b;
    } else {
      b = bounded(r.z, (uint32_t)kj);
    }
    const uint32_t i = (uint32_t)__ldg(list_i + (int64_t)u * ki + a);
    const uint32_t j = (uint32_t)__ldg(list_j + (int64_t)u * kj + b);
    keys[c] = (i != j) ? make_key(u, i, j, m) : MFCD_KEY_NONE;
  }
}

// ---------------------------------------------------------------------------
// sequential-accept semantics through a stable sort
// ---------------------------------------------------------------------------
static inline size_t align_up(size_t x) { return (x + 255) & ~size_t(255); }

struct UniqueLayout {
  size_t k_in, k_out, i_in, i_out, flag, pos, cub_temp, total, cub_bytes;
};

static UniqueLayout unique_layout(int64_t n_seen, int64_t count) {
  UniqueLayout L;
  const int64_t T = n_seen + count;
  size_t off = 0;
  L.k_in = off;  off += align_up(sizeof(uint64_t) * T);
  L.k_out = off; off += align_up(sizeof(uint64_t) * T);
  L.i_in = off;  off += align_up(sizeof(uint32_t) * T);
  L.i_out = off; off += align_up(sizeof(uint32_t) * T);
  L.flag = off;  off += align_up(sizeof(int32_t) * (count + 1));
  L.pos = off;   off += align_up(sizeof(int32_t) * (count + 1));
  size_t a = 0, b = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr,
                                  (uint32_t*)nullptr, T, 0, 64, (cudaStream_t)0);
  cub::DeviceScan::ExclusiveSum(nullptr, b, (const int32_t*)nullptr, (int32_t*)nullptr, count + 1, (cudaStream_t)0);
  L.cub_bytes = a > b ? a : b;
  L.cub_temp = off; off += align_up(L.cub_bytes);
  L.total = off;
  return L;
}

__global__ void k_unique_fill(const uint64_t* __restrict__ seen, int64_t n_seen, const uint64_t* __restrict__ keys,
                              int64_t count, uint64_t* __restrict__ k_in, uint32_t* __restrict__ i_in) {
  const int64_t T = n_seen + count;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < T; t += (int64_t)gridDim.x * blockDim.x) {
    k_in[t] = (t < n_seen) ? seen[t] : keys[t - n_seen];
    i_in[t] = (uint32_t)t;
  }
}

__global__ void k_unique_flag(const uint64_t* __restrict__ k_sorted, const uint32_t* __restrict__ i_sorted,
                              int64_t n_seen, int64_t T, int32_t* __restrict__ flag) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < T; t += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t src = i_sorted[t];
    if (src < (uint32_t)n_seen) continue;
    const uint64_t k = k_sorted[t];
    const bool head = (t == 0) || (k_sorted[t - 1] != k);      // stable sort: a run's head is its earliest entry
    flag[src - n_seen] = (head && k != MFCD_KEY_NONE) ? 1 : 0;
  }
}

__global__ void k_unique_emit(const uint64_t* __restrict__ keys, const int32_t* __restrict__ flag,
                              const int32_t* __restrict__ pos, int64_t count, int64_t want, uint64_t* __restrict__ out,
                              int64_t* __restrict__ n_out) {
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < count; c += (int64_t)gridDim.x * blockDim.x) {
    if (flag[c] && pos[c] < want) out[pos[c]] = keys[c];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const int64_t total = pos[count];                          // exclusive sum over count+1 entries
    *n_out = total < want ? total : want;
  }
}

// ---------------------------------------------------------------------------
// K8 labels
// ---------------------------------------------------------------------------
__device__ __forceinline__ float philox_word_uniform(uint64_t seed, uint64_t w) {
  const uint4 r = Philox::run(seed, w >> 2, 0x4C41424Cu /* 'LABL' stream */);
  const uint32_t sel = (uint32_t)(w & 3u);
  const uint32_t v = sel == 0 ? r.x : (sel == 1 ? r.y : (sel == 2 ? r.z : r.w));
  return u01_24(v);
}

__device__ __forceinline__ void decode_key(uint64_t key, int64_t m, int& u, int& i, int& j) {
  j = (int)(key % (uint64_t)m);
  const uint64_t t = key / (uint64_t)m;
  i = (int)(t % (uint64_t)m);
  u = (int)(t / (uint64_t)m);
}

__global__ void __launch_bounds__(256)
k_btl_labels(mfcd_xview X, const uint64_t* __restrict__ keys, int64_t N, int64_t m, int K, float scale, int soft,
             uint64_t seed, const float* __restrict__ uniforms, mfcd_triplet* __restrict__ out) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < N; t += (int64_t)gridDim.x * blockDim.x) {
    int u, i, j;
    decode_key(keys[t], m, u, i, j);
    const float q = sigmoidf_ref(scale * (xview_at(X, u, i) - xview_at(X, u, j)));   // structure.py:509
    float acc = 0.f;
    for (int k = 0; k < K; ++k) {
      const uint64_t w = (uint64_t)t * (uint64_t)K + (uint64_t)k;
      const float r = uniforms ? uniforms[w] : philox_word_uniform(seed, w);
      const float lab = (r < q) ? 1.f : 0.f;                   // torch.bernoulli: uniform < p
      if (soft) {
        acc += lab;
      } else {
        int4 o; o.x = u; o.y = i; o.z = j; o.w = __float_as_int(lab);
        reinterpret_cast<int4*>(out)[w] = o;                   // K consecutive copies (structure.py:516-518)
      }
    }
    if (soft) {
      int4 o; o.x = u; o.y = i; o.z = j; o.w = __float_as_int(acc / (float)K);   // torch.mean of K draws (:512)
      reinterpret_cast<int4*>(out)[t] = o;
    }
  }
}

__global__ void k_philox_uniforms(uint64_t seed, uint64_t w0, int64_t count, float* __restrict__ out) {
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < count; c += (int64_t)gridDim.x * blockDim.x)
    out[c] = philox_word_uniform(seed, w0 + (uint64_t)c);
}

}  // namespace mfcd

using namespace mfcd;

static int check_nm(const char* fn, int64_t n, int64_t m, int64_t count, const void* keys) {
  MFCD_REQUIRE(n >= 1 && m >= 2 && count >= 0, "%s: need n >= 1, m >= 2, count >= 0", fn);
  MFCD_REQUIRE(n < (int64_t(1) << 31) && m < (int64_t(1) << 31), "%s: n, m must fit int32", fn);
  MFCD_REQUIRE((double)n * (double)m * (double)m < 1.8e19, "%s: n*m*m overflows the 64-bit triplet key", fn);
  MFCD_REQUIRE(count == 0 || keys != nullptr, "%s: keys is NULL", fn);
  return MFCD_OK;
}

extern "C" int mfcd_sample_random(int64_t n, int64_t m, int64_t count, uint64_t seed, uint64_t counter0,
                                  uint64_t* keys, void* stream) {
  int rc = check_nm("mfcd_sample_random", n, m, count, keys);
  if (rc != MFCD_OK || count == 0) return rc;
  k_sample_random<<<grid_for(count, 256, 8), 256, 0, as_stream(stream)>>>(n, m, count, seed, counter0, keys);
  MFCD_CHECK_LAUNCH();
  return MFCD_OK;
}

extern "C" int mfcd_sample_margin(int64_t n, int64_t m, int64_t count, uint64_t seed, uint64_t counter0,
                                  const mfcd_xview* X, float margin, uint64_t* keys, void* stream) {
  int rc = check_nm("mfcd_sample_margin", n, m, count, keys);
  if (rc != MFCD_OK || count == 0) return rc;
  MFCD_REQUIRE(X && (X->X || (X->A && X->B && X->dx >= 1)), "mfcd_sample_margin: bad xview");
  k_sample_margin<<<grid_for(count, 256, 8), 256, 0, as_stream(stream)>>>(n, m, count, seed, counter0, *X, margin,
                                                                          keys);
  MFCD_CHECK_LAUNCH();
  return MFCD_OK;
}

extern "C" int mfcd_sample_popularity(int64_t n, int64_t m, int64_t count, uint64_t seed, uint64_t counter0,
                                      const double* cdf, uint64_t* keys, void* stream) {
  int rc = check_nm("mfcd_sample_popularity", n, m, count, keys);
  if (rc != MFCD_OK || count == 0) return rc;
  MFCD_REQUIRE(cdf != nullptr, "mfcd_sample_popularity: cdf is NULL");
  k_sample_popularity<<<grid_for(count, 256, 8), 256, 0, as_stream(stream)>>>(n, m, count, seed, counter0, cdf, keys);
  MFCD_CHECK_LAUNCH();
  return MFCD_OK;
}

extern "C" int mfcd_sample_block(int64_t m, int64_t count, uint64_t seed, uint64_t counter0,
                                 const int32_t* top_users, int64_t n_top_users, const int32_t* top_items,
                                 int64_t n_top_items, uint64_t* keys, void* stream) {
  MFCD_REQUIRE(m >= 2 && count >= 0, "mfcd_sample_block: bad sizes");
  MFCD_REQUIRE(n_top_users >= 1 && n_top_items >= 2, "mfcd_sample_block: need >= 1 user and >= 2 items");
  if (count == 0) return MFCD_OK;
  MFCD_REQUIRE(top_users && top_items && keys, "mfcd_sample_block: NULL pointer");
  k_sample_block<<<grid_for(count, 256, 8), 256, 0, as_stream(stream)>>>(m, count, seed, counter0, top_users,
                                                                         n_top_users, top_items, n_top_items, keys);
  MFCD_CHECK_LAUNCH();
  return MFCD_OK;
}

extern "C" int mfcd_sample_lists(int64_t n, int64_t m, int64_t count, uint64_t seed, uint64_t counter0,
                                 const int32_t* list_i, int32_t ki, const int32_t* list_j, int32_t kj,
                                 int32_t same_list, uint64_t* keys, void* stream) {
  int rc = check_nm("mfcd_sample_lists", n, m, count, keys);
  if (rc != MFCD_OK || count == 0) return rc;
  MFCD_REQUIRE(list_i && list_j && ki >= 1 && kj >= 1, "mfcd_sample_lists: bad lists");
  MFCD_REQUIRE(!same_list || (ki == kj && ki >= 2), "mfcd_sample_lists: same_list needs one list of >= 2 items");
  k_sample_lists<<<grid_for(count, 256, 8), 256, 0, as_stream(stream)>>>(n, m, count, seed, counter0, list_i, ki,
                                                                         list_j, kj, same_list, keys);
  MFCD_CHECK_LAUNCH();
  return MFCD_OK;
}

extern "C" int mfcd_unique_workspace_bytes(int64_t n_seen, int64_t count, size_t* bytes) {
  MFCD_REQUIRE(bytes && n_seen >= 0 && count >= 0, "mfcd_unique_workspace_bytes: bad argument");
  MFCD_REQUIRE(n_seen + count < (int64_t(1) << 32), "mfcd_unique_workspace_bytes: more than 2^32 keys per round");
  *bytes = unique_layout(n_seen, count).total;
  return MFCD_OK;
}

extern "C" int mfcd_unique_accept(const uint64_t* seen, int64_t n_seen, const uint64_t* keys, int64_t count,
                                  int64_t want, uint64_t* out, int64_t* n_out, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  MFCD_REQUIRE(n_seen >= 0 && count >= 0 && want >= 0, "mfcd_unique_accept: negative size");
  MFCD_REQUIRE(n_out != nullptr, "mfcd_unique_accept: n_out is NULL");
  MFCD_REQUIRE(n_seen + count < (int64_t(1) << 32), "mfcd_unique_accept: more than 2^32 keys per round");
  cudaStream_t st = as_stream(stream);
  if (count == 0) {
    MFCD_CUDA(cudaMemsetAsync(n_out, 0, sizeof(int64_t), st));
    return MFCD_OK;
  }
  MFCD_REQUIRE(keys && out && (n_seen == 0 || seen), "mfcd_unique_accept: NULL pointer");
  const UniqueLayout L = unique_layout(n_seen, count);
  if (workspace == nullptr || workspace_bytes < L.total) {
    set_error("mfcd_unique_accept: workspace too small (%zu < %zu bytes)", workspace_bytes, L.total);
    return MFCD_ERR_WORKSPACE;
  }
  char* base = static_cast<char*>(workspace);
  uint64_t* k_in = reinterpret_cast<uint64_t*>(base + L.k_in);
  uint64_t* k_out = reinterpret_cast<uint64_t*>(base + L.k_out);
  uint32_t* i_in = reinterpret_cast<uint32_t*>(base + L.i_in);
  uint32_t* i_out = reinterpret_cast<uint32_t*>(base + L.i_out);
  int32_t* flag = reinterpret_cast<int32_t*>(base + L.flag);
  int32_t* pos = reinterpret_cast<int32_t*>(base + L.pos);
  const int64_t T = n_seen + count;
  const int grid = grid_for(T, 256, 8);
  k_unique_fill<<<grid, 256, 0, st>>>(seen, n_seen, keys, count, k_in, i_in);
  MFCD_CHECK_LAUNCH();
  size_t cb = L.cub_bytes;
  MFCD_CUDA(cub::DeviceRadixSort::SortPairs(base + L.cub_temp, cb, (const uint64_t*)k_in, k_out,
                                            (const uint32_t*)i_in, i_out, T, 0, 64, st));
  MFCD_CUDA(cudaMemsetAsync(flag, 0, sizeof(int32_t) * (count + 1), st));
  k_unique_flag<<<grid, 256, 0, st>>>(k_out, i_out, n_seen, T, flag);
  MFCD_CHECK_LAUNCH();
  cb = L.cub_bytes;
  MFCD_CUDA(cub::DeviceScan::ExclusiveSum(base + L.cub_temp, cb, (const int32_t*)flag, pos, count + 1, st));
  k_unique_emit<<<grid_for(count, 256, 8), 256, 0, st>>>(keys, flag, pos, count, want, out, n_out);
  MFCD_CHECK_LAUNCH();
  return MFCD_OK;
}

extern "C" int mfcd_btl_labels(const mfcd_xview* X, const uint64_t* keys, int64_t N, int64_t m, int32_t K,
                               float scale, int32_t soft, uint64_t seed, const float* uniforms, mfcd_triplet* out,
                               void* stream) {
  MFCD_REQUIRE(X && (X->X || (X->A && X->B && X->dx >= 1)), "mfcd_btl_labels: bad xview");
  MFCD_REQUIRE(N >= 0 && m >= 2 && K >= 1, "mfcd_btl_labels: bad sizes");
  if (N == 0) return MFCD_OK;
  MFCD_REQUIRE(keys && out, "mfcd_btl_labels: NULL pointer");
  k_btl_labels<<<grid_for(N, 256, 8), 256, 0, as_stream(stream)>>>(*X, keys, N, m, K, scale, soft, seed, uniforms, out);
  MFCD_CHECK_LAUNCH();
  return MFCD_OK;
}

extern "C" int mfcd_philox_uniforms(uint64_t seed, uint64_t w0, int64_t count, float* out, void* stream) {
  MFCD_REQUIRE(count >= 0, "mfcd_philox_uniforms: count < 0");
  if (count == 0) return MFCD_OK;
  MFCD_REQUIRE(out != nullptr, "mfcd_philox_uniforms: out is NULL");
  k_philox_uniforms<<<grid_for(count, 256, 8), 256, 0, as_stream(stream)>>>(seed, w0, count, out);
  MFCD_CHECK_LAUNCH();
  return MFCD_OK;
}

// ---------------------------------------------------------------------------
// Batch layout for K1's RUNS mode: inside every batch of `batch_size` consecutive records, bring the
// triplets of one user together (stable: a user's triplets keep their relative order).  Which triplets a
// batch holds does not change, so the optimiser step is the same sum in a different order.
// One stable radix sort (cub, library) of (batch index << 32 | user) keys per chunk of whole batches.
// ---------------------------------------------------------------------------
namespace mfcd {
constexpr int64_t kGroupChunk = int64_t(1) << 25;

struct GroupLayout {
  size_t k_in, k_out, i_in, i_out, tmp, cub_temp, total, cub_bytes;
  int64_t chunk;
};

static GroupLayout group_layout(int64_t N, int64_t batch) {
  GroupLayout L;
  int64_t chunk = batch >= kGroupChunk ? batch : (kGroupChunk / batch) * batch;
  if (chunk > N) chunk = N;
  L.chunk = chunk;
  auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
  size_t off = 0;
  L.k_in = off; off += up(sizeof(uint64_t) * chunk);
  L.k_out = off; off += up(sizeof(uint64_t) * chunk);
  L.i_in = off; off += up(sizeof(uint32_t) * chunk);
  L.i_out = off; off += up(sizeof(uint32_t) * chunk);
  L.tmp = off; off += up(sizeof(mfcd_triplet) * chunk);
  size_t cb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cb, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr,
                                  (uint32_t*)nullptr, chunk, 0, 64);
  L.cub_bytes = cb;
  L.cub_temp = off; off += up(cb);
  L.total = off;
  return L;
}

__global__ void __launch_bounds__(256)
k_group_keys(const mfcd_triplet* __restrict__ rec, int64_t count, int64_t batch, uint64_t* __restrict__ keys,
             uint32_t* __restrict__ idx) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < count; k += (int64_t)gridDim.x * blockDim.x) {
    keys[k] = ((uint64_t)(k / batch) << 32) | (uint32_t)__ldg(&rec[k].u);
    idx[k] = (uint32_t)k;
  }
}

__global__ void __launch_bounds__(256)
k_group_gather(const mfcd_triplet* __restrict__ rec, const uint32_t* __restrict__ idx, int64_t count,
               mfcd_triplet* __restrict__ out) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < count; k += (int64_t)gridDim.x * blockDim.x)
    reinterpret_cast<int4*>(out)[k] = __ldg(reinterpret_cast<const int4*>(rec) + idx[k]);
}
}  // namespace mfcd

extern "C" int mfcd_group_by_user_workspace(int64_t N, int64_t batch_size, size_t* bytes) {
  MFCD_REQUIRE(bytes != nullptr && N >= 0 && batch_size >= 1, "mfcd_group_by_user_workspace: bad argument");
  *bytes = N == 0 ? 0 : mfcd::group_layout(N, batch_size).total;
  return MFCD_OK;
}

extern "C" int mfcd_group_by_user(mfcd_triplet* rec, int64_t N, int64_t batch_size, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  using namespace mfcd;
  MFCD_REQUIRE(N >= 0 && batch_size >= 1, "mfcd_group_by_user: bad sizes");
  if (N == 0) return MFCD_OK;
  MFCD_REQUIRE(rec != nullptr, "mfcd_group_by_user: rec is NULL");
  MFCD_REQUIRE(batch_size < (int64_t(1) << 32), "mfcd_group_by_user: batch too large");
  const GroupLayout L = group_layout(N, batch_size);
  if (workspace == nullptr || workspace_bytes < L.total) {
    set_error("mfcd_group_by_user: workspace too small (%zu < %zu bytes)", workspace_bytes, L.total);
    return MFCD_ERR_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  char* base = static_cast<char*>(workspace);
  uint64_t* k_in = reinterpret_cast<uint64_t*>(base + L.k_in);
  uint64_t* k_out = reinterpret_cast<uint64_t*>(base + L.k_out);
  uint32_t* i_in = reinterpret_cast<uint32_t*>(base + L.i_in);
  uint32_t* i_out = reinterpret_cast<uint32_t*>(base + L.i_out);
  mfcd_triplet* tmp = reinterpret_cast<mfcd_triplet*>(base + L.tmp);
  for (int64_t c0 = 0; c0 < N; c0 += L.chunk) {
    const int64_t count = (N - c0) < L.chunk ? (N - c0) : L.chunk;
    const int64_t nb = (count + batch_size - 1) / batch_size;
    int hi_bits = 0;
    while ((int64_t(1) << hi_bits) < nb) ++hi_bits;
    const int grid = grid_for(count, 256, 8);
    k_group_keys<<<grid, 256, 0, st>>>(rec + c0, count, batch_size, k_in, i_in);
    MFCD_CHECK_LAUNCH();
    size_t cb = L.cub_bytes;
    MFCD_CUDA(cub::DeviceRadixSort::SortPairs(base + L.cub_temp, cb, (const uint64_t*)k_in, k_out,
                                              (const uint32_t*)i_in, i_out, count, 0, 32 + hi_bits, st));
    k_group_gather<<<grid, 256, 0, st>>>(rec + c0, i_out, count, tmp);
    MFCD_CHECK_LAUNCH();
    MFCD_CUDA(cudaMemcpyAsync(rec + c0, tmp, sizeof(mfcd_triplet) * count, cudaMemcpyDeviceToDevice, st));
  }
  return MFCD_OK;
}
