// Library plumbing: error strings, device info, record packing / gathering.
#include <vector>
#include "common.cuh"
#include "internal.h"

namespace mfcd {

// ---- K1 launch timing (mfcd_profile_k1): CUDA event pairs on the launching stream ------------------------
static thread_local bool g_prof_on = false;
static thread_local std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_events;

K1Timer::K1Timer(cudaStream_t st) : st_(st), on_(g_prof_on) {
  if (!on_) return;
  cudaEvent_t a = nullptr, b = nullptr;
  if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) { on_ = false; return; }
  g_prof_events.emplace_back(a, b);
  cudaEventRecord(a, st_);
}
K1Timer::~K1Timer() {
  if (on_) cudaEventRecord(g_prof_events.back().second, st_);
}

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
  return (int)e;
}

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cached = v;
    cached_dev = dev;
  }
  return cached;
}

__global__ void k_pack(const int64_t* __restrict__ u, const int64_t* __restrict__ i,
                       const int64_t* __restrict__ j, const double* __restrict__ z, int64_t N,
                       mfcd_triplet* __restrict__ out) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < N; k += (int64_t)gridDim.x * blockDim.x) {
    int4 r;
    r.x = (int)u[k]; r.y = (int)i[k]; r.z = (int)j[k];
    r.w = __float_as_int((float)z[k]);          // z.float(), structure.py:849
    reinterpret_cast<int4*>(out)[k] = r;
  }
}

__global__ void k_unpack(const mfcd_triplet* __restrict__ rec, int64_t N, int64_t* u, int64_t* i, int64_t* j,
                         double* z) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < N; k += (int64_t)gridDim.x * blockDim.x) {
    int4 r = __ldg(reinterpret_cast<const int4*>(rec) + k);
    u[k] = r.x; i[k] = r.y; j[k] = r.z; z[k] = (double)__int_as_float(r.w);
  }
}

__global__ void k_gather(const mfcd_triplet* __restrict__ rec, const int32_t* __restrict__ perm, int64_t N,
                         mfcd_triplet* __restrict__ out) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < N; k += (int64_t)gridDim.x * blockDim.x) {
    reinterpret_cast<int4*>(out)[k] = __ldg(reinterpret_cast<const int4*>(rec) + perm[k]);
  }
}

// 8-byte wire format for hard-labelled comparisons (host <-> device transfers are PCIe-bound at 16 B/triplet):
// bits [0,1) label, [1,21) j, [21,41) i, [41,64) u  =>  n_users <= 2^23, n_items <= 2^20.
__global__ void k_pack8(const mfcd_triplet* __restrict__ rec, int64_t N, unsigned long long* __restrict__ out,
                        int* __restrict__ bad) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < N; k += (int64_t)gridDim.x * blockDim.x) {
    const int4 r = __ldg(reinterpret_cast<const int4*>(rec) + k);
    const float z = __int_as_float(r.w);
    if ((z != 0.f && z != 1.f) || (unsigned)r.x >= (1u << 23) || (unsigned)r.y >= (1u << 20) || (unsigned)r.z >= (1u << 20))
      atomicExch(bad, 1);
    out[k] = ((unsigned long long)(unsigned)r.x << 41) | ((unsigned long long)(unsigned)r.y << 21) |
             ((unsigned long long)(unsigned)r.z << 1) | (z != 0.f ? 1ull : 0ull);
  }
}

__global__ void k_unpack8(const unsigned long long* __restrict__ in, int64_t N, mfcd_triplet* __restrict__ out) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < N; k += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long v = __ldg(in + k);
    int4 r;
    r.x = (int)(v >> 41);
    r.y = (int)((v >> 21) & 0xFFFFFu);
    r.z = (int)((v >> 1) & 0xFFFFFu);
    r.w = __float_as_int((v & 1ull) ? 1.f : 0.f);
    reinterpret_cast<int4*>(out)[k] = r;
  }
}

}  // namespace mfcd

using namespace mfcd;

extern "C" int mfcd_pack_triplets8(const mfcd_triplet* rec, int64_t N, uint64_t* out, int32_t* bad, void* stream) {
  MFCD_REQUIRE(N >= 0, "mfcd_pack_triplets8: N < 0");
  if (N == 0) return MFCD_OK;
  MFCD_REQUIRE(rec && out && bad, "mfcd_pack_triplets8: NULL pointer");
  k_pack8<<<grid_for(N, 256, 8), 256, 0, as_stream(stream)>>>(rec, N, reinterpret_cast<unsigned long long*>(out), bad);
  MFCD_CHECK_LAUNCH();
  return MFCD_OK;
}

extern "C" int mfcd_unpack_triplets8(const uint64_t* packed, int64_t N, mfcd_triplet* out, void* stream) {
  MFCD_REQUIRE(N >= 0, "mfcd_unpack_triplets8: N < 0");
  if (N == 0) return MFCD_OK;
  MFCD_REQUIRE(packed && out, "mfcd_unpack_triplets8: NULL pointer");
  k_unpack8<<<grid_for(N, 256, 8), 256, 0, as_stream(stream)>>>(reinterpret_cast<const unsigned long long*>(packed), N, out);
  MFCD_CHECK_LAUNCH();
  return MFCD_OK;
}

extern "C" int mfcd_profile_k1(int32_t enable) {
  g_prof_on = enable != 0;
  return MFCD_OK;
}

extern "C" int mfcd_profile_k1_read(double* total_ms, int64_t* launches) {
  MFCD_REQUIRE(total_ms && launches, "mfcd_profile_k1_read: NULL pointer");
  double tot = 0.0;
  int64_t n = 0;
  for (auto& ev : g_prof_events) {
    float ms = 0.f;
    if (cudaEventSynchronize(ev.second) == cudaSuccess && cudaEventElapsedTime(&ms, ev.first, ev.second) == cudaSuccess) {
      tot += ms;
      ++n;
    }
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  g_prof_events.clear();
  *total_ms = tot;
  *launches = n;
  return MFCD_OK;
}

extern "C" int mfcd_abi_version(void) { return MFCD_ABI_VERSION; }
extern "C" const char* mfcd_last_error(void) { return g_err; }

extern "C" int mfcd_device_sm_count(int* out) {
  MFCD_REQUIRE(out != nullptr, "mfcd_device_sm_count: out is NULL");
  int dev = 0;
  MFCD_CUDA(cudaGetDevice(&dev));
  MFCD_CUDA(cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, dev));
  return MFCD_OK;
}

extern "C" int mfcd_pack_triplets(const int64_t* u, const int64_t* i, const int64_t* j, const double* z,
                                  int64_t N, mfcd_triplet* out, void* stream) {
  MFCD_REQUIRE(N >= 0, "mfcd_pack_triplets: N < 0");
  if (N == 0) return MFCD_OK;
  MFCD_REQUIRE(u && i && j && z && out, "mfcd_pack_triplets: NULL pointer");
  k_pack<<<grid_for(N, 256, 8), 256, 0, as_stream(stream)>>>(u, i, j, z, N, out);
  MFCD_CHECK_LAUNCH();
  return MFCD_OK;
}

extern "C" int mfcd_unpack_triplets(const mfcd_triplet* rec, int64_t N, int64_t* u, int64_t* i, int64_t* j,
                                    double* z, void* stream) {
  MFCD_REQUIRE(N >= 0, "mfcd_unpack_triplets: N < 0");
  if (N == 0) return MFCD_OK;
  MFCD_REQUIRE(rec && u && i && j && z, "mfcd_unpack_triplets: NULL pointer");
  k_unpack<<<grid_for(N, 256, 8), 256, 0, as_stream(stream)>>>(rec, N, u, i, j, z);
  MFCD_CHECK_LAUNCH();
  return MFCD_OK;
}

extern "C" int mfcd_gather_triplets(const mfcd_triplet* rec, const int32_t* perm, int64_t N, mfcd_triplet* out,
                                    void* stream) {
  MFCD_REQUIRE(N >= 0, "mfcd_gather_triplets: N < 0");
  if (N == 0) return MFCD_OK;
  MFCD_REQUIRE(rec && perm && out, "mfcd_gather_triplets: NULL pointer");
  k_gather<<<grid_for(N, 256, 8), 256, 0, as_stream(stream)>>>(rec, perm, N, out);
  MFCD_CHECK_LAUNCH();
  return MFCD_OK;
}
