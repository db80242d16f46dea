// Run-length wire format for one user-grouped batch of HARD-labelled comparisons (host <-> device staging).
//
// The end-to-end step is PCIe-bound: 16-byte records move at ~3.4e9 triplets/s, the 8-byte format at
// ~5.7e9, while K1 + K3 run at > 8e9.  A batch that keeps each user's triplets adjacent
// (mfcd_group_by_user) does not need the user id per triplet: the wire carries
//     header[4]      u32: n_runs, B, 0, 0
//     word_run0[nw]  u32: number of run starts before triplet 32 w
//     zbits[nw]      u32: label bit of triplet 32 w + p at bit p
//     nbits[nw]      u32: 1 where a triplet's user differs from its predecessor's (bit 0 of the batch is 1)
//     ij[B]          u32: i | j << 16            (n_items <= 65536)
//     users[n_runs]  u32: the user of every run, in order
// = 4.375 bytes per triplet + 4 bytes per run (config 4: ~4.5 bytes per triplet instead of 16).
// K1 (k_fwd_bwd_lean<..., WIRE>) reads this format directly: a warp tile is one bit word, the user of lane p
// is users[word_run0[w] + popc(nbits[w] & mask(p)) - 1].  mfcd_unpack_wire rebuilds 16-byte records for the
// other consumers.
#include <cub/device/device_scan.cuh>
#include "internal.h"

namespace mfcd {

constexpr int kWireBlockWords = 256;                         // bit words per block
constexpr int kWireBlock = kWireBlockWords * 32;             // triplets per block

struct WireLayout {
  int64_t nw, nb, run0, zbits, nbits, ij, users, fixed_words;   // run0 = word_run0; offsets in u32 words
};

static WireLayout wire_layout(int64_t B) {
  WireLayout L;
  L.nw = (B + 31) / 32;
  L.nb = (L.nw + kWireBlockWords - 1) / kWireBlockWords;
  L.run0 = 4;
  L.zbits = L.run0 + L.nw;
  L.nbits = L.zbits + L.nw;
  L.ij = L.nbits + L.nw;
  L.users = L.ij + B;
  L.fixed_words = L.users;
  return L;
}

// bit words + ij + per-block run-start counts.  One thread per triplet, 256-thread CTAs walk whole blocks.
__global__ void __launch_bounds__(256)
k_wire_bits(const mfcd_triplet* __restrict__ rec, int64_t B, WireLayout L, uint32_t* __restrict__ wire,
            uint32_t* __restrict__ blk_count, int* __restrict__ bad) {
  __shared__ uint32_t s_cnt[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t b = blockIdx.x; b < L.nb; b += gridDim.x) {
    uint32_t cnt = 0;
    for (int w0 = warp; w0 < kWireBlockWords; w0 += 8) {
      const int64_t w = b * kWireBlockWords + w0;
      const int64_t k = w * 32 + lane;
      bool zb = false, nbit = false;
      if (k < B) {
        const int4 r = __ldg(reinterpret_cast<const int4*>(rec) + k);
        const float z = __int_as_float(r.w);
        if ((z != 0.f && z != 1.f) || (unsigned)r.y >= 65536u || (unsigned)r.z >= 65536u || r.x < 0) atomicExch(bad, 1);
        zb = z != 0.f;
        nbit = (k == 0) || (__ldg(&rec[k - 1].u) != r.x);
        wire[L.ij + k] = ((uint32_t)r.y & 0xffffu) | ((uint32_t)r.z << 16);
      }
      const uint32_t zw = __ballot_sync(0xffffffffu, zb), nwd = __ballot_sync(0xffffffffu, nbit);
      if (lane == 0 && w < L.nw) {
        wire[L.zbits + w] = zw;
        wire[L.nbits + w] = nwd;
        cnt += __popc(nwd);
      }
    }
    if (lane == 0) s_cnt[warp] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t t = 0;
      for (int q = 0; q < 8; ++q) t += s_cnt[q];
      blk_count[b] = t;
    }
    __syncthreads();
  }
}

// exclusive scan of the 256 per-word run-start counts of a block: thread t -> starts in words before t
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* s_warp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  uint32_t before = 0;
  for (int q = 0; q < warp; ++q) before += s_warp[q];
  __syncthreads();
  return before + incl - v;
}

__global__ void __launch_bounds__(256)
k_wire_users(const mfcd_triplet* __restrict__ rec, int64_t B, WireLayout L, const uint32_t* __restrict__ block_run0,
             uint32_t* __restrict__ wire) {
  __shared__ uint32_t s_warp[8];
  __shared__ uint32_t s_base[kWireBlockWords];
  for (int64_t b = blockIdx.x; b < L.nb; b += gridDim.x) {
    const int64_t w = b * kWireBlockWords + threadIdx.x;
    const uint32_t nbw = w < L.nw ? wire[L.nbits + w] : 0u;
    const uint32_t base = block_run0[b] + block_excl_scan_256(__popc(nbw), s_warp);
    s_base[threadIdx.x] = base;
    if (w < L.nw) wire[L.run0 + w] = base;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int w0 = warp; w0 < kWireBlockWords; w0 += 8) {
      const int64_t ww = b * kWireBlockWords + w0;
      if (ww >= L.nw) break;
      const uint32_t bits = wire[L.nbits + ww];
      if ((bits >> lane) & 1u) {
        const uint32_t run = s_base[w0] + __popc(bits & ((1u << lane) - 1u));
        wire[L.users + run] = (uint32_t)__ldg(&rec[ww * 32 + lane].u);
      }
    }
    __syncthreads();
  }
}

__global__ void k_wire_header(WireLayout L, int64_t B, const uint32_t* __restrict__ block_run0,
                              const uint32_t* __restrict__ blk_count, uint32_t* __restrict__ wire) {
  wire[0] = block_run0[L.nb - 1] + blk_count[L.nb - 1];
  wire[1] = (uint32_t)B;
  wire[2] = 0; wire[3] = 0;
}

__global__ void __launch_bounds__(256)
k_wire_unpack(const uint32_t* __restrict__ wire, int64_t B, WireLayout L, mfcd_triplet* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t w = warp0; w < L.nw; w += nwarps) {
    const int64_t k = w * 32 + lane;
    if (k < B) {
      const uint32_t bits = __ldg(wire + L.nbits + w);
      const uint32_t zw = __ldg(wire + L.zbits + w);
      const uint32_t ij = __ldg(wire + L.ij + k);
      // run of this triplet = starts up to and including its own bit, minus one
      const uint32_t run = __ldg(wire + L.run0 + w) + __popc(bits & (0xffffffffu >> (31 - lane))) - 1u;
      int4 r;
      r.x = (int)__ldg(wire + L.users + run);
      r.y = (int)(ij & 0xffffu);
      r.z = (int)(ij >> 16);
      r.w = __float_as_int(((zw >> lane) & 1u) ? 1.f : 0.f);
      reinterpret_cast<int4*>(out)[k] = r;
    }
  }
}

}  // namespace mfcd

using namespace mfcd;

static size_t wire_cub_bytes(int64_t nb) {
  size_t cb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, cb, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)(nb > 0 ? nb : 1));
  return (cb + 255) & ~size_t(255);
}

extern "C" int mfcd_wire_layout(int64_t B, int64_t* fixed_words, int64_t* capacity_words, size_t* workspace_bytes) {
  MFCD_REQUIRE(B >= 0, "mfcd_wire_layout: B < 0");
  const WireLayout L = wire_layout(B);
  if (fixed_words) *fixed_words = L.fixed_words;
  if (capacity_words) *capacity_words = L.fixed_words + B;          // worst case: every triplet its own run
  if (workspace_bytes) *workspace_bytes = wire_cub_bytes(L.nb) + 2 * sizeof(uint32_t) * (size_t)(L.nb > 0 ? L.nb : 1);
  return MFCD_OK;
}

extern "C" int mfcd_pack_wire(const mfcd_triplet* rec, int64_t B, uint32_t* wire, int64_t capacity_words, int32_t* bad,
                              void* workspace, size_t workspace_bytes, void* stream) {
  MFCD_REQUIRE(B >= 1 && B < (int64_t(1) << 31), "mfcd_pack_wire: bad batch size");
  MFCD_REQUIRE(rec && wire && bad, "mfcd_pack_wire: NULL pointer");
  const WireLayout L = wire_layout(B);
  MFCD_REQUIRE(capacity_words >= L.fixed_words + B, "mfcd_pack_wire: wire buffer too small");
  const size_t cb_al = wire_cub_bytes(L.nb);
  if (workspace == nullptr || workspace_bytes < cb_al + 2 * sizeof(uint32_t) * (size_t)L.nb) {
    set_error("mfcd_pack_wire: workspace too small");
    return MFCD_ERR_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  uint32_t* blk_count = reinterpret_cast<uint32_t*>(static_cast<char*>(workspace) + cb_al);
  uint32_t* block_run0 = blk_count + L.nb;
  const int grid = (int)(L.nb < 148 * 8 ? L.nb : 148 * 8);
  k_wire_bits<<<grid, 256, 0, st>>>(rec, B, L, wire, blk_count, bad);
  MFCD_CHECK_LAUNCH();
  size_t cb = cb_al;
  MFCD_CUDA(cub::DeviceScan::ExclusiveSum(workspace, cb, (const uint32_t*)blk_count, block_run0, (int)L.nb, st));
  k_wire_users<<<grid, 256, 0, st>>>(rec, B, L, block_run0, wire);
  MFCD_CHECK_LAUNCH();
  k_wire_header<<<1, 1, 0, st>>>(L, B, block_run0, blk_count, wire);
  MFCD_CHECK_LAUNCH();
  return MFCD_OK;
}

extern "C" int mfcd_unpack_wire(const uint32_t* wire, int64_t B, mfcd_triplet* out, void* stream) {
  MFCD_REQUIRE(B >= 1 && B < (int64_t(1) << 31), "mfcd_unpack_wire: bad batch size");
  MFCD_REQUIRE(wire && out, "mfcd_unpack_wire: NULL pointer");
  const WireLayout L = wire_layout(B);
  k_wire_unpack<<<grid_for(B, 256, 8), 256, 0, as_stream(stream)>>>(wire, B, L, out);
  MFCD_CHECK_LAUNCH();
  return MFCD_OK;
}
