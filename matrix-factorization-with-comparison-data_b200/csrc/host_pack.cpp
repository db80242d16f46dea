// Host-side packer of the 8-byte staging format (include/mfcd_b200.h: mfcd_host_pack_triplets8).
//
// The reference's loader hands the training step int64 (u, i, j) and a float label per sample
// (structure.py:845-846, 28 bytes per sample through `x.to(device)`); this library's host record is 16 bytes.
// PCIe is the limit of a host-fed step (16 B x 2^22 triplets = 67 MB = 1.2 ms against 0.46 ms of GPU work), so a
// host that wants to feed the GPU faster packs each batch to 8 bytes per triplet on its own cores WHILE the
// previous batch is in flight.  This file is that packer: a small persistent thread pool (no per-call thread
// creation), AVX-512 kernel with streaming stores when the CPU has it, portable scalar loop otherwise.  The bits
// are those of the device packer mfcd_pack_triplets8 (wire.cu) and of hostpack.pack8 (numpy): bit 0 label,
// [1,21) j, [21,41) i, [41,64) u.
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "../../include/mfcd_b200.h"

namespace {

// ---- one chunk of records -> packed words; returns non-zero if a record does not fit the format -------------
inline uint32_t pack_scalar(const mfcd_triplet* rec, int64_t n, uint64_t* out) {
    uint32_t bad = 0;
    for (int64_t k = 0; k < n; ++k) {
        uint32_t u = (uint32_t)rec[k].u, i = (uint32_t)rec[k].i, j = (uint32_t)rec[k].j, zb;
        std::memcpy(&zb, &rec[k].z, 4);
        // labels: +0.0f, -0.0f -> 0, 1.0f -> 1; anything else is a soft label
        bad |= (u >> 23) | (i >> 20) | (j >> 20) | (uint32_t)((zb & 0x7fffffffu) != 0 && zb != 0x3f800000u);
        out[k] = ((uint64_t)u << 41) | ((uint64_t)i << 21) | ((uint64_t)j << 1) | (uint64_t)(zb == 0x3f800000u);
    }
    return bad;
}

#if defined(__x86_64__)
#define MFCD_AVX512 __attribute__((target("avx512f,avx512bw,avx512dq")))
// Four records (one zmm).  A record is two 64-bit lanes: even = u | i << 32, odd = j | zbits << 32.
// Returns the packed word of record r in BOTH lanes 2r and 2r+1; ORs format violations into `bad`.
MFCD_AVX512 static inline __m512i pack4_avx512(const __m512i v, __mmask8& bad) {
    const __m512i zero = _mm512_setzero_si512();
    const __m512i low = _mm512_and_si512(v, _mm512_set1_epi64(0xffffffffll)), high = _mm512_srli_epi64(v, 32);
    // even lanes: u << 41 | i << 21; odd lanes: j << 1 | (z == 1.0f)
    const __m512i ev = _mm512_or_si512(_mm512_slli_epi64(low, 41), _mm512_slli_epi64(high, 21));
    const __mmask8 is_one = _mm512_cmpeq_epi64_mask(high, _mm512_set1_epi64(0x3f800000ll));
    const __m512i j2 = _mm512_slli_epi64(low, 1);
    const __m512i od = _mm512_mask_add_epi64(j2, is_one, j2, _mm512_set1_epi64(1));
    const __mmask8 bu = _mm512_cmpneq_epi64_mask(_mm512_srli_epi64(low, 23), zero);
    const __mmask8 bi = _mm512_cmpneq_epi64_mask(_mm512_srli_epi64(high, 20), zero);
    const __mmask8 bj = _mm512_cmpneq_epi64_mask(_mm512_srli_epi64(low, 20), zero);
    const __mmask8 soft = (__mmask8)(_mm512_cmpneq_epi64_mask(_mm512_and_si512(high, _mm512_set1_epi64(0x7fffffffll)), zero) & ~is_one);
    bad |= (__mmask8)(((bu | bi) & 0x55) | ((bj | soft) & 0xaa));
    const __m512i w = _mm512_mask_blend_epi64(0xaa, ev, od);
    return _mm512_or_si512(w, _mm512_shuffle_epi32(w, (_MM_PERM_ENUM)0x4e));
}

MFCD_AVX512 uint32_t pack_avx512(const mfcd_triplet* rec, int64_t n, uint64_t* out) {
    const __m512i pick = _mm512_setr_epi64(0, 2, 4, 6, 8, 10, 12, 14);
    __mmask8 bad = 0;
    int64_t k = 0;
    const bool aligned = (((uintptr_t)out) & 63) == 0;
    for (; k + 8 <= n; k += 8) {
        const __m512i a = pack4_avx512(_mm512_loadu_si512((const void*)(rec + k)), bad);
        const __m512i b = pack4_avx512(_mm512_loadu_si512((const void*)(rec + k + 4)), bad);
        const __m512i r = _mm512_permutex2var_epi64(a, pick, b);
        if (aligned) _mm512_stream_si512((__m512i*)(out + k), r);      // the words are read next by the DMA engine
        else _mm512_storeu_si512((void*)(out + k), r);
    }
    const uint32_t tail_bad = pack_scalar(rec + k, n - k, out + k);
    _mm_sfence();
    return (uint32_t)bad | tail_bad;
}
#endif

typedef uint32_t (*pack_fn)(const mfcd_triplet*, int64_t, uint64_t*);

pack_fn choose_pack() {
    const char* force = std::getenv("MFCD_HOST_PACK_ISA");          // "scalar": skip the vector kernel (tests)
    if (force && std::strcmp(force, "scalar") == 0) return pack_scalar;
#if defined(__x86_64__)
    if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512dq"))
        return pack_avx512;
#endif
    return pack_scalar;
}

// ---- a persistent pool: workers sleep on a condition variable between jobs, chunks are claimed atomically ----
struct Job {
    const mfcd_triplet* rec = nullptr;
    uint64_t* out = nullptr;
    int64_t n = 0, chunk = 0, n_chunks = 0;
    pack_fn fn = nullptr;
    std::atomic<int64_t> next{0};
    std::atomic<uint32_t> bad{0};
};

class Pool {
public:
    ~Pool() {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    // runs the job on `threads` threads in total (the caller is one of them)
    uint32_t run(Job& job, int threads) {
        std::lock_guard<std::mutex> serial(api_);                 // one job at a time
        grow(threads - 1);
        {
            std::lock_guard<std::mutex> g(m_);
            job_ = &job;
            wanted_ = threads - 1;
            active_ = 0;
            ++generation_;
        }
        cv_.notify_all();
        work(job);
        std::unique_lock<std::mutex> g(m_);
        wanted_ = 0;                                              // late wakers find nothing to join
        done_.wait(g, [&] { return active_ == 0; });
        job_ = nullptr;
        return job.bad.load();
    }

private:
    static void work(Job& job) {
        uint32_t bad = 0;
        for (;;) {
            const int64_t c = job.next.fetch_add(1, std::memory_order_relaxed);
            if (c >= job.n_chunks) break;
            const int64_t k0 = c * job.chunk, k1 = (k0 + job.chunk < job.n) ? k0 + job.chunk : job.n;
            bad |= job.fn(job.rec + k0, k1 - k0, job.out + k0);
        }
        if (bad) job.bad.fetch_or(bad);
    }
    void grow(int n) {
        while ((int)workers_.size() < n) workers_.emplace_back([this] { loop(); });
    }
    void loop() {
        uint64_t seen = 0;
        std::unique_lock<std::mutex> g(m_);
        for (;;) {
            cv_.wait(g, [&] { return stop_ || (generation_ != seen && wanted_ > 0); });
            if (stop_) return;
            seen = generation_;
            --wanted_;
            ++active_;
            Job* job = job_;
            g.unlock();
            work(*job);
            g.lock();
            if (--active_ == 0) done_.notify_all();
        }
    }
    std::mutex api_, m_;
    std::condition_variable cv_, done_;
    std::vector<std::thread> workers_;
    Job* job_ = nullptr;
    uint64_t generation_ = 0;
    int wanted_ = 0, active_ = 0;
    bool stop_ = false;
};

Pool& pool() {
    static Pool* p = new Pool();       // never destroyed: no join at process exit from a foreign runtime
    return *p;
}

}  // namespace

extern "C" int mfcd_host_pack_triplets8(const mfcd_triplet* rec, int64_t N, uint64_t* out, int32_t threads,
                                        int32_t* bad) {
    if (N < 0 || (N > 0 && (!rec || !out))) return MFCD_ERR_ARG;
    static const pack_fn fn = choose_pack();
    int hw = (int)std::thread::hardware_concurrency();
    if (hw < 1) hw = 1;
    if (threads < 1 || threads > hw) threads = hw;
    if (threads > 64) threads = 64;
    uint32_t flag;
    const int64_t CHUNK = 1 << 15;                                // 512 KB of records per claim, 64-byte aligned output
    if (threads == 1 || N <= 2 * CHUNK) {
        flag = fn(rec, N, out);
    } else {
        Job job;
        job.rec = rec, job.out = out, job.n = N, job.chunk = CHUNK, job.n_chunks = (N + CHUNK - 1) / CHUNK, job.fn = fn;
        flag = pool().run(job, threads);
    }
    if (bad && flag) *bad = 1;
    return MFCD_OK;
}

extern "C" int mfcd_host_pack_isa(void) {
#if defined(__x86_64__)
    return choose_pack() == pack_avx512 ? 512 : 0;
#else
    return 0;
#endif
}
