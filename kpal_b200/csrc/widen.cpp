// Host side of the narrow device->host copy of a count profile.
//
// Profile.counts is int64[4^k] (reference kpal/klib.py:170), but the counts of a
// large-k profile are small: moving them over PCIe as uint16 and widening on the
// host takes a quarter of the bytes of the int64 copy, which is the longest single
// piece of `kpal count` end to end (134 MB at k = 12).  The copy runs in chunks;
// this file holds the workers that widen chunk c into the caller's int64 array
// while chunk c+1 is still in flight (cabi.cu drives the copy and publishes the
// chunks as they land).
//
// The workers are a small persistent pool (created on first use, never joined: the
// library may be unloaded at interpreter exit while they sleep on the condition
// variable), so a call costs a wake-up, not a thread spawn per chunk.
#include "../../include/kpal_b200.h"

#if defined(__x86_64__) || defined(__i386__)
#define KPAL_WIDEN_X86 1
#include <immintrin.h>
#else
#define KPAL_WIDEN_X86 0        // e.g. the aarch64 host of a GB200: plain loops (the compiler emits NEON)
#endif
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace kpal {

#if KPAL_WIDEN_X86
// src[0..n) uint16 -> dst[0..n) int64.  Streaming (non-temporal) stores where dst is
// 16-byte aligned: the array is written once and is larger than the caches, so the
// read-for-ownership of an ordinary store would double the memory traffic.
static void widen_u16_range(const uint16_t *src, int64_t *dst, uint64_t n)
{
    uint64_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 15u)) { dst[i] = src[i]; ++i; }
    const __m128i zero = _mm_setzero_si128();
    for (; i + 8 <= n; i += 8) {
        const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i));   // 8 x u16
        const __m128i lo = _mm_unpacklo_epi16(v, zero), hi = _mm_unpackhi_epi16(v, zero);   // 4 x u32 each
        __m128i *out = reinterpret_cast<__m128i *>(dst + i);
        _mm_stream_si128(out + 0, _mm_unpacklo_epi32(lo, zero));
        _mm_stream_si128(out + 1, _mm_unpackhi_epi32(lo, zero));
        _mm_stream_si128(out + 2, _mm_unpacklo_epi32(hi, zero));
        _mm_stream_si128(out + 3, _mm_unpackhi_epi32(hi, zero));
    }
    for (; i < n; ++i) dst[i] = src[i];
    _mm_sfence();
}

// src[0..n) uint8 -> dst[0..n) int64, same store discipline.
static void widen_u8_range(const uint8_t *src, int64_t *dst, uint64_t n)
{
    uint64_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 15u)) { dst[i] = src[i]; ++i; }
    const __m128i zero = _mm_setzero_si128();
    for (; i + 16 <= n; i += 16) {
        const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i));   // 16 x u8
        const __m128i h[2] = {_mm_unpacklo_epi8(v, zero), _mm_unpackhi_epi8(v, zero)};      // 8 x u16 each
        __m128i *out = reinterpret_cast<__m128i *>(dst + i);
        for (int q = 0; q < 2; ++q) {
            const __m128i lo = _mm_unpacklo_epi16(h[q], zero), hi = _mm_unpackhi_epi16(h[q], zero);
            _mm_stream_si128(out + 4 * q + 0, _mm_unpacklo_epi32(lo, zero));
            _mm_stream_si128(out + 4 * q + 1, _mm_unpackhi_epi32(lo, zero));
            _mm_stream_si128(out + 4 * q + 2, _mm_unpacklo_epi32(hi, zero));
            _mm_stream_si128(out + 4 * q + 3, _mm_unpackhi_epi32(hi, zero));
        }
    }
    for (; i < n; ++i) dst[i] = src[i];
    _mm_sfence();
}

// AVX-512 forms (chosen at run time): one zero-extending load and ONE 64-byte streaming
// store per cache line of the destination instead of four 16-byte ones.
__attribute__((target("avx512f"))) static void widen_u16_range_512(const uint16_t *src, int64_t *dst, uint64_t n)
{
    uint64_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 63u)) { dst[i] = src[i]; ++i; }
    for (; i + 8 <= n; i += 8)
        _mm512_stream_si512(reinterpret_cast<__m512i *>(dst + i),
                            _mm512_cvtepu16_epi64(_mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i))));
    for (; i < n; ++i) dst[i] = src[i];
    _mm_sfence();
}
__attribute__((target("avx512f"))) static void widen_u8_range_512(const uint8_t *src, int64_t *dst, uint64_t n)
{
    uint64_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 63u)) { dst[i] = src[i]; ++i; }
    for (; i + 8 <= n; i += 8)
        _mm512_stream_si512(reinterpret_cast<__m512i *>(dst + i),
                            _mm512_cvtepu8_epi64(_mm_loadl_epi64(reinterpret_cast<const __m128i *>(src + i))));
    for (; i < n; ++i) dst[i] = src[i];
    _mm_sfence();
}
static const bool g_have_avx512 = [] {
    const char *e = getenv("KPAL_NO_AVX512");
    return __builtin_cpu_supports("avx512f") && !(e && e[0] == '1');
}();

#else
// Portable form for non-x86 hosts: zero-extending loops the compiler vectorises.
static void widen_u16_range(const uint16_t *src, int64_t *dst, uint64_t n)
{
    for (uint64_t i = 0; i < n; ++i) dst[i] = src[i];
}
static void widen_u8_range(const uint8_t *src, int64_t *dst, uint64_t n)
{
    for (uint64_t i = 0; i < n; ++i) dst[i] = src[i];
}
static void widen_u16_range_512(const uint16_t *src, int64_t *dst, uint64_t n) { widen_u16_range(src, dst, n); }
static void widen_u8_range_512(const uint8_t *src, int64_t *dst, uint64_t n) { widen_u8_range(src, dst, n); }
static const bool g_have_avx512 = false;
#endif

static void widen_range(const void *src, int width, uint64_t b, int64_t *dst, uint64_t n)
{
    if (width == 2) {
        const uint16_t *s = static_cast<const uint16_t *>(src) + b;
        if (g_have_avx512) widen_u16_range_512(s, dst + b, n); else widen_u16_range(s, dst + b, n);
    } else {
        const uint8_t *s = static_cast<const uint8_t *>(src) + b;
        if (g_have_avx512) widen_u8_range_512(s, dst + b, n); else widen_u8_range(s, dst + b, n);
    }
}

struct WidenJob {
    const void *src = nullptr;
    int width = 2;                  // bytes per source element: 2 (uint16) or 1 (uint8)
    int64_t *dst = nullptr;
    uint64_t n = 0;                 // elements
    uint64_t chunk = 0;             // elements per published chunk
    uint64_t piece = 0;             // elements per work item (a chunk is cut into pieces)
    uint64_t n_items = 0;
    std::atomic<uint64_t> next{0};          // next work item to hand out
    std::atomic<uint64_t> done{0};          // finished work items
    std::atomic<uint64_t> ready{0};         // elements that have landed in src
    std::atomic<int> abort{0};
    // a job of another kind (the FASTA segment packers of cabi.cu): every worker calls it once
    void (*custom)(void *) = nullptr;
    void *custom_arg = nullptr;
};

class WidenPool {
public:
    static WidenPool &get()
    {
        static WidenPool *pool = new WidenPool();       // intentionally leaked (see header)
        return *pool;
    }

    void start(WidenJob *job)
    {
        {
            std::lock_guard<std::mutex> lock(m_);
            job_ = job;
            ++generation_;
        }
        cv_.notify_all();
    }

    // the caller works too; returns when every item is done
    void finish(WidenJob *job)
    {
        work(job);
        while (job->done.load(std::memory_order_acquire) < job->n_items) std::this_thread::yield();
        std::unique_lock<std::mutex> lock(m_);
        job_ = nullptr;
        // no worker may still hold a pointer to the job when the caller's frame goes away
        idle_cv_.wait(lock, [&] { return active_ == 0; });
    }

    unsigned workers() const { return unsigned(threads_.size()); }

private:
    WidenPool()
    {
        unsigned hw = std::thread::hardware_concurrency();
        if (hw == 0) hw = 1;
        unsigned n = hw > 16 ? 15 : (hw > 1 ? hw - 1 : 0);
        for (unsigned t = 0; t < n; ++t) {
            threads_.emplace_back([this] { loop(); });
            threads_.back().detach();
        }
    }

    void loop()
    {
        uint64_t seen = 0;
        for (;;) {
            WidenJob *job = nullptr;
            {
                std::unique_lock<std::mutex> lock(m_);
                cv_.wait(lock, [&] { return generation_ != seen; });
                seen = generation_;
                job = job_;
                if (job) ++active_;
            }
            if (!job) continue;
            work(job);
            {
                std::lock_guard<std::mutex> lock(m_);
                --active_;
            }
            idle_cv_.notify_all();
        }
    }

    static void work(WidenJob *job)
    {
        if (job->custom) { job->custom(job->custom_arg); return; }
        for (;;) {
            const uint64_t item = job->next.fetch_add(1, std::memory_order_relaxed);
            if (item >= job->n_items) return;
            const uint64_t b = item * job->piece;
            const uint64_t e = b + job->piece < job->n ? b + job->piece : job->n;
            // the chunk holding [b, e) must have landed
            const uint64_t need = ((e + job->chunk - 1) / job->chunk) * job->chunk;
            const uint64_t need_c = need < job->n ? need : job->n;
            while (job->ready.load(std::memory_order_acquire) < need_c) {
                if (job->abort.load(std::memory_order_relaxed)) break;
                std::this_thread::yield();
            }
            if (!job->abort.load(std::memory_order_relaxed)) widen_range(job->src, job->width, b, job->dst, e - b);
            job->done.fetch_add(1, std::memory_order_release);
        }
    }

    std::mutex m_;
    std::condition_variable cv_, idle_cv_;
    std::vector<std::thread> threads_;
    WidenJob *job_ = nullptr;
    uint64_t generation_ = 0;
    unsigned active_ = 0;
};

// ---- interface used by cabi.cu --------------------------------------------
// begin: wake the workers on (src -> dst); publish: the first `elements` of src are
// valid; end: join in, wait for completion (abort = 1 drops the remaining work, the
// destination is then unspecified and the caller rewrites it).
struct WidenHandle { WidenJob job; };
static std::mutex g_widen_one_job;          // the pool serves one job at a time: begin .. end

WidenHandle *widen_begin(const void *src, int width, int64_t *dst, uint64_t n, uint64_t chunk)
{
    g_widen_one_job.lock();
    WidenHandle *h = new WidenHandle();
    WidenJob &j = h->job;
    j.src = src; j.width = width; j.dst = dst; j.n = n;
    j.chunk = chunk ? chunk : n;
    // pieces of <= 64 K elements that divide a chunk: fine-grained enough for 16 workers
    // on the first chunk, coarse enough that handing them out costs nothing
    uint64_t piece = 1u << 16;
    while (piece > 1024 && j.chunk % piece) piece >>= 1;
    if (j.chunk % piece) piece = j.chunk;
    j.piece = piece;
    j.n_items = (n + piece - 1) / piece;
    WidenPool::get().start(&j);
    return h;
}

void widen_publish(WidenHandle *h, uint64_t elements)
{
    h->job.ready.store(elements, std::memory_order_release);
}

void widen_end(WidenHandle *h, int abort)
{
    if (abort) h->job.abort.store(1);
    WidenPool::get().finish(&h->job);
    delete h;
    g_widen_one_job.unlock();
}

unsigned widen_workers() { return WidenPool::get().workers() + 1; }

// The same pool for another kind of work: fn(arg) runs once on every worker (begin) and on
// the caller (end), which returns when all of them are back.  One job at a time, like the
// widening: begin .. end holds the pool.
WidenHandle *pool_run_begin(void (*fn)(void *), void *arg)
{
    g_widen_one_job.lock();
    WidenHandle *h = new WidenHandle();
    h->job.custom = fn;
    h->job.custom_arg = arg;
    WidenPool::get().start(&h->job);
    return h;
}

void pool_run_end(WidenHandle *h)
{
    WidenPool::get().finish(&h->job);
    delete h;
    g_widen_one_job.unlock();
}

}  // namespace kpal

// Host only: the widening stage on its own (include/kpal_b200.h).  The chunks are
// published one after the other, as the device->host copy does.
extern "C" int kpal_widen_u16(const uint16_t *narrow, uint64_t n, uint64_t chunk, int64_t *counts_out)
{
    if (n == 0) return KPAL_OK;
    if (!narrow || !counts_out) return KPAL_EINVAL;
    if (chunk == 0 || chunk > n) chunk = n;
    kpal::WidenHandle *h = kpal::widen_begin(narrow, 2, counts_out, n, chunk);
    for (uint64_t at = chunk; ; at += chunk) {
        kpal::widen_publish(h, at < n ? at : n);
        if (at >= n) break;
    }
    kpal::widen_end(h, 0);
    return KPAL_OK;
}

extern "C" int kpal_widen_u8(const uint8_t *narrow, uint64_t n, uint64_t chunk, int64_t *counts_out)
{
    if (n == 0) return KPAL_OK;
    if (!narrow || !counts_out) return KPAL_EINVAL;
    if (chunk == 0 || chunk > n) chunk = n;
    kpal::WidenHandle *h = kpal::widen_begin(narrow, 1, counts_out, n, chunk);
    for (uint64_t at = chunk; ; at += chunk) {
        kpal::widen_publish(h, at < n ? at : n);
        if (at >= n) break;
    }
    kpal::widen_end(h, 0);
    return KPAL_OK;
}
