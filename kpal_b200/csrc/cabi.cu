// extern "C" surface of libkpal_b200.so (see include/kpal_b200.h).
//
// The "host" entry points own the H2D / D2H traffic and the scratch memory;
// the "dev" entry points are thin launch wrappers for callers that already
// hold device memory (bench harness, multi-GPU driver in kpal_b200/multigpu.py).
#include "common.cuh"
#include "slotted.h"

#include <algorithm>
#include <chrono>
#include <stdio.h>
#include <mutex>
#include <memory>
#include <numeric>
#include <thread>
#include <string.h>
#include <stdlib.h>
#include <unistd.h>
#include <vector>

namespace kpal {

// launchers (count.cu / distance.cu)
int launch_count(const uint32_t *, const uint32_t *, uint64_t, int, void *, int, cudaStream_t, bool zero_table = false);
int launch_finalize(const void *, int, int, int, int64_t *, cudaStream_t);
int launch_finalize_narrow(const void *, int, int, int, uint16_t *, uint8_t *, int64_t *, uint64_t, unsigned int *,
                           cudaStream_t, void *d_list8 = nullptr, unsigned int cap8 = 0, void *d_list16 = nullptr,
                           unsigned int cap16 = 0);
// widen.cpp: host workers that widen the uint16 form of a profile to int64
struct WidenHandle;
WidenHandle *widen_begin(const void *src, int width, int64_t *dst, uint64_t n, uint64_t chunk);
void widen_publish(WidenHandle *h, uint64_t elements);
void widen_end(WidenHandle *h, int abort);
int launch_balance(const int64_t *, int64_t *, int, cudaStream_t);
int launch_accumulate(const void *, void *, int, uint64_t, cudaStream_t);
int launch_by_record_u16(const uint32_t *, const uint32_t *, const uint64_t *, uint64_t, uint64_t, int, int,
                         uint16_t *, cudaStream_t);
int launch_by_record(const uint32_t *, const uint32_t *, const uint64_t *, uint64_t, uint64_t, int,
                     int, int64_t *, cudaStream_t);
int launch_prepare(const int64_t *, uint64_t, int, int, int, double *, double *, uint32_t *,
                   double *, double *, unsigned long long *, cudaStream_t);
int launch_distance_tiles(const double *, const double *, const uint32_t *, const double *,
                          const double *, const int32_t *, uint64_t, int, int, int, int, int,
                          uint64_t, uint64_t, double *, uint32_t *, double *, cudaStream_t,
                          double *d_packed = nullptr);
int launch_distance_unpack(const double *, const double *, const double *, const int32_t *, uint64_t, int, int,
                           int, uint64_t, uint64_t, int, double *, cudaStream_t);
uint64_t distance_num_tiles(uint64_t n);
uint64_t distance_tile_elems();
// distance_gram.cu: euclidean / cosine through an exact integer Gram matrix on the tensor cores
uint64_t gram_row_stride(int k);
int launch_gram_prepare(const int64_t *, uint64_t, int, int, uint8_t *, unsigned long long *, unsigned long long *,
                        unsigned int *, cudaStream_t);
int launch_gram_norm_max(const unsigned long long *, uint64_t, unsigned long long *, cudaStream_t);
int launch_gram_distances(const uint8_t *, const unsigned long long *, const unsigned long long *, unsigned long long,
                          uint64_t, int, int, int, int, long long *, double *, cudaStream_t);
uint64_t prepared_stride_host(int k);
uint64_t fasta_scratch_bytes(uint64_t n_bytes);
void set_exact_div(bool on);
void set_count_path(int v);
void set_tiled_finalize(int v);
void set_by_record_path(int v);
void set_radix_payload_bits(int bits);
void set_radix_debug(int v);
void set_radix_shape(int v);
void set_radix_max_buckets(int v);
void set_pair_fused(int v);
void set_pair_flush_every(int v);
int launch_fasta_pack(const uint8_t *, uint64_t, uint32_t *, uint32_t *, void *, cudaStream_t);
int launch_fasta_pack_begin(uint64_t, uint32_t *, uint32_t *, void *, cudaStream_t);
int launch_fasta_pack_tiles(const uint8_t *, uint64_t, uint64_t, uint64_t, uint32_t *, uint32_t *, void *,
                            cudaStream_t, uint64_t layout_bytes = 0);
WidenHandle *pool_run_begin(void (*fn)(void *), void *arg);
void pool_run_end(WidenHandle *h);
unsigned widen_workers();
uint64_t fasta_tile_bytes();
uint64_t split_length(int k);
uint64_t split_scratch_bytes(int k);
int launch_split(const int64_t *, int, int64_t *, int64_t *, void *, cudaStream_t);
int launch_show_balance(const int64_t *, int, double *, unsigned long long *, cudaStream_t);
int launch_positive_pair(int64_t *, int64_t *, int, cudaStream_t);
uint64_t peer_inbox_bytes(int k, int counter_bits, int world);
int launch_reduce_push(const void *, int, int, int, int, void *const *, cudaStream_t);
int launch_reduce_collect(const void *, int, int, int, int, void *, cudaStream_t);
int launch_count_push(const uint32_t *, const uint32_t *, uint64_t, int, void *, int, int, int,
                      void *const *, cudaStream_t, int *);
uint64_t slice_inbox_bytes(int k, int world);
uint64_t slice_begin_host(int k, int o, int world);
int launch_slice_push(const void *, int, int, int, int, void *const *, unsigned long long, int, unsigned int *, cudaStream_t);
int launch_slice_signal(int, int, int, void *const *, unsigned long long, int, cudaStream_t);
int launch_slice_collect(int, int, int, void *const *, unsigned long long, int, int64_t *, uint16_t *, uint8_t *,
                         unsigned int *, cudaStream_t);

static thread_local char t_error[512] = "";
std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_host_fasta{-1};       // -1: from the environment (KPAL_HOST_FASTA=1)
static std::atomic<int> g_fasta_split{0};         // 1: a large FASTA text is cut in two parts, the first counted while the second uploads (fasta_gpu_count; measured: no gain yet, so off)
static std::atomic<int> g_fasta_hybrid_share{0};  // percent of the text the host packs (0: adapted from call to call)
static std::atomic<int> g_fasta_hybrid{1};        // 1: idle host threads pack segments from the end of a large text while its head uploads raw (fasta_hybrid_count)
static std::atomic<uint64_t> g_last_h2d_bytes{0}, g_last_host_text_bytes{0};   // kpal_last_upload
static std::atomic<int> g_fasta_chunks{0};      // 0 = automatic (one chunk per ~6 MB, at most 16), else 1 .. 32
static std::atomic<int> g_dma_share{0};         // sixteenths of a narrow-copied profile the DMA engine moves as int64 (pinned destinations)
static std::atomic<int> g_gram{1};              // 1: euclidean / cosine matrices take the tensor-core Gram form when the counts allow it
static std::atomic<int> g_narrow_lists{1};      // 1: counts that do not fit the narrow form travel in a side list (0: all-or-nothing, as before)
static std::atomic<int> g_narrow_d2h{1};        // 1: large profiles leave the device as uint8 / uint16, 2: uint16 only (see finalize_to_host)

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
}

int sm_count()
{
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 148;
        cached = prop.multiProcessorCount;
        cached_dev = dev;
    }
    return cached;
}

// RAII device / pinned buffers for the host-level entry points
struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes)
    {
        cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
        if (e != cudaSuccess) {
            p = nullptr;
            set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
            cudaGetLastError();
            return e == cudaErrorMemoryAllocation ? KPAL_ENOMEM : KPAL_ECUDA;
        }
        return KPAL_OK;
    }
    template <typename T> T *as() const { return static_cast<T *>(p); }
};
struct PinBuf {
    void *p = nullptr;
    ~PinBuf() { if (p) cudaFreeHost(p); }
    int alloc(size_t bytes)
    {
        cudaError_t e = cudaMallocHost(&p, bytes ? bytes : 1);
        if (e != cudaSuccess) {
            p = nullptr;
            set_error("cudaMallocHost(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
            cudaGetLastError();
            return e == cudaErrorMemoryAllocation ? KPAL_ENOMEM : KPAL_ECUDA;
        }
        return KPAL_OK;
    }
    template <typename T> T *as() const { return static_cast<T *>(p); }
};

// Grow-only per-device workspace of the host-level counting entry points, so
// that repeated calls (kpal count over many files, the bench's end-to-end loop)
// do not pay cudaMalloc / cudaMallocHost every time.  One call at a time.
struct GrowDev {
    void *p = nullptr; size_t cap = 0;
    unsigned char *as_bytes() const { return static_cast<unsigned char *>(p); }
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return KPAL_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {
            p = nullptr; cudaGetLastError();
            set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
            return e == cudaErrorMemoryAllocation ? KPAL_ENOMEM : KPAL_ECUDA;
        }
        cap = bytes;
        return KPAL_OK;
    }
};
struct GrowPin {
    void *p = nullptr; size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return KPAL_OK;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMallocHost(&p, bytes);
        if (e != cudaSuccess) {
            p = nullptr; cudaGetLastError();
            set_error("cudaMallocHost(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
            return e == cudaErrorMemoryAllocation ? KPAL_ENOMEM : KPAL_ECUDA;
        }
        cap = bytes;
        return KPAL_OK;
    }
};
struct CountWorkspace {
    int device = -1;
    GrowDev codes, valid, table, counts, text, fscratch, counts16, counts8, overflow, rows16[2], wide_flag;
    GrowDev list8, list16;                       // side lists of the narrow profile copy: (index, value) of the large counts
    GrowDev br_codes, br_valid, br_starts;       // by-record: the packed chunks and record starts of the call
    // hybrid upload (fasta_hybrid_count): the host's share of the text, adapted from call to call on this device
    double hybrid_share = 0.45;
    cudaEvent_t hybrid_ev[2] = {};               // raw upload: start / done (timing events)
    double hybrid_prev[3] = {0, 0, 0};           // previous call: packers' seconds, host text bytes, raw bytes
    GrowPin pcodes, pvalid, pstatus, pnarrow, pflag, prows16[2], plist;
    cudaEvent_t rows_done[2] = {};               // by-record: a batch of uint16 rows has landed
    cudaStream_t copy_stream = nullptr;          // H2D of the FASTA text, chunk by chunk
    cudaStream_t count_stream = nullptr;         // count kernels of the first of two parts (fasta_gpu_count)
    cudaEvent_t part_packed = nullptr, part_counted = nullptr;
    cudaEvent_t chunk_done[32] = {};
    cudaEvent_t d2h_done[16] = {};               // chunks of the narrow D2H of the profile
    cudaEvent_t flag_done = nullptr;             // ... and the flag words ahead of them
};
static std::mutex g_count_mutex;                 // held for the whole host-level call
static std::vector<CountWorkspace *> g_count_ws;

static int get_count_ws(CountWorkspace **out)
{
    int dev = 0;
    KPAL_CUDA(cudaGetDevice(&dev));
    for (auto *w : g_count_ws) if (w->device == dev) { *out = w; return KPAL_OK; }
    CountWorkspace *w = new CountWorkspace();
    w->device = dev;
    g_count_ws.push_back(w);
    *out = w;
    return KPAL_OK;
}

static int require_device()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (%s); libkpal_b200 has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return KPAL_ECUDA;
    }
    return KPAL_OK;
}

// Scratch of the dev-level distance entry point (acc / cnt matrices), cached per (device,
// stream) and grown on demand.  Calls on one stream are ordered by the stream, calls on
// different streams (other host threads) get buffers of their own, so no two calls in
// flight ever share an accumulator; a buffer is only replaced after its stream has drained.
// The host-level entry points (matrix sessions, pair distances) own their accumulators.
struct DistScratch {
    int device = -1;
    cudaStream_t stream = nullptr;
    uint64_t n = 0;
    double *acc = nullptr;
    uint32_t *cnt = nullptr;
};
static std::mutex g_scratch_mutex;
static std::vector<DistScratch> g_scratch;

static int get_dist_scratch(uint64_t n, cudaStream_t stream, double **acc, uint32_t **cnt)
{
    int dev = 0;
    KPAL_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_scratch_mutex);
    DistScratch *s = nullptr;
    for (auto &x : g_scratch) if (x.device == dev && x.stream == stream) s = &x;
    if (!s) { g_scratch.push_back(DistScratch()); s = &g_scratch.back(); s->device = dev; s->stream = stream; }
    if (s->n < n) {
        if (s->acc || s->cnt) KPAL_CUDA(cudaStreamSynchronize(stream));     // earlier calls may still use them
        if (s->acc) cudaFree(s->acc);
        if (s->cnt) cudaFree(s->cnt);
        s->acc = nullptr; s->cnt = nullptr; s->n = 0;
        KPAL_CUDA(cudaMalloc(&s->acc, n * n * sizeof(double)));
        KPAL_CUDA(cudaMalloc(&s->cnt, n * n * sizeof(uint32_t)));
        s->n = n;
    }
    *acc = s->acc; *cnt = s->cnt;
    return KPAL_OK;
}

// order of the profiles by total: ascending (the scaled profile A has the
// smaller total), descending with `down` (A has the larger total).  Stable, so
// ties keep input order (either role gives the same value for equal totals).
static int make_order(const double *d_totals, uint64_t n, int down, int32_t *d_order,
                      cudaStream_t stream)
{
    std::vector<double> tot(n);
    KPAL_CUDA(cudaMemcpyAsync(tot.data(), d_totals, n * sizeof(double), cudaMemcpyDeviceToHost, stream));
    KPAL_CUDA(cudaStreamSynchronize(stream));
    std::vector<int32_t> ord(n);
    std::iota(ord.begin(), ord.end(), 0);
    // NaN-free: totals are finite non-negative
    if (down) std::stable_sort(ord.begin(), ord.end(), [&](int32_t x, int32_t y) { return tot[x] > tot[y]; });
    else std::stable_sort(ord.begin(), ord.end(), [&](int32_t x, int32_t y) { return tot[x] < tot[y]; });
    KPAL_CUDA(cudaMemcpyAsync(d_order, ord.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
    KPAL_CUDA(cudaStreamSynchronize(stream));
    return KPAL_OK;
}

}  // namespace kpal

using namespace kpal;

// ------------------------------------------------------------------ misc
extern "C" int kpal_abi_version(void) { return KPAL_B200_ABI_VERSION; }
extern "C" const char *kpal_last_error(void) { return t_error; }

extern "C" int kpal_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int kpal_set_device(int device)
{
    KPAL_CHECK(require_device());
    KPAL_CUDA(cudaSetDevice(device));
    return KPAL_OK;
}

extern "C" int kpal_get_device(void)
{
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return -1; }
    return dev;
}

extern "C" void *kpal_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        set_error("cudaMallocHost(%zu) failed", bytes);
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
extern "C" void kpal_host_free(void *p) { if (p) cudaFreeHost(p); }

extern "C" void *kpal_dev_alloc(size_t bytes)
{
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
extern "C" void kpal_dev_free(void *p) { if (p) cudaFree(p); }

extern "C" int kpal_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream)
{
    KPAL_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return KPAL_OK;
}
extern "C" int kpal_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream)
{
    KPAL_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return KPAL_OK;
}
extern "C" int kpal_dev_memset(void *dst, int value, size_t bytes, void *stream)
{
    KPAL_CUDA(cudaMemsetAsync(dst, value, bytes, (cudaStream_t)stream));
    return KPAL_OK;
}
extern "C" int kpal_stream_sync(void *stream)
{
    KPAL_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return KPAL_OK;
}

extern "C" int kpal_set_option(const char *name, int value)
{
    if (!name) return bad_arg("null option name");
    if (!strcmp(name, "host_fasta")) { g_host_fasta.store(value ? 1 : 0); return KPAL_OK; }
    if (!strcmp(name, "fasta_chunks")) {
        if (value < 0 || value > 32) return bad_arg("fasta_chunks must be 0 (auto) .. 32");
        g_fasta_chunks.store(value); return KPAL_OK;
    }
    if (!strcmp(name, "fasta_split")) { g_fasta_split.store(value ? 1 : 0); return KPAL_OK; }
    if (!strcmp(name, "fasta_hybrid")) {
        if (value < 0 || value > 64) return bad_arg("fasta_hybrid must be 0 (off), 1 (on) or 2 .. 64 (on, that many host packers at most)");
        g_fasta_hybrid.store(value); return KPAL_OK;
    }
    if (!strcmp(name, "fasta_hybrid_share")) {
        if (value < 0 || value > 95) return bad_arg("fasta_hybrid_share must be 0 (adaptive) .. 95 percent of the text");
        g_fasta_hybrid_share.store(value); return KPAL_OK;
    }
    if (!strcmp(name, "exact_div")) { set_exact_div(value != 0); return KPAL_OK; }
    if (!strcmp(name, "gram")) { g_gram.store(value ? 1 : 0); return KPAL_OK; }
    if (!strcmp(name, "narrow_lists")) { g_narrow_lists.store(value ? 1 : 0); return KPAL_OK; }
    if (!strcmp(name, "narrow_d2h")) {
        if (value < 0 || value > 2) return bad_arg("narrow_d2h must be 0 (int64), 1 (uint8 / uint16) or 2 (uint16)");
        g_narrow_d2h.store(value); return KPAL_OK;
    }
    if (!strcmp(name, "dma_share")) {
        if (value < 0 || value > 8) return bad_arg("dma_share must be 0 .. 8 (sixteenths of the profile)");
        g_dma_share.store(value); return KPAL_OK;
    }
    if (!strcmp(name, "count_path")) {
        if (value < 0 || value > 3) return bad_arg("count_path must be 0 (auto), 1 (RED), 2 (radix) or 3 (one-window radix)");
        set_count_path(value); return KPAL_OK;
    }
    if (!strcmp(name, "tiled_finalize")) { set_tiled_finalize(value != 0); return KPAL_OK; }
    if (!strcmp(name, "by_record_path")) {
        if (value < 0 || value > 1) return bad_arg("by_record_path must be 0 (shared-memory slabs) or 1 (RED rows)");
        set_by_record_path(value); return KPAL_OK;
    }
    if (!strcmp(name, "radix_shape")) {
        if (value < 0 || value > 2) return bad_arg("radix_shape must be 0 (auto), 1 (1024 x 1 CTA/SM) or 2 (512 x 2)");
        set_radix_shape(value); return KPAL_OK;
    }
    if (!strcmp(name, "radix_max_buckets")) {
        if (value != 1024 && value != 2048) return bad_arg("radix_max_buckets must be 1024 or 2048");
        set_radix_max_buckets(value); return KPAL_OK;
    }
    if (!strcmp(name, "pair_flush_every")) {
        if (value < 0 || value > 6) return bad_arg("pair_flush_every must be 0 (auto) .. 6 tiles");
        set_pair_flush_every(value); return KPAL_OK;
    }
    if (!strcmp(name, "pair_fused")) { set_pair_fused(value); return KPAL_OK; }
    if (!strcmp(name, "radix_debug")) { set_radix_debug(value); return KPAL_OK; }   // timing experiments
    if (!strcmp(name, "radix_payload_bits")) {
        if (value < 0 || value > 15) return bad_arg("radix_payload_bits must be 0 (auto) .. 15");
        set_radix_payload_bits(value); return KPAL_OK;
    }
    set_error("unknown option '%s'", name);
    return KPAL_EINVAL;
}

extern "C" void kpal_last_upload(uint64_t *h2d_bytes, uint64_t *host_packed_text_bytes)
{
    if (h2d_bytes) *h2d_bytes = g_last_h2d_bytes.load();
    if (host_packed_text_bytes) *host_packed_text_bytes = g_last_host_text_bytes.load();
}

extern "C" uint64_t kpal_fasta_scratch_bytes(uint64_t n_bytes) { return fasta_scratch_bytes(n_bytes); }

extern "C" int kpal_dev_fasta_pack(const void *d_text, uint64_t n_bytes, uint32_t *d_codes,
                                   uint32_t *d_valid, void *d_scratch, void *stream)
{
    if ((!d_text && n_bytes) || !d_codes || !d_valid || !d_scratch) return bad_arg("null device pointer");
    return launch_fasta_pack(static_cast<const uint8_t *>(d_text), n_bytes, d_codes, d_valid, d_scratch,
                             (cudaStream_t)stream);
}

extern "C" uint64_t kpal_kernel_launches(void) { return g_launches.load(); }
extern "C" void kpal_reset_kernel_launches(void) { g_launches.store(0); }

// ---------------------------------------------------- counting: device API
extern "C" int kpal_dev_count_packed(const uint32_t *d_codes, const uint32_t *d_valid,
                                     uint64_t n_bases, int k, void *d_table, int counter_bits,
                                     void *stream)
{
    if (!d_table || ((!d_codes || !d_valid) && n_bases)) return bad_arg("null device pointer");
    return launch_count(d_codes, d_valid, n_bases, k, d_table, counter_bits, (cudaStream_t)stream);
}

extern "C" int kpal_dev_count_packed_fresh(const uint32_t *d_codes, const uint32_t *d_valid,
                                           uint64_t n_bases, int k, void *d_table, int counter_bits,
                                           void *stream)
{
    if (!d_table || ((!d_codes || !d_valid) && n_bases)) return bad_arg("null device pointer");
    return launch_count(d_codes, d_valid, n_bases, k, d_table, counter_bits, (cudaStream_t)stream, true);
}

extern "C" int kpal_dev_finalize_counts(const void *d_table, int counter_bits, int k, int balance,
                                        int64_t *d_counts, void *stream)
{
    if (!d_table || !d_counts) return bad_arg("null device pointer");
    return launch_finalize(d_table, counter_bits, k, balance, d_counts, (cudaStream_t)stream);
}

extern "C" int kpal_dev_balance(const int64_t *d_in, int64_t *d_out, int k, void *stream)
{
    if (!d_in || !d_out) return bad_arg("null device pointer");
    return launch_balance(d_in, d_out, k, (cudaStream_t)stream);
}

extern "C" int kpal_dev_count_by_record(const uint32_t *d_codes, const uint32_t *d_valid,
                                        const uint64_t *d_rec_starts, uint64_t first, uint64_t n,
                                        int k, int balance, int64_t *d_rows, void *stream)
{
    if (n && (!d_codes || !d_valid || !d_rec_starts || !d_rows)) return bad_arg("null device pointer");
    return launch_by_record(d_codes, d_valid, d_rec_starts, first, n, k, balance, d_rows,
                            (cudaStream_t)stream);
}

// ------------------------------------------------------ counting: host API
// packed host stream (pinned workspace) -> device table (accumulated)
static int upload_and_count(CountWorkspace *w, uint64_t n_bases, int k, void *d_table, int bits,
                            cudaStream_t st)
{
    uint64_t cw, vw;
    kpal_packed_words(n_bases, &cw, &vw);
    KPAL_CHECK(w->codes.ensure(cw * 4));
    KPAL_CHECK(w->valid.ensure(vw * 4));
    KPAL_CUDA(cudaMemcpyAsync(w->codes.p, w->pcodes.p, cw * 4, cudaMemcpyHostToDevice, st));
    KPAL_CUDA(cudaMemcpyAsync(w->valid.p, w->pvalid.p, vw * 4, cudaMemcpyHostToDevice, st));
    return launch_count(static_cast<uint32_t *>(w->codes.p), static_cast<uint32_t *>(w->valid.p),
                        n_bases, k, d_table, bits, st);
}

// Device scalars written by the GPU FASTA packer (fasta.cu: FastaScratch).
struct FastaStatus {
    unsigned long long first_header, total_bases;
    unsigned int flags, pad;
};

// KPAL_TRACE=1: host timestamps (us since the call began) of the phases of the host-level
// count, printed to stderr at the end of the call.  A measuring aid, off by default.
struct CallTrace {
    bool on = false;
    std::chrono::steady_clock::time_point t0;
    char line[512]; size_t at = 0;
    void begin()
    {
        static const bool enabled = [] { const char *e = getenv("KPAL_TRACE"); return e && e[0] == '1'; }();
        on = enabled; at = 0;
        if (on) t0 = std::chrono::steady_clock::now();
    }
    void mark(const char *what)
    {
        if (!on || at + 48 > sizeof line) return;
        const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
        at += size_t(snprintf(line + at, sizeof line - at, " %s=%.0f", what, us));
    }
    // device side: timing events recorded on the streams of the call, printed relative to the first
    cudaEvent_t dev[8] = {}; const char *dev_name[8] = {}; int n_dev = 0;
    void dev_mark(const char *what, cudaStream_t st)
    {
        if (!on || n_dev >= 8) return;
        if (!dev[n_dev] && cudaEventCreate(&dev[n_dev]) != cudaSuccess) return;
        dev_name[n_dev] = what;
        cudaEventRecord(dev[n_dev++], st);
    }
    void end()
    {
        if (!on) return;
        for (int i = 1; i < n_dev && at + 48 <= sizeof line; ++i) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, dev[0], dev[i]) == cudaSuccess)
                at += size_t(snprintf(line + at, sizeof line - at, " dev:%s=%.0f", dev_name[i], ms * 1e3));
        }
        cudaGetLastError();
        n_dev = 0;
        fprintf(stderr, "[kpal trace us]%s\n", line);
    }
};
static CallTrace g_trace;       // used under g_count_mutex

// The narrow forms of `split` counts (w->counts8 / w->counts16, flags in w->overflow, an int64
// tail of `tail` counts in w->counts) -> counts_out on the host: flags first, the first uint8
// chunk speculatively behind them, then the chunks of the narrowest form that holds every
// count, widened by the host workers while the next chunk is in flight.  *done = false when a
// count exceeds 65535 (nothing usable was written: the caller copies int64) or when the FASTA
// packer turned the text down (*text_flags).
struct NarrowListEntry { unsigned long long index, value; };

// counts_out[index] = value for the n listed bins (a few threads when the list is long)
static void apply_narrow_list(const NarrowListEntry *list, uint64_t n, int64_t *counts_out)
{
    auto run = [&](uint64_t a, uint64_t b) { for (uint64_t i = a; i < b; ++i) counts_out[list[i].index] = int64_t(list[i].value); };
    if (n < 16384) { run(0, n); return; }
    const unsigned n_threads = 8;
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < n_threads; ++t) pool.emplace_back(run, n * t / n_threads, n * (t + 1) / n_threads);
    for (auto &t : pool) t.join();
}

static int narrow_copy_out(CountWorkspace *w, uint64_t split, uint64_t tail, uint64_t piece, bool try8,
                       int64_t *counts_out, cudaStream_t st, const FastaStatus *pending, unsigned *text_flags,
                       bool *done, unsigned int cap8 = 0, unsigned int cap16 = 0)
{
    *done = false;
    KPAL_CUDA(cudaMemcpyAsync(w->pflag.p, w->overflow.p, 16, cudaMemcpyDeviceToHost, st));
    KPAL_CUDA(cudaEventRecord(w->flag_done, st));
    // chunks of whole pieces, >= 1 MiB each (the last one may be shorter)
    auto chunk_of = [&](int width) {
        uint64_t per = 1;
        while (per * piece * uint64_t(width) < (1ull << 20) && per < 16) per *= 2;
        return per * piece;
    };
    auto copy_chunk = [&](int width, uint64_t chunk, uint64_t c) -> cudaError_t {
        const unsigned char *src = static_cast<const unsigned char *>(width == 1 ? w->counts8.p : w->counts16.p);
        const uint64_t at = c * chunk, len = std::min(chunk, split - at);
        cudaError_t e = cudaMemcpyAsync(static_cast<unsigned char *>(w->pnarrow.p) + at * width, src + at * width,
                                        len * width, cudaMemcpyDeviceToHost, st);
        return e != cudaSuccess ? e : cudaEventRecord(w->d2h_done[c], st);
    };
    int width = try8 ? 1 : 2;
    uint64_t chunk = chunk_of(width);
    if (try8) KPAL_CUDA(copy_chunk(1, chunk, 0));               // speculative, behind the flags
    g_trace.mark("queued");
    KPAL_CUDA(cudaEventSynchronize(w->flag_done));
    g_trace.mark("flags");
    const volatile unsigned int *flag = static_cast<const volatile unsigned int *>(w->pflag.p);
    if (pending && (pending[0].flags | pending[1].flags)) {
        *text_flags = pending[0].flags | pending[1].flags;
        KPAL_CUDA(cudaStreamSynchronize(st));
        *done = true;                                               // (nothing to copy: the caller re-packs)
        return KPAL_OK;
    }
    // the narrowest form that holds every count, or holds all but a short list of them
    const unsigned int n8 = flag[2], n16 = flag[3];
    const bool fits8 = try8 && (flag[1] == 0 || (cap8 && n8 <= cap8));
    const bool fits16 = flag[0] == 0 || (cap16 && n16 <= cap16);
    if (fits8 || fits16) {
        uint64_t first = 1;                                     // chunk 0 is already on its way
        if (!fits8) { width = 2; chunk = chunk_of(2); first = 0; }
        if (!try8) first = 0;
        const uint64_t n_listed = fits8 ? (flag[1] ? n8 : 0) : (flag[0] ? n16 : 0);
        const uint64_t n_chunks = (split + chunk - 1) / chunk;
        for (uint64_t c = first; c < n_chunks; ++c) KPAL_CUDA(copy_chunk(width, chunk, c));
        if (n_listed) {
            KPAL_CHECK(w->plist.ensure(n_listed * sizeof(NarrowListEntry)));
            KPAL_CUDA(cudaMemcpyAsync(w->plist.p, fits8 ? w->list8.p : w->list16.p, n_listed * sizeof(NarrowListEntry),
                                      cudaMemcpyDeviceToHost, st));
        }
        if (tail) KPAL_CUDA(cudaMemcpyAsync(counts_out + split, w->counts.p, tail * 8, cudaMemcpyDeviceToHost, st));
        // from here on the workers are awake: every exit goes through widen_end
        WidenHandle *h = widen_begin(w->pnarrow.p, width, counts_out, split, chunk);
        cudaError_t err = cudaSuccess;
        for (uint64_t c = 0; c < n_chunks; ++c) {
            err = cudaEventSynchronize(w->d2h_done[c]);
            if (err != cudaSuccess) break;
            widen_publish(h, std::min((c + 1) * chunk, split));
        }
        g_trace.mark(width == 1 ? "d2h_u8" : "d2h_u16");
        widen_end(h, err != cudaSuccess ? 1 : 0);
        g_trace.mark("widened");
        if (err == cudaSuccess && (tail || n_listed)) err = cudaStreamSynchronize(st);
        if (err == cudaSuccess && n_listed)
            apply_narrow_list(static_cast<const NarrowListEntry *>(w->plist.p), n_listed, counts_out);
        g_trace.mark("tail");
        if (err != cudaSuccess) {
            set_error("narrow D2H of the profile failed: %s", cudaGetErrorString(err));
            cudaGetLastError();
            return KPAL_ECUDA;
        }
        *done = true;
        return KPAL_OK;
    }
    return KPAL_OK;
}

// Counter table on the device -> the caller's int64 profile on the host
// (widen + optional balance, then D2H).
//
// From 4^10 bins on, the profile leaves the device narrow: as uint8 when every count fits
// (an eighth of the PCIe bytes of the int64 array, the longest single piece of the
// host-level call: 134 MB at k = 12), else as uint16.  The finalize kernel writes both
// forms and two flag words saying which of them hold every count; the flags travel ahead
// of the data and the first uint8 chunk is copied speculatively behind them, so the
// decision costs no bubble on the copy engine.  The copy runs in up to 16 chunks into
// pinned staging; host workers (widen.cpp) widen chunk c into `counts_out` while chunk c+1
// is in flight, so the caller's array -- pageable or pinned -- is written exactly once, by
// the CPU.  A count above 65535 (a large table needs a very repetitive input for that)
// sends the call down the int64 finalize + copy instead.  Exact either way.
//
// `pending` (optional): the two status slots of the GPU FASTA packer (one per part of the
// text, fasta_gpu_count) whose D2H copies are already queued on `st`.  They are looked at together with the flags; if the packer turned the
// text down (*text_flags != 0 on return) nothing is written and the caller redoes the
// file through the host packer.
static int finalize_to_host(CountWorkspace *w, const void *d_table, int bits, int k, int balance,
                            int64_t *counts_out, cudaStream_t st, const FastaStatus *pending = nullptr,
                            unsigned *text_flags = nullptr)
{
    const uint64_t bins = 1ull << (2 * k);
    const int narrow = g_narrow_d2h.load();
    if (text_flags) *text_flags = 0;
    if (narrow && bins >= (1ull << 20)) {
        const bool try8 = narrow == 1;
        // The host threads' store bandwidth, not PCIe, bounds the narrow copy (8 bytes written
        // per 1 or 2 received).  When the caller's array is pinned, the DMA engine therefore
        // takes the last `share`/16 of the profile as plain int64, written straight into the
        // array behind the narrow chunks while the host threads are still widening.
        uint64_t share = uint64_t(g_dma_share.load());
        if (share) {
            cudaPointerAttributes attr;
            if (cudaPointerGetAttributes(&attr, counts_out) != cudaSuccess || attr.type != cudaMemoryTypeHost) share = 0;
            cudaGetLastError();
        }
        const uint64_t piece = bins / 16;                           // >= 65536 elements
        const uint64_t split = piece * (16 - share), tail = bins - split;
        KPAL_CHECK(w->counts16.ensure(split * 2));
        if (try8) KPAL_CHECK(w->counts8.ensure(split));
        if (tail) KPAL_CHECK(w->counts.ensure(tail * 8));
        KPAL_CHECK(w->overflow.ensure(16));
        KPAL_CHECK(w->pnarrow.ensure(split * 2));
        KPAL_CHECK(w->pflag.ensure(16));
        if (!w->d2h_done[0]) {
            for (auto &e : w->d2h_done) KPAL_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            KPAL_CUDA(cudaEventCreateWithFlags(&w->flag_done, cudaEventDisableTiming));
        }
        // side lists: up to 1/32 of the bins above 255 (uint8 form), 1/256 above 65535 (uint16 form)
        const bool lists = g_narrow_lists.load() != 0;
        const unsigned int cap8 = lists ? unsigned(std::max<uint64_t>(4096, bins / 32)) : 0u;
        const unsigned int cap16 = lists ? unsigned(std::max<uint64_t>(4096, bins / 256)) : 0u;
        KPAL_CHECK(w->list8.ensure(size_t(cap8) * sizeof(NarrowListEntry)));
        KPAL_CHECK(w->list16.ensure(size_t(cap16) * sizeof(NarrowListEntry)));
        KPAL_CUDA(cudaMemsetAsync(w->overflow.p, 0, 16, st));
        KPAL_CHECK(launch_finalize_narrow(d_table, bits, k, balance, static_cast<uint16_t *>(w->counts16.p),
                                          try8 ? static_cast<uint8_t *>(w->counts8.p) : nullptr,
                                          static_cast<int64_t *>(w->counts.p), split,
                                          static_cast<unsigned int *>(w->overflow.p), st, lists ? w->list8.p : nullptr,
                                          cap8, lists ? w->list16.p : nullptr, cap16));
        g_trace.dev_mark("finalized", st);
        bool done = false;
        KPAL_CHECK(narrow_copy_out(w, split, tail, piece, try8, counts_out, st, pending, text_flags, &done, cap8, cap16));
        if (done) return KPAL_OK;
        KPAL_CUDA(cudaStreamSynchronize(st));                       // drain the speculative chunk
    } else if (pending) {
        KPAL_CUDA(cudaStreamSynchronize(st));
        if (pending[0].flags | pending[1].flags) { *text_flags = pending[0].flags | pending[1].flags; return KPAL_OK; }
    }
    KPAL_CHECK(w->counts.ensure(bins * 8));
    KPAL_CHECK(launch_finalize(d_table, bits, k, balance, static_cast<int64_t *>(w->counts.p), st));
    KPAL_CUDA(cudaMemcpyAsync(counts_out, w->counts.p, bins * 8, cudaMemcpyDeviceToHost, st));
    KPAL_CUDA(cudaStreamSynchronize(st));
    g_trace.mark("d2h_i64");
    return KPAL_OK;
}

static int count_packed_to_host(CountWorkspace *w, uint64_t n_bases, int k, int balance,
                                int64_t *counts_out)
{
    const uint64_t bins = 1ull << (2 * k);
    const int bits = (n_bases >= (1ull << 32)) ? 64 : 32;
    KPAL_CHECK(w->table.ensure(bins * (bits / 8)));
    cudaStream_t st = 0;
    KPAL_CUDA(cudaMemsetAsync(w->table.p, 0, bins * (bits / 8), st));
    KPAL_CHECK(upload_and_count(w, n_bases, k, w->table.p, bits, st));
    return finalize_to_host(w, w->table.p, bits, k, balance, counts_out, st);
}

static int check_k_host(int k)
{
    if (k < 1 || k > KPAL_MAX_K) {
        set_error("k-mer length %d out of range [1, %d]", k, KPAL_MAX_K);
        return KPAL_EINVAL;
    }
    return KPAL_OK;
}

static int pack_sequences_ws(CountWorkspace *w, const char *text, const uint64_t *offsets,
                             uint64_t n_records, uint64_t *n_bases)
{
    KPAL_CHECK(kpal_pack_sequences(text, offsets, n_records, nullptr, nullptr, nullptr, n_bases));
    uint64_t cw, vw;
    kpal_packed_words(*n_bases, &cw, &vw);
    KPAL_CHECK(w->pcodes.ensure(cw * 4));
    KPAL_CHECK(w->pvalid.ensure(vw * 4));
    return kpal_pack_sequences(text, offsets, n_records, static_cast<uint32_t *>(w->pcodes.p),
                               static_cast<uint32_t *>(w->pvalid.p), nullptr, n_bases);
}

static int pack_fasta_ws(CountWorkspace *w, const char *fasta, uint64_t n_bytes, uint64_t *n_bases)
{
    uint64_t n_rec = 0, name_bytes = 0;
    KPAL_CHECK(kpal_fasta_scan(fasta, n_bytes, &n_rec, n_bases, &name_bytes));
    uint64_t cw, vw;
    kpal_packed_words(*n_bases, &cw, &vw);
    KPAL_CHECK(w->pcodes.ensure(cw * 4));
    KPAL_CHECK(w->pvalid.ensure(vw * 4));
    return kpal_fasta_pack(fasta, n_bytes, static_cast<uint32_t *>(w->pcodes.p),
                           static_cast<uint32_t *>(w->pvalid.p), nullptr, nullptr);
}

extern "C" int kpal_count_sequences(const char *text, const uint64_t *offsets, uint64_t n_records,
                                    int k, int balance, int64_t *counts_out)
{
    if (!counts_out || !offsets) return bad_arg("null pointer");
    KPAL_CHECK(check_k_host(k));
    KPAL_CHECK(require_device());
    std::lock_guard<std::mutex> lock(g_count_mutex);
    CountWorkspace *w;
    KPAL_CHECK(get_count_ws(&w));
    uint64_t n_bases = 0;
    KPAL_CHECK(pack_sequences_ws(w, text, offsets, n_records, &n_bases));
    return count_packed_to_host(w, n_bases, k, balance, counts_out);
}

// A header line ('>' at a line start) near three quarters of the text, or 0: where
// kpal_count_fasta cuts the file in two so that the first part is counted while the second
// is still on the bus.  Looks at 1 MB only ('>' does not occur inside sequence lines, so
// reads have one every few hundred bytes; a genome without a header there is not cut).
static uint64_t fasta_split_point(const char *fasta, uint64_t n_bytes)
{
    if (g_fasta_split.load() == 0 || n_bytes < (16ull << 20)) return 0;
    uint64_t at = n_bytes / 4 * 3;
    const uint64_t end = std::min<uint64_t>(n_bytes, at + (1ull << 20));
    while (at < end) {
        const char *hit = static_cast<const char *>(memchr(fasta + at, '>', end - at));
        if (!hit) return 0;
        at = uint64_t(hit - fasta);
        if (fasta[at - 1] == '\n') return at;
        ++at;
    }
    return 0;
}

// ---------------------------------------------------------------- hybrid upload
// The upload of the text is 60 % of a large kpal_count_fasta call and the host has nothing
// to do meanwhile.  So the text is cut at header lines into segments of ~512 KB: the head of
// the text goes to the device raw (chunked upload + device packer, as before) while the
// pool's threads pack the segments of the tail (pack.cpp: fasta_pack_segment, 32 bytes per
// step) into pinned staging, and only their 0.375 B/base cross the bus behind the raw text.
// The packed stream is SLOTTED: segment j owns the bases [slot_j, slot_{j+1}) with
// slot_j = align64(cut_j) + 64 j (a byte emits at most one base, so slot_j lies behind
// everything the text before cut_j can emit), unused slot ends are invalid, and one count
// launch takes the whole stream -- the count kernels already run over one position per text
// byte (the device packer's upper bound).  The host's share of the text follows the measured
// rates of the previous call (host packing vs. raw upload) so that the packers finish just
// before the raw text has gone up; exact whatever the share.
struct HybridState {
    const unsigned char *text = nullptr;
    uint64_t n_bytes = 0;
    int k = 0;
    SlottedPlan plan;
    uint32_t *pcodes = nullptr, *pvalid = nullptr;    // pinned staging, laid out like the device stream from slot[first] on
    uint64_t first = 0;                           // host segments [first, m)
    std::atomic<uint64_t> next{0};
    std::unique_ptr<std::atomic<unsigned char>[]> done;
    std::atomic<uint64_t> host_bases{0};
    std::atomic<unsigned> joined{0};
    unsigned max_packers = 0;
    std::atomic<uint64_t> busy_ns{0};             // summed over the packers that got a segment
    std::atomic<unsigned> active{0};
};

static void hybrid_packer(void *arg)
{
    HybridState *h = static_cast<HybridState *>(arg);
    if (h->joined.fetch_add(1) >= h->max_packers) return;
    const uint64_t base = h->plan.slot[h->first];
    const auto t0 = std::chrono::steady_clock::now();
    bool worked = false;
    for (;;) {
        const uint64_t j = h->next.fetch_add(1, std::memory_order_relaxed);
        if (j >= h->plan.m) break;
        worked = true;
        const uint64_t n = slotted_pack_segment(h->plan, h->text, h->n_bytes, j, h->first, h->k, h->pcodes, h->pvalid, base);
        h->host_bases.fetch_add(n, std::memory_order_relaxed);
        h->done[j].store(1, std::memory_order_release);
    }
    if (worked) {       // the packing rate proper (no wake-up latency): what the next call's share is computed from
        h->busy_ns.fetch_add(uint64_t(std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count()));
        h->active.fetch_add(1);
    }
}

static unsigned hybrid_packers()
{
    const int opt = g_fasta_hybrid.load();
    if (opt <= 0 || !fasta_segment_fast()) return 0;
    unsigned n = widen_workers() - 1;               // the caller feeds the copy engine
    if (opt > 1) n = std::min<unsigned>(n, unsigned(opt));
    else {
        // One process per GPU (torchrun): the ranks of a node share its cores.  With few cores per
        // rank the packers take from the copy engines what they save them (measured, 8 ranks on 32
        // vCPUs: 6.39 ms per call with 4 packers each, 6.05 ms without; 2 ranks on 24 vCPUs: 2.97
        // vs 3.37 ms), so below 12 cores per rank -- the smallest share measured to pay -- the text
        // goes up raw.
        const char *e = getenv("LOCAL_WORLD_SIZE");
        const unsigned ranks = e ? unsigned(std::max(1, atoi(e))) : 1u;
        const long online = sysconf(_SC_NPROCESSORS_ONLN);      // (not the affinity mask: a rank may be bound to its GPU's node)
        const unsigned hw = online > 0 ? unsigned(online) : std::thread::hardware_concurrency();
        if (ranks > 1 && hw) {
            if (hw / ranks < 12) return 0;
            n = std::min(n, hw / ranks);
        }
    }
    return n;
}

// The host's share of the text for the next call on this device, from the rates of the previous one.
static void hybrid_adapt(CountWorkspace *w)
{
    if (w->hybrid_prev[0] <= 0 || w->hybrid_prev[1] <= 0 || w->hybrid_prev[2] <= 0 || !w->hybrid_ev[0]) return;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, w->hybrid_ev[0], w->hybrid_ev[1]) != cudaSuccess || ms <= 0.f) { cudaGetLastError(); return; }
    const double host_rate = w->hybrid_prev[1] / w->hybrid_prev[0], bus_rate = w->hybrid_prev[2] / (ms * 1e-3);
    // The bus carries (1 - x) of the text raw and x of it packed (0.375 B/base, a little more
    // with the slot ends and the shorter copies: 0.41); the packers should be done a little
    // before the bus is:  x / host_rate = 0.93 ((1 - x) + 0.41 x) / bus_rate
    const double x = 0.93 * host_rate / (bus_rate + 0.93 * 0.59 * host_rate);
    w->hybrid_share = std::min(0.75, std::max(0.05, 0.3 * w->hybrid_share + 0.7 * x));
    w->hybrid_prev[0] = 0;
}

// Returns KPAL_OK with *taken = false when the text is not one for the hybrid form.
static int fasta_hybrid_count(CountWorkspace *w, const char *fasta, uint64_t n_bytes, int k,
                              void *d_table, int bits, cudaStream_t st, unsigned *flags,
                              uint64_t *n_bases, bool *taken)
{
    *taken = false;
    const unsigned packers = hybrid_packers();
    if (packers < 2 || n_bytes < (32ull << 20) || n_bytes >= (1ull << 32)) return KPAL_OK;
    HybridState h;
    if (!slotted_plan(fasta, n_bytes, 512u << 10, h.plan) || h.plan.m < 8) return KPAL_OK;
    *taken = true;
    hybrid_adapt(w);
    const int forced = g_fasta_hybrid_share.load();
    const double share = forced > 0 ? forced / 100.0 : w->hybrid_share;
    const uint64_t m = h.plan.m;
    const uint64_t first = std::min<uint64_t>(m - 1, std::max<uint64_t>(1, uint64_t(double(m) * (1.0 - share) + 0.5)));
    h.text = reinterpret_cast<const unsigned char *>(fasta);
    h.n_bytes = n_bytes;
    h.k = k;
    h.first = first;
    h.next.store(first);
    const uint64_t stream_bases = h.plan.stream_bases(first), host_base = h.plan.slot[first], raw_len = h.plan.cut[first];
    uint64_t cw = 0, vw = 0, cw_max = 0, vw_max = 0;
    kpal_packed_words(stream_bases, &cw, &vw);
    kpal_packed_words(h.plan.stream_bases(0), &cw_max, &vw_max);
    // (sized for any share: the share moves from call to call and a pinned reallocation costs milliseconds)
    KPAL_CHECK(w->text.ensure(n_bytes + 32));
    KPAL_CHECK(w->codes.ensure(cw_max * 4));
    KPAL_CHECK(w->valid.ensure(vw_max * 4));
    KPAL_CHECK(w->pcodes.ensure(cw_max * 4));
    KPAL_CHECK(w->pvalid.ensure(vw_max * 4));
    KPAL_CHECK(w->fscratch.ensure(fasta_scratch_bytes(n_bytes)));
    KPAL_CHECK(w->pstatus.ensure(2 * sizeof(FastaStatus)));
    FastaStatus *status = static_cast<FastaStatus *>(w->pstatus.p);
    memset(status, 0, 2 * sizeof(FastaStatus));
    if (!w->copy_stream) {
        KPAL_CUDA(cudaStreamCreateWithFlags(&w->copy_stream, cudaStreamNonBlocking));
        for (auto &e : w->chunk_done) KPAL_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    if (!w->hybrid_ev[0]) for (auto &e : w->hybrid_ev) KPAL_CUDA(cudaEventCreate(&e));
    h.pcodes = static_cast<uint32_t *>(w->pcodes.p);
    h.pvalid = static_cast<uint32_t *>(w->pvalid.p);
    h.done.reset(new std::atomic<unsigned char>[m]);
    for (uint64_t j = 0; j < m; ++j) h.done[j].store(0, std::memory_order_relaxed);
    h.max_packers = packers;
    uint8_t *d_text = w->text.as_bytes();
    uint32_t *d_codes = static_cast<uint32_t *>(w->codes.p), *d_valid = static_cast<uint32_t *>(w->valid.p);
    void *d_scratch = w->fscratch.p;
    const uint64_t tile = fasta_tile_bytes();

    const auto t_host0 = std::chrono::steady_clock::now();
    WidenHandle *pool = pool_run_begin(hybrid_packer, &h);           // the packers start on the tail
    int rc = KPAL_OK;
    cudaError_t err = cudaSuccess;
    auto fail = [&](cudaError_t e) { if (err == cudaSuccess) err = e; };
    // zero the stream up to the host's part (the device packer ORs into it; the slots of the
    // host's part are written whole by the copies) and the halo behind the stream
    fail(cudaMemsetAsync(d_codes, 0, host_base / 4, st));
    fail(cudaMemsetAsync(d_valid, 0, host_base / 8, st));
    fail(cudaMemsetAsync(d_codes + stream_bases / 16, 0, (cw - stream_bases / 16) * 4, st));
    fail(cudaMemsetAsync(d_valid + stream_bases / 32, 0, (vw - stream_bases / 32) * 4, st));
    fail(cudaMemsetAsync(d_scratch, 0, sizeof(FastaStatus) + 16, st));
    fail(cudaMemsetAsync(d_scratch, 0xff, sizeof(unsigned long long), st));      // first_header = ~0
    g_trace.dev_mark("start", w->copy_stream);
    fail(cudaEventRecord(w->hybrid_ev[0], w->copy_stream));
    // the head: raw chunks on the copy stream, each packed as soon as it has landed
    uint64_t n_chunks = std::min<uint64_t>(std::max<uint64_t>(raw_len / (6ull << 20), 1), 16);
    if (g_fasta_chunks.load() > 0) n_chunks = std::min<uint64_t>(uint64_t(g_fasta_chunks.load()), 30);
    const uint64_t chunk = ((raw_len + n_chunks - 1) / n_chunks + tile - 1) / tile * tile;
    uint64_t event = 0;
    for (uint64_t off = 0; off < raw_len && err == cudaSuccess && rc == KPAL_OK; off += chunk) {
        const uint64_t len = std::min(chunk, raw_len - off);
        fail(cudaMemcpyAsync(d_text + off, fasta + off, len, cudaMemcpyHostToDevice, w->copy_stream));
        fail(cudaEventRecord(w->chunk_done[event], w->copy_stream));
        fail(cudaStreamWaitEvent(st, w->chunk_done[event], 0));
        ++event;
        if (err == cudaSuccess)
            rc = launch_fasta_pack_tiles(d_text, raw_len, off / tile, (off + len + tile - 1) / tile, d_codes, d_valid, d_scratch, st);
    }
    fail(cudaEventRecord(w->hybrid_ev[1], w->copy_stream));
    g_trace.dev_mark("raw_h2d", w->copy_stream);
    if (err == cudaSuccess && rc == KPAL_OK)
        fail(cudaMemcpyAsync(status, d_scratch, sizeof(FastaStatus), cudaMemcpyDeviceToHost, st));
    // The device's part is counted as soon as it is packed, i.e. while the packed tail is still on
    // the bus: it ends in >= 64 invalid bases, so no window of it reaches into the host's slots.
    if (err == cudaSuccess && rc == KPAL_OK) rc = launch_count(d_codes, d_valid, host_base, k, d_table, bits, st);
    g_trace.mark("raw_queued");
    // the tail: the finished segments go up in runs, in order, behind the raw text
    uint64_t up = first;
    double host_s = 0;
    while (up < m) {
        uint64_t r = up;
        while (r < m && h.done[r].load(std::memory_order_acquire)) ++r;
        if (r == m && host_s == 0) host_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_host0).count();
        if (r - up >= 16 || (r == m && r > up)) {
            // (the junction records lie behind the last slot and are complete with the last segment)
            const uint64_t b0 = h.plan.slot[up], b1 = r == m ? stream_bases : h.plan.slot[r];
            if (err == cudaSuccess) {
                fail(cudaMemcpyAsync(d_codes + b0 / 16, h.pcodes + (b0 - host_base) / 16, (b1 - b0) / 4, cudaMemcpyHostToDevice, w->copy_stream));
                fail(cudaMemcpyAsync(d_valid + b0 / 32, h.pvalid + (b0 - host_base) / 32, (b1 - b0) / 8, cudaMemcpyHostToDevice, w->copy_stream));
            }
            up = r;
        } else {
            std::this_thread::yield();
        }
    }
    pool_run_end(pool);
    g_trace.mark("host_packed");
    if (err != cudaSuccess) {
        set_error("hybrid FASTA upload failed: %s", cudaGetErrorString(err));
        cudaGetLastError();
        cudaStreamSynchronize(w->copy_stream);
        cudaStreamSynchronize(st);
        return KPAL_ECUDA;
    }
    if (rc != KPAL_OK) { cudaStreamSynchronize(w->copy_stream); cudaStreamSynchronize(st); return rc; }
    KPAL_CUDA(cudaEventRecord(w->chunk_done[31], w->copy_stream));
    g_trace.dev_mark("h2d", w->copy_stream);
    // the second count launch into the same table: the host's slots and the junction records
    KPAL_CUDA(cudaStreamWaitEvent(st, w->chunk_done[31], 0));
    g_trace.dev_mark("packed", st);
    KPAL_CHECK(launch_count(d_codes + host_base / 16, d_valid + host_base / 32, stream_bases - host_base, k, d_table, bits, st));
    g_trace.dev_mark("counted", st);
    g_trace.mark("count_queued");
    if (h.active.load()) host_s = double(h.busy_ns.load()) * 1e-9 / double(h.active.load());
    w->hybrid_prev[0] = host_s; w->hybrid_prev[1] = double(n_bytes - raw_len); w->hybrid_prev[2] = double(raw_len);
    g_last_h2d_bytes.store(raw_len + (stream_bases - host_base) / 8 * 3);
    g_last_host_text_bytes.store(n_bytes - raw_len);
    if (g_trace.on && g_trace.at + 48 <= sizeof g_trace.line)
        g_trace.at += size_t(snprintf(g_trace.line + g_trace.at, sizeof g_trace.line - g_trace.at, " host_share_pct=%.0f",
                                      100.0 * double(n_bytes - raw_len) / double(n_bytes)));
    if (!flags) return KPAL_OK;
    KPAL_CUDA(cudaStreamSynchronize(st));
    *flags = status[0].flags;
    if (n_bases) *n_bases = status[0].total_bases + h.host_bases.load();
    return KPAL_OK;
}

// Raw FASTA bytes (host) -> device text -> GPU scan/pack -> windows accumulated
// into d_table.  *flags gets bit 0 when the text holds bytes the GPU packer
// does not handle (tabs & co on sequence lines): the caller then redoes the
// file through the host packer.  Synchronises the stream -- unless `flags` is null:
// then the status copies are only queued and the caller reads the two FastaStatus
// slots of w->pstatus after its own synchronisation (finalize_to_host does, together
// with its flag words).
//
// The text goes up in chunks on a copy stream; the packer runs on the tiles of a chunk as
// soon as it has landed, so scan/pack hides behind the PCIe transfer.  With the option
// "fasta_split" a large text is cut at a header line into two parts that are packed and
// counted independently (a part that begins with a header is a FASTA file of its own: no
// window spans the cut): the first part's count kernels run while the second part is still
// being uploaded, so only the last quarter of the counting is left once the last byte has
// arrived.  Exact, but not faster yet (profiles/r01_e2e_trace.log: the count does leave
// the critical path, 290 -> 125 us after the last pack, but the scan/pack chain of the
// second part's chunks falls 150 us behind the first part's count kernels), hence off.
static int fasta_gpu_count(CountWorkspace *w, const char *fasta, uint64_t n_bytes, int k,
                           void *d_table, int bits, cudaStream_t st, unsigned *flags,
                           uint64_t *n_bases)
{
    if (g_fasta_split.load() == 0) {
        bool taken = false;
        const int rc = fasta_hybrid_count(w, fasta, n_bytes, k, d_table, bits, st, flags, n_bases, &taken);
        if (rc != KPAL_OK || taken) return rc;
    }
    g_last_h2d_bytes.store(n_bytes);
    g_last_host_text_bytes.store(0);
    const uint64_t split = fasta_split_point(fasta, n_bytes);
    const uint64_t part_len[2] = {split ? split : n_bytes, split ? n_bytes - split : 0};
    const uint64_t text_off[2] = {0, (part_len[0] + 255) / 256 * 256};      // 16-byte loads need alignment
    uint64_t cw[2] = {0, 0}, vw[2] = {0, 0};
    kpal_packed_words(part_len[0], &cw[0], &vw[0]);                          // capacity: one base per input byte
    if (split) kpal_packed_words(part_len[1], &cw[1], &vw[1]);
    KPAL_CHECK(w->text.ensure(text_off[1] + part_len[1] + 32));
    KPAL_CHECK(w->codes.ensure((cw[0] + cw[1]) * 4));
    KPAL_CHECK(w->valid.ensure((vw[0] + vw[1]) * 4));
    const uint64_t scratch_off[2] = {0, (fasta_scratch_bytes(part_len[0]) + 255) / 256 * 256};
    KPAL_CHECK(w->fscratch.ensure(scratch_off[1] + (split ? fasta_scratch_bytes(part_len[1]) : 0)));
    KPAL_CHECK(w->pstatus.ensure(2 * sizeof(FastaStatus)));
    FastaStatus *status = static_cast<FastaStatus *>(w->pstatus.p);
    memset(status, 0, 2 * sizeof(FastaStatus));
    const uint64_t tile = fasta_tile_bytes();
    // (what remains after the last byte has landed is the scan/pack of ONE chunk: 80 us with 16
    // chunks of the 109 MB of config 2, 45 us with 32 -- but every chunk costs ~5 us on the
    // bus, so 32 chunks end later than 16: profiles/r01_e2e_trace.log)
    uint64_t n_chunks = n_bytes / (6ull << 20);
    n_chunks = std::min<uint64_t>(std::max<uint64_t>(n_chunks, 1), 16);
    if (g_fasta_chunks.load() > 0) n_chunks = std::min<uint64_t>(uint64_t(g_fasta_chunks.load()), 32);
    if (split) n_chunks = std::max<uint64_t>(n_chunks, 2);
    const bool piped = n_chunks > 1;
    if (piped && !w->copy_stream) {
        KPAL_CUDA(cudaStreamCreateWithFlags(&w->copy_stream, cudaStreamNonBlocking));
        for (auto &e : w->chunk_done) KPAL_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    if (split && !w->count_stream) {
        KPAL_CUDA(cudaStreamCreateWithFlags(&w->count_stream, cudaStreamNonBlocking));
        KPAL_CUDA(cudaEventCreateWithFlags(&w->part_packed, cudaEventDisableTiming));
        KPAL_CUDA(cudaEventCreateWithFlags(&w->part_counted, cudaEventDisableTiming));
    }
    // the caller's work queued on `st` before this call (the table memset) comes first
    if (split) {
        KPAL_CUDA(cudaEventRecord(w->part_packed, st));
        KPAL_CUDA(cudaStreamWaitEvent(w->count_stream, w->part_packed, 0));
    }
    // Both parts are initialised before the first byte is queued: the memsets may run on the
    // copy engine, where they would wait behind every upload chunk already in its queue.
    for (int part = 0; part < (split ? 2 : 1); ++part)
        KPAL_CHECK(launch_fasta_pack_begin(part_len[part], static_cast<uint32_t *>(w->codes.p) + (part ? cw[0] : 0),
                                           static_cast<uint32_t *>(w->valid.p) + (part ? vw[0] : 0),
                                           w->fscratch.as_bytes() + scratch_off[part], st));
    g_trace.dev_mark("start", piped ? w->copy_stream : st);
    uint64_t event = 0;
    for (int part = 0; part < (split ? 2 : 1); ++part) {
        const char *src = fasta + (part ? split : 0);
        const uint64_t len_p = part_len[part];
        uint8_t *d_text = w->text.as_bytes() + text_off[part];
        uint32_t *d_codes = static_cast<uint32_t *>(w->codes.p) + (part ? cw[0] : 0);
        uint32_t *d_valid = static_cast<uint32_t *>(w->valid.p) + (part ? vw[0] : 0);
        void *d_scratch = w->fscratch.as_bytes() + scratch_off[part];
        // this part's share of the chunks, at least one
        uint64_t chunks_p = split ? std::max<uint64_t>(1, (n_chunks * len_p + n_bytes / 2) / n_bytes) : n_chunks;
        chunks_p = std::min<uint64_t>(chunks_p, 32 - event - (split && part == 0 ? 1 : 0));
        const uint64_t chunk = ((len_p + chunks_p - 1) / chunks_p + tile - 1) / tile * tile;
        for (uint64_t off = 0; off < len_p; off += chunk) {
            const uint64_t len = std::min(chunk, len_p - off);
            if (piped) {
                KPAL_CUDA(cudaMemcpyAsync(d_text + off, src + off, len, cudaMemcpyHostToDevice, w->copy_stream));
                KPAL_CUDA(cudaEventRecord(w->chunk_done[event], w->copy_stream));
                KPAL_CUDA(cudaStreamWaitEvent(st, w->chunk_done[event], 0));
                ++event;
            } else {
                KPAL_CUDA(cudaMemcpyAsync(d_text + off, src + off, len, cudaMemcpyHostToDevice, st));
            }
            KPAL_CHECK(launch_fasta_pack_tiles(d_text, len_p, off / tile, (off + len + tile - 1) / tile, d_codes,
                                               d_valid, d_scratch, st));
        }
        if (part == (split ? 1 : 0)) {
            g_trace.dev_mark("h2d", piped ? w->copy_stream : st);
            g_trace.dev_mark("packed", st);
        }
        KPAL_CUDA(cudaMemcpyAsync(status + part, d_scratch, sizeof(FastaStatus), cudaMemcpyDeviceToHost, st));
        // len_p is an upper bound of the packed length; the tail is all-invalid padding.
        // The first of two parts is counted on a stream of its own: queued on `st` its
        // count kernels would hold up the scan/pack of the second part's chunks, which then
        // finish long after the last byte has arrived (measured: 290 instead of 80 us).
        if (split && part == 0) {
            KPAL_CUDA(cudaEventRecord(w->part_packed, st));
            KPAL_CUDA(cudaStreamWaitEvent(w->count_stream, w->part_packed, 0));
            KPAL_CHECK(launch_count(d_codes, d_valid, len_p, k, d_table, bits, w->count_stream));
            KPAL_CUDA(cudaEventRecord(w->part_counted, w->count_stream));
        } else {
            // one table, one radix staging area: the second count follows the first
            if (split) KPAL_CUDA(cudaStreamWaitEvent(st, w->part_counted, 0));
            KPAL_CHECK(launch_count(d_codes, d_valid, len_p, k, d_table, bits, st));
        }
    }
    g_trace.dev_mark("counted", st);
    g_trace.mark("count_queued");
    if (!flags) return KPAL_OK;
    KPAL_CUDA(cudaStreamSynchronize(st));
    *flags = status[0].flags | status[1].flags;
    if (n_bases) *n_bases = status[0].total_bases + status[1].total_bases;
    return KPAL_OK;
}


static bool use_gpu_fasta()
{
    int v = g_host_fasta.load();
    if (v < 0) {
        const char *e = getenv("KPAL_HOST_FASTA");
        v = (e && e[0] == '1') ? 1 : 0;
        g_host_fasta.store(v);
    }
    return v == 0;
}

extern "C" int kpal_count_fasta(const char *fasta, uint64_t n_bytes, int k, int balance,
                                int64_t *counts_out)
{
    if (!counts_out || (!fasta && n_bytes)) return bad_arg("null pointer");
    KPAL_CHECK(check_k_host(k));
    KPAL_CHECK(require_device());
    std::lock_guard<std::mutex> lock(g_count_mutex);
    CountWorkspace *w;
    KPAL_CHECK(get_count_ws(&w));
    if (use_gpu_fasta() && n_bytes > 0) {
        const uint64_t bins = 1ull << (2 * k);
        const int bits = (n_bytes >= (1ull << 32)) ? 64 : 32;
        KPAL_CHECK(w->table.ensure(bins * (bits / 8)));
        cudaStream_t st = 0;
        KPAL_CUDA(cudaMemsetAsync(w->table.p, 0, bins * (bits / 8), st));
        g_trace.begin();
        KPAL_CHECK(fasta_gpu_count(w, fasta, n_bytes, k, w->table.p, bits, st, nullptr, nullptr));
        unsigned flags = 0;
        const int rc = finalize_to_host(w, w->table.p, bits, k, balance, counts_out, st,
                                        static_cast<const FastaStatus *>(w->pstatus.p), &flags);
        g_trace.end();
        if (rc != KPAL_OK || !flags) return rc;
        // exotic whitespace: fall through to the host packer (exact rstrip semantics)
    }
    uint64_t n_bases = 0;
    KPAL_CHECK(pack_fasta_ws(w, fasta, n_bytes, &n_bases));
    return count_packed_to_host(w, n_bases, k, balance, counts_out);
}

// Host FASTA bytes -> accumulate into a caller-owned DEVICE table (multi-GPU
// driver: every rank counts its shard, then the tables are reduced with NCCL).
extern "C" int kpal_count_fasta_to_dev(const char *fasta, uint64_t n_bytes, int k, void *d_table,
                                       int counter_bits, void *stream, uint64_t *n_bases_out)
{
    if (!d_table || (!fasta && n_bytes)) return bad_arg("null pointer");
    KPAL_CHECK(check_k_host(k));
    KPAL_CHECK(require_device());
    std::lock_guard<std::mutex> lock(g_count_mutex);
    CountWorkspace *w;
    KPAL_CHECK(get_count_ws(&w));
    uint64_t n_bases = 0;
    if (use_gpu_fasta() && n_bytes > 0) {
        // The GPU packer may turn out to be unusable for this text only after the
        // windows were accumulated, so count into a scratch table first.
        const uint64_t bins = 1ull << (2 * k);
        KPAL_CHECK(w->table.ensure(bins * (counter_bits / 8)));
        cudaStream_t st = (cudaStream_t)stream;
        KPAL_CUDA(cudaMemsetAsync(w->table.p, 0, bins * (counter_bits / 8), st));
        unsigned flags = 0;
        KPAL_CHECK(fasta_gpu_count(w, fasta, n_bytes, k, w->table.p, counter_bits, st, &flags, &n_bases));
        if (!flags) {
            KPAL_CHECK(launch_accumulate(w->table.p, d_table, counter_bits, bins, st));
            KPAL_CUDA(cudaStreamSynchronize(st));
            if (n_bases_out) *n_bases_out = n_bases;
            return KPAL_OK;
        }
    }
    KPAL_CHECK(pack_fasta_ws(w, fasta, n_bytes, &n_bases));
    if (n_bases_out) *n_bases_out = n_bases;
    KPAL_CHECK(upload_and_count(w, n_bases, k, d_table, counter_bits, (cudaStream_t)stream));
    // the pinned staging buffers are reused by the next call: wait for the copies
    KPAL_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return KPAL_OK;
}

// Host FASTA bytes -> this call's counter table in the library's workspace (no copy into a
// caller-owned table): what the multi-GPU driver hands to kpal_dev_slice_push.  The pointer
// stays valid until the next host-level counting call on this device.  Synchronises the stream.
extern "C" int kpal_count_fasta_dev_table(const char *fasta, uint64_t n_bytes, int k, void **d_table_out,
                                          int *counter_bits_out, void *stream)
{
    if (!d_table_out || !counter_bits_out || (!fasta && n_bytes)) return bad_arg("null pointer");
    KPAL_CHECK(check_k_host(k));
    KPAL_CHECK(require_device());
    std::lock_guard<std::mutex> lock(g_count_mutex);
    CountWorkspace *w;
    KPAL_CHECK(get_count_ws(&w));
    const uint64_t bins = 1ull << (2 * k);
    const int bits = (n_bytes >= (1ull << 32)) ? 64 : 32;
    KPAL_CHECK(w->table.ensure(bins * (bits / 8)));
    cudaStream_t st = (cudaStream_t)stream;
    *d_table_out = w->table.p;
    *counter_bits_out = bits;
    if (use_gpu_fasta() && n_bytes > 0) {
        KPAL_CUDA(cudaMemsetAsync(w->table.p, 0, bins * (bits / 8), st));
        unsigned flags = 0;
        uint64_t n_bases = 0;
        KPAL_CHECK(fasta_gpu_count(w, fasta, n_bytes, k, w->table.p, bits, st, &flags, &n_bases));
        if (!flags) return KPAL_OK;
    }
    uint64_t n_bases = 0;
    KPAL_CUDA(cudaMemsetAsync(w->table.p, 0, bins * (bits / 8), st));
    KPAL_CHECK(pack_fasta_ws(w, fasta, n_bytes, &n_bases));
    KPAL_CHECK(upload_and_count(w, n_bases, k, w->table.p, bits, st));
    KPAL_CUDA(cudaStreamSynchronize(st));
    return KPAL_OK;
}

// A device counter table (this GPU's, or the sum of all ranks' tables after the
// multi-GPU reduce) -> the caller's int64 profile on the host: widen + optional
// balance + the narrow D2H of finalize_to_host.  Synchronises the stream.
extern "C" int kpal_dev_table_to_host(const void *d_table, int counter_bits, int k, int balance,
                                      int64_t *counts_out, void *stream)
{
    if (!d_table || !counts_out) return bad_arg("null pointer");
    if (counter_bits != 32 && counter_bits != 64) return bad_arg("counter_bits must be 32 or 64");
    KPAL_CHECK(check_k_host(k));
    KPAL_CHECK(require_device());
    std::lock_guard<std::mutex> lock(g_count_mutex);
    CountWorkspace *w;
    KPAL_CHECK(get_count_ws(&w));
    return finalize_to_host(w, d_table, counter_bits, k, balance, counts_out, (cudaStream_t)stream);
}

// Dense int64 rows on the host for records [first, first + n).  When no record of the call can
// overflow 16 bits (the usual case: reads, contigs below 65536 bases) the rows leave the device
// as uint16 -- a quarter of the PCIe bytes -- in batches through two pinned buffers, and the
// host threads of widen.cpp widen batch b into the caller's array while batch b + 1 is being
// counted and copied; the caller's array may be pageable.  Otherwise: int64 rows, copied as
// they are.
extern "C" int kpal_count_by_record(const uint32_t *codes, const uint32_t *valid, uint64_t n_bases,
                                    const uint64_t *rec_starts, uint64_t first, uint64_t n, int k,
                                    int balance, int64_t *rows_out)
{
    if (k < 1 || k > KPAL_MAX_K) { set_error("k-mer length %d out of range [1, %d]", k, KPAL_MAX_K); return KPAL_EINVAL; }
    if (n == 0) return KPAL_OK;
    if (!codes || !valid || !rec_starts || !rows_out) return bad_arg("null pointer");
    KPAL_CHECK(require_device());
    const uint64_t bins = 1ull << (2 * k);
    // upload only the chunks the requested records touch
    const uint64_t b0 = rec_starts[first], b1 = rec_starts[first + n];
    if (b1 > n_bases || b0 > b1) return bad_arg("record starts outside the packed stream");
    const uint64_t c0 = b0 / 64, c1 = (b1 + 63) / 64 + 1;       // + halo chunk
    std::vector<uint64_t> rs(n + 1);
    uint64_t longest = 0;
    for (uint64_t r = 0; r <= n; ++r) {
        rs[r] = rec_starts[first + r] - c0 * 64;
        if (r) longest = std::max(longest, rs[r] - rs[r - 1]);
    }
    // (device buffers from the grow-only workspace: a cudaMalloc / cudaFree pair per buffer and call
    // costs more than the kernel when a --by-record run makes hundreds of calls)
    std::lock_guard<std::mutex> lock(g_count_mutex);
    CountWorkspace *w;
    KPAL_CHECK(get_count_ws(&w));
    KPAL_CHECK(w->br_codes.ensure((c1 - c0) * 16));
    KPAL_CHECK(w->br_valid.ensure((c1 - c0) * 8));
    KPAL_CHECK(w->br_starts.ensure((n + 1) * 8));
    struct Ref { void *p; uint32_t *u32() const { return static_cast<uint32_t *>(p); } uint64_t *u64() const { return static_cast<uint64_t *>(p); } };
    const Ref d_codes{w->br_codes.p}, d_valid{w->br_valid.p}, d_rs{w->br_starts.p};
    DevBuf d_rows;
    cudaStream_t st = 0;
    KPAL_CUDA(cudaMemcpyAsync(d_codes.p, codes + c0 * 4, (c1 - c0) * 16, cudaMemcpyHostToDevice, st));
    KPAL_CUDA(cudaMemcpyAsync(d_valid.p, valid + c0 * 2, (c1 - c0) * 8, cudaMemcpyHostToDevice, st));
    KPAL_CUDA(cudaMemcpyAsync(d_rs.p, rs.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st));

    if (g_narrow_d2h.load() && k >= 2 && longest < (balance ? 32768ull : 65536ull)) {
        // Batches of ~16 MB of uint16 rows (128 MB of int64 on the host side), at most 128 MB: several
        // per call, so that the copy of batch b + 1 runs while the host threads widen batch b.
        const uint64_t batch = std::max<uint64_t>(1, std::min<uint64_t>(n, (16ull << 20) / (bins * 2)));
        for (int i = 0; i < 2; ++i) {
            KPAL_CHECK(w->rows16[i].ensure(batch * bins * 2));
            KPAL_CHECK(w->prows16[i].ensure(batch * bins * 2));
            if (!w->rows_done[i]) KPAL_CUDA(cudaEventCreateWithFlags(&w->rows_done[i], cudaEventDisableTiming));
        }
        const uint64_t n_batches = (n + batch - 1) / batch;
        auto enqueue = [&](uint64_t b) -> int {
            const uint64_t r = b * batch, m = std::min(batch, n - r);
            uint16_t *d = static_cast<uint16_t *>(w->rows16[b & 1].p);
            KPAL_CHECK(launch_by_record_u16(d_codes.u32(), d_valid.u32(), d_rs.u64(),
                                            r, m, k, balance, d, st));
            KPAL_CUDA(cudaMemcpyAsync(w->prows16[b & 1].p, d, m * bins * 2, cudaMemcpyDeviceToHost, st));
            KPAL_CUDA(cudaEventRecord(w->rows_done[b & 1], st));
            return KPAL_OK;
        };
        int rc = enqueue(0);
        if (rc == KPAL_OK && n_batches > 1) rc = enqueue(1);
        for (uint64_t b = 0; b < n_batches && rc == KPAL_OK; ++b) {
            const uint64_t r = b * batch, m = std::min(batch, n - r);
            cudaError_t e = cudaEventSynchronize(w->rows_done[b & 1]);
            if (e != cudaSuccess) {
                set_error("narrow D2H of the rows failed: %s", cudaGetErrorString(e));
                cudaGetLastError();
                rc = KPAL_ECUDA;
                break;
            }
            WidenHandle *h = widen_begin(w->prows16[b & 1].p, 2, rows_out + r * bins, m * bins, m * bins);
            widen_publish(h, m * bins);
            widen_end(h, 0);
            if (b + 2 < n_batches) rc = enqueue(b + 2);     // its buffers are free again
        }
        cudaStreamSynchronize(st);                           // nothing of this call stays queued
        return rc;
    }

    // rows in batches that fit comfortably on the device
    const uint64_t batch = std::max<uint64_t>(1, std::min<uint64_t>(n, (4ull << 30) / (bins * 8)));
    KPAL_CHECK(d_rows.alloc(batch * bins * 8));
    for (uint64_t r = 0; r < n; r += batch) {
        const uint64_t m = std::min(batch, n - r);
        KPAL_CHECK(launch_by_record(d_codes.u32(), d_valid.u32(), d_rs.u64(),
                                    r, m, k, balance, d_rows.as<int64_t>(), st));
        KPAL_CUDA(cudaMemcpyAsync(rows_out + r * bins, d_rows.p, m * bins * 8, cudaMemcpyDeviceToHost, st));
        KPAL_CUDA(cudaStreamSynchronize(st));
    }
    return KPAL_OK;
}

extern "C" int kpal_balance(int64_t *counts, int k)
{
    if (!counts) return bad_arg("null pointer");
    if (k < 1 || k > KPAL_MAX_K) { set_error("k-mer length %d out of range [1, %d]", k, KPAL_MAX_K); return KPAL_EINVAL; }
    KPAL_CHECK(require_device());
    const uint64_t bins = 1ull << (2 * k);
    DevBuf d_in, d_out;
    KPAL_CHECK(d_in.alloc(bins * 8));
    KPAL_CHECK(d_out.alloc(bins * 8));
    cudaStream_t st = 0;
    KPAL_CUDA(cudaMemcpyAsync(d_in.p, counts, bins * 8, cudaMemcpyHostToDevice, st));
    KPAL_CHECK(launch_balance(d_in.as<int64_t>(), d_out.as<int64_t>(), k, st));
    KPAL_CUDA(cudaMemcpyAsync(counts, d_out.p, bins * 8, cudaMemcpyDeviceToHost, st));
    KPAL_CUDA(cudaStreamSynchronize(st));
    return KPAL_OK;
}

// --------------------------------------------------- distances: device API
extern "C" uint64_t kpal_prepared_stride(int k) { return prepared_stride_host(k); }
extern "C" uint64_t kpal_distance_num_tiles(uint64_t n) { return distance_num_tiles(n); }

extern "C" int kpal_dev_profiles_prepare(const int64_t *d_counts, uint64_t n, int k, int do_balance,
                                         int do_scale, double *d_F, double *d_R, uint32_t *d_bitmap,
                                         double *d_totals, double *d_norm2, void *stream)
{
    if (n && (!d_counts || !d_F || !d_bitmap || !d_totals || !d_norm2)) return bad_arg("null device pointer");
    DevBuf tot;
    KPAL_CHECK(tot.alloc(n * 8));
    int r = launch_prepare(d_counts, n, k, do_balance, do_scale, d_F, d_R, d_bitmap, d_totals, d_norm2,
                           tot.as<unsigned long long>(), (cudaStream_t)stream);
    if (r == KPAL_OK) KPAL_CUDA(cudaStreamSynchronize((cudaStream_t)stream));   // tot is freed on return
    return r;
}

extern "C" int kpal_dev_order_by_total(const double *d_totals, uint64_t n, int down, int32_t *d_order,
                                       void *stream)
{
    if (!d_totals || !d_order) return bad_arg("null device pointer");
    return make_order(d_totals, n, down, d_order, (cudaStream_t)stream);
}

extern "C" int kpal_dev_distance_tiles(const double *d_F, const double *d_R, const uint32_t *d_bitmap,
                                       const double *d_totals, const double *d_norm2,
                                       const int32_t *d_order, uint64_t n, int k, int metric,
                                       int pairwise, int do_scale, int down, uint64_t tile_begin,
                                       uint64_t tile_end, double *d_out, void *stream)
{
    if (!d_F || !d_bitmap || !d_totals || !d_norm2 || !d_out) return bad_arg("null device pointer");
    double *acc; uint32_t *cnt;
    KPAL_CHECK(get_dist_scratch(n, (cudaStream_t)stream, &acc, &cnt));
    return launch_distance_tiles(d_F, d_R, d_bitmap, d_totals, d_norm2, d_order, n, k, metric, pairwise,
                                 do_scale, down, tile_begin, tile_end, acc, cnt, d_out,
                                 (cudaStream_t)stream);
}

// Multi-GPU form of kpal_dev_distance_tiles: the finished values of tiles [tile_begin,
// tile_end) as a compact [tile][kpal_distance_tile_elems()] array, ready for one gather.
extern "C" uint64_t kpal_distance_tile_elems(void) { return distance_tile_elems(); }

extern "C" int kpal_dev_distance_tiles_packed(const double *d_F, const double *d_R, const uint32_t *d_bitmap,
                                              const double *d_totals, const double *d_norm2,
                                              const int32_t *d_order, uint64_t n, int k, int metric,
                                              int pairwise, int do_scale, int down, uint64_t tile_begin,
                                              uint64_t tile_end, double *d_packed, void *stream)
{
    if (!d_F || !d_bitmap || !d_totals || !d_norm2 || !d_packed) return bad_arg("null device pointer");
    double *acc; uint32_t *cnt;
    KPAL_CHECK(get_dist_scratch(n, (cudaStream_t)stream, &acc, &cnt));
    return launch_distance_tiles(d_F, d_R, d_bitmap, d_totals, d_norm2, d_order, n, k, metric, pairwise,
                                 do_scale, down, tile_begin, tile_end, acc, cnt, nullptr,
                                 (cudaStream_t)stream, d_packed);
}

extern "C" int kpal_dev_distance_unpack_tiles(const double *d_packed, const double *d_totals,
                                              const double *d_norm2, const int32_t *d_order, uint64_t n,
                                              int metric, int pairwise, int do_scale, uint64_t tile_begin,
                                              uint64_t tile_end, int diagonal, double *d_out, void *stream)
{
    if (!d_packed || !d_totals || !d_norm2 || !d_out) return bad_arg("null device pointer");
    return launch_distance_unpack(d_packed, d_totals, d_norm2, d_order, n, metric, pairwise, do_scale,
                                  tile_begin, tile_end, diagonal, d_out, (cudaStream_t)stream);
}

// Euclidean / cosine through the exact integer Gram matrix (distance_gram.cu), device API.
extern "C" uint64_t kpal_gram_row_stride(int k) { return (k < 1 || k > KPAL_MAX_K) ? 0 : gram_row_stride(k); }

extern "C" int kpal_dev_gram_prepare(const int64_t *d_counts, uint64_t n, int k, int do_balance, uint8_t *d_rows_u8,
                                     uint64_t *d_totals, uint64_t *d_norms, uint32_t *d_flags, void *stream)
{
    if (n && (!d_counts || !d_rows_u8 || !d_totals || !d_norms || !d_flags)) return bad_arg("null device pointer");
    return launch_gram_prepare(d_counts, n, k, do_balance, d_rows_u8,
                               reinterpret_cast<unsigned long long *>(d_totals),
                               reinterpret_cast<unsigned long long *>(d_norms), d_flags, (cudaStream_t)stream);
}

extern "C" int kpal_dev_gram_distances(const uint8_t *d_rows_u8, const uint64_t *d_totals, const uint64_t *d_norms,
                                       uint64_t norm_max, uint64_t n, int k, int metric, int do_scale, int down,
                                       int64_t *d_gram, double *d_out, void *stream)
{
    if (!d_rows_u8 || !d_totals || !d_norms || !d_gram || !d_out) return bad_arg("null device pointer");
    return launch_gram_distances(d_rows_u8, reinterpret_cast<const unsigned long long *>(d_totals),
                                 reinterpret_cast<const unsigned long long *>(d_norms), norm_max, n, k, metric,
                                 do_scale, down, reinterpret_cast<long long *>(d_gram), d_out, (cudaStream_t)stream);
}

// ----------------------------------------------------- distances: host API
// A matrix session takes the profile set in slabs (kpal_matrix_push) so the caller never
// has to hold all N x 4^k int64 counts in host memory the way kpal/kmer.py:694-698 does:
// every pushed slab is uploaded into one of two device slabs on a copy stream and turned
// into the prepared fp64 arrays on the compute stream while the next slab uploads.
namespace kpal {
struct MatrixSession {
    int device = 0;
    uint64_t n = 0, d = 0, stride = 0, slab_rows = 0, pushed = 0;
    int k = 0, metric = 0, pairwise = 0, do_balance = 0, do_scale = 0, down = 0;
    bool need_r = false;
    int cur = 0;
    DevBuf F, R, bitmap, totals, norm2, order, tot_i64, slab[2], d_out, acc, cnt;   // acc / cnt: this session's accumulators
    // euclidean / cosine with option "gram": u8 rows, exact totals / norms, range flags { flags, pad, norm_max }
    bool gram = false;
    DevBuf x8, g_totals, g_norms, g_flags, g_matrix;
    cudaStream_t copy = nullptr, compute = nullptr;
    cudaEvent_t copied[2] = {nullptr, nullptr}, prepared[2] = {nullptr, nullptr};
    ~MatrixSession()
    {
        for (int i = 0; i < 2; ++i) {
            if (copied[i]) cudaEventDestroy(copied[i]);
            if (prepared[i]) cudaEventDestroy(prepared[i]);
        }
        if (copy) cudaStreamDestroy(copy);
        if (compute) cudaStreamDestroy(compute);
    }
};
}  // namespace kpal

static int matrix_open(MatrixSession *s)
{
    KPAL_CUDA(cudaGetDevice(&s->device));
    s->d = 1ull << (2 * s->k);
    s->stride = prepared_stride_host(s->k);
    s->need_r = (s->metric == KPAL_METRIC_MULTISET && s->pairwise == KPAL_PAIRWISE_PROD);
    const uint64_t n = s->n;
    KPAL_CHECK(s->F.alloc(n * s->stride * 8));
    if (s->need_r) KPAL_CHECK(s->R.alloc(n * s->stride * 8));
    KPAL_CHECK(s->bitmap.alloc(n * (s->stride / 32) * 4));
    KPAL_CHECK(s->totals.alloc(n * 8));
    KPAL_CHECK(s->norm2.alloc(n * 8));
    KPAL_CHECK(s->order.alloc(n * 4));
    KPAL_CHECK(s->d_out.alloc(n * n * 8));
    s->gram = g_gram.load() != 0 && n > 12 &&
              (s->metric == KPAL_METRIC_EUCLIDEAN || s->metric == KPAL_METRIC_COSINE);
    if (s->gram) {
        KPAL_CHECK(s->x8.alloc(n * gram_row_stride(s->k)));
        KPAL_CHECK(s->g_totals.alloc(n * 8));
        KPAL_CHECK(s->g_norms.alloc(n * 8));
        KPAL_CHECK(s->g_flags.alloc(16));
        KPAL_CUDA(cudaMemset(s->g_flags.p, 0, 16));
    }
    // raw int64 profiles pass through two bounded slabs (<= 512 MiB each)
    s->slab_rows = std::max<uint64_t>(1, std::min<uint64_t>(std::min<uint64_t>(n, 65535), (512ull << 20) / (s->d * 8)));
    for (int i = 0; i < 2; ++i) KPAL_CHECK(s->slab[i].alloc(s->slab_rows * s->d * 8));
    KPAL_CHECK(s->tot_i64.alloc(s->slab_rows * 8));
    KPAL_CUDA(cudaStreamCreateWithFlags(&s->copy, cudaStreamNonBlocking));
    KPAL_CUDA(cudaStreamCreateWithFlags(&s->compute, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        KPAL_CUDA(cudaEventCreateWithFlags(&s->copied[i], cudaEventDisableTiming));
        KPAL_CUDA(cudaEventCreateWithFlags(&s->prepared[i], cudaEventDisableTiming));
        KPAL_CUDA(cudaEventRecord(s->prepared[i], s->compute));
    }
    return KPAL_OK;
}

static int matrix_push(MatrixSession *s, const int64_t *rows, uint64_t m)
{
    if (s->pushed + m > s->n) return bad_arg("more profiles pushed than the session was opened for");
    for (uint64_t r = 0; r < m; r += s->slab_rows) {
        const uint64_t c = std::min(s->slab_rows, m - r), at = s->pushed;
        const int b = s->cur;
        KPAL_CUDA(cudaStreamWaitEvent(s->copy, s->prepared[b], 0));       // slab b is free again
        KPAL_CUDA(cudaMemcpyAsync(s->slab[b].p, rows + r * s->d, c * s->d * 8, cudaMemcpyHostToDevice, s->copy));
        KPAL_CUDA(cudaEventRecord(s->copied[b], s->copy));
        KPAL_CUDA(cudaStreamWaitEvent(s->compute, s->copied[b], 0));
        KPAL_CHECK(launch_prepare(s->slab[b].as<int64_t>(), c, s->k, s->do_balance, s->do_scale,
                                  s->F.as<double>() + at * s->stride,
                                  s->need_r ? s->R.as<double>() + at * s->stride : nullptr,
                                  s->bitmap.as<uint32_t>() + at * (s->stride / 32),
                                  s->totals.as<double>() + at, s->norm2.as<double>() + at,
                                  s->tot_i64.as<unsigned long long>(), s->compute));
        if (s->gram)
            KPAL_CHECK(launch_gram_prepare(s->slab[b].as<int64_t>(), c, s->k, s->do_balance,
                                           s->x8.as<uint8_t>() + at * gram_row_stride(s->k),
                                           s->g_totals.as<unsigned long long>() + at,
                                           s->g_norms.as<unsigned long long>() + at,
                                           s->g_flags.as<unsigned int>(), s->compute));
        KPAL_CUDA(cudaEventRecord(s->prepared[b], s->compute));
        s->pushed += c;
        s->cur ^= 1;
    }
    // the caller may reuse `rows` as soon as this returns
    KPAL_CUDA(cudaStreamSynchronize(s->copy));
    return KPAL_OK;
}

static int matrix_finish(MatrixSession *s, double *out)
{
    if (s->pushed != s->n) return bad_arg("fewer profiles pushed than the session was opened for");
    const uint64_t n = s->n;
    cudaStream_t st = s->compute;
    if (s->gram) {
        // the Gram form is exact while every count fits 8 bits; the device says whether it does
        struct { unsigned int flags, pad; unsigned long long norm_max; } h;
        KPAL_CHECK(launch_gram_norm_max(s->g_norms.as<unsigned long long>(), n,
                                        reinterpret_cast<unsigned long long *>(s->g_flags.as<unsigned char>() + 8), st));
        KPAL_CUDA(cudaMemcpyAsync(&h, s->g_flags.p, 16, cudaMemcpyDeviceToHost, st));
        KPAL_CUDA(cudaStreamSynchronize(st));
        if (h.flags == 0) {
            if (!s->g_matrix.p) KPAL_CHECK(s->g_matrix.alloc(n * n * 8));
            KPAL_CHECK(launch_gram_distances(s->x8.as<uint8_t>(), s->g_totals.as<unsigned long long>(),
                                             s->g_norms.as<unsigned long long>(), h.norm_max, n, s->k, s->metric,
                                             s->do_scale, s->down, s->g_matrix.as<long long>(),
                                             s->d_out.as<double>(), st));
            KPAL_CUDA(cudaMemcpyAsync(out, s->d_out.p, n * n * 8, cudaMemcpyDeviceToHost, st));
            KPAL_CUDA(cudaStreamSynchronize(st));
            return KPAL_OK;
        }
        // a count above 255: the element-wise fp64 kernel below
    }
    const int32_t *d_order = nullptr;
    if (s->do_scale) {
        KPAL_CHECK(make_order(s->totals.as<double>(), n, s->down, s->order.as<int32_t>(), st));
        d_order = s->order.as<int32_t>();
    }
    if (!s->acc.p) KPAL_CHECK(s->acc.alloc(n * n * sizeof(double)));
    if (!s->cnt.p) KPAL_CHECK(s->cnt.alloc(n * n * sizeof(uint32_t)));
    double *acc = s->acc.as<double>();
    uint32_t *cnt = s->cnt.as<uint32_t>();
    KPAL_CHECK(launch_distance_tiles(s->F.as<double>(), s->R.as<double>(), s->bitmap.as<uint32_t>(),
                                     s->totals.as<double>(), s->norm2.as<double>(), d_order, n, s->k,
                                     s->metric, s->pairwise, s->do_scale, s->down, 0,
                                     distance_num_tiles(n), acc, cnt, s->d_out.as<double>(), st));
    KPAL_CUDA(cudaMemcpyAsync(out, s->d_out.p, n * n * 8, cudaMemcpyDeviceToHost, st));
    KPAL_CUDA(cudaStreamSynchronize(st));
    return KPAL_OK;
}

extern "C" int kpal_matrix_open(uint64_t n, int k, int metric, int pairwise, int do_balance,
                                int do_scale, int down, void **session)
{
    if (!session) return bad_arg("null pointer");
    *session = nullptr;
    if (n < 1) return bad_arg("need at least one profile");
    if (k < 1 || k > KPAL_MAX_K) { set_error("k-mer length %d out of range [1, %d]", k, KPAL_MAX_K); return KPAL_EINVAL; }
    if (metric < 0 || metric > KPAL_METRIC_COSINE || pairwise < 0 || pairwise > KPAL_PAIRWISE_SUM)
        return bad_arg("unknown metric / pairwise selector");
    KPAL_CHECK(require_device());
    MatrixSession *s = new MatrixSession();
    s->n = n; s->k = k; s->metric = metric; s->pairwise = pairwise;
    s->do_balance = do_balance; s->do_scale = do_scale; s->down = down;
    const int r = matrix_open(s);
    if (r != KPAL_OK) { delete s; return r; }
    *session = s;
    return KPAL_OK;
}

extern "C" int kpal_matrix_push(void *session, const int64_t *rows, uint64_t m)
{
    if (!session) return bad_arg("null session");
    if (m == 0) return KPAL_OK;
    if (!rows) return bad_arg("null pointer");
    return matrix_push(static_cast<MatrixSession *>(session), rows, m);
}

extern "C" int kpal_matrix_finish(void *session, double *out)
{
    if (!session || !out) return bad_arg("null pointer");
    return matrix_finish(static_cast<MatrixSession *>(session), out);
}

extern "C" void kpal_matrix_close(void *session)
{
    delete static_cast<MatrixSession *>(session);
}

extern "C" int kpal_distance_matrix(const int64_t *profiles, uint64_t n, int k, int metric,
                                    int pairwise, int do_balance, int do_scale, int down, double *out)
{
    if (!profiles || !out) return bad_arg("null pointer");
    void *session = nullptr;
    KPAL_CHECK(kpal_matrix_open(n, k, metric, pairwise, do_balance, do_scale, down, &session));
    int r = kpal_matrix_push(session, profiles, n);
    if (r == KPAL_OK) r = kpal_matrix_finish(session, out);
    kpal_matrix_close(session);
    return r;
}

extern "C" int kpal_pair_distance(const int64_t *left, const int64_t *right, int k, int metric,
                                  int pairwise, int do_balance, int do_scale, int down, double *out)
{
    if (!left || !right || !out) return bad_arg("null pointer");
    if (k < 1 || k > KPAL_MAX_K) { set_error("k-mer length %d out of range [1, %d]", k, KPAL_MAX_K); return KPAL_EINVAL; }
    const uint64_t d = 1ull << (2 * k);
    void *session = nullptr;
    KPAL_CHECK(kpal_matrix_open(2, k, metric, pairwise, do_balance, do_scale, down, &session));
    double m[4];
    int r = kpal_matrix_push(session, left, 1);
    if (r == KPAL_OK) r = kpal_matrix_push(session, right, 1);
    if (r == KPAL_OK) r = kpal_matrix_finish(session, m);
    kpal_matrix_close(session);
    (void)d;
    if (r == KPAL_OK) *out = m[1];
    return r;
}

// ProfileDistance.distance with do_positive (kpal/kdistlib.py:139-157): balance, then the
// pair-dependent mask, then scale factors from the MASKED totals, then the metric.
extern "C" int kpal_pair_distance_positive(const int64_t *left, const int64_t *right, int k,
                                           int metric, int pairwise, int do_balance, int do_scale,
                                           int down, double *out)
{
    if (!left || !right || !out) return bad_arg("null pointer");
    if (k < 1 || k > KPAL_MAX_K) { set_error("k-mer length %d out of range [1, %d]", k, KPAL_MAX_K); return KPAL_EINVAL; }
    if (metric < 0 || metric > KPAL_METRIC_COSINE || pairwise < 0 || pairwise > KPAL_PAIRWISE_SUM)
        return bad_arg("unknown metric / pairwise selector");
    KPAL_CHECK(require_device());
    const uint64_t d = 1ull << (2 * k), stride = prepared_stride_host(k);
    const bool need_r = (metric == KPAL_METRIC_MULTISET && pairwise == KPAL_PAIRWISE_PROD);
    DevBuf raw, work, F, R, bitmap, totals, norm2, order, tot_i64, d_out, d_acc, d_cnt;
    KPAL_CHECK(raw.alloc(2 * d * 8));
    KPAL_CHECK(work.alloc(2 * d * 8));
    KPAL_CHECK(F.alloc(2 * stride * 8));
    if (need_r) KPAL_CHECK(R.alloc(2 * stride * 8));
    KPAL_CHECK(bitmap.alloc(2 * (stride / 32) * 4));
    KPAL_CHECK(totals.alloc(16));
    KPAL_CHECK(norm2.alloc(16));
    KPAL_CHECK(order.alloc(8));
    KPAL_CHECK(tot_i64.alloc(16));
    KPAL_CHECK(d_out.alloc(4 * 8));
    KPAL_CHECK(d_acc.alloc(4 * 8));
    KPAL_CHECK(d_cnt.alloc(4 * 4));
    cudaStream_t st = 0;
    KPAL_CUDA(cudaMemcpyAsync(raw.p, left, d * 8, cudaMemcpyHostToDevice, st));
    KPAL_CUDA(cudaMemcpyAsync(raw.as<int64_t>() + d, right, d * 8, cudaMemcpyHostToDevice, st));
    int64_t *pair = raw.as<int64_t>();
    if (do_balance) {
        KPAL_CHECK(launch_balance(raw.as<int64_t>(), work.as<int64_t>(), k, st));
        KPAL_CHECK(launch_balance(raw.as<int64_t>() + d, work.as<int64_t>() + d, k, st));
        pair = work.as<int64_t>();
    }
    KPAL_CHECK(launch_positive_pair(pair, pair + d, k, st));
    KPAL_CHECK(launch_prepare(pair, 2, k, 0, do_scale, F.as<double>(), need_r ? R.as<double>() : nullptr,
                              bitmap.as<uint32_t>(), totals.as<double>(), norm2.as<double>(),
                              tot_i64.as<unsigned long long>(), st));
    const int32_t *d_order = nullptr;
    if (do_scale) {
        KPAL_CHECK(make_order(totals.as<double>(), 2, down, order.as<int32_t>(), st));
        d_order = order.as<int32_t>();
    }
    KPAL_CHECK(launch_distance_tiles(F.as<double>(), R.as<double>(), bitmap.as<uint32_t>(),
                                     totals.as<double>(), norm2.as<double>(), d_order, 2, k, metric,
                                     pairwise, do_scale, down, 0, distance_num_tiles(2), d_acc.as<double>(),
                                     d_cnt.as<uint32_t>(),
                                     d_out.as<double>(), st));
    double m[4];
    KPAL_CUDA(cudaMemcpyAsync(m, d_out.p, 32, cudaMemcpyDeviceToHost, st));
    KPAL_CUDA(cudaStreamSynchronize(st));
    *out = m[1];
    return KPAL_OK;
}

// ------------------------------------------------ split / showbalance: host API
extern "C" uint64_t kpal_split_length(int k)
{
    return (k < 1 || k > KPAL_MAX_K) ? 0 : split_length(k);
}

extern "C" int kpal_split(const int64_t *counts, int k, int64_t *forward, int64_t *reverse)
{
    if (!counts || !forward || !reverse) return bad_arg("null pointer");
    if (k < 1 || k > KPAL_MAX_K) { set_error("k-mer length %d out of range [1, %d]", k, KPAL_MAX_K); return KPAL_EINVAL; }
    KPAL_CHECK(require_device());
    const uint64_t bins = 1ull << (2 * k), half = split_length(k);
    DevBuf d_in, d_f, d_r, scratch;
    KPAL_CHECK(d_in.alloc(bins * 8));
    KPAL_CHECK(d_f.alloc(half * 8));
    KPAL_CHECK(d_r.alloc(half * 8));
    KPAL_CHECK(scratch.alloc(split_scratch_bytes(k)));
    cudaStream_t st = 0;
    KPAL_CUDA(cudaMemcpyAsync(d_in.p, counts, bins * 8, cudaMemcpyHostToDevice, st));
    KPAL_CHECK(launch_split(d_in.as<int64_t>(), k, d_f.as<int64_t>(), d_r.as<int64_t>(), scratch.p, st));
    KPAL_CUDA(cudaMemcpyAsync(forward, d_f.p, half * 8, cudaMemcpyDeviceToHost, st));
    KPAL_CUDA(cudaMemcpyAsync(reverse, d_r.p, half * 8, cudaMemcpyDeviceToHost, st));
    KPAL_CUDA(cudaStreamSynchronize(st));
    return KPAL_OK;
}

extern "C" int kpal_show_balance(const int64_t *counts, int k, double *out)
{
    if (!counts || !out) return bad_arg("null pointer");
    if (k < 1 || k > KPAL_MAX_K) { set_error("k-mer length %d out of range [1, %d]", k, KPAL_MAX_K); return KPAL_EINVAL; }
    KPAL_CHECK(require_device());
    const uint64_t bins = 1ull << (2 * k);
    DevBuf d_in, res;
    KPAL_CHECK(d_in.alloc(bins * 8));
    KPAL_CHECK(res.alloc(16));
    cudaStream_t st = 0;
    KPAL_CUDA(cudaMemcpyAsync(d_in.p, counts, bins * 8, cudaMemcpyHostToDevice, st));
    KPAL_CHECK(launch_show_balance(d_in.as<int64_t>(), k, res.as<double>(),
                                   res.as<unsigned long long>() + 1, st));
    struct { double sum; unsigned long long nz; } h;
    KPAL_CUDA(cudaMemcpyAsync(&h, res.p, 16, cudaMemcpyDeviceToHost, st));
    KPAL_CUDA(cudaStreamSynchronize(st));
    *out = h.sum / double(h.nz + 1);                       // kpal/metrics.py:123
    return KPAL_OK;
}

// ------------------------------------------- multi-GPU: peer-memory table reduce
extern "C" int kpal_ipc_export(const void *d_ptr, void *handle64)
{
    if (!d_ptr || !handle64) return bad_arg("null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    KPAL_CUDA(cudaIpcGetMemHandle(&h, const_cast<void *>(d_ptr)));
    memcpy(handle64, &h, 64);
    return KPAL_OK;
}

extern "C" int kpal_ipc_open(const void *handle64, void **d_peer_ptr)
{
    if (!handle64 || !d_peer_ptr) return bad_arg("null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    *d_peer_ptr = nullptr;
    KPAL_CUDA(cudaIpcOpenMemHandle(d_peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return KPAL_OK;
}

extern "C" int kpal_ipc_close(void *d_peer_ptr)
{
    if (!d_peer_ptr) return KPAL_OK;
    KPAL_CUDA(cudaIpcCloseMemHandle(d_peer_ptr));
    return KPAL_OK;
}

extern "C" uint64_t kpal_peer_inbox_bytes(int k, int counter_bits, int world)
{
    if (k < 1 || k > KPAL_MAX_K || world < 1 || (counter_bits != 32 && counter_bits != 64)) return 0;
    return peer_inbox_bytes(k, counter_bits, world);
}

extern "C" int kpal_dev_reduce_push(const void *d_table, int counter_bits, int k, int rank, int world,
                                    void *const *inbox_ptrs, void *stream)
{
    return launch_reduce_push(d_table, counter_bits, k, rank, world, inbox_ptrs, (cudaStream_t)stream);
}

extern "C" int kpal_dev_reduce_collect(const void *d_inbox, int counter_bits, int k, int rank, int world,
                                       void *d_root_table, void *stream)
{
    return launch_reduce_collect(d_inbox, counter_bits, k, rank, world, d_root_table, (cudaStream_t)stream);
}

extern "C" int kpal_dev_count_packed_push(const uint32_t *d_codes, const uint32_t *d_valid,
                                          uint64_t n_bases, int k, void *d_table, int counter_bits,
                                          int rank, int world, void *const *inbox_ptrs, void *stream,
                                          int *fused_out)
{
    if (!d_table || (n_bases && (!d_codes || !d_valid))) return bad_arg("null device pointer");
    return launch_count_push(d_codes, d_valid, n_bases, k, d_table, counter_bits, rank, world,
                             inbox_ptrs, (cudaStream_t)stream, fused_out);
}

// ------------------------- multi-GPU: balance + narrow reduce-scatter + distributed finalize
extern "C" uint64_t kpal_slice_inbox_bytes(int k, int world)
{
    if (k < 6 || k > KPAL_MAX_K || world < 1 || world > 16) return 0;
    return slice_inbox_bytes(k, world);
}

extern "C" uint64_t kpal_slice_begin(int k, int rank, int world)
{
    if (k < 1 || k > KPAL_MAX_K || world < 1 || rank < 0) return 0;
    return slice_begin_host(k, rank, world);
}

extern "C" int kpal_dev_slice_push(const void *d_table, int counter_bits, int k, int rank, int world,
                                   void *const *inbox_ptrs, uint64_t epoch, int wide_rows, void *stream)
{
    if (!d_table) return bad_arg("null device pointer");
    if (epoch < 1) return bad_arg("epochs count from 1 (a zeroed inbox means epoch 0)");
    KPAL_CHECK(require_device());
    CountWorkspace *w;
    {
        std::lock_guard<std::mutex> lock(g_count_mutex);
        KPAL_CHECK(get_count_ws(&w));
        if (!w->wide_flag.p) {          // the kernel's counters: zero once, its last CTA resets them
            KPAL_CHECK(w->wide_flag.ensure(32));
            KPAL_CUDA(cudaMemset(w->wide_flag.p, 0, 32));
        }
    }
    return launch_slice_push(d_table, counter_bits, k, rank, world, inbox_ptrs, epoch, wide_rows,
                             static_cast<unsigned int *>(w->wide_flag.p), (cudaStream_t)stream);
}

extern "C" int kpal_dev_slice_signal(int k, int rank, int world, void *const *inbox_ptrs, uint64_t epoch,
                                     int wide_rows, void *stream)
{
    if (epoch < 1) return bad_arg("epochs count from 1 (a zeroed inbox means epoch 0)");
    KPAL_CHECK(require_device());
    return launch_slice_signal(k, rank, world, inbox_ptrs, epoch, wide_rows, (cudaStream_t)stream);
}

extern "C" int kpal_dev_slice_collect(void *const *inbox_ptrs, int k, int rank, int world, uint64_t epoch,
                                      int signal, int64_t *d_slice_out, void *stream)
{
    KPAL_CHECK(require_device());
    return launch_slice_collect(k, rank, world, inbox_ptrs, epoch, signal, d_slice_out, nullptr, nullptr, nullptr,
                                (cudaStream_t)stream);
}

// kpal_dev_slice_collect + the device->host copy of the slice in the narrow form of finalize_to_host
// (uint8 / uint16 over PCIe, widened into slice_out by the host workers).  slice_out may be any host
// memory, e.g. this rank's part of a profile in memory shared between the ranks' processes.
extern "C" int kpal_dev_slice_collect_to_host(void *const *inbox_ptrs, int k, int rank, int world, uint64_t epoch,
                                              int signal, int64_t *slice_out, void *stream)
{
    if (!inbox_ptrs || !slice_out) return bad_arg("null pointer");
    KPAL_CHECK(check_k_host(k));
    KPAL_CHECK(require_device());
    if (rank < 0 || rank >= world) return bad_arg("rank outside the world");
    std::lock_guard<std::mutex> lock(g_count_mutex);
    CountWorkspace *w;
    KPAL_CHECK(get_count_ws(&w));
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t n = slice_begin_host(k, rank + 1, world) - slice_begin_host(k, rank, world);
    KPAL_CHECK(w->counts.ensure(n * 8));
    const int narrow = g_narrow_d2h.load();
    if (narrow && n >= (1ull << 16)) {
        const bool try8 = narrow == 1;
        KPAL_CHECK(w->counts16.ensure(n * 2));
        if (try8) KPAL_CHECK(w->counts8.ensure(n));
        KPAL_CHECK(w->overflow.ensure(16));
        KPAL_CHECK(w->pnarrow.ensure(n * 2));
        KPAL_CHECK(w->pflag.ensure(16));
        if (!w->d2h_done[0]) {
            for (auto &e : w->d2h_done) KPAL_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            KPAL_CUDA(cudaEventCreateWithFlags(&w->flag_done, cudaEventDisableTiming));
        }
        KPAL_CHECK(launch_slice_collect(k, rank, world, inbox_ptrs, epoch, signal, static_cast<int64_t *>(w->counts.p),
                                        static_cast<uint16_t *>(w->counts16.p),
                                        try8 ? static_cast<uint8_t *>(w->counts8.p) : nullptr,
                                        static_cast<unsigned int *>(w->overflow.p), st));
        bool done = false;
        KPAL_CHECK(narrow_copy_out(w, n, 0, ((n + 15) / 16 + 15) / 16 * 16, try8, slice_out, st, nullptr, nullptr, &done));
        if (done) return KPAL_OK;
        KPAL_CUDA(cudaStreamSynchronize(st));
    } else {
        KPAL_CHECK(launch_slice_collect(k, rank, world, inbox_ptrs, epoch, signal, static_cast<int64_t *>(w->counts.p), nullptr,
                                        nullptr, nullptr, st));
    }
    KPAL_CUDA(cudaMemcpyAsync(slice_out, w->counts.p, n * 8, cudaMemcpyDeviceToHost, st));
    KPAL_CUDA(cudaStreamSynchronize(st));
    return KPAL_OK;
}
