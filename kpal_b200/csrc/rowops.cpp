// Host side of writing many per-record profiles (kpal count --by-record, BASELINE configs[2]):
// Profile.save (reference kpal/klib.py:227-256) stores, per profile, the gzip-compressed counts
// and six statistics -- total, non_zero, mean, median, std (kpal/klib.py:192-225: NumPy sum,
// count_nonzero, mean, median, std).  At 100 000 profiles the NumPy calls (1.8 ms per profile)
// and the per-chunk zlib calls from Python (1.2 ms) are the run time, not the GPU.  These two
// entry points do that work for a whole batch of rows on all host threads:
//
//   kpal_row_stats       the five statistics of every row, bit-identical to NumPy's (the same
//                        operations in the same order: exact integer sums, the mean as
//                        double(sum) / n, the median as the mean of the two middle order
//                        statistics, the standard deviation as sqrt(pairwise_sum((x - mean)^2) / n)
//                        with NumPy's pairwise summation: blocks of 128, eight accumulators);
//   kpal_deflate_chunks  zlib streams (what HDF5's deflate filter stores) of equal-sized chunks.
//
// Host only, no GPU.
#include "../../include/kpal_b200.h"

#include <math.h>
#include <stdint.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <utility>
#include <vector>

namespace {

// numpy/core/src/umath/loops_utils.h.src: pairwise sum of n doubles
double pairwise_sum(const double *a, size_t n)
{
    if (n < 8) {
        double res = 0.0;
        for (size_t i = 0; i < n; ++i) res += a[i];
        return res;
    }
    if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        size_t i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    size_t n2 = n / 2;
    n2 -= n2 % 8;
    return pairwise_sum(a, n2) + pairwise_sum(a + n2, n - n2);
}

unsigned worker_count(uint64_t items)
{
    unsigned n = std::thread::hardware_concurrency();
    if (n == 0) n = 4;
    if (n > 64) n = 64;
    if (uint64_t(n) > items) n = unsigned(items ? items : 1);
    return n;
}

template <typename F>
void parallel_for(uint64_t items, F &&body)
{
    const unsigned n = worker_count(items);
    std::atomic<uint64_t> next{0};
    auto run = [&] { for (uint64_t i; (i = next.fetch_add(1)) < items;) body(i); };
    if (n <= 1) { run(); return; }
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < n; ++t) pool.emplace_back(run);
    for (auto &t : pool) t.join();
}

}  // namespace

// stats_out: [n_rows][5] doubles = total, non_zero, mean, median, std (total and non_zero are
// integers below 2^53 for any profile this library counts)
// pairwise_sum of (double(x[i]) - mean)^2 without materialising the squares: the same
// operations in the same order as np.std (subtract, multiply, NumPy's pairwise summation)
static double pairwise_sq(const int64_t *x, double mean, size_t n)
{
    auto term = [&](size_t i) { const double d = double(x[i]) - mean; return d * d; };
    if (n < 8) {
        double res = 0.0;
        for (size_t i = 0; i < n; ++i) res += term(i);
        return res;
    }
    if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = term(j);
        size_t i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += term(i + j);
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += term(i);
        return res;
    }
    size_t n2 = n / 2;
    n2 -= n2 % 8;
    return pairwise_sq(x, mean, n2) + pairwise_sq(x + n2, mean, n - n2);
}

static void one_row_stats(const int64_t *x, uint64_t n_cols, double *out)
{
    int64_t total = 0, lowest = 0;
    uint64_t non_zero = 0;
    for (uint64_t i = 0; i < n_cols; ++i) { total += x[i]; non_zero += x[i] != 0; lowest = std::min(lowest, x[i]); }
    const double mean = double(total) / double(n_cols);
    const double var = pairwise_sq(x, mean, n_cols) / double(n_cols);
    // median: mean of the two middle order statistics (one for an odd length), as np.median.
    // Counts are not negative: when more than half of them are zero, so are the middle ones.
    const uint64_t mid = n_cols / 2;
    double median = 0.0;
    if (lowest < 0 || n_cols - non_zero < mid + 1) {
        std::vector<int64_t> copy(x, x + n_cols);
        std::nth_element(copy.begin(), copy.begin() + mid, copy.end());
        median = double(copy[mid]);
        if (n_cols % 2 == 0) {
            const int64_t below = *std::max_element(copy.begin(), copy.begin() + mid);
            median = (double(below) + double(copy[mid])) / 2.0;        // np.mean of the two: add, then divide
        }
    }
    out[0] = double(total); out[1] = double(non_zero); out[2] = mean; out[3] = median; out[4] = sqrt(var);
}

extern "C" int kpal_row_stats(const int64_t *rows, uint64_t n_rows, uint64_t n_cols, double *stats_out)
{
    if ((!rows || !stats_out) && n_rows) return KPAL_EINVAL;
    if (n_cols == 0) return KPAL_EINVAL;
    parallel_for(n_rows, [&](uint64_t r) { one_row_stats(rows + r * n_cols, n_cols, stats_out + r * 5); });
    return KPAL_OK;
}

extern "C" uint64_t kpal_deflate_bound(uint64_t chunk_bytes) { return compressBound(uLong(chunk_bytes)); }

// chunk c = data[c * chunk_bytes, (c + 1) * chunk_bytes) -> out[c * slot_bytes ...], sizes[c] bytes.
extern "C" int kpal_deflate_chunks(const void *data, uint64_t n_chunks, uint64_t chunk_bytes, int level,
                                   void *out, uint64_t slot_bytes, uint32_t *sizes)
{
    if ((!data || !out || !sizes) && n_chunks) return KPAL_EINVAL;
    if (slot_bytes < compressBound(uLong(chunk_bytes)) || level < 0 || level > 9) return KPAL_EINVAL;
    std::atomic<int> failed{0};
    // a work item = up to 32 neighbouring chunks (one k = 8 profile)
    const uint64_t per = 32, items = (n_chunks + per - 1) / per;
    parallel_for(items, [&](uint64_t item) {
        for (uint64_t c = item * per; c < std::min(n_chunks, (item + 1) * per); ++c) {
            uLongf len = uLongf(slot_bytes);
            const int rc = compress2(static_cast<Bytef *>(out) + c * slot_bytes, &len,
                                     static_cast<const Bytef *>(data) + c * chunk_bytes, uLong(chunk_bytes), level);
            if (rc != Z_OK) { failed.store(1); len = 0; }
            sizes[c] = uint32_t(len);
        }
    });
    return failed.load() ? KPAL_EINVAL : KPAL_OK;
}

// ---------------------------------------------------------------------------
// Deflate for sparse count rows
// ---------------------------------------------------------------------------
// A per-record profile is almost all zero bytes (1 kbp at k = 8: <= 993 counts in 65536, each
// an int64 with one non-zero byte), and zlib spends 2.8 ms per row on finding that out with
// its hash chains -- 85 % of a `kpal count --by-record` run.  This encoder writes a valid zlib
// stream (RFC 1950 / 1951, what HDF5's deflate filter and any inflate read) in one pass over
// the bytes: non-zero bytes as literals, runs of zero bytes as one literal 0 followed by
// matches of distance 1 (length up to 258), in ONE dynamic-Huffman block whose code is fixed
// in advance for this kind of data: match-258 1 bit, literal 0 2 bits, literals 1..228 10
// bits, everything else 11 bits (Kraft sum exactly 1), a single 1-bit distance code.  An
// all-zero 64 KiB chunk is 2 bits per 258 bytes.  Dense data (more than a quarter of the
// bytes non-zero) and anything that would not fit the slot goes to zlib's compress2.
namespace {

struct SparseCode {
    uint16_t code[286];          // bit-reversed (deflate sends Huffman codes most significant bit first)
    uint8_t len[286];
    std::vector<std::pair<uint32_t, int>> header;        // (bits, count) of the block header, in order
    SparseCode()
    {
        for (int i = 0; i < 286; ++i) len[i] = 11;
        len[0] = 2;
        for (int i = 1; i <= 228; ++i) len[i] = 10;
        len[285] = 1;
        canonical(len, 286, code);
        // ---- block header: BFINAL = 1, BTYPE = 10 (dynamic), HLIT = 29, HDIST = 0, HCLEN
        // code-length alphabet: symbols 1, 2, 10, 11 (3 bits each) and 16 = "repeat previous 3..6" (1 bit)
        uint8_t cl_len[19] = {0};
        cl_len[16] = 1; cl_len[1] = cl_len[2] = cl_len[10] = cl_len[11] = 3;
        uint16_t cl_code[19];
        canonical(cl_len, 19, cl_code);
        static const int order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        const int hclen = 18;                              // up to symbol 1
        put(1, 1); put(2, 2); put(286 - 257, 5); put(0, 5); put(hclen - 4, 4);
        for (int i = 0; i < hclen; ++i) put(cl_len[order[i]], 3);
        // the 287 code lengths (286 literal / length + 1 distance), run-length coded
        std::vector<int> all(len, len + 286);
        all.push_back(1);
        for (size_t i = 0; i < all.size();) {
            const int v = all[i];
            size_t run = 1;
            while (i + run < all.size() && all[i + run] == v) ++run;
            put(cl_code[v], cl_len[v]);
            size_t left = run - 1;
            while (left >= 3) {
                const size_t rep = left >= 9 || left <= 6 ? std::min<size_t>(left, 6) : left - 3;   // never leave 1 or 2 behind
                put(cl_code[16], cl_len[16]); put(uint32_t(rep - 3), 2);
                left -= rep;
            }
            for (; left; --left) put(cl_code[v], cl_len[v]);
            i += run;
        }
    }
    void put(uint32_t bits, int n) { header.emplace_back(bits, n); }
    static void canonical(const uint8_t *lens, int n, uint16_t *codes)
    {
        int count[16] = {0}, next[16] = {0};
        for (int i = 0; i < n; ++i) ++count[lens[i]];
        count[0] = 0;
        int c = 0;
        for (int b = 1; b < 16; ++b) { c = (c + count[b - 1]) << 1; next[b] = c; }
        for (int i = 0; i < n; ++i) {
            if (!lens[i]) { codes[i] = 0; continue; }
            int v = next[lens[i]]++, r = 0;
            for (int b = 0; b < lens[i]; ++b) r |= ((v >> b) & 1) << (lens[i] - 1 - b);
            codes[i] = uint16_t(r);
        }
    }
};

struct BitOut {
    unsigned char *p, *end;
    uint64_t acc = 0;
    int n = 0;
    bool overflow = false;
    BitOut(unsigned char *out, uint64_t cap) : p(out), end(out + cap) {}
    inline void put(uint32_t bits, int count)
    {
        acc |= uint64_t(bits) << n;
        n += count;
        if (n >= 32) {
            if (p + 4 > end) { overflow = true; n = 0; acc = 0; return; }
            memcpy(p, &acc, 4);                           // little endian: the low bits leave first
            p += 4; acc >>= 32; n -= 32;
        }
    }
    void flush()
    {
        while (n > 0) {
            if (p >= end) { overflow = true; return; }
            *p++ = (unsigned char)acc; acc >>= 8; n -= 8;
        }
        n = 0;
    }
};

// length 3..257 -> (code 257..284, extra bits, count); RFC 1951 3.2.5
inline void length_symbol(unsigned length, unsigned &sym, unsigned &extra, int &extra_bits)
{
    static const unsigned short base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59,
                                            67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const unsigned char bits[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    int i = 0;
    while (i < 28 && base[i + 1] <= length) ++i;
    sym = 257 + unsigned(i); extra = length - base[i]; extra_bits = bits[i];
}

// Returns the stream's size, or 0 when the data is not sparse / does not fit (caller: compress2).
uint64_t sparse_deflate(const unsigned char *src, uint64_t n, unsigned char *out, uint64_t cap)
{
    static const SparseCode code;
    if (cap < 16) return 0;
    BitOut bits(out + 2, cap - 2 - 4);
    out[0] = 0x78; out[1] = 0x5E;
    for (const auto &h : code.header) bits.put(h.first, h.second);
    uint64_t p = 0, literals = 0;
    const uint64_t dense = n / 4 + 64;
    // Adler-32 alongside: a zero byte leaves `a` alone and adds it to `b`, so a run of R zeros is b += R * a
    uint64_t ad_a = 1, ad_b = 0;
    const uint64_t mod = 65521;
    while (p < n) {
        if (src[p]) {
            bits.put(code.code[src[p]], code.len[src[p]]);
            ad_a += src[p]; if (ad_a >= mod) ad_a -= mod;
            ad_b += ad_a; if (ad_b >= mod) ad_b -= mod;
            ++p;
            if (++literals > dense) return 0;
            continue;
        }
        // a run of zero bytes: 8 at a time, then the tail
        uint64_t q = p;
        while (q < n && (q & 7) && src[q] == 0) ++q;
        if (q < n && !(q & 7) && src[q] == 0) {
            while (q + 32 <= n) {
                uint64_t w[4];
                memcpy(w, src + q, 32);
                if (w[0] | w[1] | w[2] | w[3]) break;
                q += 32;
            }
            while (q + 8 <= n) { uint64_t w; memcpy(&w, src + q, 8); if (w) break; q += 8; }
            while (q < n && src[q] == 0) ++q;
        }
        uint64_t run = q - p;
        ad_b = (ad_b + (run % mod) * ad_a) % mod;
        if (p == 0 || src[p - 1] != 0) { bits.put(code.code[0], code.len[0]); --run; }      // something to copy from
        while (run >= 258) { bits.put(code.code[285], 2); run -= 258; }                       // match 258 + the distance code (1 bit, 0)
        if (run >= 3) {
            unsigned sym, extra; int eb;
            length_symbol(unsigned(run), sym, extra, eb);
            bits.put(code.code[sym], code.len[sym]);
            if (eb) bits.put(extra, eb);
            bits.put(0, 1);                                                                     // distance 1
        } else {
            for (; run; --run) bits.put(code.code[0], code.len[0]);
        }
        p = q;
        if (bits.overflow) return 0;
    }
    bits.put(code.code[256], code.len[256]);
    bits.flush();
    if (bits.overflow) return 0;
    const uLong adler = uLong((ad_b << 16) | ad_a);
    unsigned char *t = bits.p;
    t[0] = (unsigned char)(adler >> 24); t[1] = (unsigned char)(adler >> 16);
    t[2] = (unsigned char)(adler >> 8); t[3] = (unsigned char)adler;
    return uint64_t(t + 4 - out);
}

}  // namespace

// As kpal_deflate_chunks, with the one-pass encoder for sparse chunks (same zlib container;
// the streams inflate to the same bytes but are not the ones zlib itself would write).
extern "C" int kpal_deflate_chunks_sparse(const void *data, uint64_t n_chunks, uint64_t chunk_bytes, int level,
                                          void *out, uint64_t slot_bytes, uint32_t *sizes)
{
    if ((!data || !out || !sizes) && n_chunks) return KPAL_EINVAL;
    if (slot_bytes < compressBound(uLong(chunk_bytes)) || level < 0 || level > 9 || chunk_bytes >= (1ull << 31)) return KPAL_EINVAL;
    std::atomic<int> failed{0};
    const uint64_t per = 32, items = (n_chunks + per - 1) / per;
    parallel_for(items, [&](uint64_t item) {
        for (uint64_t c = item * per; c < std::min(n_chunks, (item + 1) * per); ++c) {
            unsigned char *dst = static_cast<unsigned char *>(out) + c * slot_bytes;
            const unsigned char *src = static_cast<const unsigned char *>(data) + c * chunk_bytes;
            uint64_t size = level > 0 ? sparse_deflate(src, chunk_bytes, dst, slot_bytes) : 0;
            if (!size) {
                uLongf len = uLongf(slot_bytes);
                if (compress2(dst, &len, src, uLong(chunk_bytes), level) != Z_OK) { failed.store(1); len = 0; }
                size = len;
            }
            sizes[c] = uint32_t(size);
        }
    });
    return failed.load() ? KPAL_EINVAL : KPAL_OK;
}

// The packed form without the slot array: every worker appends its streams to an arena of its
// own, `finish` copies them into the caller's buffer in chunk order.  (The slot form needs
// n_chunks x compressBound(chunk) bytes -- as large as the rows themselves.)
namespace {
struct PackedJob {
    std::vector<std::vector<unsigned char>> arena;
    std::vector<uint32_t> owner;            // arena of chunk c
    std::vector<uint64_t> offset;           // ... and where in it
    std::vector<uint32_t> size;
};
}  // namespace

static int deflate_packed(const void *data, uint64_t n_chunks, uint64_t chunk_bytes, int level, int sparse,
                          uint32_t *sizes, void **handle_out, uint64_t *total_out, uint64_t per, uint64_t n_cols,
                          double *stats_out);

extern "C" int kpal_deflate_packed_begin(const void *data, uint64_t n_chunks, uint64_t chunk_bytes, int level,
                                         int sparse, uint32_t *sizes, void **handle_out, uint64_t *total_out)
{
    return deflate_packed(data, n_chunks, chunk_bytes, level, sparse, sizes, handle_out, total_out, 32, 0, nullptr);
}

// The rows of a --by-record batch in ONE pass: a work item is a row -- its statistics (as
// kpal_row_stats) and then the streams of its chunks, while the row is still in the core's cache.
extern "C" int kpal_rows_stats_deflate_begin(const int64_t *rows, uint64_t n_rows, uint64_t n_cols, uint64_t chunk_bytes,
                                             int level, int sparse, double *stats_out, uint32_t *sizes,
                                             void **handle_out, uint64_t *total_out)
{
    if (!stats_out || n_cols == 0 || chunk_bytes == 0 || (n_cols * 8) % chunk_bytes) return KPAL_EINVAL;
    const uint64_t per = n_cols * 8 / chunk_bytes;
    return deflate_packed(rows, n_rows * per, chunk_bytes, level, sparse, sizes, handle_out, total_out, per, n_cols, stats_out);
}

static int deflate_packed(const void *data, uint64_t n_chunks, uint64_t chunk_bytes, int level, int sparse,
                          uint32_t *sizes, void **handle_out, uint64_t *total_out, uint64_t per, uint64_t n_cols,
                          double *stats_out)
{
    if (!handle_out || !total_out || ((!data || !sizes) && n_chunks)) return KPAL_EINVAL;
    if (level < 0 || level > 9 || chunk_bytes == 0 || chunk_bytes >= (1ull << 31)) return KPAL_EINVAL;
    const uint64_t bound = compressBound(uLong(chunk_bytes));
    const uint64_t items = (n_chunks + per - 1) / per;
    const unsigned n_threads = worker_count(items);
    PackedJob *job = new PackedJob();
    job->arena.resize(n_threads);
    job->owner.resize(n_chunks);
    job->offset.resize(n_chunks);
    job->size.resize(n_chunks);
    std::atomic<uint64_t> next{0};
    std::atomic<int> failed{0};
    auto run = [&](unsigned t) {
        std::vector<unsigned char> &arena = job->arena[t];
        std::vector<unsigned char> scratch(bound);
        for (uint64_t item; (item = next.fetch_add(1)) < items;) {
            if (stats_out) one_row_stats(static_cast<const int64_t *>(data) + item * n_cols, n_cols, stats_out + item * 5);
            for (uint64_t c = item * per; c < std::min(n_chunks, (item + 1) * per); ++c) {
                const unsigned char *src = static_cast<const unsigned char *>(data) + c * chunk_bytes;
                uint64_t len = (sparse && level > 0) ? sparse_deflate(src, chunk_bytes, scratch.data(), bound) : 0;
                if (!len) {
                    uLongf z = uLongf(bound);
                    if (compress2(scratch.data(), &z, src, uLong(chunk_bytes), level) != Z_OK) { failed.store(1); z = 0; }
                    len = z;
                }
                job->owner[c] = t;
                job->offset[c] = arena.size();
                job->size[c] = uint32_t(len);
                arena.insert(arena.end(), scratch.data(), scratch.data() + len);
            }
        }
    };
    if (n_threads <= 1) run(0);
    else {
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < n_threads; ++t) pool.emplace_back(run, t);
        for (auto &t : pool) t.join();
    }
    if (failed.load()) { delete job; return KPAL_EINVAL; }
    uint64_t total = 0;
    for (uint64_t c = 0; c < n_chunks; ++c) { sizes[c] = job->size[c]; total += job->size[c]; }
    *handle_out = job;
    *total_out = total;
    return KPAL_OK;
}

// out (NULL: discard) receives the streams back to back, in chunk order; frees the job.
extern "C" int kpal_deflate_packed_finish(void *handle, void *out)
{
    if (!handle) return KPAL_EINVAL;
    PackedJob *job = static_cast<PackedJob *>(handle);
    if (out) {
        uint64_t at = 0;
        for (size_t c = 0; c < job->size.size(); ++c) {
            memcpy(static_cast<unsigned char *>(out) + at, job->arena[job->owner[c]].data() + job->offset[c], job->size[c]);
            at += job->size[c];
        }
    }
    delete job;
    return KPAL_OK;
}

// The streams of kpal_deflate_chunks packed back to back: out[0 .. sum(sizes)) (returned).
extern "C" uint64_t kpal_compact_slots(const void *slots, uint64_t slot_bytes, const uint32_t *sizes,
                                       uint64_t n_chunks, void *out)
{
    uint64_t at = 0;
    for (uint64_t c = 0; c < n_chunks; ++c) {
        if (out) memcpy(static_cast<unsigned char *>(out) + at, static_cast<const unsigned char *>(slots) + c * slot_bytes, sizes[c]);
        at += sizes[c];
    }
    return at;
}
