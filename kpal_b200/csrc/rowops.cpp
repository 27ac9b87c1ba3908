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
extern "C" int kpal_row_stats(const int64_t *rows, uint64_t n_rows, uint64_t n_cols, double *stats_out)
{
    if ((!rows || !stats_out) && n_rows) return KPAL_EINVAL;
    if (n_cols == 0) return KPAL_EINVAL;
    parallel_for(n_rows, [&](uint64_t r) {
        const int64_t *x = rows + r * n_cols;
        int64_t total = 0;
        uint64_t non_zero = 0;
        for (uint64_t i = 0; i < n_cols; ++i) { total += x[i]; non_zero += x[i] != 0; }
        const double mean = double(total) / double(n_cols);
        std::vector<double> work(n_cols);
        for (uint64_t i = 0; i < n_cols; ++i) { const double d = double(x[i]) - mean; work[i] = d * d; }
        const double var = pairwise_sum(work.data(), n_cols) / double(n_cols);
        // median: mean of the two middle order statistics (one for an odd length), as np.median
        std::vector<int64_t> copy(x, x + n_cols);
        const uint64_t mid = n_cols / 2;
        std::nth_element(copy.begin(), copy.begin() + mid, copy.end());
        double median = double(copy[mid]);
        if (n_cols % 2 == 0) {
            const int64_t below = *std::max_element(copy.begin(), copy.begin() + mid);
            median = (double(below) + double(copy[mid])) / 2.0;        // np.mean of the two: add, then divide
        }
        double *out = stats_out + r * 5;
        out[0] = double(total); out[1] = double(non_zero); out[2] = mean; out[3] = median; out[4] = sqrt(var);
    });
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
