// Host-side text writer for the distance matrix (the step after the distance
// kernels): the lower triangle in the exact format of kdistlib.distance_matrix
// (kpal/kdistlib.py:179-186) -- row i = 1..n-1 holds d(p_i, p_j) for j < i, each
// value as Python's '{0:.{precision}f}', separated by one space, '\n' per row.
//
// Python formats a float with a correctly rounded (round-half-even on the
// exact binary value) fixed notation; std::to_chars(fixed, precision) is
// specified the same way, so the digits are identical.  Differences handled
// here: Python prints every NaN as "nan" (to_chars / printf give "-nan" for a
// negative one).  "inf" / "-inf" and "-0.000" agree.
//
// 8.4 M values (4096 profiles) take seconds through str.format; here rows are
// split over the host cores by equal value counts.
#include "../../include/kpal_b200.h"

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace kpal {
void set_error(const char *fmt, ...);
}

namespace {

inline void append_value(std::string &out, double v, int precision)
{
    if (std::isnan(v)) { out.append("nan"); return; }
    char buf[512];
    auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::fixed, precision);
    if (r.ec == std::errc()) { out.append(buf, r.ptr - buf); return; }
    // > 500 characters (huge magnitude and/or precision): rare, take the slow road
    std::vector<char> big(400 + size_t(precision));
    r = std::to_chars(big.data(), big.data() + big.size(), v, std::chars_format::fixed, precision);
    out.append(big.data(), r.ptr - big.data());
}

void format_rows(const double *values, uint64_t ld, uint64_t row_begin, uint64_t row_end,
                 int precision, std::string &out)
{
    uint64_t cells = 0;
    for (uint64_t i = row_begin; i < row_end; ++i) cells += i;
    out.reserve(cells * (size_t(precision) + 3) + 16);
    for (uint64_t i = row_begin; i < row_end; ++i) {
        const double *row = values + i * ld;
        for (uint64_t j = 0; j < i; ++j) {
            if (j) out.push_back(' ');
            append_value(out, row[j], precision);
        }
        out.push_back('\n');
    }
}

}  // namespace

extern "C" int kpal_format_matrix(const double *values, uint64_t n, uint64_t ld, int precision,
                                  char *text, uint64_t capacity, uint64_t *length)
{
    if (!length || (n > 1 && !values)) { kpal::set_error("null pointer"); return KPAL_EINVAL; }
    if (precision < 0 || precision > 1000) { kpal::set_error("precision %d out of range [0, 1000]", precision); return KPAL_EINVAL; }
    if (ld < n) { kpal::set_error("leading dimension smaller than the number of profiles"); return KPAL_EINVAL; }
    *length = 0;
    if (n < 2) return KPAL_OK;
    // rows 1..n-1 cut into pieces of about equal cell count, one per host thread
    const uint64_t cells = n * (n - 1) / 2;
    unsigned hw = std::thread::hardware_concurrency();
    uint64_t pieces = std::max<uint64_t>(1, std::min<uint64_t>(hw ? hw : 1, cells / 4096));
    pieces = std::min<uint64_t>(pieces, 64);
    std::vector<uint64_t> cut(pieces + 1, n);
    cut[0] = 1;
    for (uint64_t p = 1; p < pieces; ++p) {
        // first row r with r(r-1)/2 >= p * cells / pieces
        const double target = double(cells) * double(p) / double(pieces);
        uint64_t r = uint64_t(0.5 + std::sqrt(0.25 + 2.0 * target));
        cut[p] = std::min<uint64_t>(std::max<uint64_t>(r, cut[p - 1]), n);
    }
    std::vector<std::string> parts(pieces);
    if (pieces == 1) {
        format_rows(values, ld, 1, n, precision, parts[0]);
    } else {
        std::vector<std::thread> workers;
        for (uint64_t p = 0; p < pieces; ++p)
            workers.emplace_back(format_rows, values, ld, cut[p], cut[p + 1], precision, std::ref(parts[p]));
        for (auto &w : workers) w.join();
    }
    uint64_t total = 0;
    for (auto &s : parts) total += s.size();
    *length = total;
    if (total > capacity || !text) {
        kpal::set_error("matrix text needs %llu bytes, the buffer holds %llu",
                        (unsigned long long)total, (unsigned long long)capacity);
        return KPAL_EOVERFLOW;
    }
    uint64_t at = 0;
    for (auto &s : parts) { memcpy(text + at, s.data(), s.size()); at += s.size(); }
    return KPAL_OK;
}
