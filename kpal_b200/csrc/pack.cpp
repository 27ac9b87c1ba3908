// Host side of the packed-sequence format: sequence list / FASTA text ->
// 2-bit codes + validity bits + record starts (see include/kpal_b200.h).
//
// Replaces, for the device path, the text handling the reference does with
// Bio.SeqIO.parse + str(record.seq) (kpal/klib.py:111,131) and the regex split
// on [^AaCcGgTt] (kpal/klib.py:152): every other byte becomes an *invalid*
// base, and one invalid base separates records.
//
// Multi-threaded (std::thread): the input is cut at line / record boundaries,
// each worker sizes its part, a prefix sum fixes every part's position in the
// bit streams, then the workers pack concurrently (the only shared words are
// the first/last word of a part, merged with atomic OR).
#include <algorithm>
#include <atomic>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/kpal_b200.h"

namespace kpal {
void set_error(const char *fmt, ...);

static uint8_t g_lut[256];
static std::once_flag g_lut_once;

static void init_lut()
{
    for (int i = 0; i < 256; ++i) g_lut[i] = 4;     // 4 = invalid
    g_lut[(int)'A'] = g_lut[(int)'a'] = 0;            // kpal/klib.py:43-48
    g_lut[(int)'C'] = g_lut[(int)'c'] = 1;
    g_lut[(int)'G'] = g_lut[(int)'g'] = 2;
    g_lut[(int)'T'] = g_lut[(int)'t'] = 3;
}

static unsigned worker_count(uint64_t bytes)
{
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 1;
    if (hw > 32) hw = 32;
    uint64_t by_size = bytes / (1u << 20) + 1;        // >= 1 MiB per worker
    return (unsigned)std::min<uint64_t>(hw, by_size);
}

template <typename F>
static void run_parallel(unsigned n, F &&f)
{
    if (n <= 1) { f(0u); return; }
    std::vector<std::thread> th;
    th.reserve(n);
    for (unsigned t = 0; t < n; ++t) th.emplace_back([&f, t] { f(t); });
    for (auto &x : th) x.join();
}

// Appends bases to the two bit streams starting at an arbitrary base offset.
class BitWriter {
public:
    BitWriter(uint32_t *codes, uint32_t *valid, uint64_t pos)
        : codes_(codes), valid_(valid), pos_(pos), first_word_(pos / 32), c_(0), v_(0) {}

    inline void push(unsigned code4)   // 0..3 valid base, 4 invalid
    {
        const unsigned g = (unsigned)(pos_ & 31u);
        if (code4 < 4) {
            c_ |= (uint64_t)code4 << (62 - 2 * g);
            v_ |= 1u << (31 - g);
        }
        ++pos_;
        if ((pos_ & 31u) == 0) flush((pos_ - 1) / 32);
    }
    void finish()
    {
        if (pos_ & 31u) flush(pos_ / 32, /*last=*/true);
    }
    uint64_t pos() const { return pos_; }

private:
    inline void flush(uint64_t word, bool last = false)
    {
        const uint32_t hi = (uint32_t)(c_ >> 32), lo = (uint32_t)c_;
        if (word == first_word_ || last) {          // may be shared with a neighbour part
            __atomic_fetch_or(&codes_[2 * word], hi, __ATOMIC_RELAXED);
            __atomic_fetch_or(&codes_[2 * word + 1], lo, __ATOMIC_RELAXED);
            __atomic_fetch_or(&valid_[word], v_, __ATOMIC_RELAXED);
        } else {
            codes_[2 * word] = hi;
            codes_[2 * word + 1] = lo;
            valid_[word] = v_;
        }
        c_ = 0;
        v_ = 0;
    }
    uint32_t *codes_, *valid_;
    uint64_t pos_, first_word_;
    uint64_t c_;
    uint32_t v_;
};

}  // namespace kpal

using namespace kpal;

extern "C" void kpal_packed_words(uint64_t n_bases, uint64_t *code_words, uint64_t *valid_words)
{
    // one extra 64-base chunk so that chunk+1 (the halo) is always readable
    const uint64_t chunks = (n_bases + 63) / 64 + 1;
    if (code_words) *code_words = chunks * 4;
    if (valid_words) *valid_words = chunks * 2;
}

extern "C" int kpal_pack_sequences(const char *text, const uint64_t *offsets, uint64_t n_records,
                                   uint32_t *codes, uint32_t *valid, uint64_t *rec_starts,
                                   uint64_t *n_bases_out)
{
    std::call_once(g_lut_once, init_lut);
    if (!offsets) { set_error("invalid argument: null offsets"); return KPAL_EINVAL; }
    const uint64_t total = (n_records ? offsets[n_records] - offsets[0] : 0) + n_records;
    if (n_bases_out) *n_bases_out = total;
    if (!codes || !valid) return KPAL_OK;              // size query only
    if (!text && total != n_records) { set_error("invalid argument: null text"); return KPAL_EINVAL; }
    uint64_t cw, vw;
    kpal_packed_words(total, &cw, &vw);
    memset(codes, 0, cw * sizeof(uint32_t));
    memset(valid, 0, vw * sizeof(uint32_t));
    if (rec_starts)
        for (uint64_t r = 0; r <= n_records; ++r) rec_starts[r] = offsets[r] - offsets[0] + r;

    // Workers take contiguous, 32-base aligned ranges of the OUTPUT stream, so a
    // single huge record (a chromosome) is packed in parallel as well.  Output
    // position q belongs to the record r with start(r) <= q < start(r+1), where
    // start(r) = offsets[r] - offsets[0] + r; its last position is the separator.
    if (total == 0) return KPAL_OK;
    const unsigned nw = worker_count(total);
    auto start_of = [&](uint64_t r) { return offsets[r] - offsets[0] + r; };
    run_parallel(nw, [&](unsigned t) {
        uint64_t q0 = (total / nw * t) & ~uint64_t(31);
        uint64_t q1 = (t + 1 == nw) ? total : ((total / nw * (t + 1)) & ~uint64_t(31));
        if (q0 >= q1) return;
        // record containing q0: largest r with start(r) <= q0
        uint64_t lo = 0, hi = n_records;              // invariant: start(lo) <= q0 < start(hi)
        while (hi - lo > 1) {
            const uint64_t mid = lo + (hi - lo) / 2;
            if (start_of(mid) <= q0) lo = mid; else hi = mid;
        }
        uint64_t r = lo;
        BitWriter w(codes, valid, q0);
        while (w.pos() < q1) {
            const uint64_t rs = start_of(r);
            const uint64_t len = offsets[r + 1] - offsets[r];
            uint64_t i = w.pos() - rs;                  // index inside record r (== len: separator)
            const unsigned char *p = (const unsigned char *)text + offsets[r];
            const uint64_t stop = std::min<uint64_t>(len, q1 - rs);
            for (; i < stop; ++i) w.push(g_lut[p[i]]);
            if (w.pos() < q1) { w.push(4); ++r; }       // record separator
        }
        w.finish();
    });
    return KPAL_OK;
}

// ---------------------------------------------------------------------------
// FASTA
// ---------------------------------------------------------------------------
namespace {

inline bool is_space(unsigned char c)   // Python str.rstrip() whitespace (ASCII part)
{
    return c == ' ' || (c >= 9 && c <= 13) || (c >= 28 && c <= 31);
}

struct Part {
    uint64_t begin = 0, end = 0;       // byte range, begins at a line start
    uint64_t headers = 0;              // '>' lines
    uint64_t bases_before = 0;         // kept sequence bytes before the part's first header
    uint64_t bases_after = 0;          // kept sequence bytes after it (0 if no header)
    uint64_t name_bytes = 0;           // incl. terminating '\0' per header
};

// Walk the lines of [begin, end).  on_header(name_ptr, name_len), on_base(byte).
template <typename H, typename B>
inline void walk(const unsigned char *s, uint64_t begin, uint64_t end, H &&on_header, B &&on_base)
{
    uint64_t p = begin;
    while (p < end) {
        const unsigned char *nl = (const unsigned char *)memchr(s + p, '\n', end - p);
        const uint64_t le = nl ? (uint64_t)(nl - s) : end;       // line = [p, le)
        if (s[p] == '>' ) {
            uint64_t a = p + 1, b = le;
            while (b > a && is_space(s[b - 1])) --b;              // title.rstrip()
            while (a < b && is_space(s[a])) ++a;                  // split(None, 1)[0]
            uint64_t c = a;
            while (c < b && !is_space(s[c])) ++c;
            on_header(s + a, c - a);
        } else {
            uint64_t b = le;
            while (b > p && is_space(s[b - 1])) --b;              // line.rstrip()
            for (uint64_t q = p; q < b; ++q) {
                const unsigned char ch = s[q];
                if (ch == ' ' || ch == '\r') continue;            // .replace(" ", "").replace("\r", "")
                on_base(ch);
            }
        }
        p = le + 1;
    }
}

std::vector<Part> make_parts(const unsigned char *s, uint64_t n)
{
    const unsigned nw = worker_count(n);
    std::vector<Part> parts(nw);
    uint64_t prev = 0;
    for (unsigned t = 0; t < nw; ++t) {
        uint64_t e = (t + 1 == nw) ? n : n / nw * (t + 1);
        if (e < prev) e = prev;
        if (t + 1 != nw && e < n) {                                // snap to the next line start
            const unsigned char *nl = (const unsigned char *)memchr(s + e, '\n', n - e);
            e = nl ? (uint64_t)(nl - s) + 1 : n;
        }
        parts[t].begin = prev;
        parts[t].end = e;
        prev = e;
    }
    run_parallel(nw, [&](unsigned t) {
        Part &pt = parts[t];
        walk(s, pt.begin, pt.end,
             [&](const unsigned char *, uint64_t len) { ++pt.headers; pt.name_bytes += len + 1; },
             [&](unsigned char) { if (pt.headers) ++pt.bases_after; else ++pt.bases_before; });
    });
    return parts;
}

}  // namespace

extern "C" int kpal_fasta_scan(const char *fasta, uint64_t n_bytes, uint64_t *n_records,
                               uint64_t *n_bases, uint64_t *name_bytes)
{
    if (!fasta && n_bytes) { set_error("invalid argument: null fasta"); return KPAL_EINVAL; }
    const auto parts = make_parts((const unsigned char *)fasta, n_bytes);
    uint64_t recs = 0, bases = 0, names = 0;
    for (const Part &p : parts) {
        if (recs) bases += p.bases_before;       // text before the first header is skipped
        recs += p.headers;
        bases += p.bases_after;
        names += p.name_bytes;
    }
    if (n_records) *n_records = recs;
    if (n_bases) *n_bases = bases + recs;        // + one separator per record
    if (name_bytes) *name_bytes = names;
    return KPAL_OK;
}

extern "C" int kpal_fasta_pack(const char *fasta, uint64_t n_bytes, uint32_t *codes, uint32_t *valid,
                               uint64_t *rec_starts, char *names)
{
    std::call_once(g_lut_once, init_lut);
    if ((!fasta && n_bytes) || !codes || !valid) {
        set_error("invalid argument: null buffer");
        return KPAL_EINVAL;
    }
    const unsigned char *s = (const unsigned char *)fasta;
    const auto parts = make_parts(s, n_bytes);
    const unsigned nw = (unsigned)parts.size();
    // Packed layout: record r = its bases followed by one separator.  A part
    // first emits the tail of the record that was open when it begins
    // (bases_before), then for each header: separator of the previous record
    // (none before the very first record), then the new record's bases.  The
    // final separator is written at the end.
    std::vector<uint64_t> base0(nw), rec0(nw), name0(nw);
    uint64_t recs = 0, pos = 0, names_pos = 0;
    for (unsigned t = 0; t < nw; ++t) {
        base0[t] = pos; rec0[t] = recs; name0[t] = names_pos;
        const Part &p = parts[t];
        if (recs) pos += p.bases_before;
        // separators emitted inside this part: one per header except the global first
        pos += p.headers - ((recs == 0 && p.headers) ? 1 : 0);
        pos += p.bases_after;
        recs += p.headers;
        names_pos += p.name_bytes;
    }
    const uint64_t total = pos + (recs ? 1 : 0);
    uint64_t cw, vw;
    kpal_packed_words(total, &cw, &vw);
    memset(codes, 0, cw * sizeof(uint32_t));
    memset(valid, 0, vw * sizeof(uint32_t));

    run_parallel(nw, [&](unsigned t) {
        const Part &pt = parts[t];
        BitWriter w(codes, valid, base0[t]);
        uint64_t rec = rec0[t];
        char *np = names ? names + name0[t] : nullptr;
        const bool skip_leading = (rec0[t] == 0);        // nothing open yet: skip until a header
        bool open = !skip_leading;
        walk(s, pt.begin, pt.end,
             [&](const unsigned char *name, uint64_t len) {
                 if (rec > 0) w.push(4);                // close the previous record
                 if (rec_starts) rec_starts[rec] = w.pos();
                 if (np) { memcpy(np, name, len); np[len] = '\0'; np += len + 1; }
                 ++rec;
                 open = true;
             },
             [&](unsigned char ch) { if (open) w.push(g_lut[ch]); });
        if (t + 1 == nw && recs) {                       // separator after the last record
            w.push(4);
            if (rec_starts) rec_starts[recs] = w.pos();
        }
        w.finish();
    });
    return KPAL_OK;
}
