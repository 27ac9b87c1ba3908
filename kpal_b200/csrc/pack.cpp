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
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/kpal_b200.h"
#include "slotted.h"

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

// ---------------------------------------------------------------------------
// FASTA segments and the slotted stream (slotted.h) for the hybrid upload of
// kpal_count_fasta (cabi.cu)
// ---------------------------------------------------------------------------
// While the copy engine moves the head of a text to the device raw, idle host threads pack
// segments of its tail and only their 0.375 B/base go over the bus.  A segment begins at a
// line start and is written into a SLOT of the packed stream that belongs to it alone;
// everything behind the emitted bases up to the slot's end is invalid.  The record separator
// is the invalid base emitted at each header, as the device packer (fasta.cu) does.  Same
// FastaIterator rules as walk() above; the fast form handles 32 bytes per step (AVX2 + BMI2)
// and hands lines with blanks / control bytes to the scalar rules, so every text is packed
// exactly.
namespace {

// The slot is written with streaming (non-temporal) stores: the words are read next by the
// copy engine, and a PCIe read of lines that sit modified in the cores' caches runs at a
// fraction of the rate of a read from memory (measured: 20 instead of 53 GB/s).
struct SlotWriter {
    uint32_t *codes, *valid;       // the slot's first words (8-byte aligned: slots begin on 64-base boundaries)
    uint64_t pos = 0;              // bases emitted
    uint64_t c = 0;                // the open 32-base word
    uint32_t v = 0;
    SlotWriter(uint32_t *cw, uint32_t *vw) : codes(cw), valid(vw) {}
    inline void store(uint64_t w, uint64_t cc, uint32_t vv)
    {
        // codes[2w] = high half (first 16 bases), codes[2w + 1] = low half
#if defined(__x86_64__)
        _mm_stream_si64(reinterpret_cast<long long *>(codes + 2 * w), (long long)((cc << 32) | (cc >> 32)));
        _mm_stream_si32(reinterpret_cast<int *>(valid + w), int(vv));
#else
        codes[2 * w] = uint32_t(cc >> 32);
        codes[2 * w + 1] = uint32_t(cc);
        valid[w] = vv;
#endif
    }
    // n bases (1 .. 32): codes in the top 2n bits of cc, validity in the top n bits of vv, rest 0
    inline void append(uint64_t cc, uint32_t vv, unsigned n)
    {
        const unsigned g = unsigned(pos & 31u);
        c |= cc >> (2 * g);
        v |= vv >> g;
        pos += n;
        if (g + n >= 32) {
            store((pos >> 5) - 1, c, v);
            c = g ? cc << (64 - 2 * g) : 0;
            v = g ? vv << (32 - g) : 0;
        }
    }
    inline void push(unsigned code4)
    {
        if (code4 < 4) append(uint64_t(code4) << 62, 0x80000000u, 1); else append(0, 0, 1);
    }
    // close the open word and zero the slot up to `cap_bases` (a multiple of 64)
    void finish(uint64_t cap_bases)
    {
        uint64_t w = pos >> 5;
        if (pos & 31u) store(w++, c, v);
        for (const uint64_t words = cap_bases >> 5; w < words; ++w) store(w, 0, 0);
#if defined(__x86_64__)
        _mm_sfence();
#endif
    }
};

// rest of a sequence line from p (no blank or control byte between the line's start and p):
// rstrip, drop ' ' and '\r', every other byte is a base.  Returns the next line's start.
inline uint64_t scalar_line_rest(const unsigned char *s, uint64_t p, uint64_t end, SlotWriter &w)
{
    const unsigned char *nl = (const unsigned char *)memchr(s + p, '\n', end - p);
    const uint64_t le = nl ? uint64_t(nl - s) : end;
    uint64_t b = le;
    while (b > p && is_space(s[b - 1])) --b;
    for (uint64_t q = p; q < b; ++q) {
        const unsigned char ch = s[q];
        if (ch == ' ' || ch == '\r') continue;
        w.push(g_lut[ch]);
    }
    return le + 1;
}

inline uint64_t skip_line(const unsigned char *s, uint64_t p, uint64_t end)
{
    const unsigned char *nl = (const unsigned char *)memchr(s + p, '\n', end - p);
    return nl ? uint64_t(nl - s) + 1 : end;
}

// `open`: a record is open at p (a header line lies before the segment)
uint64_t pack_segment_scalar(const unsigned char *s, uint64_t p, uint64_t end, SlotWriter &w, bool open)
{
    while (p < end) {
        if (s[p] == '>') { w.push(4); open = true; p = skip_line(s, p, end); }
        else if (!open) p = skip_line(s, p, end);
        else p = scalar_line_rest(s, p, end, w);
    }
    return w.pos;
}

#if defined(__x86_64__)
__attribute__((target("avx2,bmi2")))
uint64_t pack_segment_avx2(const unsigned char *s, uint64_t p, uint64_t end, uint64_t n_total, SlotWriter &w, bool open)
{
    const __m256i rev16 = _mm256_setr_epi8(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0,
                                           15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0);
    const __m256i blank = _mm256_set1_epi8(0x20), upper = _mm256_set1_epi8((char)0xDF);
    const __m256i cA = _mm256_set1_epi8('A'), cC = _mm256_set1_epi8('C'), cG = _mm256_set1_epi8('G'),
                  cT = _mm256_set1_epi8('T');
    while (p < end) {
        if (s[p] == '>') { w.append(0, 0, 1); open = true; p = skip_line(s, p, end); continue; }
        if (!open) { p = skip_line(s, p, end); continue; }
        for (;;) {                                   // a sequence line, 32 bytes per step
            if (p + 32 > n_total) { p = scalar_line_rest(s, p, end, w); break; }
            const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + p));
            // bytes <= 0x20: the newline, blanks, control bytes
            const uint32_t ctl = uint32_t(_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_min_epu8(x, blank), x)));
            const unsigned n = ctl ? unsigned(__builtin_ctz(ctl)) : 32u;
            if (n) {
                // byte 0 -> bit 31 of the masks: reverse the bytes first
                const __m256i r = _mm256_permute4x64_epi64(_mm256_shuffle_epi8(x, rev16), 0x4E);
                const __m256i u = _mm256_and_si256(r, upper);
                const __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(u, cA), _mm256_cmpeq_epi8(u, cC)),
                                                   _mm256_or_si256(_mm256_cmpeq_epi8(u, cG), _mm256_cmpeq_epi8(u, cT)));
                uint32_t vm = uint32_t(_mm256_movemask_epi8(ok));
                // code = (b2, b1 ^ b2) of the ASCII byte: A 00, C 01, G 10, T 11 (either case)
                uint32_t hi = uint32_t(_mm256_movemask_epi8(_mm256_slli_epi16(r, 5)));
                uint32_t lo = uint32_t(_mm256_movemask_epi8(_mm256_slli_epi16(_mm256_xor_si256(r, _mm256_srli_epi16(r, 1)), 6)));
                if (n < 32) vm &= ~(0xffffffffu >> n);
                hi &= vm; lo &= vm;
                const uint64_t cc = _pdep_u64(hi, 0xAAAAAAAAAAAAAAAAull) | _pdep_u64(lo, 0x5555555555555555ull);
                w.append(cc, vm, n);
            }
            p += n;
            if (!ctl) continue;
            if (s[p] == '\n') { ++p; break; }
            p = scalar_line_rest(s, p, end, w);      // a blank or control byte inside the line
            break;
        }
    }
    return w.pos;
}
static const bool g_pack_avx2 = [] {
    const char *e = getenv("KPAL_NO_AVX2");
    return __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2") && !(e && e[0] == '1');
}();
#else
static const bool g_pack_avx2 = false;
#endif

// The last `want` positions emitted by text[.., end) (end = a line start), nearest first:
// out[0] is the position right before the cut.  Stops at the first invalid position (no
// window reaches across it; the rest is filled with invalid) and at the first header line.
void context_before(const unsigned char *s, uint64_t end, uint64_t first_header, int want, unsigned char *out)
{
    int got = 0;
    uint64_t p = end;
    while (got < want && p > first_header) {
        const uint64_t le = p - 1;                                       // s[le] == '\n'
        const unsigned char *nl = le ? (const unsigned char *)memrchr(s, '\n', le) : nullptr;
        const uint64_t ls = nl ? uint64_t(nl - s) + 1 : 0;
        if (ls < first_header) break;                                    // text before the first header: nothing emitted
        if (s[ls] == '>') { out[got++] = 4; break; }                     // the record separator
        uint64_t b = le;
        while (b > ls && is_space(s[b - 1])) --b;
        bool stop = false;
        for (uint64_t q = b; q > ls && got < want; --q) {
            const unsigned char ch = s[q - 1];
            if (ch == ' ' || ch == '\r') continue;
            const unsigned char code = g_lut[ch];
            out[got++] = code;
            if (code == 4) { stop = true; break; }
        }
        if (stop) break;
        p = ls;
    }
    while (got < want) out[got++] = 4;
}

// The first `want` positions emitted by text[begin, ..) (begin = a line start inside a record).
void context_after(const unsigned char *s, uint64_t begin, uint64_t n_total, int want, unsigned char *out)
{
    int got = 0;
    uint64_t p = begin;
    while (got < want && p < n_total) {
        if (s[p] == '>') { out[got++] = 4; break; }
        const unsigned char *nl = (const unsigned char *)memchr(s + p, '\n', n_total - p);
        const uint64_t le = nl ? uint64_t(nl - s) : n_total;
        uint64_t b = le;
        while (b > p && is_space(s[b - 1])) --b;
        bool stop = false;
        for (uint64_t q = p; q < b && got < want; ++q) {
            const unsigned char ch = s[q];
            if (ch == ' ' || ch == '\r') continue;
            const unsigned char code = g_lut[ch];
            out[got++] = code;
            if (code == 4) { stop = true; break; }
        }
        if (stop) break;
        p = le + 1;
    }
    while (got < want) out[got++] = 4;
}

}  // namespace

namespace kpal {
bool fasta_segment_fast() { return g_pack_avx2; }

// Packs text[begin, end) (begin = a line start) into the slot (codes / valid = its first
// words, cap_bases = its length, a multiple of 64 and >= end - begin).  `open`: a header
// line lies before `begin` (else bytes before the segment's first header are skipped).
// Returns the bases emitted.
uint64_t fasta_pack_segment(const unsigned char *s, uint64_t n_total, uint64_t begin, uint64_t end,
                            uint32_t *codes, uint32_t *valid, uint64_t cap_bases, bool open)
{
    std::call_once(g_lut_once, init_lut);
    SlotWriter w(codes, valid);
#if defined(__x86_64__)
    if (g_pack_avx2) pack_segment_avx2(s, begin, end, n_total, w, open);
    else
#endif
        pack_segment_scalar(s, begin, end, w, open);
    w.finish(cap_bases);
    return w.pos;
}

bool slotted_plan(const char *fasta, uint64_t n_bytes, uint64_t seg, SlottedPlan &plan)
{
    plan.cut.clear();
    plan.slot.clear();
    plan.m = 0;
    if (n_bytes == 0 || seg < 64) return false;
    // the first header line: usually byte 0
    uint64_t fh = ~0ull;
    if (fasta[0] == '>') fh = 0;
    else {
        const uint64_t lim = std::min<uint64_t>(n_bytes, 1u << 20);
        for (uint64_t at = 0; at < lim;) {
            const char *nl = static_cast<const char *>(memchr(fasta + at, '\n', lim - at));
            if (!nl) break;
            at = uint64_t(nl - fasta) + 1;
            if (at < n_bytes && fasta[at] == '>') { fh = at; break; }
        }
    }
    if (fh == ~0ull) return false;
    plan.first_header = fh;
    plan.cut.push_back(0);
    const uint64_t window = 64u << 10;
    unsigned missed = 0;
    for (uint64_t target = seg; target + seg / 2 < n_bytes; target += seg) {
        const uint64_t at = std::max(target, plan.cut.back() + 1);
        const uint64_t end = std::min(n_bytes, at + window);
        const char *nl = at < end ? static_cast<const char *>(memchr(fasta + at, '\n', end - at)) : nullptr;
        if (nl && uint64_t(nl - fasta) + 1 < n_bytes) plan.cut.push_back(uint64_t(nl - fasta) + 1);
        else if (++missed >= 3 && plan.cut.size() < 4) return false;
    }
    plan.cut.push_back(n_bytes);
    plan.m = plan.cut.size() - 1;
    plan.slot.resize(plan.m + 1);
    for (uint64_t j = 0; j <= plan.m; ++j) plan.slot[j] = (plan.cut[j] + 63) / 64 * 64 + 64 * j;
    return true;
}

uint64_t slotted_pack_segment(const SlottedPlan &plan, const unsigned char *text, uint64_t n_bytes, uint64_t j,
                              uint64_t first, int k, uint32_t *codes, uint32_t *valid, uint64_t base)
{
    const uint64_t begin = plan.cut[j], end = plan.cut[j + 1];
    const bool open = begin > plan.first_header;
    const uint64_t n = fasta_pack_segment(text, n_bytes, begin, end, codes + (plan.slot[j] - base) / 16,
                                          valid + (plan.slot[j] - base) / 32, plan.slot[j + 1] - plan.slot[j], open);
    if (j > 0) {
        // the junction of this segment's cut: k - 1 positions before it, k - 1 behind it
        const uint64_t js = plan.junction_slot(first, j);
        SlotWriter w(codes + (js - base) / 16, valid + (js - base) / 32);
        const int ctx = k - 1;
        if (open && ctx > 0) {
            unsigned char before[KPAL_MAX_K], after[KPAL_MAX_K];
            context_before(text, begin, plan.first_header, ctx, before);
            context_after(text, begin, n_bytes, ctx, after);
            for (int i = ctx - 1; i >= 0; --i) w.push(before[i]);
            for (int i = 0; i < ctx; ++i) w.push(after[i]);
        }
        w.finish(64);
    }
    return n;
}
}  // namespace kpal

extern "C" int kpal_fasta_pack_segment(const char *fasta, uint64_t n_bytes, uint64_t begin, uint64_t end,
                                       uint32_t *codes, uint32_t *valid, uint64_t cap_bases,
                                       uint64_t *n_bases_out)
{
    if ((!fasta && n_bytes) || !codes || !valid || begin > end || end > n_bytes || (cap_bases & 63u) ||
        cap_bases < end - begin) {
        set_error("invalid argument: FASTA segment");
        return KPAL_EINVAL;
    }
    const uint64_t n = kpal::fasta_pack_segment((const unsigned char *)fasta, n_bytes, begin, end, codes, valid,
                                                cap_bases, false);
    if (n_bases_out) *n_bases_out = n;
    return KPAL_OK;
}

extern "C" uint64_t kpal_fasta_slotted_bases(uint64_t n_bytes, uint64_t seg_bytes)
{
    if (seg_bytes < 64) seg_bytes = 64;
    const uint64_t m = n_bytes / seg_bytes + 2;                  // upper bound of the segments
    return (n_bytes + 63) / 64 * 64 + 64 * m + 64 * m;           // slots + junction records
}

extern "C" int kpal_fasta_pack_slotted(const char *fasta, uint64_t n_bytes, int k, uint64_t seg_bytes,
                                       uint32_t *codes, uint32_t *valid, uint64_t *stream_bases_out)
{
    if (!fasta || !codes || !valid || !stream_bases_out || k < 1 || k > KPAL_MAX_K) {
        set_error("invalid argument: slotted FASTA pack");
        return KPAL_EINVAL;
    }
    std::call_once(g_lut_once, init_lut);
    kpal::SlottedPlan plan;
    if (!kpal::slotted_plan(fasta, n_bytes, seg_bytes, plan)) {
        set_error("the text cannot be cut into slots (no header line in its first MB, or lines above 64 KB)");
        return KPAL_EINVAL;
    }
    const uint64_t total = plan.stream_bases(0);
    const unsigned nw = std::max(1u, std::min<unsigned>(worker_count(n_bytes), unsigned(plan.m)));
    std::atomic<uint64_t> next{0};
    run_parallel(nw, [&](unsigned) {
        for (;;) {
            const uint64_t j = next.fetch_add(1);
            if (j >= plan.m) return;
            kpal::slotted_pack_segment(plan, (const unsigned char *)fasta, n_bytes, j, 0, k, codes, valid, 0);
        }
    });
    // segment 0 has no junction: its record slot stays invalid
    memset(codes + plan.junction_slot(0, 0) / 16, 0, 16);
    memset(valid + plan.junction_slot(0, 0) / 32, 0, 8);
    *stream_bases_out = total;
    return KPAL_OK;
}
