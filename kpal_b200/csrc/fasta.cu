// FASTA text -> packed 2-bit stream ON THE DEVICE.
//
// Replaces, for the end-to-end counting path, the host work the reference
// does in Bio.SeqIO.parse + str(record.seq) (kpal/klib.py:111): the raw file
// bytes are uploaded once and classified / compacted by the GPU, so the host
// never touches the sequence (the C++ packer in pack.cpp stays as the general
// path: per-record output, sequence lists, and the rare inputs this fast path
// refuses, see below).
//
// FastaIterator rules implemented (same as pack.cpp / oracle parse_fasta):
//   * a line whose first byte is '>' is a header: dropped, and it emits ONE
//     invalid base (the record separator);
//   * bytes before the first header line are dropped;
//   * on sequence lines '\n', ' ' and '\r' are dropped, every other byte is a
//     base (valid iff ACGTacgt).
// Python's rstrip() also drops *trailing* tabs / VT / FF / FS-US while keeping
// them (as k-mer splitting bytes) in the middle of a line; that needs a
// look-ahead to the end of the line, so such bytes raise a flag instead and
// the caller re-packs the file on the host (bit-exactness is kept either way).
//
// Three streaming passes over the text (each tile = 4096 bytes = one CTA of
// 256 threads x 16 bytes):
//   1. per tile: position of the last '\n', first header position (atomicMin)
//   2. per tile: number of emitted bases (needs the line type of every byte:
//      last newline before it -> first byte of its line)
//   3. per tile: exclusive offsets -> write codes / valid bits (atomicOr on the
//      at most two words a thread's 16 bases straddle)
// with two single-CTA scans over the per-tile values in between.
#include "common.cuh"

namespace kpal {

constexpr int kTileThreads = 256;
constexpr int kBytesPerThread = 16;
constexpr int kTileBytes = kTileThreads * kBytesPerThread;

struct FastaScratch {            // device scalars
    unsigned long long first_header;   // position of the first header line ('>' at line start)
    unsigned long long total_bases;
    unsigned int flags;                // bit 0: exotic whitespace seen on a sequence line
    unsigned int pad;
    long long carry_last_nl;           // last newline before the next tile range (chunked packing)
};

__device__ __forceinline__ int64_t block_reduce_max(int64_t v, int64_t *smem)
{
    for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = v;
    __syncthreads();
    int64_t r = smem[0];
    for (int w = 1; w < kTileThreads / 32; ++w) r = max(r, smem[w]);
    __syncthreads();
    return r;
}

__device__ __forceinline__ void load16(const uint8_t *text, uint64_t n, uint64_t pos, uint8_t (&b)[16])
{
    if (pos + 16 <= n) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(text + pos));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 16; ++i) b[i] = uint8_t(w[i >> 2] >> (8 * (i & 3)));
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) b[i] = (pos + i < n) ? __ldg(text + pos + i) : uint8_t(' ');
    }
}

// pass 1
__global__ void __launch_bounds__(kTileThreads)
fasta_newlines_kernel(const uint8_t *__restrict__ text, uint64_t n, uint64_t tile0,
                      int64_t *__restrict__ tile_last_nl, FastaScratch *__restrict__ sc)
{
    __shared__ int64_t red[kTileThreads / 32];
    const uint64_t tile = tile0 + blockIdx.x;
    const uint64_t pos = tile * kTileBytes + threadIdx.x * kBytesPerThread;
    uint8_t b[16];
    load16(text, n, pos, b);
    int64_t last = -1;
    unsigned long long first_hdr = ~0ull;
    uint8_t prev = (pos == 0 || pos > n) ? uint8_t('\n') : __ldg(text + pos - 1);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        if (pos + i < n) {
            if (b[i] == '\n') last = int64_t(pos + i);
            if (b[i] == '>' && prev == '\n' && first_hdr == ~0ull) first_hdr = pos + i;
        }
        prev = b[i];
    }
    // first header of the file: one guarded atomic per tile (a read FASTA has a header every
    // few lines; one atomicMin per header on a single address serialised the whole pass)
    const int64_t fh = -block_reduce_max(first_hdr == ~0ull ? INT64_MIN + 1 : -int64_t(first_hdr), red);
    const int64_t m = block_reduce_max(last, red);
    if (threadIdx.x == 0) {
        tile_last_nl[tile] = m;
        if (fh != INT64_MAX && (unsigned long long)fh < *(volatile unsigned long long *)&sc->first_header)
            atomicMin(&sc->first_header, (unsigned long long)fh);
    }
}

// exclusive running max over the tiles (single CTA)
__global__ void __launch_bounds__(1024)
scan_max_kernel(const int64_t *__restrict__ in, int64_t *__restrict__ out, uint64_t tile0,
                uint64_t n_tiles, FastaScratch *__restrict__ sc)
{
    __shared__ int64_t warp_max[32];
    __shared__ int64_t carry_s;
    if (threadIdx.x == 0) carry_s = tile0 ? sc->carry_last_nl : -1;
    __syncthreads();
    for (uint64_t base = tile0; base < n_tiles; base += blockDim.x) {
        const uint64_t i = base + threadIdx.x;
        const int64_t v = i < n_tiles ? in[i] : -1;
        int64_t inc = v;                                   // inclusive scan in the warp
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if ((threadIdx.x & 31) >= o) inc = max(inc, t);
        }
        if ((threadIdx.x & 31) == 31) warp_max[threadIdx.x >> 5] = inc;
        __syncthreads();
        int64_t before = carry_s;
        for (unsigned w = 0; w < (threadIdx.x >> 5); ++w) before = max(before, warp_max[w]);
        int64_t excl = __shfl_up_sync(0xffffffffu, inc, 1);
        if ((threadIdx.x & 31) == 0) excl = -1;
        if (i < n_tiles) out[i] = max(before, excl);
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry_s = max(before, inc);
        __syncthreads();
    }
    if (threadIdx.x == 0) sc->carry_last_nl = carry_s;
}

// exclusive running sum over the tiles (single CTA); also stores the total
__global__ void __launch_bounds__(1024)
scan_sum_kernel(const uint32_t *__restrict__ in, unsigned long long *__restrict__ out, uint64_t tile0,
                uint64_t n_tiles, FastaScratch *__restrict__ sc)
{
    __shared__ unsigned long long warp_sum[32];
    __shared__ unsigned long long carry_s;
    if (threadIdx.x == 0) carry_s = tile0 ? sc->total_bases : 0;
    __syncthreads();
    for (uint64_t base = tile0; base < n_tiles; base += blockDim.x) {
        const uint64_t i = base + threadIdx.x;
        const unsigned long long v = i < n_tiles ? in[i] : 0ull;
        unsigned long long inc = v;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
            if ((threadIdx.x & 31) >= o) inc += t;
        }
        if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = inc;
        __syncthreads();
        unsigned long long before = carry_s;
        for (unsigned w = 0; w < (threadIdx.x >> 5); ++w) before += warp_sum[w];
        if (i < n_tiles) out[i] = before + inc - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry_s = before + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) sc->total_bases = carry_s;
}

// Classify the 16 bytes of this thread.  Returns the emit mask (bit i = byte i
// emits a base), fills codes (2 bits per EMITTED base, first emitted base in the
// most significant bits) and valid (1 bit per emitted base, same order).
__device__ __forceinline__ uint32_t classify16(const uint8_t *__restrict__ text, uint64_t n,
                                               uint64_t pos, int64_t prev_nl, uint64_t first_header,
                                               uint32_t &codes, uint32_t &valid, bool &exotic)
{
    uint8_t b[16];
    load16(text, n, pos, b);
    // type of the line we start in: first byte of that line
    int64_t line_start = prev_nl + 1;
    bool header = (uint64_t(line_start) < n) && (uint64_t(line_start) < pos
                      ? __ldg(text + line_start) == '>' : false);
    bool at_line_start = (uint64_t(line_start) == pos);
    uint32_t mask = 0, c = 0, v = 0;
    int emitted = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const uint64_t p = pos + i;
        const uint8_t ch = b[i];
        if (p >= n) break;
        if (at_line_start) { header = (ch == '>'); }
        bool emit = false, ok = false;
        uint32_t code = 0;
        if (ch == '\n') {
            at_line_start = true;
        } else {
            if (header) {
                emit = at_line_start && p >= first_header;          // the '>' itself: record separator
            } else if (p > first_header && ch != ' ' && ch != '\r') {
                emit = true;
                const uint8_t u = ch & 0xDF;                          // upper case
                ok = (u == 'A') | (u == 'C') | (u == 'G') | (u == 'T');
                code = (u == 'C') ? 1u : (u == 'G') ? 2u : (u == 'T') ? 3u : 0u;
                if (ch == 9 || ch == 11 || ch == 12 || (ch >= 28 && ch <= 31)) exotic = true;
            }
            at_line_start = false;
        }
        if (emit) {
            mask |= 1u << i;
            if (ok) {
                c |= code << (30 - 2 * emitted);
                v |= 1u << (15 - emitted);
            }
            ++emitted;
        }
    }
    codes = c;
    valid = v;
    return mask;
}

// last newline strictly before this thread's first byte (block-level exclusive max scan)
__device__ __forceinline__ int64_t thread_prev_nl(const uint8_t *__restrict__ text, uint64_t n,
                                                  uint64_t pos, int64_t tile_prev, int64_t *smem)
{
    uint8_t b[16];
    load16(text, n, pos, b);
    int64_t last = -1;
#pragma unroll
    for (int i = 0; i < 16; ++i)
        if (pos + i < n && b[i] == '\n') last = int64_t(pos + i);
    int64_t inc = last;
    for (int o = 1; o < 32; o <<= 1) {
        const int64_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if ((threadIdx.x & 31) >= o) inc = max(inc, t);
    }
    if ((threadIdx.x & 31) == 31) smem[threadIdx.x >> 5] = inc;
    __syncthreads();
    int64_t before = tile_prev;
    for (unsigned w = 0; w < (threadIdx.x >> 5); ++w) before = max(before, smem[w]);
    int64_t excl = __shfl_up_sync(0xffffffffu, inc, 1);
    if ((threadIdx.x & 31) == 0) excl = -1;
    __syncthreads();
    return max(before, excl);
}

// pass 2
__global__ void __launch_bounds__(kTileThreads)
fasta_count_kernel(const uint8_t *__restrict__ text, uint64_t n, uint64_t tile0,
                   const int64_t *__restrict__ tile_prev_nl, uint32_t *__restrict__ tile_count,
                   FastaScratch *__restrict__ sc)
{
    __shared__ int64_t red[kTileThreads / 32];
    __shared__ uint32_t sums[kTileThreads / 32];
    const uint64_t tile = tile0 + blockIdx.x;
    const uint64_t pos = tile * kTileBytes + threadIdx.x * kBytesPerThread;
    const int64_t prev = thread_prev_nl(text, n, pos, tile_prev_nl[tile], red);
    uint32_t codes, valid;
    bool exotic = false;
    const uint32_t mask = classify16(text, n, pos, prev, sc->first_header, codes, valid, exotic);
    if (exotic) atomicOr(&sc->flags, 1u);
    uint32_t cnt = __popc(mask);
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) sums[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < kTileThreads / 32; ++w) t += sums[w];
        tile_count[tile] = t;
    }
}

// pass 3
__global__ void __launch_bounds__(kTileThreads)
fasta_write_kernel(const uint8_t *__restrict__ text, uint64_t n, uint64_t tile0,
                   const int64_t *__restrict__ tile_prev_nl,
                   const unsigned long long *__restrict__ tile_base, const FastaScratch *__restrict__ sc,
                   uint32_t *__restrict__ out_codes, uint32_t *__restrict__ out_valid)
{
    __shared__ int64_t red[kTileThreads / 32];
    __shared__ uint32_t sums[kTileThreads / 32];
    const uint64_t tile = tile0 + blockIdx.x;
    const uint64_t pos = tile * kTileBytes + threadIdx.x * kBytesPerThread;
    const int64_t prev = thread_prev_nl(text, n, pos, tile_prev_nl[tile], red);
    uint32_t codes, valid;
    bool exotic = false;
    const uint32_t mask = classify16(text, n, pos, prev, sc->first_header, codes, valid, exotic);
    const uint32_t cnt = __popc(mask);
    uint32_t inc = cnt;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if ((threadIdx.x & 31) >= o) inc += t;
    }
    if ((threadIdx.x & 31) == 31) sums[threadIdx.x >> 5] = inc;
    __syncthreads();
    uint32_t before = 0;
    for (unsigned w = 0; w < (threadIdx.x >> 5); ++w) before += sums[w];
    if (cnt == 0) return;
    const uint64_t o = tile_base[tile] + before + (inc - cnt);    // first output base of this thread
    // codes: 32 bits (16 bases) left aligned, placed at base offset o
    {
        const uint64_t w = o / 16;
        const unsigned s = 2 * unsigned(o % 16);
        atomicOr(out_codes + w, codes >> s);
        if (s && (codes << (32 - s))) atomicOr(out_codes + w + 1, codes << (32 - s));
    }
    if (valid) {
        const uint32_t v32 = valid << 16;                                // 16 bits left aligned in 32
        const uint64_t w = o / 32;
        const unsigned s = unsigned(o % 32);
        atomicOr(out_valid + w, v32 >> s);
        if (s > 16 && (v32 << (32 - s))) atomicOr(out_valid + w + 1, v32 << (32 - s));
    }
}

// ---------------------------------------------------------------------------
// launcher.  Scratch layout (device, caller-provided, fasta_scratch_bytes()):
//   FastaScratch | tile_last_nl[int64] | tile_prev_nl[int64] | tile_count[u32] | tile_base[u64]
// ---------------------------------------------------------------------------
uint64_t fasta_scratch_bytes(uint64_t n_bytes)
{
    const uint64_t tiles = (n_bytes + kTileBytes - 1) / kTileBytes + 1;
    return 64 + tiles * (8 + 8 + 8 + 8);
}

// Zero the outputs and the scalars: once per text, before the first tile range.
int launch_fasta_pack_begin(uint64_t n_bytes, uint32_t *d_codes, uint32_t *d_valid, void *d_scratch,
                            cudaStream_t stream)
{
    // output capacity: one base per input byte (kpal_packed_words(n_bytes))
    uint64_t cw, vw;
    kpal_packed_words(n_bytes, &cw, &vw);
    KPAL_CUDA(cudaMemsetAsync(d_codes, 0, cw * 4, stream));
    KPAL_CUDA(cudaMemsetAsync(d_valid, 0, vw * 4, stream));
    FastaScratch *sc = static_cast<FastaScratch *>(d_scratch);
    KPAL_CUDA(cudaMemsetAsync(sc, 0, sizeof(FastaScratch), stream));
    KPAL_CUDA(cudaMemsetAsync(sc, 0xff, sizeof(unsigned long long), stream));    // first_header = ~0
    return KPAL_OK;
}

uint64_t fasta_tile_bytes() { return kTileBytes; }

// Pack the bytes of tiles [tile0, tile1) (tile = 4096 bytes).  Ranges must be submitted in
// order on one stream; only bytes below tile1 * 4096 are read, so a range can run as soon
// as its part of the text has arrived (the host-level entry points overlap the H2D copy of
// chunk c+1 with the packing of chunk c this way).
//
// layout_bytes (0 = n_bytes): the text length the scratch arrays were laid out for, when
// n_bytes -- the end of the text as the kernels see it -- is only settled with the last range
// (hybrid upload: the device packs the head of the text up to wherever the host packers got).
int launch_fasta_pack_tiles(const uint8_t *d_text, uint64_t n_bytes, uint64_t tile0, uint64_t tile1,
                            uint32_t *d_codes, uint32_t *d_valid, void *d_scratch, cudaStream_t stream,
                            uint64_t layout_bytes)
{
    const uint64_t tiles_now = (n_bytes + kTileBytes - 1) / kTileBytes;
    if (layout_bytes < n_bytes) layout_bytes = n_bytes;
    const uint64_t tiles = (layout_bytes + kTileBytes - 1) / kTileBytes;
    if (tile1 > tiles_now) tile1 = tiles_now;
    if (tile0 >= tile1) return KPAL_OK;
    if (tiles > 0x7fffffffull) return bad_arg("FASTA text too large for one launch");
    FastaScratch *sc = static_cast<FastaScratch *>(d_scratch);
    unsigned char *base = static_cast<unsigned char *>(d_scratch) + 64;
    int64_t *tile_last = reinterpret_cast<int64_t *>(base);
    int64_t *tile_prev = tile_last + tiles + 1;
    unsigned long long *tile_base = reinterpret_cast<unsigned long long *>(tile_prev + tiles + 1);
    uint32_t *tile_count = reinterpret_cast<uint32_t *>(tile_base + tiles + 1);
    const unsigned grid = unsigned(tile1 - tile0);

    fasta_newlines_kernel<<<grid, kTileThreads, 0, stream>>>(d_text, n_bytes, tile0, tile_last, sc);
    KPAL_LAUNCH_CHECK("fasta_newlines_kernel");
    scan_max_kernel<<<1, 1024, 0, stream>>>(tile_last, tile_prev, tile0, tile1, sc);
    KPAL_LAUNCH_CHECK("scan_max_kernel");
    fasta_count_kernel<<<grid, kTileThreads, 0, stream>>>(d_text, n_bytes, tile0, tile_prev, tile_count, sc);
    KPAL_LAUNCH_CHECK("fasta_count_kernel");
    scan_sum_kernel<<<1, 1024, 0, stream>>>(tile_count, tile_base, tile0, tile1, sc);
    KPAL_LAUNCH_CHECK("scan_sum_kernel");
    fasta_write_kernel<<<grid, kTileThreads, 0, stream>>>(d_text, n_bytes, tile0, tile_prev, tile_base,
                                                          sc, d_codes, d_valid);
    KPAL_LAUNCH_CHECK("fasta_write_kernel");
    return KPAL_OK;
}

int launch_fasta_pack(const uint8_t *d_text, uint64_t n_bytes, uint32_t *d_codes, uint32_t *d_valid,
                      void *d_scratch, cudaStream_t stream)
{
    KPAL_CHECK(launch_fasta_pack_begin(n_bytes, d_codes, d_valid, d_scratch, stream));
    return launch_fasta_pack_tiles(d_text, n_bytes, 0, ~0ull, d_codes, d_valid, d_scratch, stream, 0);
}

}  // namespace kpal
