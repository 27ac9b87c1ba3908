// k-mer counting kernels for sm_100a.
//
// Replaces the per-base Python loop of Profile.from_sequences
// (reference kpal/klib.py:154-168) and Profile.balance (kpal/klib.py:285-298).
//
// Input is the packed stream described in include/kpal_b200.h: 2-bit codes
// (16 bases / u32, first base in the most significant bits) + 1-bit validity
// (32 bases / u32).  One thread consumes one 64-base chunk with a single
// 128-bit code load and a 64-bit validity load; the k-1 look-ahead bases come
// from the next lane by warp shuffle (lane 31 re-reads one halo word).
//
//   count_global_kernel   any k: RED.ADD into a table resident in L2 / HBM
//   count_smem_kernel     k <= 7: per-CTA privatised shared-memory histogram,
//                         lane-replicated for tiny k, flushed with RED
//   finalize_kernel       u32/u64 table -> int64 profile, fused with balance
//   by_record_kernel      one CTA per record: the row is built slab by slab in shared
//                         memory and streamed out once (RED rows for very long records)
#include "common.cuh"

namespace kpal {

// ---------------------------------------------------------------------------
// Per-chunk front end shared by all count kernels.
// ---------------------------------------------------------------------------
struct Chunk {
    uint32_t w[5];      // 64 bases of codes + 16 look-ahead bases
    uint64_t starts;    // bit (63 - o) set <=> the window starting at base o is all-valid
};

__device__ __forceinline__ void shl128(uint64_t hi, uint64_t lo, int s, uint64_t &ohi, uint64_t &olo)
{
    // 0 < s < 64
    ohi = (hi << s) | (lo >> (64 - s));
    olo = lo << s;
}

// `active` may differ between lanes, but every lane of the warp must call this
// (it shuffles).  codes/valid are padded so chunk+1 is always readable.
__device__ __forceinline__ Chunk load_chunk(const uint4 *__restrict__ codes,
                                            const uint2 *__restrict__ valid,
                                            uint64_t chunk, bool active, int k)
{
    Chunk c;
    uint4 cw = make_uint4(0, 0, 0, 0);
    uint2 vw = make_uint2(0, 0);
    if (active) {
        cw = __ldg(codes + chunk);
        vw = __ldg(valid + chunk);
    }
    const unsigned lane = threadIdx.x & 31u;
    uint32_t next_c = __shfl_down_sync(0xffffffffu, cw.x, 1);
    uint32_t next_v = __shfl_down_sync(0xffffffffu, vw.x, 1);
    if (lane == 31u && active) {   // halo of the warp's last lane: one extra word each
        next_c = __ldg(reinterpret_cast<const uint32_t *>(codes + chunk + 1));
        next_v = __ldg(reinterpret_cast<const uint32_t *>(valid + chunk + 1));
    }
    c.w[0] = cw.x; c.w[1] = cw.y; c.w[2] = cw.z; c.w[3] = cw.w; c.w[4] = next_c;

    // run mask: A_L(p) = AND_{t<L} V(p+t), built by the binary method on k.
    const uint64_t vhi = (uint64_t(vw.x) << 32) | vw.y;
    const uint64_t vlo = uint64_t(next_v) << 32;
    uint64_t ahi = vhi, alo = vlo;
    int len = 1;
    for (int b = 30 - __clz(k); b >= 0; --b) {       // bits of k below its top bit
        uint64_t shi, slo;
        shl128(ahi, alo, len, shi, slo);
        ahi &= shi; alo &= slo; len <<= 1;
        if ((k >> b) & 1) {
            shl128(vhi, vlo, len, shi, slo);
            ahi &= shi; alo &= slo; len += 1;
        }
    }
    c.starts = ahi;
    return c;
}

// k-mer index of the window starting at base o (compile-time) of the chunk.
template <int O>
__device__ __forceinline__ uint32_t window_index(const Chunk &c, int shift)
{
    constexpr int j = O / 16, r = O % 16;
    const uint32_t x = (r == 0) ? c.w[j] : __funnelshift_l(c.w[j + 1], c.w[j], 2 * r);
    return x >> shift;
}

template <int O, typename F>
__device__ __forceinline__ void for_each_window_from(const Chunk &c, int shift, F &&f)
{
    if constexpr (O < 64) {
        if ((c.starts >> (63 - O)) & 1ull) f(window_index<O>(c, shift));
        for_each_window_from<O + 1>(c, shift, f);
    }
}

template <typename F>
__device__ __forceinline__ void for_each_window(const Chunk &c, int shift, F &&f)
{
    for_each_window_from<0>(c, shift, f);
}

// ---------------------------------------------------------------------------
// Any k: RED.ADD into the global table (L2 resident for k <= 12 with u32).
// ---------------------------------------------------------------------------
template <typename CounterT>
__global__ void __launch_bounds__(256)
count_global_kernel(const uint4 *__restrict__ codes, const uint2 *__restrict__ valid,
                    uint64_t n_chunks, int k, CounterT *__restrict__ table)
{
    const int shift = 32 - 2 * k;
    const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    // warp-uniform trip count so the halo shuffle always has 32 participants
    const uint64_t first = uint64_t(blockIdx.x) * blockDim.x + (threadIdx.x & ~31u);
    for (uint64_t base = first; base < n_chunks; base += stride) {
        const uint64_t chunk = base + (threadIdx.x & 31u);
        const Chunk c = load_chunk(codes, valid, chunk, chunk < n_chunks, k);
        for_each_window(c, shift, [&](uint32_t idx) { atomicAdd(table + idx, CounterT(1)); });
    }
}

// ---------------------------------------------------------------------------
// k <= 7: shared-memory privatised histogram.  `rep` lane-indexed copies of the
// histogram (layout [bin][copy], copy = lane % rep) remove same-address
// serialisation for tiny k (k=1: 4 bins would otherwise take 8-way conflicts).
// ---------------------------------------------------------------------------
template <typename CounterT>
__global__ void __launch_bounds__(256)
count_smem_kernel(const uint4 *__restrict__ codes, const uint2 *__restrict__ valid,
                  uint64_t n_chunks, int k, int rep_log2, CounterT *__restrict__ table)
{
    extern __shared__ uint32_t hist[];
    const uint32_t bins = 1u << (2 * k);
    const uint32_t words = bins << rep_log2;
    for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) hist[i] = 0;
    __syncthreads();

    const int shift = 32 - 2 * k;
    const uint32_t copy = threadIdx.x & ((1u << rep_log2) - 1u);
    const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    const uint64_t first = uint64_t(blockIdx.x) * blockDim.x + (threadIdx.x & ~31u);
    for (uint64_t base = first; base < n_chunks; base += stride) {
        const uint64_t chunk = base + (threadIdx.x & 31u);
        const Chunk c = load_chunk(codes, valid, chunk, chunk < n_chunks, k);
        for_each_window(c, shift, [&](uint32_t idx) {
            atomicAdd(&hist[(idx << rep_log2) + copy], 1u);
        });
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < bins; b += blockDim.x) {
        uint32_t s = 0;
        for (uint32_t r = 0; r < (1u << rep_log2); ++r) s += hist[(b << rep_log2) + r];
        if (s) atomicAdd(table + b, CounterT(s));
    }
}

// ---------------------------------------------------------------------------
// table -> int64 profile, fused with balance: out[i] = t[i] + t[rc(i)]
// (pairs get the sum, palindromes are doubled: kpal/klib.py:293-298).
// ---------------------------------------------------------------------------
// OutT = int64_t *: the API's profile.  OutT = NarrowOut: the forms the host entry points
// move over PCIe -- bins below `split` as uint16 and (o8 != nullptr) also as uint8, a
// quarter / an eighth of the bytes, widened to int64 by host threads while the copy is
// still running; bins from `split` on as int64 (o64[i - split]), copied by the DMA engine
// straight into the caller's array while the host threads are busy widening.  The host
// copies the narrowest form that holds every count below `split`: flags[0] is raised by a
// count above 65535 (the caller then redoes the finalize in int64), flags[1] by one above 255.
// A count that does not fit its narrow form is ALSO appended to a side list (index, value):
// a profile with a few large counts (poly-A, satellites) still travels narrow and the host
// patches the listed bins after widening.  flags[2] / flags[3] count the entries of the two
// lists (for the uint8 form: counts above 255; for the uint16 form: above 65535); entries
// beyond a list's capacity are dropped, the host then takes the next wider form.
struct NarrowEntry { unsigned long long index, value; };
struct NarrowOut {
    uint16_t *o16;
    uint8_t *o8;
    int64_t *o64;
    uint64_t split;
    unsigned int *flags;
    NarrowEntry *list8, *list16;
    unsigned int cap8, cap16;
};

__device__ __forceinline__ void put_count(int64_t *__restrict__ out, uint64_t i, unsigned long long v)
{
    out[i] = int64_t(v);
}
__device__ __forceinline__ void put_count(const NarrowOut &out, uint64_t i, unsigned long long v)
{
    if (i >= out.split) { out.o64[i - out.split] = int64_t(v); return; }
    if (v > 0xffull) {                          // same value from every writer: a plain store is enough
        out.flags[1] = 1u;
        if (out.list8) {
            const unsigned int at = atomicAdd(out.flags + 2, 1u);
            if (at < out.cap8) out.list8[at] = NarrowEntry{i, v};
        }
        if (v > 0xffffull) {
            out.flags[0] = 1u;
            if (out.list16) {
                const unsigned int at = atomicAdd(out.flags + 3, 1u);
                if (at < out.cap16) out.list16[at] = NarrowEntry{i, v};
            }
        }
    }
    out.o16[i] = uint16_t(v);
    if (out.o8) out.o8[i] = uint8_t(v);
}

template <typename CounterT, typename OutT>
__global__ void __launch_bounds__(256)
finalize_kernel(const CounterT *__restrict__ table, int k, int balance, const OutT out)
{
    const uint32_t n = 1u << (2 * k);
    const int shift = 32 - 2 * k;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        unsigned long long v = table[i];
        if (balance) v += __ldg(table + rc_index(i, shift));
        put_count(out, i, v);
    }
}

// Balanced finalize without the scattered gather.  Index = [h : 3 bases][m : k-6
// bases][l : 3 bases]; rc(index) = [rc(l)][rc(m)][rc(h)].  A CTA owns the 64 x 64
// tile of one m together with the tile of rc(m): both are read once in 256-byte
// rows, transposed through shared memory and written once as int64 -- table
// read once, profile written once (the plain kernel reads 4 scattered bytes
// per 32-byte sector for the partner).
// two neighbouring counts: one 16-byte store for the int64 profile
__device__ __forceinline__ void put_count2(int64_t *__restrict__ out, uint64_t i, unsigned long long v0,
                                           unsigned long long v1)
{
    *reinterpret_cast<ulonglong2 *>(out + i) = make_ulonglong2(v0, v1);
}
__device__ __forceinline__ void put_count2(const NarrowOut &out, uint64_t i, unsigned long long v0,
                                           unsigned long long v1)
{
    put_count(out, i, v0);
    put_count(out, i + 1, v1);
}

template <typename CounterT, typename OutT>
__global__ void __launch_bounds__(256)
finalize_balance_tiled_kernel(const CounterT *__restrict__ table, int k, const OutT out)
{
    extern __shared__ __align__(16) unsigned char fin_smem[];
    CounterT *A = reinterpret_cast<CounterT *>(fin_smem);      // [64][65] tile of m
    CounterT *B = A + 64 * 65;                                  // [64][65] tile of rc(m)
    const uint32_t m = blockIdx.x;
    const int mid_bits = 2 * (k - 6);
    const uint32_t mr = mid_bits ? ((~rev2(m)) >> (32 - mid_bits)) : 0u;
    if (m > mr) return;                                         // done by the CTA of rc(m)
    const int hshift = 2 * k - 6;
    // Both tiles are read with 16-byte loads, ALL of them issued before the first use (8 per
    // thread with 32-bit counters: 32 KB in flight per CTA) -- the first form of this kernel
    // waited for one 4-byte load per loop trip and ran at a third of the HBM rate.
    constexpr int V = 16 / int(sizeof(CounterT));               // counters per 16-byte vector
    constexpr int NV = 4096 / V / 256;                          // vectors per thread and tile
    uint4 va[NV], vb[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        const uint32_t v = threadIdx.x + 256u * q, h = v / (64 / V), l = (v % (64 / V)) * V;
        va[q] = *reinterpret_cast<const uint4 *>(table + ((uint64_t(h) << hshift) | (uint64_t(m) << 6) | l));
    }
    if (m != mr) {
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const uint32_t v = threadIdx.x + 256u * q, h = v / (64 / V), l = (v % (64 / V)) * V;
            vb[q] = *reinterpret_cast<const uint4 *>(table + ((uint64_t(h) << hshift) | (uint64_t(mr) << 6) | l));
        }
    }
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        const uint32_t v = threadIdx.x + 256u * q, h = v / (64 / V), l = (v % (64 / V)) * V;
        const CounterT *ea = reinterpret_cast<const CounterT *>(&va[q]), *eb = reinterpret_cast<const CounterT *>(&vb[q]);
#pragma unroll
        for (int j = 0; j < V; ++j) {
            A[h * 65 + l + j] = ea[j];
            if (m != mr) B[h * 65 + l + j] = eb[j];
        }
    }
    __syncthreads();
    const CounterT *partner = (m != mr) ? B : A;
    for (uint32_t e = threadIdx.x; e < 2048; e += 256) {        // two neighbouring counts per thread
        const uint32_t h = e >> 5, l = (e & 31u) * 2;
        const uint32_t rh = rc_index(h, 26);
        const uint32_t t0 = rc_index(l, 26) * 65 + rh, t1 = rc_index(l + 1, 26) * 65 + rh;
        put_count2(out, (uint64_t(h) << hshift) | (uint64_t(m) << 6) | l,
                   (unsigned long long)(A[h * 65 + l]) + (unsigned long long)(partner[t0]),
                   (unsigned long long)(A[h * 65 + l + 1]) + (unsigned long long)(partner[t1]));
        if (m != mr)
            put_count2(out, (uint64_t(h) << hshift) | (uint64_t(mr) << 6) | l,
                       (unsigned long long)(B[h * 65 + l]) + (unsigned long long)(A[t0]),
                       (unsigned long long)(B[h * 65 + l + 1]) + (unsigned long long)(A[t1]));
    }
}

// ---------------------------------------------------------------------------
// Per-record profiles (Profile.from_fasta_by_record, kpal/klib.py:114-133).
// The output is one dense int64 row of 4^k counts per record: the path is bound by
// the HBM write of the rows (8 * 4^k bytes per record).
//
// by_record_kernel: one CTA per record builds the row slab by slab in shared memory
// (a slab = up to 32768 bins as 16-bit counters, 64 KB -> three CTAs per SM, so the
// zero / count / write-out phases of different records overlap): zero the slab, count
// the record's windows that fall into it with shared-memory atomics, then stream the slab
// out widened to int64 -- every row byte is written exactly once, with full-line
// streaming stores, and nothing is read back.  Records whose counts could exceed 16 bits
// (65535 windows or more; half that with balance) take the RED path below instead.
//
// by_record_red_row (the first implementation, kept for such records): zero-fill the
// row in global memory, then RED the windows into it.  With ~1000 CTAs in flight the
// rows (512 KB each at k = 8) leave the L2 before the REDs arrive, which then cost a
// sector read-modify-write each: measured 3.6 TB/s of row writes, 55 % of the copy peak.
// ---------------------------------------------------------------------------
constexpr int kRowThreads = 512;
constexpr uint32_t kSlabBins = 32768;

template <typename F>
__device__ __forceinline__ void for_record_windows(const uint4 *__restrict__ codes, const uint2 *__restrict__ valid,
                                                   uint64_t b0, uint64_t b1, int k, F &&f)
{
    const int shift = 32 - 2 * k;
    const uint64_t c0 = b0 / kChunkBases, c1 = (b1 + kChunkBases - 1) / kChunkBases;
    for (uint64_t base = c0 + (threadIdx.x & ~31u); base < c1; base += blockDim.x) {
        const uint64_t chunk = base + (threadIdx.x & 31u);
        Chunk c = load_chunk(codes, valid, chunk, chunk < c1, k);
        // keep only windows that start inside this record
        const uint64_t p0 = chunk * kChunkBases;
        uint64_t keep = ~0ull;
        if (p0 < b0) keep &= (b0 - p0 >= 64) ? 0ull : (~0ull >> (b0 - p0));
        if (p0 + 64 > b1) keep &= (b1 <= p0) ? 0ull : ~(~0ull >> (b1 - p0));
        c.starts &= keep;
        for_each_window(c, shift, f);
    }
}

// The same walk with the windows of every chunk spread over the warp: a chunk's words are
// broadcast from the lane that loaded it and lane l takes the windows starting at bases l and
// l + 32.  A read of a few hundred bases occupies only a few lanes of one warp in the walk
// above (64 windows in sequence per lane); here its windows are binned 32 at a time.
template <typename F>
__device__ __forceinline__ void for_record_windows_spread(const uint4 *__restrict__ codes,
                                                          const uint2 *__restrict__ valid,
                                                          uint64_t b0, uint64_t b1, int k, F &&f)
{
    const int shift = 32 - 2 * k;
    const unsigned lane = threadIdx.x & 31u;
    const bool upper = lane >= 16u;
    const int rot = 2 * int(lane & 15u);
    const uint64_t c0 = b0 / kChunkBases, c1 = (b1 + kChunkBases - 1) / kChunkBases;
    for (uint64_t base = c0 + (threadIdx.x & ~31u); base < c1; base += blockDim.x) {
        const uint64_t chunk = base + lane;
        Chunk c = load_chunk(codes, valid, chunk, chunk < c1, k);
        const uint64_t p0 = chunk * kChunkBases;
        uint64_t keep = ~0ull;
        if (p0 < b0) keep &= (b0 - p0 >= 64) ? 0ull : (~0ull >> (b0 - p0));
        if (p0 + 64 > b1) keep &= (b1 <= p0) ? 0ull : ~(~0ull >> (b1 - p0));
        c.starts &= keep;
        unsigned todo = __ballot_sync(0xffffffffu, c.starts != 0ull);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1u;
            const uint32_t w0 = __shfl_sync(0xffffffffu, c.w[0], src), w1 = __shfl_sync(0xffffffffu, c.w[1], src);
            const uint32_t w2 = __shfl_sync(0xffffffffu, c.w[2], src), w3 = __shfl_sync(0xffffffffu, c.w[3], src);
            const uint32_t w4 = __shfl_sync(0xffffffffu, c.w[4], src);
            const uint32_t s_hi = __shfl_sync(0xffffffffu, uint32_t(c.starts >> 32), src);
            const uint32_t s_lo = __shfl_sync(0xffffffffu, uint32_t(c.starts), src);
            if ((s_hi >> (31u - lane)) & 1u)                                       // window at base `lane`
                f(__funnelshift_l(upper ? w2 : w1, upper ? w1 : w0, rot) >> shift);
            if ((s_lo >> (31u - lane)) & 1u)                                       // ... at base 32 + `lane`
                f(__funnelshift_l(upper ? w4 : w3, upper ? w3 : w2, rot) >> shift);
        }
    }
}

__device__ __forceinline__ void by_record_red_row(const uint4 *__restrict__ codes, const uint2 *__restrict__ valid,
                                                  uint64_t b0, uint64_t b1, int k, int balance,
                                                  unsigned long long *__restrict__ row)
{
    const uint64_t bins = 1ull << (2 * k);
    const int shift = 32 - 2 * k;
    // 1. zero fill (16-byte stores; bins*8 is a multiple of 16 for k >= 1)
    ulonglong2 *row2 = reinterpret_cast<ulonglong2 *>(row);
    for (uint64_t i = threadIdx.x; i < bins / 2; i += blockDim.x)
        row2[i] = make_ulonglong2(0ull, 0ull);
    __threadfence();
    __syncthreads();
    // 2. windows of bases [b0, b1)
    for_record_windows(codes, valid, b0, b1, k, [&](uint32_t idx) {
        atomicAdd(row + idx, 1ull);
        if (balance) atomicAdd(row + rc_index(idx, shift), 1ull);
    });
    __syncthreads();
}

// RowT = int64_t: the API's rows.  RowT = uint16_t: the slab itself, for the host entry point's
// narrow copy (kpal_count_by_record checks that no record of the call can overflow 16 bits).
template <bool RED_ONLY, typename RowT>
__global__ void __launch_bounds__(kRowThreads)
by_record_kernel(const uint4 *__restrict__ codes, const uint2 *__restrict__ valid,
                 const uint64_t *__restrict__ rec_starts, uint64_t first_rec, uint64_t n_rec,
                 int k, int balance, RowT *__restrict__ rows)
{
    extern __shared__ __align__(16) uint32_t slab[];            // 16-bit counters, two per word
    const uint64_t bins = 1ull << (2 * k);
    const int shift = 32 - 2 * k;
    const uint32_t slab_bins = bins < kSlabBins ? uint32_t(bins) : kSlabBins;
    for (uint64_t r = blockIdx.x; r < n_rec; r += gridDim.x) {
        const uint64_t b0 = rec_starts[first_rec + r], b1 = rec_starts[first_rec + r + 1];
        if constexpr (sizeof(RowT) == 8) {
            if (RED_ONLY || (b1 - b0) >= (balance ? 32768ull : 65536ull)) {      // CTA-uniform
                by_record_red_row(codes, valid, b0, b1, k, balance,
                                  reinterpret_cast<unsigned long long *>(rows + r * bins));
                continue;
            }
        }
        for (uint64_t s0 = 0; s0 < bins; s0 += slab_bins) {
            if (slab_bins >= 8) {
                for (uint32_t i = threadIdx.x * 4; i < slab_bins / 2; i += kRowThreads * 4)
                    *reinterpret_cast<uint4 *>(slab + i) = make_uint4(0, 0, 0, 0);
            } else if (threadIdx.x < 2) {
                slab[threadIdx.x] = 0;
            }
            __syncthreads();
            for_record_windows_spread(codes, valid, b0, b1, k, [&](uint32_t idx) {
                const uint32_t a = idx - uint32_t(s0);
                if (a < slab_bins) atomicAdd(slab + (a >> 1), 1u << (16 * (a & 1)));
                if (balance) {
                    const uint32_t b = rc_index(idx, shift) - uint32_t(s0);
                    if (b < slab_bins) atomicAdd(slab + (b >> 1), 1u << (16 * (b & 1)));
                }
            });
            __syncthreads();
            // write-out: consecutive lanes write consecutive 16 bytes (two int64 counts from one
            // shared word), so every store instruction of a warp covers 512 contiguous bytes
            if constexpr (sizeof(RowT) == 8) {
                ulonglong2 *out = reinterpret_cast<ulonglong2 *>(rows + r * bins + s0);
                for (uint32_t i = threadIdx.x; i < slab_bins / 2; i += kRowThreads) {
                    const uint32_t w = slab[i];
                    __stcs(out + i, make_ulonglong2(w & 0xffffu, w >> 16));
                }
            } else {                                    // 4^k >= 16: whole 16-byte pieces
                uint4 *out = reinterpret_cast<uint4 *>(rows + r * bins + s0);
                for (uint32_t i = threadIdx.x; i < slab_bins / 8; i += kRowThreads)
                    __stcs(out + i, *reinterpret_cast<const uint4 *>(slab + 4 * i));
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(256)
balance_i64_kernel(const int64_t *__restrict__ in, int k, int64_t *__restrict__ out)
{
    const uint32_t n = 1u << (2 * k);
    const int shift = 32 - 2 * k;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        out[i] = in[i] + __ldg(in + rc_index(i, shift));
}

template <typename CounterT>
__global__ void __launch_bounds__(256)
accumulate_kernel(const CounterT *__restrict__ src, CounterT *__restrict__ dst, uint64_t n)
{
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += uint64_t(gridDim.x) * blockDim.x)
        dst[i] += src[i];
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
// count_radix.cu
bool radix_supported(int k);
int launch_count_radix(const uint32_t *, const uint32_t *, uint64_t, int, void *, int, cudaStream_t,
                       const PeerOut *peer = nullptr);
bool radix_peer_supported(int k, int world);
// count_pairs.cu
bool pairs_supported(int k);
int launch_count_pairs(const uint32_t *, const uint32_t *, uint64_t, int, void *, int, cudaStream_t);
// peer_reduce.cu
int launch_reduce_push(const void *, int, int, int, int, void *const *, cudaStream_t);
int peer_check_args(int k, int counter_bits, int rank, int world);

// run-time switches (kpal_set_option): "count_path" 0 = automatic, 1 = always the
// scattered-RED kernel, 2 = the radix-partitioned path wherever it is supported (two windows
// per payload for k <= 12, count_pairs.cu), 3 = the one-window radix path (count_radix.cu);
// "tiled_finalize" 0 = plain gather kernel for the balanced finalize.
static std::atomic<int> g_count_path{0};
static std::atomic<int> g_tiled_finalize{1};
static std::atomic<int> g_by_record_path{0};        // 0 = shared-memory slabs, 1 = zero-fill + RED rows
void set_by_record_path(int v) { g_by_record_path.store(v); }
void set_count_path(int v) { g_count_path.store(v); }
void set_tiled_finalize(int v) { g_tiled_finalize.store(v); }

static bool use_radix_path(int k, uint64_t n_bases)
{
    const int mode = g_count_path.load();
    if (mode == 1 || !radix_supported(k)) return false;
    if (mode >= 2) return true;
    // The two passes cost a fixed ~2 x 4^k x 4 bytes of table traffic; below these sizes
    // the RED kernel wins.  From k = 13 on the table no longer fits in L2 and the RED
    // rate drops ~7x, so the switch comes earlier relative to the table size.
    const uint64_t threshold = k <= 12 ? (16ull << 20) : (4ull << 20) << (2 * (k - 13));
    return n_bases >= threshold;
}

int launch_accumulate(const void *d_src, void *d_dst, int counter_bits, uint64_t n, cudaStream_t stream)
{
    uint64_t want = (n + 255) / 256;
    const uint64_t cap = uint64_t(sm_count()) * 32;
    const unsigned grid = unsigned(want < cap ? want : cap);
    if (counter_bits == 32)
        accumulate_kernel<uint32_t><<<grid, 256, 0, stream>>>(static_cast<const uint32_t *>(d_src),
                                                              static_cast<uint32_t *>(d_dst), n);
    else
        accumulate_kernel<unsigned long long><<<grid, 256, 0, stream>>>(
            static_cast<const unsigned long long *>(d_src), static_cast<unsigned long long *>(d_dst), n);
    KPAL_LAUNCH_CHECK("accumulate_kernel");
    return KPAL_OK;
}

static int check_k(int k)
{
    if (k < 1 || k > KPAL_MAX_K) {
        set_error("k-mer length %d out of range [1, %d]", k, KPAL_MAX_K);
        return KPAL_EINVAL;
    }
    return KPAL_OK;
}

// zero_table: the table is zeroed first (a memset on the stream) instead of by the caller.
// (Folding the memset into the first count kernel -- every CTA zeroing its share while it
// bins -- was built and measured: 0.2703 vs 0.2706 ms per step, no gain, so it went again.)
int launch_count(const uint32_t *d_codes, const uint32_t *d_valid, uint64_t n_bases, int k,
                 void *d_table, int counter_bits, cudaStream_t stream, bool zero_table)
{
    KPAL_CHECK(check_k(k));
    if (counter_bits != 32 && counter_bits != 64) return bad_arg("counter_bits must be 32 or 64");
    const bool pairs = n_bases > 0 && k > 7 && use_radix_path(k, n_bases) && pairs_supported(k) &&
                       g_count_path.load() != 3 && !(counter_bits == 32 && n_bases >= (1ull << 32));
    if (zero_table)
        KPAL_CUDA(cudaMemsetAsync(d_table, 0, (size_t(1) << (2 * k)) * size_t(counter_bits / 8), stream));
    if (n_bases == 0) return KPAL_OK;
    if (counter_bits == 32 && n_bases >= (1ull << 32)) {
        set_error("%llu bases would overflow 32-bit counters; use counter_bits=64",
                  (unsigned long long)n_bases);
        return KPAL_EOVERFLOW;
    }
    const uint64_t n_chunks = n_chunks_of(n_bases);
    const uint4 *codes = reinterpret_cast<const uint4 *>(d_codes);
    const uint2 *valid = reinterpret_cast<const uint2 *>(d_valid);
    const int sms = sm_count();
    if (k <= 7) {
        // copies: as many as fit in 64 KB, at most one per lane
        int rep_log2 = 0;
        while (rep_log2 < 5 && ((4ull << (2 * k)) << (rep_log2 + 1)) <= 65536ull) ++rep_log2;
        const size_t smem = (size_t(4) << (2 * k)) << rep_log2;
        uint64_t want = (n_chunks + 255) / 256;
        const unsigned grid = unsigned(want < uint64_t(sms) * 3 ? want : uint64_t(sms) * 3);
        if (counter_bits == 32) {
            KPAL_CUDA(cudaFuncSetAttribute(count_smem_kernel<uint32_t>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
            count_smem_kernel<uint32_t><<<grid, 256, smem, stream>>>(
                codes, valid, n_chunks, k, rep_log2, static_cast<uint32_t *>(d_table));
        } else {
            KPAL_CUDA(cudaFuncSetAttribute(count_smem_kernel<unsigned long long>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
            count_smem_kernel<unsigned long long><<<grid, 256, smem, stream>>>(
                codes, valid, n_chunks, k, rep_log2, static_cast<unsigned long long *>(d_table));
        }
        KPAL_LAUNCH_CHECK("count_smem_kernel");
    } else if (use_radix_path(k, n_bases)) {
        if (pairs)
            return launch_count_pairs(d_codes, d_valid, n_bases, k, d_table, counter_bits, stream);
        return launch_count_radix(d_codes, d_valid, n_bases, k, d_table, counter_bits, stream);
    } else {
        uint64_t want = (n_chunks + 255) / 256;
        const uint64_t cap = uint64_t(sms) * 8 * 4;      // a few waves of 8 CTAs/SM
        const unsigned grid = unsigned(want < cap ? want : cap);
        if (counter_bits == 32)
            count_global_kernel<uint32_t><<<grid, 256, 0, stream>>>(
                codes, valid, n_chunks, k, static_cast<uint32_t *>(d_table));
        else
            count_global_kernel<unsigned long long><<<grid, 256, 0, stream>>>(
                codes, valid, n_chunks, k, static_cast<unsigned long long *>(d_table));
        KPAL_LAUNCH_CHECK("count_global_kernel");
    }
    return KPAL_OK;
}

template <typename OutT>
static int launch_finalize_as(const void *d_table, int counter_bits, int k, int balance, const OutT d_out,
                              cudaStream_t stream)
{
    KPAL_CHECK(check_k(k));
    if (counter_bits != 32 && counter_bits != 64) return bad_arg("counter_bits must be 32 or 64");
    const uint64_t n = 1ull << (2 * k);
    uint64_t want = (n + 255) / 256;
    const uint64_t cap = uint64_t(sm_count()) * 32;
    const unsigned grid = unsigned(want < cap ? want : cap);
    if (balance && k >= 6 && g_tiled_finalize.load()) {
        const unsigned tiles = 1u << (2 * (k - 6));
        const size_t smem = size_t(2) * 64 * 65 * (counter_bits / 8);
        if (counter_bits == 32) {
            finalize_balance_tiled_kernel<uint32_t, OutT><<<tiles, 256, smem, stream>>>(
                static_cast<const uint32_t *>(d_table), k, d_out);
        } else {
            KPAL_CUDA(cudaFuncSetAttribute(finalize_balance_tiled_kernel<unsigned long long, OutT>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
            finalize_balance_tiled_kernel<unsigned long long, OutT><<<tiles, 256, smem, stream>>>(
                static_cast<const unsigned long long *>(d_table), k, d_out);
        }
        KPAL_LAUNCH_CHECK("finalize_balance_tiled_kernel");
        return KPAL_OK;
    }
    if (counter_bits == 32)
        finalize_kernel<uint32_t, OutT><<<grid, 256, 0, stream>>>(
            static_cast<const uint32_t *>(d_table), k, balance, d_out);
    else
        finalize_kernel<unsigned long long, OutT><<<grid, 256, 0, stream>>>(
            static_cast<const unsigned long long *>(d_table), k, balance, d_out);
    KPAL_LAUNCH_CHECK("finalize_kernel");
    return KPAL_OK;
}

int launch_finalize(const void *d_table, int counter_bits, int k, int balance, int64_t *d_counts,
                    cudaStream_t stream)
{
    return launch_finalize_as<int64_t *>(d_table, counter_bits, k, balance, d_counts, stream);
}

// Narrow forms for the host entry points (see NarrowOut): bins [0, split) as uint16 and
// optionally uint8, bins [split, 4^k) as int64 into d_tail64 (split = 4^k: none);
// d_flags[2] are device words the caller zeroed.
int launch_finalize_narrow(const void *d_table, int counter_bits, int k, int balance, uint16_t *d_counts16,
                           uint8_t *d_counts8, int64_t *d_tail64, uint64_t split, unsigned int *d_flags,
                           cudaStream_t stream, void *d_list8, unsigned int cap8, void *d_list16, unsigned int cap16)
{
    if (!d_flags || !d_counts16) return bad_arg("null narrow buffer");
    if (split > (1ull << (2 * k)) || (split < (1ull << (2 * k)) && !d_tail64)) return bad_arg("bad split");
    NarrowOut out;
    out.o16 = d_counts16; out.o8 = d_counts8; out.o64 = d_tail64; out.split = split; out.flags = d_flags;
    out.list8 = d_counts8 ? static_cast<NarrowEntry *>(d_list8) : nullptr; out.cap8 = cap8;
    out.list16 = static_cast<NarrowEntry *>(d_list16); out.cap16 = cap16;
    return launch_finalize_as<NarrowOut>(d_table, counter_bits, k, balance, out, stream);
}

int launch_balance(const int64_t *d_in, int64_t *d_out, int k, cudaStream_t stream)
{
    KPAL_CHECK(check_k(k));
    if (d_in == d_out) return bad_arg("kpal_dev_balance needs in != out");
    const uint64_t n = 1ull << (2 * k);
    uint64_t want = (n + 255) / 256;
    const uint64_t cap = uint64_t(sm_count()) * 32;
    balance_i64_kernel<<<unsigned(want < cap ? want : cap), 256, 0, stream>>>(d_in, k, d_out);
    KPAL_LAUNCH_CHECK("balance_i64_kernel");
    return KPAL_OK;
}

int launch_by_record(const uint32_t *d_codes, const uint32_t *d_valid, const uint64_t *d_rec_starts,
                     uint64_t first, uint64_t n, int k, int balance, int64_t *d_rows,
                     cudaStream_t stream)
{
    KPAL_CHECK(check_k(k));
    if (n == 0) return KPAL_OK;
    const uint64_t bins = 1ull << (2 * k);
    const size_t smem = size_t(bins < kSlabBins ? bins : kSlabBins) * 2;
    const auto cd = reinterpret_cast<const uint4 *>(d_codes);
    const auto vd = reinterpret_cast<const uint2 *>(d_valid);
    if (g_by_record_path.load() == 1) {                 // the RED path for every record (A/B measurements)
        const uint64_t cap = uint64_t(sm_count()) * 4;
        by_record_kernel<true, int64_t><<<unsigned(n < cap ? n : cap), kRowThreads, 0, stream>>>(
            cd, vd, d_rec_starts, first, n, k, balance, d_rows);
    } else {
        KPAL_CUDA(cudaFuncSetAttribute(by_record_kernel<false, int64_t>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        // resident CTAs per SM: three 64 KB slabs fit, and at most 2048 threads
        const uint64_t per_sm = smem > 32768 ? 3 : 4;
        const uint64_t cap = uint64_t(sm_count()) * per_sm;
        by_record_kernel<false, int64_t><<<unsigned(n < cap ? n : cap), kRowThreads, smem, stream>>>(
            cd, vd, d_rec_starts, first, n, k, balance, d_rows);
    }
    KPAL_LAUNCH_CHECK("by_record_kernel");
    return KPAL_OK;
}

// Rows as uint16 (k >= 2).  The caller guarantees that no record of [first, first + n) has
// 65536 bases or more (32768 with balance): the counts then fit, see by_record_kernel.
int launch_by_record_u16(const uint32_t *d_codes, const uint32_t *d_valid, const uint64_t *d_rec_starts,
                         uint64_t first, uint64_t n, int k, int balance, uint16_t *d_rows,
                         cudaStream_t stream)
{
    KPAL_CHECK(check_k(k));
    if (k < 2) return bad_arg("uint16 rows need k >= 2");
    if (n == 0) return KPAL_OK;
    const uint64_t bins = 1ull << (2 * k);
    const size_t smem = size_t(bins < kSlabBins ? bins : kSlabBins) * 2;
    KPAL_CUDA(cudaFuncSetAttribute(by_record_kernel<false, uint16_t>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    const uint64_t cap = uint64_t(sm_count()) * (smem > 32768 ? 3 : 4);
    by_record_kernel<false, uint16_t><<<unsigned(n < cap ? n : cap), kRowThreads, smem, stream>>>(
        reinterpret_cast<const uint4 *>(d_codes), reinterpret_cast<const uint2 *>(d_valid),
        d_rec_starts, first, n, k, balance, d_rows);
    KPAL_LAUNCH_CHECK("by_record_kernel");
    return KPAL_OK;
}

}  // namespace kpal

namespace kpal {

// Multi-GPU: count this rank's stream and deliver every slice of the resulting table
// (d_table must be zeroed by the caller, as for launch_count) to the inbox of its owner.
// On the radix path the all-to-all is fused into pass 2 (the histogram kernel stores
// into the inboxes); otherwise the table is counted locally and pushed by a copy kernel.
int launch_count_push(const uint32_t *d_codes, const uint32_t *d_valid, uint64_t n_bases, int k,
                      void *d_table, int counter_bits, int rank, int world,
                      void *const *inbox_ptrs, cudaStream_t stream, int *fused_out)
{
    KPAL_CHECK(check_k(k));
    KPAL_CHECK(peer_check_args(k, counter_bits, rank, world));
    if (!inbox_ptrs) return bad_arg("null pointer");
    if (counter_bits == 32 && n_bases >= (1ull << 32)) {
        set_error("%llu bases would overflow 32-bit counters; use counter_bits=64", (unsigned long long)n_bases);
        return KPAL_EOVERFLOW;
    }
    const bool fused = n_bases > 0 && k > 7 && use_radix_path(k, n_bases) && radix_peer_supported(k, world);
    if (fused_out) *fused_out = fused ? 1 : 0;
    if (fused) {
        PeerOut peer;
        peer.rank = rank; peer.world = world;
        for (int i = 0; i < kMaxPeers; ++i) peer.inbox[i] = i < world ? inbox_ptrs[i] : nullptr;
        for (int i = 0; i < world; ++i) if (!peer.inbox[i]) return bad_arg("null inbox pointer");
        return launch_count_radix(d_codes, d_valid, n_bases, k, d_table, counter_bits, stream, &peer);
    }
    KPAL_CHECK(launch_count(d_codes, d_valid, n_bases, k, d_table, counter_bits, stream, false));
    return launch_reduce_push(d_table, counter_bits, k, rank, world, inbox_ptrs, stream);
}

}  // namespace kpal
