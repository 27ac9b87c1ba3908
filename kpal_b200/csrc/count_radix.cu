// Two-pass radix-partitioned k-mer counting for large tables (9 <= k <= 15).
//
// Same contract as count_global_kernel (count.cu): accumulate the windows of a
// packed stream (kpal/klib.py:154-168) into a table of 4^k counters.  A scattered
// RED per window tops out at the L2 atomic rate (~0.2 T/s, profiles/r01_microbench)
// and collapses once the table leaves L2 (k = 13); shared-memory atomics run
// at ~1.6 T/s.  So the index space is cut into `nb` buckets by the high index
// bits, the low `P` bits (the payload, a u16) are staged bucket-major in HBM,
// and every bucket is then histogrammed in shared memory:
//
//   pass 1  radix_partition_kernel   one persistent CTA per SM walks its contiguous
//           share of the stream in tiles of 1024 x 32 windows.  A window takes one
//           returning shared-memory atomic (its rank in the bucket's slot) and one
//           16-bit shared store.  After each tile every slot writes its complete
//           32-byte groups (16 payloads) to the CTA's private region of that bucket;
//           the remainder (< 16) stays in the slot for the next tile, so every global
//           store is a full, aligned sector.
//   pass 2  radix_histogram_kernel   one CTA per bucket: shared-memory histogram of 2^P
//           bins over the bucket's regions (streamed 16 bytes per lane), then one
//           coalesced read-modify-write of the bucket's slice of the table.
//
// Anything that does not fit -- a slot that overflows inside a tile (skewed or
// repetitive sequence) or a region that fills up -- is counted with a plain
// RED on the table instead, so the result is exact for every input; only the
// speed depends on the distribution.
//
// Staging memory: 3 x the mean region size, i.e. ~6 bytes per base, owned by a
// grow-only per-device workspace (HBM is 180 GB; a 100 Mbp call takes 0.6 GB).
#include "radix_common.cuh"

#include <mutex>
#include <vector>

namespace kpal {

constexpr int kHistThreadsMax = 1024;

// ---------------------------------------------------------------------------
// pass 1
// ---------------------------------------------------------------------------
struct RadixParams {
    const uint2 *codes;         // 32 bases per uint2
    const uint32_t *valid;      // 32 bases per word
    uint64_t unit_begin, unit_end;   // units [begin, end) of this launch
    uint64_t n_units;           // units readable in the stream (loads are clamped to this)
    int k, P, nb, cap;          // payload bits, buckets (power of two), slot capacity
    int bucket_lo, nb_active;   // this launch bins buckets [bucket_lo, bucket_lo + nb_active) only
    uint32_t region_groups;     // capacity of one (CTA, bucket) region in groups
    uint16_t *staging;          // [grid][nb][region_groups * 16]
    uint32_t *region_fill;      // [grid][nb] payloads stored per region
    int debug;                  // timing experiments only: 1 = no flush, 2 = no slot stores, 4 = no atomics
};

struct BinCtx {
    uint32_t cnt_s, slots_s;        // shared addresses of cnt[] and slots[]
    uint32_t dummy_cnt_s;           // per-lane counter that absorbs invalid windows
    uint32_t dummy_slot_s;          // per-lane halfword that absorbs their stores
    uint32_t cap;
    uint32_t bucket_lo, nb_active;
    int shift, P;
    bool no_store;
};

// Eight windows at a time, branch-free: an invalid window increments a per-lane
// dummy counter and stores into a per-lane dummy halfword, so the atomics of a
// batch are all in flight before the first dependent store.  The stored
// halfword is the low 16 index bits; pass 2 masks it to P bits.
template <typename CounterT, bool SWEEP, int O0>
__device__ __forceinline__ void bin_eight(const Unit &u, const BinCtx &c, CounterT *table)
{
    // all selects are bit masks (0 / ~0): with 16 live booleans the compiler ran out of
    // predicate registers and spent more instructions shuffling them than on the work
    uint32_t idx[8], rank[8], okm[8];
    idx[0] = unit_window<O0 + 0>(u, c.shift); idx[1] = unit_window<O0 + 1>(u, c.shift);
    idx[2] = unit_window<O0 + 2>(u, c.shift); idx[3] = unit_window<O0 + 3>(u, c.shift);
    idx[4] = unit_window<O0 + 4>(u, c.shift); idx[5] = unit_window<O0 + 5>(u, c.shift);
    idx[6] = unit_window<O0 + 6>(u, c.shift); idx[7] = unit_window<O0 + 7>(u, c.shift);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        okm[j] = uint32_t(int32_t(u.starts << (O0 + j)) >> 31);
        uint32_t bl = idx[j] >> c.P;
        if constexpr (SWEEP) {          // several launches share the buckets: keep ours only
            bl -= c.bucket_lo;
            okm[j] &= bl < c.nb_active ? ~0u : 0u;
        }
        const uint32_t real = c.cnt_s + 4u * bl;
        rank[j] = atoms_add(c.dummy_cnt_s ^ ((c.dummy_cnt_s ^ real) & okm[j]), 1u);
    }
    uint32_t overflow = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        // rank < cap  <=>  rank - cap is negative (ranks stay far below 2^31)
        const uint32_t inm = uint32_t(int32_t(rank[j] - c.cap) >> 31) & okm[j];
        overflow |= okm[j] & ~inm;
        const uint32_t bl = SWEEP ? (idx[j] >> c.P) - c.bucket_lo : (idx[j] >> c.P);
        const uint32_t real = c.slots_s + 2u * (bl * c.cap + rank[j]);
        if (!c.no_store) sts_u16(c.dummy_slot_s ^ ((c.dummy_slot_s ^ real) & inm), idx[j]);
    }
    if (overflow) {                 // slot full (skewed / repetitive sequence): count directly
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (okm[j] && rank[j] >= c.cap) atomicAdd(table + idx[j], CounterT(1));
    }
}

template <typename CounterT, bool SWEEP, int O0>
__device__ __forceinline__ void bin_from(const Unit &u, const BinCtx &c, CounterT *table)
{
    if constexpr (O0 < kUnitBases) {
        bin_eight<CounterT, SWEEP, O0>(u, c, table);
        bin_from<CounterT, SWEEP, O0 + 8>(u, c, table);
    }
}

// rare path: count the (first n of the 8) payloads of one 16-byte piece directly
template <typename CounterT>
__device__ __noinline__ void red_piece(uint4 a, uint32_t hi, uint32_t pmask, int n,
                                       CounterT *__restrict__ table)
{
    const uint32_t v[4] = {a.x, a.y, a.z, a.w};
    for (int j = 0; j < n; ++j) {
        const uint32_t e = (v[j >> 1] >> (16 * (j & 1))) & pmask;
        atomicAdd(table + (hi | e), CounterT(1));
    }
}

// THREADS x TEAM: 1024 x 8 = one CTA per SM with ~200 KB of slots; 512 x 4 = two CTAs per
// SM with ~105 KB each, so that the store-bound flush of one CTA overlaps the
// shared-memory-bound binning of the other (the flush moves 64 KB per tile through the
// SM's store path and costs as much as a third of the binning when it runs alone).
template <typename CounterT, int THREADS, int TEAM, bool SWEEP>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS)
radix_partition_kernel(const RadixParams p, CounterT *__restrict__ table)
{
    constexpr int kRadixThreads = THREADS;
    extern __shared__ __align__(16) unsigned char radix_smem[];
    const int na = p.nb_active;                                          // buckets binned by this launch
    uint32_t *cnt = reinterpret_cast<uint32_t *>(radix_smem);           // [na] payloads in the slot (+32 dummies)
    uint32_t *fillg = cnt + na + 32;                                     // [na] groups already stored
    uint16_t *slots = reinterpret_cast<uint16_t *>(fillg + na);          // [na][cap] (+ 64 B pad / dummies)

    const int tid = threadIdx.x;
    const unsigned lane = tid & 31u;
    for (int b = tid; b < na + 32; b += kRadixThreads) cnt[b] = 0;
    for (int b = tid; b < na; b += kRadixThreads) fillg[b] = 0;
    __syncthreads();

    // contiguous share of the units, a multiple of the tile so warps stay aligned
    const uint64_t total = p.unit_end - p.unit_begin;
    uint64_t per = (total + gridDim.x - 1) / gridDim.x;
    per = (per + kRadixThreads - 1) / kRadixThreads * kRadixThreads;
    const uint64_t u0 = p.unit_begin + uint64_t(blockIdx.x) * per;
    const uint64_t u1 = (u0 + per < p.unit_end) ? u0 + per : p.unit_end;

    BinCtx ctx;
    ctx.cnt_s = smem_u32(cnt);
    ctx.slots_s = smem_u32(slots);
    ctx.dummy_cnt_s = ctx.cnt_s + 4u * (uint32_t(na) + lane);
    ctx.dummy_slot_s = ctx.slots_s + 2u * (uint32_t(na) * uint32_t(p.cap) + lane);
    ctx.cap = uint32_t(p.cap);
    ctx.shift = 32 - 2 * p.k;
    ctx.P = p.P;
    ctx.bucket_lo = uint32_t(p.bucket_lo);
    ctx.nb_active = uint32_t(p.nb_active);
    ctx.no_store = (p.debug & 2) != 0;

    uint16_t *my_regions = p.staging + uint64_t(blockIdx.x) * p.nb * p.region_groups * kGroup;
    const int team = tid / TEAM, tl = tid % TEAM;      // flush teams of TEAM consecutive lanes

    // software pipeline: the words of the next tile (and lane 31's halo words) are in
    // flight during this one
    uint2 cw_next = make_uint2(0, 0);
    uint32_t vw_next = 0, hc_next = 0, hv_next = 0;
    auto prefetch = [&](uint64_t unit, bool tile_exists) {
        cw_next = make_uint2(0, 0); vw_next = 0; hc_next = 0; hv_next = 0;
        if (tile_exists && unit < p.n_units) {
            cw_next = __ldg(p.codes + unit);
            vw_next = __ldg(p.valid + unit);
            if (lane == 31u) {                  // the stream is padded by one 64-base chunk
                hc_next = __ldg(reinterpret_cast<const uint32_t *>(p.codes + unit + 1));
                hv_next = __ldg(p.valid + unit + 1);
            }
        }
    };
    prefetch(u0 + tid, u0 < u1);

    // flush bookkeeping of this thread's team (bucket = team + 128 * iteration)
    const uint32_t fill_s = smem_u32(fillg);
    const uint32_t slot_bytes = uint32_t(p.cap) * 2u;
    uint4 *const my_regions4 = reinterpret_cast<uint4 *>(my_regions);
    const uint32_t region_v4 = p.region_groups * 2u;            // 16-byte pieces per region

    for (uint64_t t0 = u0; t0 < u1; t0 += kRadixThreads) {
        // ---- A: bin this tile's windows into the bucket slots
        const uint64_t unit = t0 + tid;
        const uint2 cw = cw_next;
        const uint32_t vw = vw_next;
        uint32_t next_c = __shfl_down_sync(0xffffffffu, cw.x, 1);
        uint32_t next_v = __shfl_down_sync(0xffffffffu, vw, 1);
        if (lane == 31u) { next_c = hc_next; next_v = hv_next; }
        prefetch(unit + kRadixThreads, t0 + kRadixThreads < u1);
        Unit u;
        u.w[0] = cw.x; u.w[1] = cw.y; u.w[2] = next_c;
        {   // run mask by the binary method on k (as load_chunk in count.cu)
            const uint64_t v = (uint64_t(vw) << 32) | next_v;
            uint64_t a = v;
            int len = 1;
            for (int bit = 30 - __clz(p.k); bit >= 0; --bit) {
                a &= a << len; len <<= 1;
                if ((p.k >> bit) & 1) { a &= v << len; len += 1; }
            }
            u.starts = (unit < u1) ? uint32_t(a >> 32) : 0u;
        }
        if (u.starts && !(p.debug & 4)) bin_from<CounterT, SWEEP, 0>(u, ctx, table);
        __syncthreads();

        // ---- B: every slot stores its complete groups and keeps the remainder.  A bucket
        // is flushed by a team of TEAM consecutive lanes, one 16-byte piece per lane, so a
        // warp store covers a few contiguous runs instead of 32 scattered sectors (the L1
        // takes one cycle per distinct line of a store instruction).  The team is
        // inside one warp: the slot bookkeeping needs no CTA barrier.
        if (p.debug & 1) { for (int b = tid; b < na; b += kRadixThreads) cnt[b] = 0; }
        else for (int b = team; b - team < na; b += kRadixThreads / TEAM) {     // warp-uniform trip count
            const bool have = b < na;
            const uint32_t bg = uint32_t(b + p.bucket_lo);                      // bucket id in the table
            const uint32_t cnt_a = ctx.cnt_s + 4u * uint32_t(b), fill_a = fill_s + 4u * uint32_t(b);
            uint32_t n = 0, f = 0;
            if (have) { n = min(lds_u32(cnt_a), ctx.cap); f = lds_u32(fill_a); }
            __syncwarp();
            const uint32_t g = n / kGroup;
            if (g) {
                const uint32_t slot_a = ctx.slots_s + uint32_t(b) * slot_bytes;
                const uint32_t dst0 = bg * region_v4 + 2u * f;
                for (uint32_t piece = tl; piece < 2 * g; piece += TEAM) {
                    const uint4 x = lds_v4(slot_a + 16u * piece);
                    if (f + piece / 2 < p.region_groups) __stcs(my_regions4 + dst0 + piece, x);
                    else red_piece<CounterT>(x, bg << p.P, (1u << p.P) - 1u, 8, table);   // region full
                }
                // remainder to the front: lanes 0/1 read pieces 0/1 themselves, nobody else does
                if (tl < 2) { const uint4 x = lds_v4(slot_a + 16u * (2 * g + tl)); sts_v4(slot_a + 16u * tl, x); }
                if (tl == 0) { sts_u32(cnt_a, n - g * kGroup); sts_u32(fill_a, min(f + g, p.region_groups)); }
            }
        }
        __syncthreads();
    }

    // ---- remainders (< 16 per bucket) and the per-region totals
    for (int b0 = 0; b0 < na; b0 += kRadixThreads / TEAM) {
        const int b = b0 + team;
        if (b >= na || tl >= 2) continue;
        const uint32_t bg = uint32_t(b + p.bucket_lo);
        const uint32_t n = cnt[b], f = fillg[b];
        uint32_t stored = f * kGroup;
        if (n) {
            const uint4 x = reinterpret_cast<const uint4 *>(slots + uint32_t(b) * p.cap)[tl];
            if (f < p.region_groups) {
                uint4 *dst = reinterpret_cast<uint4 *>(my_regions + uint64_t(bg) * p.region_groups * kGroup);
                dst[2 * f + tl] = x;                     // payloads beyond n are never read
                stored += n;
            } else {
                const int mine = int(n) - 8 * tl;
                red_piece<CounterT>(x, bg << p.P, (1u << p.P) - 1u, mine < 8 ? mine : 8, table);
            }
        }
        if (tl == 0) p.region_fill[uint64_t(blockIdx.x) * p.nb + bg] = stored;
    }
}

// ---------------------------------------------------------------------------
// pass 2
// ---------------------------------------------------------------------------
// PEER: multi-GPU mode -- instead of updating the local table, the bucket's slice
// (local table + histogram) is stored straight into the inbox slot of the rank that
// owns it (16-byte peer stores over NVLink, see peer_reduce.cu), so the all-to-all of
// the table reduce rides on this kernel instead of following it.
template <typename CounterT, bool PEER>
__global__ void __launch_bounds__(kHistThreadsMax)
radix_histogram_kernel(const uint16_t *__restrict__ staging, const uint32_t *__restrict__ region_fill,
                       int n_part_ctas, int nb, int P, uint32_t region_groups,
                       CounterT *__restrict__ table, const PeerOut peer)
{
    extern __shared__ __align__(16) uint32_t radix_hist[];
    const uint32_t bins = 1u << P;
    const int b = blockIdx.x;
    for (uint32_t i = threadIdx.x * 4; i < bins; i += blockDim.x * 4)
        *reinterpret_cast<uint4 *>(radix_hist + i) = make_uint4(0, 0, 0, 0);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    const uint32_t pmask = bins - 1u;

    // A warp walks its regions (c = warp, warp + n_warps, ...) in chunks of 128 vectors of
    // 16 bytes; the loads of chunk j+1 are issued before the atomics of chunk j.
    int c = warp;
    uint32_t n = c < n_part_ctas ? __ldg(region_fill + uint64_t(c) * nb + b) : 0u;
    uint32_t n_ahead = c + n_warps < n_part_ctas ? __ldg(region_fill + uint64_t(c + n_warps) * nb + b) : 0u;
    uint32_t i0 = 0;
    struct Chunk4 { uint4 v[4]; uint32_t at, n; bool ok; };
    auto fetch = [&](Chunk4 &k) {
        while (c < n_part_ctas && i0 * 8u >= n) {           // next non-empty region
            c += n_warps;
            n = n_ahead; i0 = 0;
            n_ahead = c + n_warps < n_part_ctas ? __ldg(region_fill + uint64_t(c + n_warps) * nb + b) : 0u;
        }
        k.ok = c < n_part_ctas;
        if (!k.ok) return;
        const uint4 *src = reinterpret_cast<const uint4 *>(staging + (uint64_t(c) * nb + b) * region_groups * kGroup);
        const uint32_t nv = (n + 7u) / 8u;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t i = i0 + lane + 32 * q;
            k.v[q] = make_uint4(0, 0, 0, 0);
            if (i < nv) k.v[q] = __ldcs(src + i);
        }
        k.at = (i0 + lane) * 8u; k.n = n;
        i0 += 128;
    };
    Chunk4 cur, nxt;
    fetch(cur);
    while (cur.ok) {
        fetch(nxt);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t at = cur.at + 256u * q;
            const uint32_t rem = at < cur.n ? cur.n - at : 0u;
            const uint32_t w[4] = {cur.v[q].x, cur.v[q].y, cur.v[q].z, cur.v[q].w};
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (uint32_t(j) < rem) atomicAdd(&radix_hist[(w[j >> 1] >> (16 * (j & 1))) & pmask], 1u);
        }
        cur = nxt;
    }
    __syncthreads();

    // table slice += histogram, four independent 16-byte read-modify-writes in flight
    const CounterT *src = table + (uint64_t(b) << P);
    CounterT *dst = table + (uint64_t(b) << P);
    if constexpr (PEER) {
        // nb is a multiple of the world size: bucket b lies inside the slice of its owner
        const uint64_t total_bins = uint64_t(nb) << P;
        const int owner = int(uint64_t(b) * uint64_t(peer.world) / uint64_t(nb));
        dst = static_cast<CounterT *>(peer.inbox[owner]) + uint64_t(peer.rank) * slot_elems(total_bins, peer.world)
              + ((uint64_t(b) << P) - slice_begin(total_bins, owner, peer.world));
    }
    for (uint32_t i0 = threadIdx.x * 4; i0 < bins; i0 += blockDim.x * 16) {
        uint4 h[4];
        if constexpr (sizeof(CounterT) == 4) {
            uint4 t[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t i = i0 + q * blockDim.x * 4;
                if (i < bins) { h[q] = *reinterpret_cast<const uint4 *>(radix_hist + i); t[q] = *reinterpret_cast<const uint4 *>(src + i); }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t i = i0 + q * blockDim.x * 4;
                if (i < bins && (PEER || (h[q].x | h[q].y | h[q].z | h[q].w))) {
                    t[q].x += h[q].x; t[q].y += h[q].y; t[q].z += h[q].z; t[q].w += h[q].w;
                    *reinterpret_cast<uint4 *>(dst + i) = t[q];
                }
            }
        } else {
            ulonglong2 t0[4], t1[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t i = i0 + q * blockDim.x * 4;
                if (i < bins) {
                    h[q] = *reinterpret_cast<const uint4 *>(radix_hist + i);
                    t0[q] = *reinterpret_cast<const ulonglong2 *>(src + i);
                    t1[q] = *reinterpret_cast<const ulonglong2 *>(src + i + 2);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t i = i0 + q * blockDim.x * 4;
                if (i < bins && (PEER || (h[q].x | h[q].y | h[q].z | h[q].w))) {
                    t0[q].x += h[q].x; t0[q].y += h[q].y; t1[q].x += h[q].z; t1[q].y += h[q].w;
                    *reinterpret_cast<ulonglong2 *>(dst + i) = t0[q];
                    *reinterpret_cast<ulonglong2 *>(dst + i + 2) = t1[q];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// workspace + launcher
// ---------------------------------------------------------------------------
struct RadixWorkspace {
    int device = -1;
    void *staging = nullptr; size_t staging_cap = 0;
    uint32_t *fill = nullptr; size_t fill_cap = 0;
};
static std::mutex g_radix_mutex;
static std::vector<RadixWorkspace *> g_radix_ws;
static std::atomic<int> g_radix_payload_bits{0};     // 0 = automatic
static std::atomic<int> g_radix_max_buckets{2048};   // buckets binned by one pass-1 launch (1024 or 2048)
void set_radix_max_buckets(int v) { g_radix_max_buckets.store(v); }
static std::atomic<int> g_radix_shape{0};            // 0 = automatic
void set_radix_shape(int v) { g_radix_shape.store(v); }
static std::atomic<int> g_radix_debug{0};
void set_radix_debug(int v) { g_radix_debug.store(v); }

void set_radix_payload_bits(int bits) { g_radix_payload_bits.store(bits); }

static int grow(void **p, size_t *cap, size_t bytes)
{
    if (bytes <= *cap) return KPAL_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) {
        *p = nullptr; cudaGetLastError();
        set_error("cudaMalloc(%zu bytes) for the radix staging failed: %s", bytes, cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? KPAL_ENOMEM : KPAL_ECUDA;
    }
    *cap = bytes;
    return KPAL_OK;
}

// Staging buffers of the current device, grown on demand (shared with count_pairs.cu).
int radix_workspace(size_t staging_bytes, size_t fill_bytes, void **staging, uint32_t **fill)
{
    int dev = 0;
    KPAL_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_radix_mutex);
    RadixWorkspace *ws = nullptr;
    for (auto *w : g_radix_ws) if (w->device == dev) ws = w;
    if (!ws) { ws = new RadixWorkspace(); ws->device = dev; g_radix_ws.push_back(ws); }
    KPAL_CHECK(grow(&ws->staging, &ws->staging_cap, staging_bytes));
    void *f = ws->fill;
    size_t fc = ws->fill_cap;
    KPAL_CHECK(grow(&f, &fc, fill_bytes));
    ws->fill = static_cast<uint32_t *>(f); ws->fill_cap = fc;
    *staging = ws->staging; *fill = ws->fill;
    return KPAL_OK;
}

bool radix_supported(int k) { return k >= 9 && k <= KPAL_MAX_K; }

// Geometry for a given k: payload bits P, buckets nb = 4^k >> P, CTA shape and slot
// capacity cap (payloads; cap % 16 == 8 keeps the 16-byte slot reads of neighbouring
// buckets on distinct shared-memory banks).
struct RadixGeometry {
    int P, nb, na, cap, threads, team;      // na: buckets binned per pass-1 launch (nb / na sweeps)
    size_t smem1;
};
static RadixGeometry radix_geometry(int k)
{
    RadixGeometry g;
    int p = g_radix_payload_bits.load();
    if (p <= 0) p = (k >= 13) ? 15 : 2 * k - 9;         // 512 buckets up to k = 12, 2^(2k-15) beyond
    if (p > 15) p = 15;
    if (p < 2 * k - 15) p = 2 * k - 15;                 // at most 32768 buckets
    g.P = p;
    g.nb = 1 << (2 * k - p);
    // Up to 2048 buckets fit the slots (40 payloads each: a slot keeps < 16 from tile to tile
    // and gains 16 per tile on average, so ~2 % of the (slot, tile) pairs spill a few windows
    // to the RED path).  Beyond that pass 1 runs nb / 2048 times over the stream, each launch
    // binning its own buckets (the stream is 0.375 B/base; re-reading it is cheap next to
    // 5 B/base of RED traffic out of L2).  Measured at k = 13, 375 Mbp: one launch of 2048
    // buckets 1.06 ms for both passes, two launches of 1024 buckets 1.53 ms.
    const int max_buckets = g_radix_max_buckets.load();
    g.na = g.nb > max_buckets ? max_buckets : g.nb;
    int shape = g_radix_shape.load();                   // 1 = 1024 threads x 1 CTA/SM, 2 = 512 x 2
    if (shape == 0) shape = 1;      // measured: two half-size CTAs per SM do not beat one (147 vs 154 us)
    if (g.na > 512) shape = 1;
    g.threads = shape == 2 ? 512 : 1024;
    // shared memory per SM: 233472 B, minus 1 KB per resident CTA
    const size_t budget = (shape == 2 ? (233472 / 2 - 1024) : (233472 - 1024 - 4096)) - (size_t(g.na) * 8 + 192);
    int c = int(budget / (2u * unsigned(g.na)));
    c = (c - 8) / 16 * 16 + 8;
    if (c > 1032) c = 1032;
    g.cap = c;
    // lanes per flush team ~ 16-byte pieces a slot gains per tile (2 per 16 payloads)
    const int mean_pieces = g.threads * kUnitBases / g.na / 8;
    g.team = mean_pieces >= 8 ? 8 : 4;
    g.smem1 = size_t(g.na) * 8 + 128 + size_t(g.na) * c * 2 + 64;
    return g;
}

template <typename CounterT>
static int launch_radix_passes(const RadixGeometry &g, RadixParams &p, int grid1, size_t smem2, int threads2,
                               CounterT *table, const PeerOut *peer, cudaStream_t stream)
{
#define KPAL_RADIX_LAUNCH(THREADS, TEAM, SWEEP)                                                       \
    do {                                                                                              \
        KPAL_CUDA(cudaFuncSetAttribute(radix_partition_kernel<CounterT, THREADS, TEAM, SWEEP>,        \
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, int(g.smem1)));   \
        radix_partition_kernel<CounterT, THREADS, TEAM, SWEEP><<<grid1, THREADS, g.smem1, stream>>>(p, table); \
    } while (0)
    for (int lo = 0; lo < g.nb; lo += g.na) {
        p.bucket_lo = lo; p.nb_active = g.na;
        if (g.na < g.nb) KPAL_RADIX_LAUNCH(1024, 4, true);
        else if (g.threads == 512 && g.team == 4) KPAL_RADIX_LAUNCH(512, 4, false);
        else if (g.threads == 512) KPAL_RADIX_LAUNCH(512, 8, false);
        else if (g.team == 4) KPAL_RADIX_LAUNCH(1024, 4, false);
        else KPAL_RADIX_LAUNCH(1024, 8, false);
        KPAL_LAUNCH_CHECK("radix_partition_kernel");
    }
#undef KPAL_RADIX_LAUNCH
    if (peer) {
        KPAL_CUDA(cudaFuncSetAttribute(radix_histogram_kernel<CounterT, true>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem2)));
        radix_histogram_kernel<CounterT, true><<<g.nb, threads2, smem2, stream>>>(
            p.staging, p.region_fill, grid1, g.nb, g.P, p.region_groups, table, *peer);
    } else {
        KPAL_CUDA(cudaFuncSetAttribute(radix_histogram_kernel<CounterT, false>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem2)));
        radix_histogram_kernel<CounterT, false><<<g.nb, threads2, smem2, stream>>>(
            p.staging, p.region_fill, grid1, g.nb, g.P, p.region_groups, table, PeerOut());
    }
    KPAL_LAUNCH_CHECK("radix_histogram_kernel");
    return KPAL_OK;
}

// Can pass 2 write into the peers' inboxes?  Only when every bucket lies inside one
// owner's slice: the bucket count must be a multiple of the world size.
bool radix_peer_supported(int k, int world)
{
    if (!radix_supported(k) || world < 1 || world > kMaxPeers) return false;
    const RadixGeometry g = radix_geometry(k);
    return g.nb % world == 0 && ((uint64_t(1) << g.P) % 4) == 0;
}

// Accumulate units of the packed stream into `table` through the two passes.  With
// `peer` (multi-GPU), the last segment's pass 2 stores table + histogram into the
// owners' inboxes instead of the table.
int launch_count_radix(const uint32_t *d_codes, const uint32_t *d_valid, uint64_t n_bases, int k,
                       void *d_table, int counter_bits, cudaStream_t stream, const PeerOut *peer)
{
    const RadixGeometry g = radix_geometry(k);
    const int P = g.P, nb = g.nb;
    const int grid1 = sm_count() * (1024 / g.threads);
    const size_t smem2 = size_t(4) << P;
    const int threads2 = P >= 15 ? 1024 : 512;

    // the stream is walked in segments so that the staging stays bounded (<= ~3 GB)
    const uint64_t n_units = 2 * n_chunks_of(n_bases);
    const uint64_t seg_units = (512ull << 20) / kUnitBases;
    for (uint64_t s0 = 0; s0 < n_units; s0 += seg_units) {
        const uint64_t s1 = (s0 + seg_units < n_units) ? s0 + seg_units : n_units;
        uint64_t per = (s1 - s0 + grid1 - 1) / grid1;
        per = (per + g.threads - 1) / g.threads * g.threads;
        const uint64_t windows_per_cta = per * kUnitBases;
        // 3 x the mean region plus slack, in groups
        const uint64_t groups = (3 * windows_per_cta / nb + 4 * kGroup + kGroup - 1) / kGroup;
        void *staging = nullptr;
        uint32_t *fill = nullptr;
        KPAL_CHECK(radix_workspace(size_t(grid1) * nb * groups * kGroup * 2, size_t(grid1) * nb * 4, &staging, &fill));
        RadixParams p;
        p.codes = reinterpret_cast<const uint2 *>(d_codes);
        p.valid = d_valid;
        p.unit_begin = s0; p.unit_end = s1; p.n_units = n_units;
        p.k = k; p.P = P; p.nb = nb; p.cap = g.cap;
        p.region_groups = uint32_t(groups);
        p.staging = static_cast<uint16_t *>(staging);
        p.region_fill = fill;
        p.debug = g_radix_debug.load();
        const PeerOut *seg_peer = (s1 == n_units) ? peer : nullptr;      // last segment only
        if (counter_bits == 32)
            KPAL_CHECK(launch_radix_passes<uint32_t>(g, p, grid1, smem2, threads2,
                                                     static_cast<uint32_t *>(d_table), seg_peer, stream));
        else
            KPAL_CHECK(launch_radix_passes<unsigned long long>(g, p, grid1, smem2, threads2,
                                                               static_cast<unsigned long long *>(d_table), seg_peer, stream));
    }
    return KPAL_OK;
}

}  // namespace kpal
