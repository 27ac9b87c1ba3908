// Radix-partitioned k-mer counting with TWO windows per payload (9 <= k <= 12).
//
// Same contract as launch_count_radix (count_radix.cu): accumulate the windows of a
// packed stream (kpal/klib.py:154-168) into a table of 4^k counters.  The one-window
// path spends one returning shared-memory atomic and one 16-bit shared store per
// window in pass 1 and runs at the rate of that scatter (profiles/r01_ncu_count_radix_raw.csv:
// 9.2 shared-memory wavefronts per 32 windows).  Two neighbouring windows share k-1 of
// their k bases, so the (k+1)-mer M that holds both is staged instead:
//
//   M = b0 b1 .. bk         window 0 = b0 .. b(k-1) = M >> 2,   window 1 = b1 .. bk = M & (4^k - 1)
//   bucket  C = b1 .. b5    (5 bases common to both windows -> 1024 buckets)
//   payload   = b0 : b6 .. bk    2 + 2 (k - 5) <= 16 bits
//
// Pairs start at the even bases of a 32-base unit.  A valid window whose partner is not
// valid (at most two per run of valid windows: ~1 % of the windows of 150-bp reads) is
// counted with a plain RED on the table.
//
//   pass 1  pair_partition_kernel    as radix_partition_kernel, per PAIR: one shared atomic
//           (rank in the bucket's slot) + one 16-bit store; full 32-byte groups are flushed
//           to the CTA's region of the bucket.  Half the shared-memory operations and half
//           the staging bytes (1 B / window) of the one-window path.
//   pass 2  pair_histogram_kernel    one CTA per (bucket, role), the role = which window of the pair:
//           role 1 (window 1 = C : R) histograms R into the bucket's contiguous table slice;
//           role 0 (window 0 = b0 : C : R >> 2) into four runs of 4^(k-6) bins.  A histogram is
//           4^(k-5) 32-bit counters in shared memory (64 KB at k = 12): three CTAs per SM, whose
//           zero / histogram / flush phases overlap.  The histogram is ADDED to the table by the
//           TMA unit (cp.reduce.async.bulk ... add.u32, SASS UBLKRED): atomic at the L2, so the
//           two roles (and the single-window REDs of pass 1) may touch the same bins from
//           different CTAs in ONE launch.  64-bit tables (>= 2^32 bases) take the older form:
//           one launch per role with plain 16-byte read-modify-writes.
//
// Slot or region overflow (skewed / repetitive sequence) falls back to REDs on the table for
// both windows of the pair: exact for every input.
#include "radix_common.cuh"

#include <type_traits>

namespace kpal {

constexpr int kPairBucketBases = 5;                 // c: bucket = bases 1 .. c of the (k+1)-mer
constexpr int kPairBuckets = 1 << (2 * kPairBucketBases);
constexpr int kPairThreads = 1024;
constexpr int kPairHistThreads = 512;

constexpr int kSlotTrash = 8;           // halfwords behind a slot's `cap` places that absorb stores of unbinned pairs
constexpr int kPairCap = 96;            // places per slot; cap + trash = 8 x 13 halfwords = 208 bytes keeps the
constexpr int kSlotHalfwords = kPairCap + kSlotTrash;       // 16-byte slot reads of neighbouring buckets on distinct bank groups
constexpr size_t kPairSmem1 = size_t(kPairBuckets + 32) * 4 + size_t(kPairBuckets) * 4 +
                              size_t(kPairBuckets) * kSlotHalfwords * 2 + 64;      // counters, fill, slots, pad

struct PairParams {
    const uint2 *codes;         // 32 bases per uint2
    const uint32_t *valid;      // 32 bases per word
    uint64_t unit_begin, unit_end;   // units [begin, end) of this launch
    uint64_t n_units;           // units readable in the stream (loads are clamped to this)
    int k;
    uint32_t region_groups;     // capacity of one (CTA, bucket) region in groups of 16 payloads
    uint16_t *staging;          // [grid][1024][region_groups * 16]
    uint32_t *region_fill;      // [grid][1024] payloads stored per region
    int flush_every;            // tiles binned between two flushes of the slots
};

// k is a template parameter of pass 1: every shift, mask and slot offset of the inner loop is an
// immediate, which frees the registers the sixteen ranks in flight need (below).
template <int K>
struct PairGeom {
    static constexpr int shift = 32 - 2 * (K + 1);                  // funnel-shifted word -> M
    static constexpr int rshift = 2 * (K - kPairBucketBases);       // M -> b0 : C
    static constexpr uint32_t rmask = (1u << rshift) - 1u;          // 4^(k-c) - 1
    static constexpr uint32_t kmask = (1u << (2 * K)) - 1u;         // 4^k - 1
};

// The pairs of a unit in two batches (12 + 4), branch-free as bin_eight in count_radix.cu.
// A pair takes a returning shared atomic (its rank in the bucket's slot) and a 16-bit store at
// slot[min(rank, cap)] -- the places from `cap` on are trash, so a full slot needs no branch;
// the unit checks once whether any rank reached `cap` and then counts those pairs with REDs.
// All atomics of a batch are issued before the first rank is used: the kernel runs at the latency
// of the shared atomics times the number a warp keeps in flight (ncu: 1.5 atomics per cycle
// and SM with eight in flight per warp, issue slots 45 % busy), so the (k+1)-mers are not kept
// in registers but extracted a second time for the stores -- registers for the ranks instead.
template <int J, int JE, typename F>
__device__ __forceinline__ void static_for(F &&f)
{
    if constexpr (J < JE) {
        f(std::integral_constant<int, J>{});
        static_for<J + 1, JE>(f);
    }
}

#ifndef KPAL_PAIR_BATCH
#define KPAL_PAIR_BATCH 8           // pairs whose atomics are in flight together (of the 16 of a unit); B200: 8 -> 0.2456, 12 -> 0.2471, 14 -> 0.2503 ms/step
#endif

// pairs [J0, J1) of the unit
template <typename CounterT, int K, int J0, int J1>
__device__ __forceinline__ void bin_pairs(const Unit &u, uint32_t both, uint32_t cnt_s, uint32_t slots_s,
                                          uint32_t dummy_cnt_s, CounterT *table)
{
    using G = PairGeom<K>;
    uint32_t rank[J1 - J0];
    static_for<J0, J1>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        const uint32_t m = unit_window<2 * j>(u, G::shift);
        const uint32_t okm = uint32_t(int32_t(both << (2 * j)) >> 31);
        const uint32_t bl4 = (m >> (G::rshift - 2)) & uint32_t(4 * (kPairBuckets - 1));
        const uint32_t real = cnt_s + bl4;
        rank[j - J0] = atoms_add(dummy_cnt_s ^ ((dummy_cnt_s ^ real) & okm), 1u);
    });
    uint32_t top = 0;
    static_for<J0, J1>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        const uint32_t m = unit_window<2 * j>(u, G::shift);
        const uint32_t okm = uint32_t(int32_t(both << (2 * j)) >> 31);
        const uint32_t bl4 = (m >> (G::rshift - 2)) & uint32_t(4 * (kPairBuckets - 1));
        const uint32_t pos = min(rank[j - J0] | ~okm, uint32_t(kPairCap));    // unbinned pair or full slot: the trash place
        top = max(top, rank[j - J0] & okm);
        // payload = b0 : R  (b0 moved down over the bucket bits; the store keeps 16 bits)
        const uint32_t pay = (m & G::rmask) | ((m >> (2 * kPairBucketBases)) & ~G::rmask);
        sts_u16(slots_s + bl4 * uint32_t(kSlotHalfwords / 2) + 2u * pos, pay);
    });
    if (top >= uint32_t(kPairCap)) {    // slot full (skewed / repetitive sequence): count both windows directly
        // Low-complexity sequence sends thousands of windows to ONE bin (poly-A reads: every
        // lane, every pair), and same-address REDs serialise in L2: the lanes that are here
        // with the same index add their number with one RED.
        const unsigned here = __activemask();
        static_for<J0, J1>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            const uint32_t m = unit_window<2 * j>(u, G::shift);
            const bool spill = ((both << (2 * j)) & 0x80000000u) && rank[j - J0] >= uint32_t(kPairCap);
            const unsigned voters = __ballot_sync(here, spill);
            if (spill) {
                const uint32_t w0 = m >> 2, w1 = m & G::kmask;
                const unsigned same0 = __match_any_sync(voters, w0);
                if ((threadIdx.x & 31u) == unsigned(__ffs(same0) - 1)) atomicAdd(table + w0, CounterT(__popc(same0)));
                const unsigned same1 = __match_any_sync(voters, w1);
                if ((threadIdx.x & 31u) == unsigned(__ffs(same1) - 1)) atomicAdd(table + w1, CounterT(__popc(same1)));
            }
        });
    }
}

template <typename CounterT, int K>
__device__ __forceinline__ void bin_unit_pairs(const Unit &u, uint32_t both, uint32_t cnt_s, uint32_t slots_s,
                                               uint32_t dummy_cnt_s, CounterT *table)
{
    constexpr int NB = KPAL_PAIR_BATCH;
    bin_pairs<CounterT, K, 0, NB>(u, both, cnt_s, slots_s, dummy_cnt_s, table);
    if constexpr (NB < 16) {
        if (both & (0xFFFFFFFFu >> (2 * NB))) bin_pairs<CounterT, K, NB, 16>(u, both, cnt_s, slots_s, dummy_cnt_s, table);
    }
}

// rare path: both windows of the (first n of the 8) payloads of one 16-byte piece of bucket C
template <typename CounterT>
__device__ __noinline__ void red_pairs(uint4 a, uint32_t C, int k, int n, CounterT *__restrict__ table)
{
    const uint32_t v[4] = {a.x, a.y, a.z, a.w};
    const int rbits = 2 * (k - kPairBucketBases);
    const uint32_t rmask = (1u << rbits) - 1u;
    for (int j = 0; j < n; ++j) {
        const uint32_t e = (v[j >> 1] >> (16 * (j & 1))) & ((4u << rbits) - 1u);
        const uint32_t b0 = e >> rbits, R = e & rmask;
        atomicAdd(table + ((b0 << (2 * (k - 1))) | (C << (rbits - 2)) | (R >> 2)), CounterT(1));
        atomicAdd(table + ((C << rbits) | R), CounterT(1));
    }
}

// One persistent 1024-thread CTA per SM; a tile is one unit (32 bases) per thread.
template <typename CounterT, int K>
__global__ void __launch_bounds__(kPairThreads, 1)
pair_partition_kernel(const PairParams p, CounterT *__restrict__ table)
{
    extern __shared__ __align__(16) unsigned char pair_smem[];
    constexpr int na = kPairBuckets;
    uint32_t *cnt = reinterpret_cast<uint32_t *>(pair_smem);            // [na] payloads in the slot (+32 dummies)
    uint32_t *fillg = cnt + na + 32;                                     // [na] groups already stored
    uint16_t *slots = reinterpret_cast<uint16_t *>(fillg + na);          // [na][cap + kSlotTrash] (+ 64 B pad)

    const int tid = threadIdx.x;
    const unsigned lane = tid & 31u;
    for (int b = tid; b < na + 32; b += kPairThreads) cnt[b] = 0;
    for (int b = tid; b < na; b += kPairThreads) fillg[b] = 0;
    __syncthreads();

    constexpr uint64_t kTile = kPairThreads;
    const uint64_t total = p.unit_end - p.unit_begin;
    uint64_t per = (total + gridDim.x - 1) / gridDim.x;
    per = (per + kTile - 1) / kTile * kTile;
    const uint64_t u0 = p.unit_begin + uint64_t(blockIdx.x) * per;
    const uint64_t u1 = (u0 + per < p.unit_end) ? u0 + per : p.unit_end;

    const uint32_t cnt_s = smem_u32(cnt), slots_s = smem_u32(slots), fill_s = smem_u32(fillg);
    const uint32_t dummy_cnt_s = cnt_s + 4u * (uint32_t(na) + lane);
    constexpr int kshift = 32 - 2 * K;
    constexpr uint32_t slot_bytes = uint32_t(kSlotHalfwords) * 2u;

    uint4 *const my_regions4 = reinterpret_cast<uint4 *>(p.staging + uint64_t(blockIdx.x) * na * p.region_groups * kGroup);
    const uint32_t region_v4 = p.region_groups * 2u;            // 16-byte pieces per region

    // software pipeline: the words of the next tile (and lane 31's halo words) are in flight
    uint2 cw_next = make_uint2(0, 0);
    uint32_t vw_next = 0, hc_next = 0, hv_next = 0;
    auto prefetch = [&](uint64_t t0) {
        const uint64_t unit = t0 + tid;
        cw_next = make_uint2(0, 0); vw_next = 0; hc_next = 0; hv_next = 0;
        if (t0 < u1 && unit < p.n_units) {
            cw_next = ldg_keep_v2(p.codes + unit);
            vw_next = ldg_keep_u32(p.valid + unit);
            if (lane == 31u) {                  // the stream is padded by one 64-base chunk
                hc_next = ldg_keep_u32(reinterpret_cast<const uint32_t *>(p.codes + unit + 1));
                hv_next = ldg_keep_u32(p.valid + unit + 1);
            }
        }
    };
    prefetch(u0);

    int since_flush = 0;
    for (uint64_t t0 = u0; t0 < u1; t0 += kTile) {
        // ---- A: bin this tile's pairs into the bucket slots
        {
            const uint64_t unit = t0 + tid;
            const uint2 cw = cw_next;
            const uint32_t vw = vw_next;
            uint32_t next_c = __shfl_down_sync(0xffffffffu, cw.x, 1);
            uint32_t next_v = __shfl_down_sync(0xffffffffu, vw, 1);
            if (lane == 31u) { next_c = hc_next; next_v = hv_next; }
            prefetch(t0 + kTile);
            Unit u;
            u.w[0] = cw.x; u.w[1] = cw.y; u.w[2] = next_c;
            {   // run mask by the binary method on k (as load_chunk in count.cu)
                const uint64_t v = (uint64_t(vw) << 32) | next_v;
                uint64_t a = v;
                int len = 1;
#pragma unroll
                for (int bit = 2; bit >= 0; --bit) {        // 9 <= K <= 12: top bit 3
                    a &= a << len; len <<= 1;
                    if ((K >> bit) & 1) { a &= v << len; len += 1; }
                }
                u.starts = (unit < u1) ? uint32_t(a >> 32) : 0u;
            }
            if (u.starts) {
                // pairs at the even bases whose two windows are both valid
                const uint32_t both = u.starts & (u.starts << 1) & 0xAAAAAAAAu;
                uint32_t singles = u.starts & ~(both | (both >> 1));
                // (a mask-free variant for units of sixteen valid pairs only pays when it is taken by
                // whole warps: with reads every warp holds record ends, and executing both
                // variants cost 45 us -- measured, profiles/README.md)
                if (both) bin_unit_pairs<CounterT, K>(u, both, cnt_s, slots_s, dummy_cnt_s, table);
                while (singles) {                   // run ends: ~1 window per run of valid windows
                    const int o = __clz(singles);
                    singles &= ~(0x80000000u >> o);
                    const uint32_t lo = (o & 16) ? u.w[1] : u.w[0], hi = (o & 16) ? u.w[2] : u.w[1];
                    atomicAdd(table + (__funnelshift_l(hi, lo, 2 * (o & 15)) >> kshift), CounterT(1));
                }
            }
        }
        // ---- B (every `flush_every` tiles and after the last one): thread b stores the complete
        // 32-byte groups of bucket b's slot to the CTA's region of the bucket and moves the
        // remainder (< 16 payloads) to the front.  One bucket per thread: no team logic, all
        // lanes busy, and the two shared-memory round trips of 32 buckets overlap per warp.
        ++since_flush;
        if (since_flush < p.flush_every && t0 + kTile < u1) continue;
        since_flush = 0;
        __syncthreads();
        {
            static_assert(kPairBuckets == kPairThreads, "one bucket per thread");
            const uint32_t b = uint32_t(tid);
            const uint32_t cnt_a = cnt_s + 4u * b, fill_a = fill_s + 4u * b;
            const uint32_t n = min(lds_u32(cnt_a), uint32_t(kPairCap)), f = lds_u32(fill_a);
            const uint32_t g = n / kGroup;
            if (g) {
                const uint32_t slot_a = slots_s + b * slot_bytes;
                uint4 *dst = my_regions4 + b * region_v4 + 2u * f;
                for (uint32_t q = 0; q < g; ++q) {
                    const uint4 x0 = lds_v4(slot_a + 32u * q), x1 = lds_v4(slot_a + 32u * q + 16u);
                    if (f + q < p.region_groups) { __stcs(dst + 2u * q, x0); __stcs(dst + 2u * q + 1u, x1); }
                    else { red_pairs<CounterT>(x0, b, K, 8, table); red_pairs<CounterT>(x1, b, K, 8, table); }
                }
                const uint4 r0 = lds_v4(slot_a + 32u * g), r1 = lds_v4(slot_a + 32u * g + 16u);   // the remainder
                sts_v4(slot_a, r0); sts_v4(slot_a + 16u, r1);
                sts_u32(cnt_a, n - g * kGroup);
                sts_u32(fill_a, min(f + g, p.region_groups));
            }
        }
        __syncthreads();
    }

    // ---- remainders (< 16 per bucket after the last flush) and the per-region totals
    {
        const uint32_t b = uint32_t(tid);
        const uint32_t n = min(cnt[b], uint32_t(kPairCap)), f = fillg[b];
        uint32_t stored = f * kGroup;
        if (n) {
            const uint4 *slot = reinterpret_cast<const uint4 *>(slots + b * kSlotHalfwords);
            const uint4 x0 = slot[0], x1 = slot[1];
            if (f < p.region_groups) {
                uint4 *dst = my_regions4 + b * region_v4 + 2u * f;
                dst[0] = x0;                             // payloads beyond n are never read
                if (n > 8) dst[1] = x1;
                stored += n;
            } else {
                red_pairs<CounterT>(x0, b, K, n < 8 ? int(n) : 8, table);
                if (n > 8) red_pairs<CounterT>(x1, b, K, int(n) - 8, table);
            }
        }
        p.region_fill[uint64_t(blockIdx.x) * na + b] = stored;
    }
}

// ---------------------------------------------------------------------------
// pass 2: one CTA per (bucket, role)
// ---------------------------------------------------------------------------
// The CTA's histogram is added to the table by the TMA unit: cp.reduce.async.bulk (SASS
// UBLKRED.G.S.ADD) reads the shared-memory histogram and performs the element-wise add
// in L2 -- one instruction per contiguous run instead of a read-modify-write loop, and
// atomic, so both roles of all buckets run in ONE launch (4.6 waves of three CTAs per SM
// whose zero / histogram / flush phases interleave) although they touch the same bins.
// FUSED = false (64-bit counters: the histogram is 32-bit, the reduce needs equal types):
// the role is the kernel argument, the flush a 16-byte read-modify-write loop, and the
// two roles are launched one after the other.
template <typename CounterT, bool FUSED>
__global__ void __launch_bounds__(kPairHistThreads, 3)
pair_histogram_kernel(const uint16_t *__restrict__ staging, const uint32_t *__restrict__ region_fill,
                      int n_part_ctas, int k, uint32_t region_groups, int role_arg,
                      CounterT *__restrict__ table)
{
    extern __shared__ __align__(128) uint32_t pair_hist[];
    __shared__ unsigned int next_group;
    if (threadIdx.x == 0) next_group = 0;
    const int rbits = 2 * (k - kPairBucketBases);
    const uint32_t bins = 1u << rbits;
    // fused: neighbouring CTAs are the two roles of one bucket, so the second read of the
    // bucket's regions finds them in L2
    const uint32_t C = FUSED ? blockIdx.x >> 1 : blockIdx.x;
    const int role = FUSED ? int(blockIdx.x & 1u) : role_arg;
    for (uint32_t i = threadIdx.x * 4; i < bins; i += blockDim.x * 4)
        *reinterpret_cast<uint4 *>(pair_hist + i) = make_uint4(0, 0, 0, 0);
    __syncthreads();

    // Eight lanes walk one region (16-byte vectors lane % 8, + 8, ...), four regions per
    // warp at a time; the next vector and the next group's fill counts are loaded ahead.
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    const uint32_t bmask = bins - 1u;
    const int eshift = role == 1 ? 0 : 2;           // window 1: R;  window 0: b0 : R >> 2
    const uint32_t hist_s = smem_u32(pair_hist);
    auto fill_of = [&](int g) -> uint32_t {
        const int c = g * 4 + (lane >> 3);
        return c < n_part_ctas ? ldg_keep_u32(region_fill + uint64_t(c) * kPairBuckets + C) : 0u;
    };
    // groups of four regions are handed out through a shared counter: with a fixed split
    // (37 groups over 16 warps) a third of the warps walked three groups, the others two, and
    // the CTA waited for them at the barrier (ncu: barrier + long scoreboard were the top stalls)
    auto grab = [&]() -> int {
        int v = 0;
        if (lane == 0) v = int(atomicAdd(&next_group, 1u));
        return __shfl_sync(0xffffffffu, v, 0);
    };
    int g = grab();
    uint32_t n_ahead = g * 4 < n_part_ctas ? fill_of(g) : 0u;
    while (g * 4 < n_part_ctas) {
        const int g_next = grab();
        const int c = g * 4 + (lane >> 3);
        const uint32_t n = n_ahead;
        n_ahead = g_next * 4 < n_part_ctas ? fill_of(g_next) : 0u;
        const uint32_t nv = (n + 7u) / 8u;
        const uint4 *src = reinterpret_cast<const uint4 *>(
            staging + (uint64_t(c < n_part_ctas ? c : 0) * kPairBuckets + C) * region_groups * kGroup);
        uint32_t i = lane & 7u;
        uint4 cur = make_uint4(0, 0, 0, 0), nxt = make_uint4(0, 0, 0, 0);
        if (i < nv) cur = ldg_keep_v4(src + i);
        if (i + 8u < nv) nxt = ldg_keep_v4(src + i + 8u);
        while (__any_sync(0xffffffffu, i < nv)) {
            uint4 far = make_uint4(0, 0, 0, 0);
            if (i + 16u < nv) far = ldg_keep_v4(src + i + 16u);
            const uint32_t at = i * 8u;
            const uint32_t rem = at < n ? n - at : 0u;
            const uint32_t w[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t e = w[j >> 1] >> (16 * (j & 1));
                const uint32_t bin = (e >> eshift) & bmask;
                if (uint32_t(j) < rem)
                    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(hist_s + 4u * bin) : "memory");
            }
            cur = nxt; nxt = far;
            i += 8u;
        }
        g = g_next;
    }
    __syncthreads();

    // table += histogram.  Role 1: one contiguous slice at C << rbits.  Role 0: histogram
    // index = b0 : (R >> 2) -> table index b0 : C : (R >> 2), four runs of bins / 4 counters.
    const uint32_t run = bins >> 2;
    if constexpr (FUSED) {
        static_assert(sizeof(CounterT) == 4, "the bulk reduce adds equal types");
        if (threadIdx.x == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // the atomics above -> the TMA unit's read
            if (role == 1) {
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.u32 [%0], [%1], %2;"
                             ::"l"(table + (uint64_t(C) << rbits)), "r"(hist_s), "r"(bins * 4u) : "memory");
            } else {
#pragma unroll
                for (uint32_t b0 = 0; b0 < 4; ++b0)
                    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.u32 [%0], [%1], %2;"
                                 ::"l"(table + ((uint64_t(b0) << (2 * (k - 1))) | (uint64_t(C) << (rbits - 2)))),
                                   "r"(hist_s + b0 * run * 4u), "r"(run * 4u) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory is released at exit
        }
    } else {
        constexpr int NQ = 2;           // read-modify-writes in flight per thread
        for (uint32_t i0 = threadIdx.x * 4; i0 < bins; i0 += blockDim.x * 4 * NQ) {
            uint4 h[NQ];
            uint32_t at[NQ];                // table index (< 4^12)
            bool any[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const uint32_t i = i0 + q * blockDim.x * 4;
                any[q] = false;
                if (i < bins) {
                    h[q] = *reinterpret_cast<const uint4 *>(pair_hist + i);
                    any[q] = (h[q].x | h[q].y | h[q].z | h[q].w) != 0;
                    if (role == 1) at[q] = (C << rbits) + i;
                    else at[q] = ((i / run) << (2 * (k - 1))) | (C << (rbits - 2)) | (i % run);
                }
            }
#pragma unroll
            for (int q = 0; q < NQ; ++q)
                if (any[q]) {
                    if constexpr (sizeof(CounterT) == 4) {
                        uint4 t = *reinterpret_cast<const uint4 *>(table + at[q]);
                        t.x += h[q].x; t.y += h[q].y; t.z += h[q].z; t.w += h[q].w;
                        *reinterpret_cast<uint4 *>(table + at[q]) = t;
                    } else {
                        ulonglong2 t0 = *reinterpret_cast<const ulonglong2 *>(table + at[q]);
                        ulonglong2 t1 = *reinterpret_cast<const ulonglong2 *>(table + at[q] + 2);
                        t0.x += h[q].x; t0.y += h[q].y; t1.x += h[q].z; t1.y += h[q].w;
                        *reinterpret_cast<ulonglong2 *>(table + at[q]) = t0;
                        *reinterpret_cast<ulonglong2 *>(table + at[q] + 2) = t1;
                    }
                }
        }
    }
}

// ---------------------------------------------------------------------------
// launcher
// ---------------------------------------------------------------------------
static std::atomic<int> g_pair_flush_every{0};  // tiles between flushes of the pass-1 slots (0 = automatic)
void set_pair_flush_every(int v) { g_pair_flush_every.store(v < 0 ? 0 : v); }
static std::atomic<int> g_pair_fused{1};        // pass 2: 1 = both roles in one launch, flushed by the TMA unit
void set_pair_fused(int v) { g_pair_fused.store(v ? 1 : 0); }

bool pairs_supported(int k) { return k >= 9 && k <= 12; }

template <typename CounterT, int K>
static int launch_partition(const PairParams &p, int grid1, CounterT *table, cudaStream_t stream)
{
    KPAL_CUDA(cudaFuncSetAttribute(pair_partition_kernel<CounterT, K>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   int(kPairSmem1)));
    pair_partition_kernel<CounterT, K><<<grid1, kPairThreads, kPairSmem1, stream>>>(p, table);
    KPAL_LAUNCH_CHECK("pair_partition_kernel");
    return KPAL_OK;
}

template <typename CounterT>
static int launch_pair_passes(const PairParams &p, int grid1, CounterT *table, cudaStream_t stream)
{
    switch (p.k) {
    case 9: KPAL_CHECK((launch_partition<CounterT, 9>(p, grid1, table, stream))); break;
    case 10: KPAL_CHECK((launch_partition<CounterT, 10>(p, grid1, table, stream))); break;
    case 11: KPAL_CHECK((launch_partition<CounterT, 11>(p, grid1, table, stream))); break;
    default: KPAL_CHECK((launch_partition<CounterT, 12>(p, grid1, table, stream))); break;
    }
    const size_t smem2 = size_t(4) << (2 * (p.k - kPairBucketBases));
    if constexpr (sizeof(CounterT) == 4) {
        if (g_pair_fused.load()) {
            KPAL_CUDA(cudaFuncSetAttribute(pair_histogram_kernel<CounterT, true>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem2)));
            pair_histogram_kernel<CounterT, true><<<2 * kPairBuckets, kPairHistThreads, smem2, stream>>>(
                p.staging, p.region_fill, grid1, p.k, p.region_groups, 0, table);
            KPAL_LAUNCH_CHECK("pair_histogram_kernel");
            return KPAL_OK;
        }
    }
    KPAL_CUDA(cudaFuncSetAttribute(pair_histogram_kernel<CounterT, false>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem2)));
    for (int role = 1; role >= 0; --role) {
        pair_histogram_kernel<CounterT, false><<<kPairBuckets, kPairHistThreads, smem2, stream>>>(
            p.staging, p.region_fill, grid1, p.k, p.region_groups, role, table);
        KPAL_LAUNCH_CHECK("pair_histogram_kernel");
    }
    return KPAL_OK;
}

int launch_count_pairs(const uint32_t *d_codes, const uint32_t *d_valid, uint64_t n_bases, int k,
                       void *d_table, int counter_bits, cudaStream_t stream)
{
    if (!pairs_supported(k)) return bad_arg("the pair path covers 9 <= k <= 12");
    const int grid1 = sm_count();
    // A slot keeps < 16 payloads over a flush and gains 16 per tile on average (uniform
    // sequence): with 96 places, 3 tiles between flushes leave 5 sigma of headroom (the surplus
    // of a fuller slot takes the RED path: exact, only slower).
    int flush_every = g_pair_flush_every.load();
    if (flush_every <= 0) flush_every = 3;
    const uint64_t tile = kPairThreads;
    const uint64_t n_units = 2 * n_chunks_of(n_bases);
    const uint64_t seg_units = (512ull << 20) / kUnitBases;
    for (uint64_t s0 = 0; s0 < n_units; s0 += seg_units) {
        const uint64_t s1 = (s0 + seg_units < n_units) ? s0 + seg_units : n_units;
        uint64_t per = (s1 - s0 + grid1 - 1) / grid1;
        per = (per + tile - 1) / tile * tile;
        const uint64_t pairs_per_cta = per * kUnitBases / 2;
        // 3 x the mean region plus slack, in groups
        const uint64_t groups = (3 * pairs_per_cta / kPairBuckets + 4 * kGroup + kGroup - 1) / kGroup;
        void *staging = nullptr;
        uint32_t *fill = nullptr;
        KPAL_CHECK(radix_workspace(size_t(grid1) * kPairBuckets * groups * kGroup * 2,
                                   size_t(grid1) * kPairBuckets * 4, &staging, &fill));
        PairParams p;
        p.codes = reinterpret_cast<const uint2 *>(d_codes);
        p.valid = d_valid;
        p.unit_begin = s0; p.unit_end = s1; p.n_units = n_units;
        p.k = k;
        p.region_groups = uint32_t(groups);
        p.staging = static_cast<uint16_t *>(staging);
        p.region_fill = fill;
        p.flush_every = flush_every;
        if (counter_bits == 32)
            KPAL_CHECK(launch_pair_passes<uint32_t>(p, grid1, static_cast<uint32_t *>(d_table), stream));
        else
            KPAL_CHECK(launch_pair_passes<unsigned long long>(p, grid1, static_cast<unsigned long long *>(d_table), stream));
    }
    return KPAL_OK;
}

}  // namespace kpal
