// Multi-GPU sum of the per-rank 4^k counter tables over NVLink peer memory
// ("counting shards sequence records across GPUs and sums
// the per-GPU count vectors"), as two kernels around one cross-GPU barrier
// instead of a library reduce:
//
//   push     every rank cuts its table into `world` contiguous slices and stores
//            slice o straight into rank o's inbox (slot = sender's rank) with
//            16-byte peer stores over NVLink -- an all-to-all in which every GPU
//            sends and receives (world-1)/world of a table at the same time, so
//            all NVSwitch ports are busy in both directions;
//   collect  (after a barrier) every rank sums the `world` slots of its inbox --
//            local HBM reads -- and stores the summed slice into the root's table
//            (peer store), where one finalize (widen + balance) follows a second
//            barrier.
//
// Per GPU the wire carries 2 x (world-1)/world table copies in total, spread
// over all links, versus a chain/tree in which the root's single link is the
// bottleneck.  The radix count path can also write its pass-2 histograms into
// the inboxes directly (count_radix.cu), which removes the push kernel.
//
// Peer pointers come from cudaIpcOpenMemHandle (one process per GPU); the
// barriers are the caller's (a 1-element NCCL all-reduce on the same stream).
#include "common.cuh"

#include <algorithm>

namespace kpal {

static int peer_check_args_fwd(int k, int counter_bits, int rank, int world);

template <typename T>
__global__ void __launch_bounds__(256)
reduce_push_kernel(const T *__restrict__ table, PeerOut peer, uint64_t bins)
{
    constexpr int V = 16 / sizeof(T);
    const int o = blockIdx.y, world = peer.world;
    const uint64_t lo = slice_begin(bins, o, world), hi = slice_begin(bins, o + 1, world);
    const uint4 *src = reinterpret_cast<const uint4 *>(table + lo);
    uint4 *dst = reinterpret_cast<uint4 *>(static_cast<T *>(peer.inbox[o]) + uint64_t(peer.rank) * slot_elems(bins, world));
    const uint64_t nv = (hi - lo) / V;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < nv;
         i += uint64_t(gridDim.x) * blockDim.x)
        dst[i] = __ldcs(src + i);
}

template <typename T>
__global__ void __launch_bounds__(256)
reduce_collect_kernel(const T *__restrict__ inbox, int rank, int world, uint64_t bins,
                      T *__restrict__ root_table)
{
    constexpr int V = 16 / sizeof(T);
    const uint64_t lo = slice_begin(bins, rank, world), hi = slice_begin(bins, rank + 1, world);
    const uint64_t slot = slot_elems(bins, world);
    const uint64_t nv = (hi - lo) / V;
    uint4 *dst = reinterpret_cast<uint4 *>(root_table + lo);
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < nv;
         i += uint64_t(gridDim.x) * blockDim.x) {
        uint4 acc = __ldcs(reinterpret_cast<const uint4 *>(inbox) + i);
        for (int s = 1; s < world; ++s) {
            const uint4 x = __ldcs(reinterpret_cast<const uint4 *>(inbox + uint64_t(s) * slot) + i);
            if constexpr (sizeof(T) == 4) {
                acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
            } else {
                unsigned long long a0 = (uint64_t(acc.y) << 32 | acc.x) + (uint64_t(x.y) << 32 | x.x);
                unsigned long long a1 = (uint64_t(acc.w) << 32 | acc.z) + (uint64_t(x.w) << 32 | x.z);
                acc = make_uint4(uint32_t(a0), uint32_t(a0 >> 32), uint32_t(a1), uint32_t(a1 >> 32));
            }
        }
        dst[i] = acc;
    }
}

// ---------------------------------------------------------------------------
// Fused form: balance + reduce-scatter + distributed finalize (profiles/README.md)
// ---------------------------------------------------------------------------
// The sum over the ranks commutes with the balance (kpal/klib.py:285-298 is linear), and the
// balanced counts of ONE rank's shard are small: so every rank balances its own table and
// sends the result narrow.
//
//   slice_push   = the tiled balanced finalize of count.cu, but each 64-count row goes, as 64
//                  BYTES, straight into the inbox of the rank that owns its table slice (peer
//                  stores over NVLink): the wire carries 1 byte per bin instead of the 4 of a
//                  u32 table.  A count of 255 or more travels as the byte 255 = "escape": its
//                  value goes, as u32, to the same position of the sender's WIDE row set in
//                  the owner's inbox (a scattered 4-byte store; rare when the shard's mean
//                  count is small).  A shard whose mean count is large sends wide rows only
//                  (`wide_rows`, chosen by the caller from bases / 4^k).  One launch either way.
//   signal       one 8-byte release store per peer: "my rows of this epoch have landed", sent by
//                  the first CTA of the NEXT kernel on the stream (the collect, or the
//                  stand-alone signal kernel): the push kernel's stores are complete, system-wide,
//                  when that kernel starts.  (A signal from inside the push kernel needs a
//                  system-scope fence per CTA: 10 us of its 30.)
//   slice_collect  on the owner: waits for the world's signals (acquire loads of its own
//                  inbox), sums the narrow (or wide) rows of all senders and writes the int64
//                  slice of the final balanced profile -- plus, for the host entry points, its
//                  uint8 / uint16 forms and the two overflow flags of finalize_to_host.
//
// Nothing is gathered on one GPU: the profile stays sharded by slice (every owner copies its
// slice to the host over its own PCIe link), so no link carries more than (world-1)/world of
// 4^k bytes.  Inboxes are double-buffered by epoch parity: a rank can be at most one step
// ahead of a peer (its collect of step s+1 needs the peer's push of step s+1, which the peer
// issues after its own collect of step s).
//
// Order of the rows inside a sender's row set.  A row = the 64 counts of one (h, m).  When the
// world is a power of two (<= 64) every slice is a whole range of h, nh = 64 / world values:
// the rows are then kept TILE-MAJOR -- row (h, m) at position m * nh + (h - h_first) -- so that
// the nh rows a tile sends to one owner are contiguous there (512 bytes at 8 ranks, 2 KB at 2)
// and leave the sender as full 16-byte-per-lane peer stores.  Any other world: in index order
// (position = row - first row of the slice), 64-byte pieces.
struct SliceInbox {
    void *base[kMaxPeers];      // inbox of rank o (both parities), mapped on THIS device
    int rank, world;
};
__host__ __device__ inline int slice_log_nh(int world)      // log2(64 / world), or -1: not tile-major
{
    for (int s = 0; s <= 6; ++s) if (world == (1 << s)) return 6 - s;
    return -1;
}

__host__ __device__ inline uint64_t slice64_begin(uint64_t bins, int o, int world)
{
    if (o >= world) return bins;
    return (bins * uint64_t(o) / uint64_t(world)) & ~uint64_t(63);      // whole 64-count rows
}
__host__ __device__ inline uint64_t slice64_cap(uint64_t bins, int world)
{
    return ((bins + world - 1) / world + 127) & ~uint64_t(63);           // >= the longest slice
}
// one parity of an inbox: [flags: 16 x u64 | pad to 256][narrow: world x cap u8][wide: world x cap u32]
__host__ __device__ inline uint64_t inbox_parity_bytes(uint64_t bins, int world)
{
    return 256 + uint64_t(world) * slice64_cap(bins, world) * 5;
}
// owner of the 64-count row that holds `index`: begins[] = the world + 1 slice boundaries (shared
// memory), bins = 4^k a power of two -- no 64-bit division on the store path (the first form of
// the push kernel spent its time there: 109 us for a 64 MB table, 16 % of the issue slots busy).
__device__ __forceinline__ int owner_of(uint64_t index, const uint64_t *begins, int k2, int world)
{
    int o = int((index * uint64_t(world)) >> k2);
    if (o + 1 < world && begins[o + 1] <= index) ++o;       // boundaries are rounded DOWN to whole rows
    return o;
}

// "my rows of epoch e have landed in your inbox": one release store per peer (threads 0 .. world-1)
__device__ __forceinline__ void slice_signal(const SliceInbox &peers, int parity, unsigned long long epoch,
                                             uint64_t bins, bool wide)
{
    const int o = threadIdx.x;
    if (o >= peers.world) return;
    const unsigned long long value = (epoch << 1) | (wide ? 1ull : 0ull);
    unsigned long long *flag = reinterpret_cast<unsigned long long *>(
        static_cast<unsigned char *>(peers.base[o]) + uint64_t(parity) * inbox_parity_bytes(bins, peers.world)) + peers.rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(value) : "memory");
}

// WIDE = false: u8 rows with escapes; WIDE = true: u32 rows.
// state[1]: a count above 32 bits (the caller's error; sticky).
//
// Index = [h : 3 bases][m : k-6 bases][l : 3 bases]; rc(index) = [rc(l)][rc(m)][rc(h)] (count.cu).
// A tile pair = the 64 x 64 tile of one m together with the tile of rc(m); only the canonical m
// (m <= rc(m)) are enumerated (canon_unrank), CTA b takes pairs b, b + grid, ...  The two tiles
// of the NEXT pair are fetched with cp.async (16 bytes per thread and request, straight into
// shared memory) while the current pair is computed: two stages.  A thread owns, in each tile, the
// 4 x 4 block of rows h = hl + 16 t (t = 0..3: the top base of h) and bins l = 4 l4 + j; the
// reverse complements of a block are four 16-byte pieces of the other tile: the partners of bin
// l + j for t = 0..3 are the four neighbouring columns (rc(hl) & ~3) + 3 - t of row rc(l + j).
// The first form of this kernel moved every count through shared memory one word at a time, did
// 64-bit divisions per store and let every thread fence: 109 us for a 64 MB table; then 35 us
// with barrier stalls on un-overlapped loads; HBM time is 13 us.
constexpr int kPushRow = 68;            // counters per shared-memory row: 64 + one 4-counter group of padding

// number of d-base words m with m <= rc(m), and the c-th of them (any fixed order)
__host__ __device__ inline uint32_t canon_count(int d)
{
    uint32_t n = (d & 1) ? 2u : 1u;
    for (int e = (d & 1) ? 3 : 2; e <= d; e += 2) n = (6u << (2 * (e - 2))) + 4u * n;
    return n;
}
__device__ inline uint32_t canon_unrank(int d, uint32_t c)
{
    // outermost bases (x, y): x < 3 - y makes m canonical whatever lies between (6 pairs), x == 3 - y
    // leaves it to the inner d - 2 bases (4 pairs), x > 3 - y never
    uint32_t m = 0;
    int hi = 2 * (d - 1), lo = 0;
    while (d >= 2) {
        const uint32_t inner = 1u << (2 * (d - 2)), free = 6u * inner;
        if (c < free) {
            const uint32_t p = c >> (2 * (d - 2));              // (0,0) (0,1) (0,2) (1,0) (1,1) (2,0)
            const uint32_t x = p < 3 ? 0u : p < 5 ? 1u : 2u, y = p < 3 ? p : p < 5 ? p - 3 : 0u;
            return m | (x << hi) | (y << lo) | ((c & (inner - 1)) << (lo + 2));
        }
        c -= free;
        const uint32_t n2 = canon_count(d - 2), x = c / n2;
        c -= x * n2;
        m |= (x << hi) | ((3u - x) << lo);
        hi -= 2; lo += 2; d -= 2;
    }
    if (d == 1) m |= c << lo;
    return m;
}

template <typename CounterT>
__device__ __forceinline__ void load4(const CounterT *p, CounterT (&out)[4])
{
    if constexpr (sizeof(CounterT) == 4) {
        const uint4 x = *reinterpret_cast<const uint4 *>(p);
        out[0] = x.x; out[1] = x.y; out[2] = x.z; out[3] = x.w;
    } else {
        const ulonglong2 x = *reinterpret_cast<const ulonglong2 *>(p), y = *reinterpret_cast<const ulonglong2 *>(p + 2);
        out[0] = x.x; out[1] = x.y; out[2] = y.x; out[3] = y.y;
    }
}
// four counters, global -> shared, asynchronously (16 bytes per request)
template <typename CounterT>
__device__ __forceinline__ void fetch4(CounterT *smem_dst, const CounterT *src)
{
    const uint32_t d = uint32_t(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
    if constexpr (sizeof(CounterT) == 8)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 16u), "l"(src + 2) : "memory");
}

template <typename CounterT, bool WIDE>
__global__ void __launch_bounds__(256, sizeof(CounterT) == 4 ? 3 : 1)
slice_push_kernel(const CounterT *__restrict__ table, int k, const SliceInbox peers, int parity,
                  unsigned long long epoch, unsigned int *__restrict__ state)
{
    const int log_nh = slice_log_nh(peers.world);               // >= 0: tile-major row sets
    extern __shared__ __align__(16) unsigned char push_smem[];
    __shared__ uint32_t begin_row_s[kMaxPeers + 1];             // slice boundaries in 64-count rows
    __shared__ ulonglong2 dst_s[2][128];                        // per stage and row of the tile pair: where its bytes / its u32 go
    constexpr int kTile = 64 * kPushRow;
    CounterT *tiles_s = reinterpret_cast<CounterT *>(push_smem);    // [stage][A | B][64][kPushRow]
    const int mid = k - 6, mid_bits = 2 * mid;
    const int hshift = 2 * k - 6;
    const uint64_t bins = 1ull << (2 * k);
    const uint32_t pairs = canon_count(mid);
    const uint64_t cap = slice64_cap(bins, peers.world);
    const uint64_t par_off = uint64_t(parity) * inbox_parity_bytes(bins, peers.world);
    if (threadIdx.x <= uint32_t(peers.world))
        begin_row_s[threadIdx.x] = uint32_t(slice64_begin(bins, int(threadIdx.x), peers.world) >> 6);
    // this thread's block and where its partners are; 4-counter groups are swizzled by bit 3 of the
    // row (group ^ 2) so that the eight rows a quarter-warp reads fall into different banks
    const uint32_t l4 = threadIdx.x & 15u, hl = threadIdx.x >> 4;
    const uint32_t own_at = hl * kPushRow + ((l4 ^ (((hl >> 3) & 1u) << 1)) << 2);            // + 16 t rows
    const uint32_t a0 = rc_index(4u * l4, 26);                                                 // row of j = 0; j: - 16 j
    const uint32_t part_at = a0 * kPushRow + (((rc_index(hl, 26) >> 2) ^ (((a0 >> 3) & 1u) << 1)) << 2);
    bool big = false;

    // start the fetch of tile pair c into a stage (and work out where its rows go); returns its m
    auto fetch = [&](uint32_t c, int stage) -> uint32_t {
        uint32_t m = 0;
        if (c < pairs) {
            m = canon_unrank(mid, c);
            const uint32_t mr = mid_bits ? ((~rev2(m)) >> (32 - mid_bits)) : 0u;
            CounterT *A = tiles_s + stage * 2 * kTile, *B = A + kTile;
#pragma unroll
            for (int t = 0; t < 4; ++t)
                fetch4(A + own_at + 16 * t * kPushRow, table + ((uint64_t(hl + 16u * t) << hshift) | (uint64_t(m) << 6) | (4u * l4)));
            if (m != mr) {
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    fetch4(B + own_at + 16 * t * kPushRow, table + ((uint64_t(hl + 16u * t) << hshift) | (uint64_t(mr) << 6) | (4u * l4)));
            }
            if (threadIdx.x < 128) {
                // destination of row (h, tile): the owner's inbox, this sender's row set.  Rows are
                // whole 64-count units and bins = 4^k: 32-bit arithmetic, no division
                const uint32_t h = threadIdx.x & 63u, mm = (threadIdx.x >> 6) ? mr : m, row = (h << mid_bits) | mm;
                int o;
                uint64_t at = uint64_t(peers.rank) * cap;
                if (log_nh >= 0) {
                    o = int(h >> log_nh);
                    at += uint64_t((mm << log_nh) | (h & ((1u << log_nh) - 1u))) << 6;
                } else {
                    o = int((row * uint32_t(peers.world)) >> hshift);
                    if (o + 1 < peers.world && begin_row_s[o + 1] <= row) ++o;     // boundaries are rounded DOWN to whole rows
                    at += uint64_t(row - begin_row_s[o]) << 6;
                }
                unsigned char *inbox = static_cast<unsigned char *>(peers.base[o]) + par_off + 256;
                dst_s[stage][threadIdx.x] = make_ulonglong2((unsigned long long)(inbox + at),
                                                            (unsigned long long)(inbox + uint64_t(peers.world) * cap + 4 * at));
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        return m;
    };

    __syncthreads();                                            // begin_row_s
    uint32_t m_next = fetch(blockIdx.x, 0);
    int stage = 0;
    for (uint32_t c = blockIdx.x; c < pairs; c += gridDim.x, stage ^= 1) {
        __syncthreads();                                        // the other stage has been read (previous trip)
        const uint32_t m = m_next;
        m_next = fetch(c + gridDim.x, stage ^ 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();                                        // this stage has landed, for every thread
        const uint32_t mr = mid_bits ? ((~rev2(m)) >> (32 - mid_bits)) : 0u;
        const bool two = m != mr;
        const CounterT *A = tiles_s + stage * 2 * kTile, *B = A + kTile;
        const bool staged = !WIDE && log_nh >= 0;               // byte rows leave through a staging area
        uint32_t bytes[2][4];                                   // [tile][t]: the four bytes of this thread's piece of row hl + 16 t
#pragma unroll
        for (int tile = 0; tile < 2; ++tile) {
            if (tile && !two) break;
            const CounterT *mine = tile ? B : A;
            const CounterT *other = (two && !tile) ? B : A;     // the tile of rc(this tile's m)
            CounterT part[4][4];                                // [j][c]: partner of (t = 3 - c, j)
#pragma unroll
            for (int j = 0; j < 4; ++j) load4(other + part_at - 16 * j * kPushRow, part[j]);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                CounterT own[4];
                load4(mine + own_at + 16 * t * kPushRow, own);
                const ulonglong2 dst = dst_s[stage][64 * tile + hl + 16 * t];
                uint32_t *wide = reinterpret_cast<uint32_t *>(dst.y) + 4u * l4;
                uint32_t out[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const CounterT sum = own[j] + part[j][3 - t];
                    bool wrapped = false;
                    if constexpr (sizeof(CounterT) == 4) wrapped = sum < own[j]; else wrapped = sum > 0xffffffffull;
                    big |= wrapped;
                    out[j] = uint32_t(sum);
                    if constexpr (!WIDE) {
                        if (sum >= CounterT(0xff) || wrapped) {     // escape: the value goes to the wide row set
                            wide[j] = out[j];
                            out[j] = 0xffu;
                        }
                    }
                }
                if constexpr (!WIDE) {
                    bytes[tile][t] = out[0] | (out[1] << 8) | (out[2] << 16) | (out[3] << 24);
                    if (!staged) *(reinterpret_cast<uint32_t *>(dst.x) + l4) = bytes[tile][t];
                } else {
                    *reinterpret_cast<uint4 *>(wide) = make_uint4(out[0], out[1], out[2], out[3]);
                }
            }
        }
        if constexpr (!WIDE) {
            if (staged) {
                // the byte rows of the pair, [tile][h][64 B], over the tiles just consumed; then every
                // thread sends 16 bytes: four lanes a row, the nh rows of an owner back to back
                uint32_t *stg = reinterpret_cast<uint32_t *>(tiles_s + stage * 2 * kTile);
                __syncthreads();                                // the tiles have been read
#pragma unroll
                for (int tile = 0; tile < 2; ++tile)
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        if (!tile || two) stg[tile * 1024 + (hl + 16 * t) * 16 + l4] = bytes[tile][t];
                __syncthreads();
#pragma unroll
                for (int tile = 0; tile < 2; ++tile) {
                    if (tile && !two) break;
                    const uint32_t h = threadIdx.x >> 2, q = threadIdx.x & 3u;
                    const uint4 v = *reinterpret_cast<const uint4 *>(stg + tile * 1024 + h * 16 + 4 * q);
                    *reinterpret_cast<uint4 *>(dst_s[stage][64 * tile + h].x + 16 * q) = v;
                }
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (big) state[1] = 1u;
}

__global__ void slice_signal_kernel(const SliceInbox peers, int parity, unsigned long long epoch, uint64_t bins, int wide)
{
    slice_signal(peers, parity, epoch, bins, wide != 0);
}

// Owner side.  out64: this rank's slice of the final profile (int64); o16 / o8 / flags (optional):
// its narrow forms for the device->host copy (flags[0]: a count above 65535, flags[1]: above 255).
// signal >= 0: the first CTA first tells the peers that this rank's rows have landed (signal = the
// wide_rows of its push, the kernel before this one on the stream).
template <bool HOST_FORMS>
__global__ void __launch_bounds__(256, 4)
slice_collect_kernel(const SliceInbox peers, int signal, uint64_t bins, int parity,
                     unsigned long long epoch, int64_t *__restrict__ out64, uint16_t *__restrict__ o16,
                     uint8_t *__restrict__ o8, unsigned int *__restrict__ flags)
{
    __shared__ unsigned int wide_s[kMaxPeers], all_narrow_s;
    const int rank = peers.rank, world = peers.world;
    const unsigned char *base = static_cast<const unsigned char *>(peers.base[rank]) + uint64_t(parity) * inbox_parity_bytes(bins, world);
    if (signal >= 0 && blockIdx.x == 0) slice_signal(peers, parity, epoch, bins, signal != 0);
    if (threadIdx.x < world) {
        const unsigned long long *flag = reinterpret_cast<const unsigned long long *>(base) + threadIdx.x;
        unsigned long long v;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
            if ((v >> 1) < epoch) __nanosleep(100);
        } while ((v >> 1) < epoch);
        wide_s[threadIdx.x] = uint32_t(v & 1ull);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t any = 0;
        for (int s = 0; s < world; ++s) any |= wide_s[s];
        all_narrow_s = any == 0;
    }
    __syncthreads();
    const uint64_t cap = slice64_cap(bins, world);
    const uint64_t lo = slice64_begin(bins, rank, world), hi = slice64_begin(bins, rank + 1, world);
    const unsigned char *narrow = base + 256;
    const uint32_t *wide = reinterpret_cast<const uint32_t *>(base + 256 + uint64_t(world) * cap);
    bool over8 = false, over16 = false;
    const uint64_t n = hi - lo;
    // A lane holds four neighbouring bins; the two lanes of a pair swap halves so that every store
    // instruction of the warp writes 512 contiguous bytes (lane 2p: bins 8p, 8p+1 then 8p+4, 8p+5;
    // lane 2p+1: 8p+2, 8p+3 then 8p+6, 8p+7).  Storing its own 32 bytes as two 16-byte pieces left
    // every 32-byte sector half written per instruction: twice the L1 -> L2 write sectors, which
    // bounded the kernel (ncu: l1tex2xbar write 54 % busy at 2.6 TB/s).  Called by whole warps.
    // position in a sender's row set -> offset in this rank's slice (tile-major row sets: SliceInbox)
    const int log_nh = slice_log_nh(world);
    const int hshift = 63 - __clzll((long long)bins) - 6;
    auto slice_offset = [&](uint64_t i) -> uint64_t {
        if (log_nh < 0) return i;
        const uint64_t p = i >> 6;
        return ((p & ((1ull << log_nh) - 1ull)) << hshift) | ((p >> log_nh) << 6) | (i & 63ull);
    };
    auto emit = [&](uint64_t at, const unsigned long long *acc, bool valid) {
        const uint64_t i = slice_offset(at);
        const bool odd = threadIdx.x & 1u;
        const unsigned long long r0 = __shfl_xor_sync(0xffffffffu, odd ? acc[0] : acc[2], 1);
        const unsigned long long r1 = __shfl_xor_sync(0xffffffffu, odd ? acc[1] : acc[3], 1);
        if (!valid) return;
        __stcs(reinterpret_cast<ulonglong2 *>(out64 + (odd ? i - 2 : i)), odd ? make_ulonglong2(r0, r1) : make_ulonglong2(acc[0], acc[1]));
        __stcs(reinterpret_cast<ulonglong2 *>(out64 + (odd ? i + 2 : i + 4)), odd ? make_ulonglong2(acc[2], acc[3]) : make_ulonglong2(r0, r1));
        if constexpr (HOST_FORMS) {
            uint32_t p16[2], p8 = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                over8 |= acc[j] > 0xffull;
                over16 |= acc[j] > 0xffffull;
                if (j & 1) p16[j >> 1] |= uint32_t(acc[j] & 0xffffull) << 16; else p16[j >> 1] = uint32_t(acc[j] & 0xffffull);
                p8 |= uint32_t(acc[j] & 0xffull) << (8 * j);
            }
            *reinterpret_cast<uint2 *>(o16 + i) = make_uint2(p16[0], p16[1]);
            if (o8) *reinterpret_cast<uint32_t *>(o8 + i) = p8;
        }
    };
    // the same for sums still packed as 16-bit halves (lo: bins 0 and 2, hi: bins 1 and 3): one
    // 32-bit exchange per chunk
    auto emit_packed = [&](uint64_t at, uint32_t lo, uint32_t hi, bool valid) {
        const uint64_t i = slice_offset(at);
        const bool odd = threadIdx.x & 1u;
        const uint32_t b01 = __byte_perm(lo, hi, 0x5410), b23 = __byte_perm(lo, hi, 0x7632);     // (bin0 | bin1 << 16), (bin2 | bin3 << 16)
        const uint32_t r = __shfl_xor_sync(0xffffffffu, odd ? b01 : b23, 1);
        if (!valid) return;
        const uint32_t a = odd ? r : b01, b = odd ? b23 : r;
        __stcs(reinterpret_cast<ulonglong2 *>(out64 + (odd ? i - 2 : i)), make_ulonglong2(a & 0xffffu, a >> 16));
        __stcs(reinterpret_cast<ulonglong2 *>(out64 + (odd ? i + 2 : i + 4)), make_ulonglong2(b & 0xffffu, b >> 16));
        if constexpr (HOST_FORMS) {
            over8 |= ((lo | hi) & 0xff00ff00u) != 0u;
            *reinterpret_cast<uint2 *>(o16 + i) = make_uint2(b01, b23);
            if (o8) *reinterpret_cast<uint32_t *>(o8 + i) = __byte_perm(b01, b23, 0x6420);
        }
    };
    // Four neighbouring bins per thread and chunk: one 4-byte (narrow) or 16-byte (wide) load per
    // sender.  The trip counts are warp-uniform (w0 = the index of the warp's first lane).
    const uint64_t stride = uint64_t(gridDim.x) * blockDim.x * 4;
    const uint32_t lane4 = (threadIdx.x & 31u) * 4u;
    const uint64_t first = (uint64_t(blockIdx.x) * blockDim.x + (threadIdx.x & ~31u)) * 4;
    if (all_narrow_s) {
        // every sender sent u8 rows: U chunks per trip, the loads of all of them (and of two
        // senders) in flight together, no branch in the loop; escapes are patched afterwards
        constexpr int U = 4;
        for (uint64_t w0 = first; w0 < n; w0 += stride * U) {
            const uint64_t i0 = w0 + lane4;
            // bytes 0 and 2 of a word are summed in the halves of lo, bytes 1 and 3 in those of hi
            // (at most 16 senders x 255: no carry between the halves)
            uint32_t lo[U], hi[U], esc = 0;
#pragma unroll
            for (int u = 0; u < U; ++u) lo[u] = hi[u] = 0u;
#pragma unroll 2
            for (int s = 0; s < world; ++s) {
                uint32_t w[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint64_t i = i0 + stride * u;
                    w[u] = i < n ? __ldcs(reinterpret_cast<const uint32_t *>(narrow + uint64_t(s) * cap + i)) : 0u;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    esc |= ((w[u] & (w[u] >> 4) & 0x0f0f0f0fu) + 0x01010101u) & 0x10101010u;       // a byte = 255
                    lo[u] += w[u] & 0x00ff00ffu;
                    hi[u] += (w[u] >> 8) & 0x00ff00ffu;
                }
            }
            if (!__any_sync(0xffffffffu, esc != 0u)) {
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (w0 + stride * u < n) emit_packed(i0 + stride * u, lo[u], hi[u], i0 + stride * u < n);
                continue;
            }
            // rare: replace every 255 by the value in the sender's wide row set
            unsigned long long acc[U][4];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                acc[u][0] = lo[u] & 0xffffu; acc[u][1] = hi[u] & 0xffffu; acc[u][2] = lo[u] >> 16; acc[u][3] = hi[u] >> 16;
            }
            if (esc) {
                for (int s = 0; s < world; ++s)
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const uint64_t i = i0 + stride * u;
                        if (i >= n) continue;
                        const uint32_t w = __ldcs(reinterpret_cast<const uint32_t *>(narrow + uint64_t(s) * cap + i));
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (((w >> (8 * j)) & 0xffu) == 0xffu)
                                acc[u][j] += (unsigned long long)(__ldcs(wide + uint64_t(s) * cap + i + j)) - 0xffull;
                    }
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (w0 + stride * u < n) emit(i0 + stride * u, acc[u], i0 + stride * u < n);
        }
    } else {
        for (uint64_t w0 = first; w0 < n; w0 += stride) {
            const uint64_t i = w0 + lane4;
            unsigned long long acc[4] = {0, 0, 0, 0};
            if (i < n) {
                for (int s = 0; s < world; ++s) {
                    if (!wide_s[s]) {
                        const uint32_t w = __ldcs(reinterpret_cast<const uint32_t *>(narrow + uint64_t(s) * cap + i));
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint32_t b = (w >> (8 * j)) & 0xffu;
                            if (b == 0xffu) b = __ldcs(wide + uint64_t(s) * cap + i + j);
                            acc[j] += b;
                        }
                    } else {
                        const uint4 x = __ldcs(reinterpret_cast<const uint4 *>(wide + uint64_t(s) * cap + i));
                        acc[0] += x.x; acc[1] += x.y; acc[2] += x.z; acc[3] += x.w;
                    }
                }
            }
            emit(i, acc, i < n);
        }
    }
    if (HOST_FORMS && flags) {
        if (over16) flags[0] = 1u;
        if (over8) flags[1] = 1u;
    }
}

uint64_t slice_inbox_bytes(int k, int world)
{
    return 2 * inbox_parity_bytes(1ull << (2 * k), world);
}
uint64_t slice_begin_host(int k, int o, int world) { return slice64_begin(1ull << (2 * k), o, world); }

static int slice_args(int k, int counter_bits, int rank, int world, void *const *inbox_ptrs, SliceInbox *out)
{
    KPAL_CHECK(peer_check_args_fwd(k, counter_bits, rank, world));
    if (k < 6) return bad_arg("the sliced reduce needs k >= 6 (64 x 64 balance tiles)");
    if ((1ull << (2 * k)) < 64ull * world) return bad_arg("table smaller than one 64-count row per rank");
    if (!inbox_ptrs) return bad_arg("null pointer");
    out->rank = rank; out->world = world;
    for (int i = 0; i < kMaxPeers; ++i) out->base[i] = i < world ? inbox_ptrs[i] : nullptr;
    for (int i = 0; i < world; ++i) if (!out->base[i]) return bad_arg("null inbox pointer");
    return KPAL_OK;
}

// balance + push of this rank's table (the last CTA signals the peers).  d_state: 4 device words,
// zero before the first launch.  wide_rows != 0: u32 rows instead of u8 rows with escapes.
int launch_slice_push(const void *d_table, int counter_bits, int k, int rank, int world, void *const *inbox_ptrs,
                      unsigned long long epoch, int wide_rows, unsigned int *d_state, cudaStream_t stream)
{
    SliceInbox peers;
    KPAL_CHECK(slice_args(k, counter_bits, rank, world, inbox_ptrs, &peers));
    if (!d_table || !d_state) return bad_arg("null pointer");
    const int parity = int(epoch & 1ull);
    const unsigned pairs = canon_count(k - 6);
    const size_t smem = size_t(2) * 2 * 64 * kPushRow * (counter_bits / 8);     // two stages of two tiles
    // a few CTAs per SM that loop over the tile pairs: one system-wide fence per CTA, at its end
    const unsigned grid = std::min<unsigned>(pairs, unsigned(sm_count()) * (counter_bits == 32 ? 3 : 1));
    static bool opted_in[4] = {false, false, false, false};     // per kernel: more than 48 KB of dynamic shared memory
    auto launch = [&](auto kernel, auto *table) {
        bool &done = opted_in[(counter_bits == 64 ? 2 : 0) + (wide_rows ? 1 : 0)];
        if (!done && cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess) return;
        done = true;
        kernel<<<grid, 256, smem, stream>>>(table, k, peers, parity, epoch, d_state);
    };
    if (counter_bits == 32) {
        const uint32_t *t = static_cast<const uint32_t *>(d_table);
        if (wide_rows) launch(slice_push_kernel<uint32_t, true>, t); else launch(slice_push_kernel<uint32_t, false>, t);
    } else {
        const unsigned long long *t = static_cast<const unsigned long long *>(d_table);
        if (wide_rows) launch(slice_push_kernel<unsigned long long, true>, t); else launch(slice_push_kernel<unsigned long long, false>, t);
    }
    KPAL_LAUNCH_CHECK("slice_push_kernel");
    return KPAL_OK;
}

int launch_slice_signal(int k, int rank, int world, void *const *inbox_ptrs, unsigned long long epoch, int wide_rows,
                        cudaStream_t stream)
{
    SliceInbox peers;
    KPAL_CHECK(slice_args(k, 32, rank, world, inbox_ptrs, &peers));
    slice_signal_kernel<<<1, 32, 0, stream>>>(peers, int(epoch & 1ull), epoch, 1ull << (2 * k), wide_rows);
    KPAL_LAUNCH_CHECK("slice_signal_kernel");
    return KPAL_OK;
}

// signal: -1 = the peers have been told already (launch_slice_signal); 0 / 1 = tell them first (the
// wide_rows of this rank's push, which must be the previous work on `stream`).
int launch_slice_collect(int k, int rank, int world, void *const *inbox_ptrs, unsigned long long epoch, int signal,
                         int64_t *d_out64, uint16_t *d_o16, uint8_t *d_o8, unsigned int *d_flags, cudaStream_t stream)
{
    SliceInbox peers;
    KPAL_CHECK(slice_args(k, 32, rank, world, inbox_ptrs, &peers));
    if (!d_out64) return bad_arg("null pointer");
    const uint64_t bins = 1ull << (2 * k);
    const uint64_t n = slice64_begin(bins, rank + 1, world) - slice64_begin(bins, rank, world);
    // a few CTAs per SM, looping: every CTA waits for the signals once
    const unsigned grid = unsigned(std::max<uint64_t>(1, std::min<uint64_t>((n / 16 + 255) / 256, uint64_t(sm_count()) * 4)));
    if (d_flags) KPAL_CUDA(cudaMemsetAsync(d_flags, 0, 8, stream));
    if (d_o16)
        slice_collect_kernel<true><<<grid, 256, 0, stream>>>(peers, signal, bins, int(epoch & 1ull), epoch, d_out64, d_o16, d_o8, d_flags);
    else
        slice_collect_kernel<false><<<grid, 256, 0, stream>>>(peers, signal, bins, int(epoch & 1ull), epoch, d_out64, d_o16, d_o8, d_flags);
    KPAL_LAUNCH_CHECK("slice_collect_kernel");
    return KPAL_OK;
}

uint64_t peer_inbox_bytes(int k, int counter_bits, int world)
{
    const uint64_t bins = 1ull << (2 * k);
    return slot_elems(bins, world) * uint64_t(world) * (counter_bits / 8);
}

int peer_check_args(int k, int counter_bits, int rank, int world)
{
    if (k < 1 || k > KPAL_MAX_K) return bad_arg("k out of range");
    if (counter_bits != 32 && counter_bits != 64) return bad_arg("counter_bits must be 32 or 64");
    if (world < 1 || world > kMaxPeers) return bad_arg("world size out of range [1, 16]");
    if (rank < 0 || rank >= world) return bad_arg("rank outside the world");
    if ((1ull << (2 * k)) < 4ull * world) return bad_arg("table smaller than 4 entries per rank");
    return KPAL_OK;
}

static int peer_check_args_fwd(int k, int counter_bits, int rank, int world)
{
    return peer_check_args(k, counter_bits, rank, world);
}

int launch_reduce_push(const void *d_table, int counter_bits, int k, int rank, int world,
                       void *const *inbox_ptrs, cudaStream_t stream)
{
    KPAL_CHECK(peer_check_args(k, counter_bits, rank, world));
    if (!d_table || !inbox_ptrs) return bad_arg("null pointer");
    PeerOut pp;
    pp.rank = rank; pp.world = world;
    for (int i = 0; i < kMaxPeers; ++i) pp.inbox[i] = i < world ? inbox_ptrs[i] : nullptr;
    for (int i = 0; i < world; ++i) if (!pp.inbox[i]) return bad_arg("null inbox pointer");
    const uint64_t bins = 1ull << (2 * k);
    const uint64_t nv = (bins / world + 4) / (16 / (counter_bits / 8));
    const unsigned gx = unsigned(std::max<uint64_t>(1, std::min<uint64_t>((nv + 255) / 256,
                                                                        uint64_t(sm_count()) * 8 / world + 1)));
    const dim3 grid(gx, unsigned(world));
    if (counter_bits == 32)
        reduce_push_kernel<uint32_t><<<grid, 256, 0, stream>>>(static_cast<const uint32_t *>(d_table), pp, bins);
    else
        reduce_push_kernel<unsigned long long><<<grid, 256, 0, stream>>>(
            static_cast<const unsigned long long *>(d_table), pp, bins);
    KPAL_LAUNCH_CHECK("reduce_push_kernel");
    return KPAL_OK;
}

int launch_reduce_collect(const void *d_inbox, int counter_bits, int k, int rank, int world,
                          void *d_root_table, cudaStream_t stream)
{
    KPAL_CHECK(peer_check_args(k, counter_bits, rank, world));
    if (!d_inbox || !d_root_table) return bad_arg("null pointer");
    const uint64_t bins = 1ull << (2 * k);
    const uint64_t nv = (bins / world + 4) / (16 / (counter_bits / 8));
    const unsigned gx = unsigned(std::max<uint64_t>(1, std::min<uint64_t>((nv + 255) / 256, uint64_t(sm_count()) * 8)));
    if (counter_bits == 32)
        reduce_collect_kernel<uint32_t><<<gx, 256, 0, stream>>>(static_cast<const uint32_t *>(d_inbox), rank, world, bins,
                                                                static_cast<uint32_t *>(d_root_table));
    else
        reduce_collect_kernel<unsigned long long><<<gx, 256, 0, stream>>>(
            static_cast<const unsigned long long *>(d_inbox), rank, world, bins,
            static_cast<unsigned long long *>(d_root_table));
    KPAL_LAUNCH_CHECK("reduce_collect_kernel");
    return KPAL_OK;
}

}  // namespace kpal
