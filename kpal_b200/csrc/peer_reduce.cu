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
//                  u32 table.  A count above 255 raises a flag; the row sets are then sent
//                  again as u32 by a second launch (a no-op otherwise).
//   signal       one 8-byte release store per peer: "my rows of this epoch have landed".
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
struct SliceInbox {
    void *base[kMaxPeers];      // inbox of rank o (both parities), mapped on THIS device
    int rank, world;
};

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
__device__ __forceinline__ int owner_of(uint64_t index, uint64_t bins, int world)
{
    int o = int(index * uint64_t(world) / bins);
    while (o + 1 < world && slice64_begin(bins, o + 1, world) <= index) ++o;
    while (o > 0 && slice64_begin(bins, o, world) > index) --o;
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
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(value) : "memory");
}

// WIDE = false: u8 rows + overflow detection; WIDE = true: u32 rows, only when wide_flag[0] != 0.
// state = { wide needed, count above 32 bits (caller's error), CTAs done (narrow), CTAs done (wide) }.
// The last CTA to finish sends the signal: of the narrow launch when no count exceeded 255,
// else of the wide launch (threadfence-reduction pattern: every CTA fences its peer stores
// system-wide before it counts itself done).
template <typename CounterT, bool WIDE>
__global__ void __launch_bounds__(256)
slice_push_kernel(const CounterT *__restrict__ table, int k, const SliceInbox peers, int parity,
                  unsigned long long epoch, unsigned int *__restrict__ state)
{
    extern __shared__ __align__(16) unsigned char push_smem[];
    __shared__ unsigned int last_s;
    if (WIDE && *reinterpret_cast<volatile unsigned int *>(state) == 0) return;
    CounterT *A = reinterpret_cast<CounterT *>(push_smem);      // [64][65] tile of m
    CounterT *B = A + 64 * 65;                                  // [64][65] tile of rc(m)
    const int mid_bits = 2 * (k - 6);
    const int hshift = 2 * k - 6;
    const uint64_t bins = 1ull << (2 * k);
    const uint32_t tiles = 1u << mid_bits;
    const uint64_t cap = slice64_cap(bins, peers.world);
    const uint64_t par_off = uint64_t(parity) * inbox_parity_bytes(bins, peers.world);
    bool big = false;
    for (uint32_t m = blockIdx.x; m < tiles; m += gridDim.x) {
        const uint32_t mr = mid_bits ? ((~rev2(m)) >> (32 - mid_bits)) : 0u;
        if (m > mr) continue;                                   // done together with rc(m)
        constexpr int V = 16 / int(sizeof(CounterT));
        constexpr int NV = 4096 / V / 256;
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
        __syncthreads();                                        // the previous tile pair has been pushed
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
        // eight neighbouring counts per thread: one 8-byte (narrow) or two 16-byte (wide) peer stores
        for (uint32_t e = threadIdx.x; e < 512; e += 256) {
            const uint32_t h = e >> 3, l = (e & 7u) * 8;
            const uint32_t rh = rc_index(h, 26);
            for (int tile = 0; tile < (m != mr ? 2 : 1); ++tile) {
                const CounterT *own = tile ? B : A, *other = tile ? A : partner;
                const uint64_t index = (uint64_t(h) << hshift) | (uint64_t(tile ? mr : m) << 6) | l;
                unsigned long long v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    v[j] = (unsigned long long)(own[h * 65 + l + j]) + (unsigned long long)(other[rc_index(l + j, 26) * 65 + rh]);
                const int o = owner_of(index, bins, peers.world);
                unsigned char *inbox = static_cast<unsigned char *>(peers.base[o]) + par_off;
                const uint64_t at = uint64_t(peers.rank) * cap + (index - slice64_begin(bins, o, peers.world));
                if constexpr (!WIDE) {
                    unsigned long long packed = 0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        big |= v[j] > 0xffull;
                        packed |= (v[j] & 0xffull) << (8 * j);
                    }
                    *reinterpret_cast<unsigned long long *>(inbox + 256 + at) = packed;
                } else {
                    uint4 *dst = reinterpret_cast<uint4 *>(inbox + 256 + uint64_t(peers.world) * cap) + at / 4;
                    dst[0] = make_uint4(uint32_t(v[0]), uint32_t(v[1]), uint32_t(v[2]), uint32_t(v[3]));
                    dst[1] = make_uint4(uint32_t(v[4]), uint32_t(v[5]), uint32_t(v[6]), uint32_t(v[7]));
#pragma unroll
                    for (int j = 0; j < 8; ++j) big |= v[j] > 0xffffffffull;
                }
            }
        }
    }
    if (big) state[WIDE ? 1 : 0] = 1u;
    // ---- done: fence the peer stores, count this CTA; the last one signals the peers
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last_s = atomicAdd(state + (WIDE ? 3 : 2), 1u) == gridDim.x - 1 ? 1u : 0u;
    }
    __syncthreads();
    if (last_s) {
        __threadfence();
        const bool wide_needed = *reinterpret_cast<volatile unsigned int *>(state) != 0;
        if (WIDE || !wide_needed) slice_signal(peers, parity, epoch, bins, WIDE);
    }
}

// Owner side.  out64: this rank's slice of the final profile (int64); o16 / o8 / flags (optional):
// its narrow forms for the device->host copy (flags[0]: a count above 65535, flags[1]: above 255).
__global__ void __launch_bounds__(256)
slice_collect_kernel(const unsigned char *__restrict__ inbox, int rank, int world, uint64_t bins, int parity,
                     unsigned long long epoch, int64_t *__restrict__ out64, uint16_t *__restrict__ o16,
                     uint8_t *__restrict__ o8, unsigned int *__restrict__ flags)
{
    __shared__ unsigned int wide_s[kMaxPeers];
    const unsigned char *base = inbox + uint64_t(parity) * inbox_parity_bytes(bins, world);
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
    const uint64_t cap = slice64_cap(bins, world);
    const uint64_t lo = slice64_begin(bins, rank, world), hi = slice64_begin(bins, rank + 1, world);
    const unsigned char *narrow = base + 256;
    const uint32_t *wide = reinterpret_cast<const uint32_t *>(base + 256 + uint64_t(world) * cap);
    bool over8 = false, over16 = false;
    for (uint64_t i = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) * 16; i < hi - lo;
         i += uint64_t(gridDim.x) * blockDim.x * 16) {
        unsigned long long acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = 0;
        for (int s = 0; s < world; ++s) {
            if (!wide_s[s]) {
                const uint4 x = __ldcs(reinterpret_cast<const uint4 *>(narrow + uint64_t(s) * cap + i));
                const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j] += (w[j >> 2] >> (8 * (j & 3))) & 0xffu;
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint4 x = __ldcs(reinterpret_cast<const uint4 *>(wide + uint64_t(s) * cap + i) + q);
                    acc[4 * q] += x.x; acc[4 * q + 1] += x.y; acc[4 * q + 2] += x.z; acc[4 * q + 3] += x.w;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 16; j += 2)
            *reinterpret_cast<ulonglong2 *>(out64 + i + j) = make_ulonglong2(acc[j], acc[j + 1]);
        if (o16) {
            uint32_t p16[8], p8[4] = {0, 0, 0, 0};
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                over8 |= acc[j] > 0xffull;
                over16 |= acc[j] > 0xffffull;
                if (j & 1) p16[j >> 1] |= uint32_t(acc[j] & 0xffffull) << 16; else p16[j >> 1] = uint32_t(acc[j] & 0xffffull);
                p8[j >> 2] |= uint32_t(acc[j] & 0xffull) << (8 * (j & 3));
            }
            reinterpret_cast<uint4 *>(o16 + i)[0] = make_uint4(p16[0], p16[1], p16[2], p16[3]);
            reinterpret_cast<uint4 *>(o16 + i)[1] = make_uint4(p16[4], p16[5], p16[6], p16[7]);
            if (o8) *reinterpret_cast<uint4 *>(o8 + i) = make_uint4(p8[0], p8[1], p8[2], p8[3]);
        }
    }
    if (o16 && flags) {
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

// balance + narrow push of this rank's table (the last CTA signals the peers).  d_wide_flag: 4 device words.
int launch_slice_push(const void *d_table, int counter_bits, int k, int rank, int world, void *const *inbox_ptrs,
                      unsigned long long epoch, unsigned int *d_wide_flag, cudaStream_t stream)
{
    SliceInbox peers;
    KPAL_CHECK(slice_args(k, counter_bits, rank, world, inbox_ptrs, &peers));
    if (!d_table || !d_wide_flag) return bad_arg("null pointer");
    const int parity = int(epoch & 1ull);
    const unsigned tiles = 1u << (2 * (k - 6));
    const size_t smem = size_t(2) * 64 * 65 * (counter_bits / 8);
    KPAL_CUDA(cudaMemsetAsync(d_wide_flag, 0, 16, stream));
    // the wide launch is a no-op unless a count exceeded 255: a small grid that loops over the tiles
    const unsigned wide_grid = std::min<unsigned>(tiles, unsigned(sm_count()) * 2);
    if (counter_bits == 32) {
        slice_push_kernel<uint32_t, false><<<tiles, 256, smem, stream>>>(static_cast<const uint32_t *>(d_table), k, peers, parity, epoch, d_wide_flag);
        KPAL_LAUNCH_CHECK("slice_push_kernel");
        slice_push_kernel<uint32_t, true><<<wide_grid, 256, smem, stream>>>(static_cast<const uint32_t *>(d_table), k, peers, parity, epoch, d_wide_flag);
        KPAL_LAUNCH_CHECK("slice_push_kernel");
    } else {
        KPAL_CUDA(cudaFuncSetAttribute(slice_push_kernel<unsigned long long, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        KPAL_CUDA(cudaFuncSetAttribute(slice_push_kernel<unsigned long long, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        slice_push_kernel<unsigned long long, false><<<tiles, 256, smem, stream>>>(static_cast<const unsigned long long *>(d_table), k, peers, parity, epoch, d_wide_flag);
        KPAL_LAUNCH_CHECK("slice_push_kernel");
        slice_push_kernel<unsigned long long, true><<<wide_grid, 256, smem, stream>>>(static_cast<const unsigned long long *>(d_table), k, peers, parity, epoch, d_wide_flag);
        KPAL_LAUNCH_CHECK("slice_push_kernel");
    }
    return KPAL_OK;
}

int launch_slice_collect(const void *d_inbox, int k, int rank, int world, unsigned long long epoch, int64_t *d_out64,
                         uint16_t *d_o16, uint8_t *d_o8, unsigned int *d_flags, cudaStream_t stream)
{
    KPAL_CHECK(peer_check_args_fwd(k, 32, rank, world));
    if (!d_inbox || !d_out64) return bad_arg("null pointer");
    const uint64_t bins = 1ull << (2 * k);
    const uint64_t n = slice64_begin(bins, rank + 1, world) - slice64_begin(bins, rank, world);
    const unsigned grid = unsigned(std::max<uint64_t>(1, std::min<uint64_t>((n / 16 + 255) / 256, uint64_t(sm_count()) * 8)));
    if (d_flags) KPAL_CUDA(cudaMemsetAsync(d_flags, 0, 8, stream));
    slice_collect_kernel<<<grid, 256, 0, stream>>>(static_cast<const unsigned char *>(d_inbox), rank, world, bins,
                                                   int(epoch & 1ull), epoch, d_out64, d_o16, d_o8, d_flags);
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
