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
