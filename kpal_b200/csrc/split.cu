// Profile.split (kpal/klib.py:300-327) and the `showbalance` figure built on it
// (kmer.get_balance, kpal/kmer.py:222-247) -- the callers next to Profile.balance
// that use the same reverse-complement permutation.
//
//   split      forward / reverse lists over the indices i <= rc(i), in index
//              order: (2 c[i], 2 c[rc(i)]) for i < rc(i), (c[i], c[i]) for a
//              palindrome.  An order-preserving compaction: per-tile counts of
//              kept indices, one single-CTA exclusive scan over the tiles, then a
//              ballot/popc ranking inside each tile.
//   balance    multiset(forward, reverse, prod) (kpal/metrics.py:101-123,160) without
//   figure     materialising the lists: sum over i <= rc(i) of |x-y| / ((x+1)(y+1))
//              in the reference's arithmetic (int64 numerator and denominator, one
//              IEEE division per element), divided by (#positions with x or y
//              non-zero) + 1.
//
// Both are one pass over int64[4^k]: HBM-bound (8 B read per entry; the rc gather
// touches every entry exactly once more, through L2).
#include "common.cuh"

namespace kpal {

constexpr int kSplitThreads = 256;
constexpr int kSplitPerThread = 8;
constexpr int kSplitTile = kSplitThreads * kSplitPerThread;      // 2048 indices per tile

__device__ __forceinline__ bool split_keep(uint32_t i, int shift, uint32_t *partner)
{
    const uint32_t r = rc_index(i, shift);
    *partner = r;
    return i <= r;
}

// kept indices per tile
__global__ void __launch_bounds__(kSplitThreads)
split_count_kernel(uint64_t bins, int k, uint32_t *__restrict__ tile_counts)
{
    const int shift = 32 - 2 * k;
    const uint64_t base = uint64_t(blockIdx.x) * kSplitTile;
    uint32_t mine = 0;
#pragma unroll
    for (int j = 0; j < kSplitPerThread; ++j) {
        const uint64_t i = base + uint64_t(j) * kSplitThreads + threadIdx.x;
        uint32_t r;
        if (i < bins && split_keep(uint32_t(i), shift, &r)) ++mine;
    }
    __shared__ uint32_t warp_sum[kSplitThreads / 32];
    for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < kSplitThreads / 32; ++w) t += warp_sum[w];
        tile_counts[blockIdx.x] = t;
    }
}

// exclusive scan of the tile counts (<= 2^19 tiles at k = 15): one CTA, 64-bit offsets
__global__ void __launch_bounds__(1024)
split_scan_kernel(const uint32_t *__restrict__ tile_counts, uint64_t n_tiles,
                  uint64_t *__restrict__ tile_offsets, uint64_t *__restrict__ total)
{
    __shared__ uint64_t part[1024];
    const uint64_t per = (n_tiles + 1023) / 1024;
    const uint64_t lo = min(n_tiles, per * threadIdx.x), hi = min(n_tiles, lo + per);
    uint64_t s = 0;
    for (uint64_t t = lo; t < hi; ++t) s += tile_counts[t];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t run = 0;
        for (int i = 0; i < 1024; ++i) { const uint64_t v = part[i]; part[i] = run; run += v; }
        *total = run;
    }
    __syncthreads();
    uint64_t run = part[threadIdx.x];
    for (uint64_t t = lo; t < hi; ++t) { tile_offsets[t] = run; run += tile_counts[t]; }
}

// rank inside the tile (index order = warp-major over the j loop) and scatter
__global__ void __launch_bounds__(kSplitThreads)
split_scatter_kernel(const int64_t *__restrict__ counts, uint64_t bins, int k,
                     const uint64_t *__restrict__ tile_offsets,
                     int64_t *__restrict__ forward, int64_t *__restrict__ reverse)
{
    const int shift = 32 - 2 * k;
    const uint64_t base = uint64_t(blockIdx.x) * kSplitTile;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    __shared__ uint32_t warp_cnt[kSplitPerThread][kSplitThreads / 32];
    uint32_t ballots[kSplitPerThread], partners[kSplitPerThread];
#pragma unroll
    for (int j = 0; j < kSplitPerThread; ++j) {
        const uint64_t i = base + uint64_t(j) * kSplitThreads + threadIdx.x;
        uint32_t r = 0;
        const bool keep = i < bins && split_keep(uint32_t(i), shift, &r);
        partners[j] = r;
        ballots[j] = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_cnt[j][warp] = __popc(ballots[j]);
    }
    __syncthreads();
    // exclusive prefix over (j, warp) in index order: index = base + j*256 + warp*32 + lane
    __shared__ uint32_t warp_off[kSplitPerThread][kSplitThreads / 32];
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int j = 0; j < kSplitPerThread; ++j)
            for (int w = 0; w < kSplitThreads / 32; ++w) { warp_off[j][w] = run; run += warp_cnt[j][w]; }
    }
    __syncthreads();
    const uint64_t tile_off = tile_offsets[blockIdx.x];
#pragma unroll
    for (int j = 0; j < kSplitPerThread; ++j) {
        if (!((ballots[j] >> lane) & 1u)) continue;
        const uint64_t i = base + uint64_t(j) * kSplitThreads + threadIdx.x;
        const uint64_t at = tile_off + warp_off[j][warp] + __popc(ballots[j] & ((1u << lane) - 1u));
        const int64_t ci = counts[i];
        if (uint32_t(i) == partners[j]) { forward[at] = ci; reverse[at] = ci; }
        else { forward[at] = 2 * ci; reverse[at] = 2 * __ldg(counts + partners[j]); }
    }
}

// multiset(forward, reverse, prod) straight from the profile
__global__ void __launch_bounds__(256)
show_balance_kernel(const int64_t *__restrict__ counts, uint64_t bins, int k,
                    double *__restrict__ sum_out, unsigned long long *__restrict__ nz_out)
{
    const int shift = 32 - 2 * k;
    double s = 0.0;
    unsigned long long nz = 0;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < bins;
         i += uint64_t(gridDim.x) * blockDim.x) {
        uint32_t r;
        if (!split_keep(uint32_t(i), shift, &r)) continue;
        const int64_t ci = counts[i];
        if (uint32_t(i) == r) { nz += ci != 0; continue; }      // palindrome: x == y, term 0
        const int64_t x = 2 * ci, y = 2 * __ldg(counts + r);
        if ((x | y) == 0) continue;
        ++nz;
        const int64_t num = x > y ? x - y : y - x;
        s += __ddiv_rn(double(num), double((x + 1) * (y + 1)));
    }
    for (int o = 16; o; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        nz += __shfl_xor_sync(0xffffffffu, nz, o);
    }
    __shared__ double ws[8];
    __shared__ unsigned long long wn[8];
    if ((threadIdx.x & 31) == 0) { ws[threadIdx.x >> 5] = s; wn[threadIdx.x >> 5] = nz; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { s += ws[w]; nz += wn[w]; }
        atomicAdd(sum_out, s);
        atomicAdd(nz_out, nz);
    }
}

// metrics.positive on both sides of a pair (kpal/kdistlib.py:143-145, kpal/metrics.py:89-98):
// keep only the positions that are non-zero in both profiles.  In place.
__global__ void __launch_bounds__(256)
positive_pair_kernel(int64_t *__restrict__ left, int64_t *__restrict__ right, uint64_t bins)
{
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < bins;
         i += uint64_t(gridDim.x) * blockDim.x) {
        const int64_t l = left[i], r = right[i];
        if (l != 0 && r == 0) left[i] = 0;
        if (r != 0 && l == 0) right[i] = 0;
    }
}

int launch_positive_pair(int64_t *d_left, int64_t *d_right, int k, cudaStream_t stream)
{
    if (k < 1 || k > KPAL_MAX_K) return bad_arg("k out of range");
    const uint64_t bins = 1ull << (2 * k);
    const uint64_t want = (bins + 255) / 256, cap = uint64_t(sm_count()) * 8;
    positive_pair_kernel<<<unsigned(want < cap ? want : cap), 256, 0, stream>>>(d_left, d_right, bins);
    KPAL_LAUNCH_CHECK("positive_pair_kernel");
    return KPAL_OK;
}

uint64_t split_length(int k)
{
    const uint64_t bins = 1ull << (2 * k);
    const uint64_t palindromes = (k % 2 == 0) ? (1ull << k) : 0;      // 4^(k/2), even k only
    return (bins + palindromes) / 2;
}

uint64_t split_scratch_bytes(int k)
{
    const uint64_t tiles = ((1ull << (2 * k)) + kSplitTile - 1) / kSplitTile;
    return tiles * 4 + tiles * 8 + 16;
}

int launch_split(const int64_t *d_counts, int k, int64_t *d_forward, int64_t *d_reverse,
                 void *d_scratch, cudaStream_t stream)
{
    if (k < 1 || k > KPAL_MAX_K) return bad_arg("k out of range");
    const uint64_t bins = 1ull << (2 * k);
    const uint64_t tiles = (bins + kSplitTile - 1) / kSplitTile;
    uint32_t *tile_counts = static_cast<uint32_t *>(d_scratch);
    uint64_t *tile_offsets = reinterpret_cast<uint64_t *>(
        static_cast<unsigned char *>(d_scratch) + ((tiles * 4 + 7) / 8) * 8);
    uint64_t *total = tile_offsets + tiles;
    split_count_kernel<<<unsigned(tiles), kSplitThreads, 0, stream>>>(bins, k, tile_counts);
    KPAL_LAUNCH_CHECK("split_count_kernel");
    split_scan_kernel<<<1, 1024, 0, stream>>>(tile_counts, tiles, tile_offsets, total);
    KPAL_LAUNCH_CHECK("split_scan_kernel");
    split_scatter_kernel<<<unsigned(tiles), kSplitThreads, 0, stream>>>(d_counts, bins, k, tile_offsets,
                                                                        d_forward, d_reverse);
    KPAL_LAUNCH_CHECK("split_scatter_kernel");
    return KPAL_OK;
}

int launch_show_balance(const int64_t *d_counts, int k, double *d_sum, unsigned long long *d_nz,
                        cudaStream_t stream)
{
    if (k < 1 || k > KPAL_MAX_K) return bad_arg("k out of range");
    const uint64_t bins = 1ull << (2 * k);
    KPAL_CUDA(cudaMemsetAsync(d_sum, 0, 8, stream));
    KPAL_CUDA(cudaMemsetAsync(d_nz, 0, 8, stream));
    const uint64_t want = (bins + 255) / 256, cap = uint64_t(sm_count()) * 8;
    show_balance_kernel<<<unsigned(want < cap ? want : cap), 256, 0, stream>>>(d_counts, bins, k, d_sum, d_nz);
    KPAL_LAUNCH_CHECK("show_balance_kernel");
    return KPAL_OK;
}

}  // namespace kpal
