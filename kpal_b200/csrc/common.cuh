// Shared helpers for libkpal_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/kpal_b200.h"

namespace kpal {

// ---- per-thread error message -------------------------------------------
void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define KPAL_CUDA(call)                                                          \
    do {                                                                         \
        cudaError_t _e = (call);                                                 \
        if (_e != cudaSuccess) {                                                 \
            kpal::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), \
                            __FILE__, __LINE__);                                 \
            return (_e == cudaErrorMemoryAllocation) ? KPAL_ENOMEM : KPAL_ECUDA; \
        }                                                                        \
    } while (0)

#define KPAL_LAUNCH_CHECK(name)                                                  \
    do {                                                                         \
        kpal::g_launches.fetch_add(1, std::memory_order_relaxed);                \
        cudaError_t _e = cudaGetLastError();                                     \
        if (_e != cudaSuccess) {                                                 \
            kpal::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e)); \
            return KPAL_ECUDA;                                                   \
        }                                                                        \
    } while (0)

#define KPAL_CHECK(expr)                                                         \
    do {                                                                         \
        int _r = (expr);                                                         \
        if (_r != KPAL_OK) return _r;                                            \
    } while (0)

inline int bad_arg(const char *what)
{
    set_error("invalid argument: %s", what);
    return KPAL_EINVAL;
}

// number of SMs of the current device (cached)
int sm_count();

// ---- packed-stream geometry ------------------------------------------------
// One thread of the count kernels consumes one CHUNK = 64 bases = one uint4 of
// codes + one uint2 of validity bits.
constexpr int kChunkBases = 64;
inline uint64_t n_chunks_of(uint64_t n_bases) { return (n_bases + kChunkBases - 1) / kChunkBases; }

// ---- device helpers -------------------------------------------------------

// Reverse the order of the sixteen 2-bit groups of a 32-bit word.
__device__ __forceinline__ uint32_t rev2(uint32_t x)
{
    x = __brev(x);
    return ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
}

// Reverse complement of a k-mer index (kpal/klib.py:394-412): complement is
// bitwise NOT (A0 C1 G2 T3), then the k 2-bit groups are reversed.
__device__ __forceinline__ uint32_t rc_index(uint32_t idx, int shift /* 32 - 2k */)
{
    return (~rev2(idx)) >> shift;
}

// ---- multi-GPU: peer-memory reduce of counter tables (peer_reduce.cu) --------
// Rank o owns the table slice [slice_begin(o), slice_begin(o+1)); every rank has an
// inbox of `world` slots of slot_elems() counters, slot s filled by rank s.
constexpr int kMaxPeers = 16;
struct PeerOut {
    void *inbox[kMaxPeers];     // device pointers valid on THIS device (IPC-mapped peers)
    int rank, world;
};
// first table index of rank o's slice: multiples of 4 elements so that every
// slice moves as whole 16-byte vectors
__host__ __device__ inline uint64_t slice_begin(uint64_t bins, int o, int world)
{
    if (o >= world) return bins;
    return (bins * uint64_t(o) / uint64_t(world)) & ~uint64_t(3);
}
__host__ __device__ inline uint64_t slot_elems(uint64_t bins, int world)
{
    return ((bins + world - 1) / world + 7) & ~uint64_t(3);        // >= the longest slice
}

}  // namespace kpal
