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

}  // namespace kpal
