// Shared-memory PTX helpers and the staging workspace of the radix count paths
// (count_radix.cu: one window per payload; count_pairs.cu: two windows per payload).
#pragma once
#include "common.cuh"

namespace kpal {

constexpr int kUnitBases = 32;          // bases (= window starts) per thread and unit
constexpr int kGroup = 16;              // payloads per 32-byte group

struct Unit {
    uint32_t w[3];      // 32 bases of codes + 16 look-ahead bases
    uint32_t starts;    // bit (31 - o) set <=> the window starting at base o is all-valid
};

template <int O>
__device__ __forceinline__ uint32_t unit_window(const Unit &u, int shift)
{
    constexpr int j = O / 16, r = O % 16;
    const uint32_t x = (r == 0) ? u.w[j] : __funnelshift_l(u.w[j + 1], u.w[j], 2 * r);
    return x >> shift;
}

// Explicit shared-state-space accesses: through generic pointers the compiler
// emitted generic ATOM / ST (+ QSPC checks) for the slot bookkeeping.
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return uint32_t(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t atoms_add(uint32_t addr, uint32_t v)
{
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v) : "memory");
}

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, const uint4 &v)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}


// Read-only global loads the compiler may not move: __ldg data is known to be immutable, so
// nvcc sinks such loads to their first use -- which turns a software prefetch (load the next
// tile's words, bin this tile, use them) back into a load-and-wait (ncu: the top stall of
// pair_partition_kernel was the first use of the "prefetched" words).  Volatile asm statements
// keep their order among themselves, and the shared-memory atomics / stores between are volatile too.
__device__ __forceinline__ uint2 ldg_keep_v2(const uint2 *p)
{
    uint2 v;
    asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ldg_keep_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ldg_keep_v4(const uint4 *p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// grow-only per-device staging of the radix paths (count_radix.cu); call under no lock
int radix_workspace(size_t staging_bytes, size_t fill_bytes, void **staging, uint32_t **fill);

}  // namespace kpal
