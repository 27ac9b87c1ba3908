// N x N profile distance matrix for sm_100a.
//
// Replaces the pair loop of kdistlib.distance_matrix (reference
// kpal/kdistlib.py:179-184) over ProfileDistance.distance
// (kpal/kdistlib.py:126-161) and metrics.multiset / euclidean /
// cosine_similarity with the scale step (kpal/metrics.py:49-86,101-147,159-162).
//
// Algebra (DESIGN.md "Distance kernel").  With S the profile totals, the
// reference scales the profile with the smaller total by s = S_big/S_small
// (or, with `down`, the larger one by 1/s).  Call A the scaled and B the
// unscaled profile of a pair, x the raw counts, F = x/S the per-profile
// frequencies, P = x + 1 and t_B = 1/S_B.  Dividing numerator and denominator
// by S_B:
//
//   prod: |sA-b| / ((sA+1)(b+1)) = |F_A - F_B| / (F_A * P_B + (F_B + t_B))
//   sum : |sA-b| / (sA+b+1)      = |F_A - F_B| / (F_A + (F_B + t_B))
//   euclidean = S_B * sqrt(sum (F_A - F_B)^2),   cosine = <F_A,F_B>/(|F_A||F_B|)
//
// Everything pair-dependent is gone: F and P are computed once per profile
// (the reference redoes copy/balance/scale per pair), the numerator is exact
// (identical profiles give exactly 0), and one element pair costs 5 fp64
// instructions + 1 MUFU:
//     n   = F_A - F_B                      DADD
//     den = fma(F_A, P_B, F_B + t_B)       DFMA   (F_B + t_B: once per B element)
//     q0  = rcp.approx(den)                MUFU.RCP64H, ~20 good bits
//     h   = fma(-den, q0, 2)               DFMA   (Newton: 1/den = q0 * h, rel. err <= 2^-39)
//     acc = fma(|n| * q0, h, acc)          DMUL + DFMA
// Unscaled: S := 1 (F = x, t = 1).  The exact IEEE division is kept as a
// run-time option (kpal_set_option("exact_div", 1)) for validation.
// Profiles are visited in order of total, so in every tile the row panel is
// the A side and the column panel the B side.
//
// Tile kernel: CTA = TA A-rows x TB B-rows, compute warps own 32 x 32 pairs
// (each thread an 8 x 4 block, accumulators in registers) + 1 producer warp that
// streams DC-element row segments into a shared-memory ring with cp.async.bulk
// (TMA bulk copies, SASS UBLKCP) completing on mbarriers.
#include "common.cuh"

#include <stdlib.h>
#include <algorithm>

namespace kpal {

// ---------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------
// (overridable at compile time for tuning experiments: -DKPAL_WI=.. etc.)
#ifndef KPAL_WI
#define KPAL_WI 4
#endif
#ifndef KPAL_WJ
#define KPAL_WJ 2
#endif
#ifndef KPAL_RI
#define KPAL_RI 8
#endif
#ifndef KPAL_RJ
#define KPAL_RJ 4
#endif
#ifndef KPAL_STAGES
#define KPAL_STAGES 3
#endif
#ifndef KPAL_PRODUCER_WARP
#define KPAL_PRODUCER_WARP 1            // 1: dedicated TMA producer warp; 0: all warps issue their share
#endif
constexpr int WI = KPAL_WI;             // compute warps along the A (row) side
constexpr int WJ = KPAL_WJ;             // compute warps along the B (column) side
constexpr int RI = KPAL_RI;             // A rows per thread
constexpr int RJ = KPAL_RJ;             // B rows per thread
constexpr int NTI = WI * 4;             // threads along i (a warp is 4 x 8 threads)
constexpr int NTJ = WJ * 8;             // threads along j
constexpr int TA = NTI * RI;            // A rows (scaled side) per tile   (default 128)
constexpr int TB = NTJ * RJ;            // B rows (unscaled side) per tile (default 64)
constexpr int DC = 32;                  // profile elements per stage
constexpr int ROW_BYTES = DC * 8 + 16;  // +16: consecutive rows land in different bank groups
constexpr int STAGES = KPAL_STAGES;
constexpr int A_BYTES = TA * ROW_BYTES;
constexpr int B_BYTES = TB * ROW_BYTES;
constexpr int COMPUTE_WARPS = WI * WJ;
constexpr bool PRODUCER_WARP = KPAL_PRODUCER_WARP != 0;
// KPAL_SETMAXNREG: the producer sits in a warpgroup of its own (3 of its 4 warps retire at
// once) and hands its registers to the compute warpgroups with setmaxnreg, so the compute
// code is no longer held to the 168 registers a 9-warp CTA allows (3 warps on one scheduler).
#ifndef KPAL_SETMAXNREG
#define KPAL_SETMAXNREG 0
#endif
constexpr bool SETMAXNREG = (KPAL_SETMAXNREG != 0) && PRODUCER_WARP;
constexpr int PRODUCER_WARPS = PRODUCER_WARP ? (SETMAXNREG ? 4 : 1) : 0;
constexpr int TILE_THREADS = (COMPUTE_WARPS + PRODUCER_WARPS) * 32;
constexpr uint64_t kStrideAlign = 128;  // prepared row stride: multiple of 128 doubles (bitmap rows 16 B aligned)
constexpr uint64_t kSliceLen = 1u << 16;  // elements of D per work item

__host__ __device__ inline uint64_t prepared_stride(int k)
{
    const uint64_t d = 1ull << (2 * k);
    return (d + kStrideAlign - 1) / kStrideAlign * kStrideAlign;
}

enum : int { M_PROD = 0, M_SUM = 1, M_EUCLID = 2, M_COSINE = 3 };

static bool g_exact_div = false;
void set_exact_div(bool on) { g_exact_div = on; }

// ---------------------------------------------------------------------------
// per-profile pre-pass
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
profile_totals_kernel(const int64_t *__restrict__ counts, uint64_t d,
                      unsigned long long *__restrict__ totals)
{
    const int64_t *row = counts + uint64_t(blockIdx.y) * d;
    long long s = 0;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < d;
         i += uint64_t(gridDim.x) * blockDim.x)
        s += row[i];
    for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    __shared__ long long ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int w = 0; w < 8; ++w) t += ws[w];
        atomicAdd(totals + blockIdx.y, (unsigned long long)t);
    }
}

// F = x/S (x if !do_scale), P = x + 1, non-zero bitmap, sum F^2.
// x = c[i] (+ c[rc(i)] when balancing, kpal/klib.py:290-298).
__global__ void __launch_bounds__(256)
profile_convert_kernel(const int64_t *__restrict__ counts, uint64_t d, uint64_t stride, int k,
                       int do_balance, int do_scale,
                       const unsigned long long *__restrict__ totals_i64,
                       double *__restrict__ F, double *__restrict__ P,
                       uint32_t *__restrict__ bitmap, double *__restrict__ totals,
                       double *__restrict__ norm2)
{
    const uint64_t p = blockIdx.y;
    const int64_t *row = counts + p * d;
    long long tot = (long long)totals_i64[p];
    if (do_balance) tot *= 2;                       // sum(c[i] + c[rc(i)]) = 2 sum(c)
    const double S = double(tot);
    if (blockIdx.x == 0 && threadIdx.x == 0) totals[p] = S;
    const int shift = 32 - 2 * k;
    double n2 = 0.0;
    // stride is a multiple of 128, the loop covers whole warps: ballot is safe
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < stride;
         i += uint64_t(gridDim.x) * blockDim.x) {
        long long x = 0;
        if (i < d) {
            x = row[i];
            if (do_balance) x += __ldg(row + rc_index(uint32_t(i), shift));
        }
        const double xd = double(x);
        double f = do_scale ? xd / S : xd;
        if (i >= d) f = 0.0;                        // padding contributes nothing
        F[p * stride + i] = f;
        if (P) P[p * stride + i] = xd + 1.0;
        n2 += f * f;
        const uint32_t bits = __ballot_sync(0xffffffffu, x != 0);
        if ((threadIdx.x & 31) == 0) bitmap[p * (stride / 32) + i / 32] = bits;
    }
    for (int o = 16; o; o >>= 1) n2 += __shfl_down_sync(0xffffffffu, n2, o);
    __shared__ double ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = n2;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += ws[w];
        atomicAdd(norm2 + p, t);
    }
}

// ---------------------------------------------------------------------------
// PTX helpers: mbarrier + bulk async copy (TMA, SASS UBLKCP)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}

// |n| / den: MUFU.RCP64H seed + one Newton step folded into the accumulate
// (3 fp64 instructions), or the IEEE division.
template <bool EXACT>
__device__ __forceinline__ double add_term(double acc, double n, double den)
{
    if constexpr (EXACT) {
        return acc + fabs(n) / den;
    } else {
        double q0;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(q0) : "d"(den));
        const double h = fma(-den, q0, 2.0);
        return fma(fabs(n) * q0, h, acc);
    }
}

// Two terms with ONE reciprocal: |n1| / d1 + |n2| / d2 = (|n1| d2 + |n2| d1) / (d1 d2).  The same
// six fp64 instructions as two add_term calls (DMUL, DMUL, DFMA for the fraction; DFMA, DMUL, DFMA
// for the Newton step and the accumulate), but half the MUFU.RCP64H -- the unit that sits right
// behind the fp64 pipe (4.55 T/s measured against 18 T fp64 instructions/s: one reciprocal per 5
// fp64 instructions is 80 % of its rate).  Denominators are >= 1 / S_B > 0 and their product is
// far inside the double range.
#ifndef KPAL_DD_UNROLL
#define KPAL_DD_UNROLL 2                 // element pairs of the inner loop unrolled (tuning: make variant DEFS=-DKPAL_DD_UNROLL=..)
#endif
#define KPAL_PRAGMA_(x) _Pragma(#x)
#define KPAL_UNROLL(n) KPAL_PRAGMA_(unroll n)
#ifndef KPAL_PAIR_RCP
#define KPAL_PAIR_RCP 1
#endif
template <bool EXACT>
__device__ __forceinline__ double add_two_terms(double acc, double n1, double d1, double n2, double d2)
{
    if constexpr (EXACT || !KPAL_PAIR_RCP) {
        return add_term<EXACT>(add_term<EXACT>(acc, n1, d1), n2, d2);
    } else {
        const double d12 = d1 * d2;
        const double num = fma(fabs(n2), d1, fabs(n1) * d2);
        double q0;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(q0) : "d"(d12));
        const double h = fma(-d12, q0, 2.0);
        return fma(num * q0, h, acc);
    }
}

// ---------------------------------------------------------------------------
// tile enumeration over the (sorted) upper triangle: row blocks I of TA sorted
// positions, column blocks J of TB; tile (I, J) holds a pair p < q iff
// TA * I < TB * (J + 1), i.e. J >= (TA * I) / TB.
// ---------------------------------------------------------------------------
__host__ __device__ inline uint64_t first_col_tile(uint64_t I) { return (uint64_t(TA) * I) / TB; }
__host__ __device__ inline uint64_t tiles_in_row(uint64_t I, uint64_t NJ)
{
    const uint64_t j0 = first_col_tile(I);
    return NJ > j0 ? NJ - j0 : 0;
}
__host__ __device__ inline uint64_t num_tiles(uint64_t n)
{
    const uint64_t NI = (n + TA - 1) / TA, NJ = (n + TB - 1) / TB;
    uint64_t t = 0;
    for (uint64_t I = 0; I < NI; ++I) t += tiles_in_row(I, NJ);
    return t;
}
__device__ inline void tile_coords(uint64_t t, uint64_t n, uint32_t &I, uint32_t &J)
{
    const uint64_t NJ = (n + TB - 1) / TB;
    uint64_t i = 0;
    while (t >= tiles_in_row(i, NJ)) { t -= tiles_in_row(i, NJ); ++i; }
    I = uint32_t(i);
    J = uint32_t(first_col_tile(i) + t);
}

struct TileArgs {
    const double *F, *P;
    const uint32_t *bitmap;
    const double *totals;
    const int32_t *order;       // sorted position -> profile index (NULL: identity)
    uint64_t n, d, stride;
    uint64_t tile_begin, n_tiles_range;
    uint64_t slice_len, n_slices;
    int do_scale;
    double *acc;                // [n][n] sums, sorted-position space (p < q)
    uint32_t *cnt;              // [n][n] union counts
};

template <int METRIC, bool EXACT>
__global__ void __launch_bounds__(TILE_THREADS, 1)
distance_tile_kernel(const TileArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bars[2 * STAGES];
    __shared__ int32_t rowsA[TA], rowsB[TB];

    constexpr bool NEED_P = (METRIC == M_PROD);
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES + (NEED_P ? B_BYTES : 0);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t item = blockIdx.x;
    const uint64_t slice = item / a.n_tiles_range;
    const uint64_t tile = a.tile_begin + item % a.n_tiles_range;
    uint32_t I, J;
    tile_coords(tile, a.n, I, J);
    const uint64_t d0 = slice * a.slice_len;
    const uint64_t d1 = min(d0 + a.slice_len, a.stride);
    const uint32_t n_iter = uint32_t((d1 - d0) / DC);

    // sorted positions -> profile rows (clamped: out-of-range rows are masked at the end)
    for (uint32_t r = threadIdx.x; r < TA + TB; r += blockDim.x) {
        const bool isA = r < TA;
        uint64_t pos = isA ? uint64_t(I) * TA + r : uint64_t(J) * TB + (r - TA);
        if (pos >= a.n) pos = a.n - 1;
        const int32_t prof = a.order ? a.order[pos] : int32_t(pos);
        if (isA) rowsA[r] = prof; else rowsB[r - TA] = prof;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(&bars[s]), 1);                       // full: producer's expect_tx
            mbar_init(smem_u32(&bars[STAGES + s]), COMPUTE_WARPS);  // empty: one arrive per warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    // Loads of one stage: one bulk async copy (TMA, UBLKCP) per row segment, issued by
    // `n_issuers` threads starting at thread `first`.  Measured (profiles/): the TMA
    // unit accepts a small copy only every ~50 cycles, so the 256 copies of a stage
    // must not sit on a compute warp's critical path.  Default: a dedicated
    // producer warp that runs ahead of the compute warps (decoupled by the
    // full/empty mbarriers).  Alternative (KPAL_PRODUCER_WARP=0, kept for tuning):
    // every warp issues its share after waiting for the slot to drain, which frees
    // the producer's register allocation but couples the warps once per stage.
    constexpr uint32_t N_ROWS = TA + TB + (NEED_P ? TB : 0);
    auto issue_stage = [&](uint32_t it, uint32_t first, uint32_t n_issuers) {
        const uint32_t s = it % STAGES;
        const uint32_t full = smem_u32(&bars[s]);
        if (threadIdx.x == first) mbar_arrive_expect_tx(full, N_ROWS * DC * 8);
        if (PRODUCER_WARP) __syncwarp();
        const uint32_t base = smem_u32(smem) + s * STAGE_BYTES;
        const uint64_t col = d0 + uint64_t(it) * DC;
        for (uint32_t row = threadIdx.x - first; row < N_ROWS; row += n_issuers) {
            const double *src;
            if (row < TA) src = a.F + uint64_t(rowsA[row]) * a.stride;
            else if (row < TA + TB) src = a.F + uint64_t(rowsB[row - TA]) * a.stride;
            else src = a.P + uint64_t(rowsB[row - TA - TB]) * a.stride;
            bulk_g2s(base + row * ROW_BYTES, src + col, DC * 8, full);
        }
    };
    if constexpr (PRODUCER_WARP) {
        if (warp >= COMPUTE_WARPS) {
            // everything the producer warpgroup ever runs is inside this branch, so that
            // the register budget after setmaxnreg.dec cannot leak into the compute code
            if constexpr (SETMAXNREG) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
            if (warp == COMPUTE_WARPS) {
                for (uint32_t it = 0; it < n_iter; ++it) {
                    mbar_wait(smem_u32(&bars[STAGES + it % STAGES]), ((it / STAGES) & 1) ^ 1);
                    issue_stage(it, COMPUTE_WARPS * 32, 32);
                }
            }
            return;
        }
        if constexpr (SETMAXNREG) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(KPAL_SETMAXNREG));
    } else {
        for (uint32_t it = 0; it < STAGES - 1 && it < n_iter; ++it) issue_stage(it, 0, TILE_THREADS);
    }

    // ===== compute warps =====
    // warp (wi, wj), lane = ti_l * 8 + tj_l.  Thread (ti, tj) owns A rows ti + NTI*r
    // (r < RI) and B rows tj + NTJ*c (c < RJ): the four quarter-warps of an LDS.128
    // read 4 consecutive A rows (4 bank groups, broadcast inside a quarter) and
    // 8 consecutive B rows (8 bank groups) -- conflict free with the row padding.
    const uint32_t ti = (warp / WJ) * 4 + (lane >> 3);
    const uint32_t tj = (warp % WJ) * 8 + (lane & 7);

    double t[RJ];
#pragma unroll
    for (int c = 0; c < RJ; ++c)
        t[c] = a.do_scale ? 1.0 / a.totals[rowsB[tj + NTJ * c]] : 1.0;

    double acc[RI][RJ];
#pragma unroll
    for (int r = 0; r < RI; ++r)
#pragma unroll
        for (int c = 0; c < RJ; ++c) acc[r][c] = 0.0;

    for (uint32_t it = 0; it < n_iter; ++it) {
        const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
        if constexpr (!PRODUCER_WARP) {
            // refill the slot everybody finished with at iteration it-1
            const uint32_t nxt = it + STAGES - 1;
            if (nxt < n_iter) {
                if (it > 0) mbar_wait(smem_u32(&bars[STAGES + (it - 1) % STAGES]), ((it - 1) / STAGES) & 1);
                issue_stage(nxt, 0, TILE_THREADS);
            }
        }
        mbar_wait(smem_u32(&bars[s]), ph);
        const unsigned char *sA = smem + s * STAGE_BYTES + ti * ROW_BYTES;
        const unsigned char *sB = smem + s * STAGE_BYTES + A_BYTES + tj * ROW_BYTES;
        KPAL_UNROLL(KPAL_DD_UNROLL)
        for (int dd = 0; dd < DC / 2; ++dd) {
            double2 av[RI], bf[RJ], bp[RJ];
#pragma unroll
            for (int r = 0; r < RI; ++r)
                av[r] = *reinterpret_cast<const double2 *>(sA + r * NTI * ROW_BYTES + dd * 16);
#pragma unroll
            for (int c = 0; c < RJ; ++c) {
                bf[c] = *reinterpret_cast<const double2 *>(sB + c * NTJ * ROW_BYTES + dd * 16);
                if constexpr (NEED_P)
                    bp[c] = *reinterpret_cast<const double2 *>(sB + B_BYTES + c * NTJ * ROW_BYTES + dd * 16);
            }
#pragma unroll
            for (int c = 0; c < RJ; ++c) {
                double gx = 0.0, gy = 0.0;
                if constexpr (METRIC == M_PROD || METRIC == M_SUM) { gx = bf[c].x + t[c]; gy = bf[c].y + t[c]; }
#pragma unroll
                for (int r = 0; r < RI; ++r) {
                    if constexpr (METRIC == M_PROD) {
                        acc[r][c] = add_two_terms<EXACT>(acc[r][c], av[r].x - bf[c].x, fma(av[r].x, bp[c].x, gx),
                                                         av[r].y - bf[c].y, fma(av[r].y, bp[c].y, gy));
                    } else if constexpr (METRIC == M_SUM) {
                        acc[r][c] = add_two_terms<EXACT>(acc[r][c], av[r].x - bf[c].x, av[r].x + gx,
                                                         av[r].y - bf[c].y, av[r].y + gy);
                    } else if constexpr (METRIC == M_EUCLID) {
                        const double nx = av[r].x - bf[c].x, ny = av[r].y - bf[c].y;
                        acc[r][c] = fma(nx, nx, acc[r][c]);
                        acc[r][c] = fma(ny, ny, acc[r][c]);
                    } else {
                        acc[r][c] = fma(av[r].x, bf[c].x, acc[r][c]);
                        acc[r][c] = fma(av[r].y, bf[c].y, acc[r][c]);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars[STAGES + s]));
    }

    // ===== epilogue: partial sums, then union counts for the multiset metrics =====
    bool okA[RI], okB[RJ];
    uint64_t p[RI], q[RJ];
#pragma unroll
    for (int r = 0; r < RI; ++r) { p[r] = uint64_t(I) * TA + ti + NTI * r; okA[r] = p[r] < a.n; }
#pragma unroll
    for (int c = 0; c < RJ; ++c) { q[c] = uint64_t(J) * TB + tj + NTJ * c; okB[c] = q[c] < a.n; }
#pragma unroll
    for (int c = 0; c < RJ; ++c)
#pragma unroll
        for (int r = 0; r < RI; ++r)
            if (okB[c] && p[r] < q[c]) atomicAdd(a.acc + p[r] * a.n + q[c], acc[r][c]);

    if constexpr (METRIC == M_PROD || METRIC == M_SUM) {
        // |{i : x_A[i] != 0 or x_B[i] != 0}| over this slice from the per-profile bitmaps
        // (kpal/metrics.py:121-123); 12 x 128-bit loads feed 32 pairs x 128 elements.
        const uint64_t words_per_row = a.stride / 32;
        const uint64_t w0 = d0 / 32, w1 = d1 / 32;       // slices are multiples of 128 elements
        const uint32_t *za[RI], *zb[RJ];
#pragma unroll
        for (int r = 0; r < RI; ++r) za[r] = a.bitmap + uint64_t(rowsA[ti + NTI * r]) * words_per_row;
#pragma unroll
        for (int c = 0; c < RJ; ++c) zb[c] = a.bitmap + uint64_t(rowsB[tj + NTJ * c]) * words_per_row;
        uint32_t u[RI][RJ];
#pragma unroll
        for (int r = 0; r < RI; ++r)
#pragma unroll
            for (int c = 0; c < RJ; ++c) u[r][c] = 0;
        for (uint64_t w = w0; w < w1; w += 4) {
            uint4 x[RI], y[RJ];
#pragma unroll
            for (int r = 0; r < RI; ++r) x[r] = __ldg(reinterpret_cast<const uint4 *>(za[r] + w));
#pragma unroll
            for (int c = 0; c < RJ; ++c) y[c] = __ldg(reinterpret_cast<const uint4 *>(zb[c] + w));
#pragma unroll
            for (int r = 0; r < RI; ++r)
#pragma unroll
                for (int c = 0; c < RJ; ++c)
                    u[r][c] += __popc(x[r].x | y[c].x) + __popc(x[r].y | y[c].y) +
                               __popc(x[r].z | y[c].z) + __popc(x[r].w | y[c].w);
        }
#pragma unroll
        for (int c = 0; c < RJ; ++c)
#pragma unroll
            for (int r = 0; r < RI; ++r)
                if (okB[c] && p[r] < q[c]) atomicAdd(a.cnt + p[r] * a.n + q[c], u[r][c]);
    }
    (void)okA;
}

// ---------------------------------------------------------------------------
// Few profiles (ProfileDistance.distance on one pair, small matrices): the
// 128 x 64 tile would be almost empty, so use one CTA per (pair, slice) with
// the threads striding over the profile elements instead.  HBM bound.
// ---------------------------------------------------------------------------
template <int METRIC, bool EXACT>
__global__ void __launch_bounds__(256)
distance_small_kernel(const TileArgs a)
{
    // pair index -> sorted positions p < q
    uint64_t pair = blockIdx.x, q = 1;
    while (pair >= q) { pair -= q; ++q; }
    const uint64_t p = pair;
    const int32_t ia = a.order ? a.order[p] : int32_t(p), ib = a.order ? a.order[q] : int32_t(q);
    const double *FA = a.F + uint64_t(ia) * a.stride, *FB = a.F + uint64_t(ib) * a.stride;
    const double *PB = (METRIC == M_PROD) ? a.P + uint64_t(ib) * a.stride : nullptr;
    const double t = a.do_scale ? 1.0 / a.totals[ib] : 1.0;
    const uint64_t d0 = uint64_t(blockIdx.y) * a.slice_len;
    const uint64_t d1 = min(d0 + a.slice_len, a.stride);
    double s = 0.0;
    uint32_t u = 0;
    for (uint64_t i = d0 + threadIdx.x; i < d1; i += blockDim.x) {
        const double fa = FA[i], fb = FB[i];
        if constexpr (METRIC == M_PROD) {
            s = add_term<EXACT>(s, fa - fb, fma(fa, PB[i], fb + t));
            u += (fa != 0.0 || fb != 0.0);
        } else if constexpr (METRIC == M_SUM) {
            s = add_term<EXACT>(s, fa - fb, fa + (fb + t));
            u += (fa != 0.0 || fb != 0.0);
        } else if constexpr (METRIC == M_EUCLID) {
            s = fma(fa - fb, fa - fb, s);
        } else {
            s = fma(fa, fb, s);
        }
    }
    for (int o = 16; o; o >>= 1) {
        s += __shfl_down_sync(0xffffffffu, s, o);
        u += __shfl_down_sync(0xffffffffu, u, o);
    }
    __shared__ double ws[8];
    __shared__ uint32_t wu[8];
    if ((threadIdx.x & 31) == 0) { ws[threadIdx.x >> 5] = s; wu[threadIdx.x >> 5] = u; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { s += ws[w]; u += wu[w]; }
        atomicAdd(a.acc + p * a.n + q, s);
        if (METRIC == M_PROD || METRIC == M_SUM) atomicAdd(a.cnt + p * a.n + q, u);
    }
}

// acc/cnt (sorted-position space) -> symmetric output in profile-index space
__global__ void __launch_bounds__(256)
distance_finalize_kernel(const double *__restrict__ acc, const uint32_t *__restrict__ cnt,
                         const double *__restrict__ totals, const double *__restrict__ norm2,
                         const int32_t *__restrict__ order, uint64_t n, int metric, int do_scale,
                         uint64_t tile_begin, uint64_t tile_end, double *__restrict__ out)
{
    // one CTA per tile of the range, threads over its pairs
    const uint64_t tile = tile_begin + blockIdx.x;
    if (tile >= tile_end) return;
    uint32_t I, J;
    tile_coords(tile, n, I, J);
    for (uint32_t e = threadIdx.x; e < TA * TB; e += blockDim.x) {
        const uint64_t p = uint64_t(I) * TA + e / TB, q = uint64_t(J) * TB + e % TB;
        if (p >= q || q >= n) continue;
        const int32_t ip = order ? order[p] : int32_t(p), iq = order ? order[q] : int32_t(q);
        const double s = acc[p * n + q];
        double v;
        if (metric == M_PROD || metric == M_SUM) {
            v = s / double(cnt[p * n + q] + 1u);                  // kpal/metrics.py:123
        } else if (metric == M_EUCLID) {
            v = sqrt(s);                                          // kpal/metrics.py:46,135
            if (do_scale) v *= totals[iq];                        // S_B (unscaled side)
        } else {
            v = s / (sqrt(norm2[ip]) * sqrt(norm2[iq]));          // kpal/metrics.py:147
        }
        out[uint64_t(ip) * n + iq] = v;
        out[uint64_t(iq) * n + ip] = v;
    }
}

// Multi-GPU result exchange: a rank's tiles leave as a compact
// [tile][TA x TB] array of finished values (one NCCL gather of N^2 / 2 doubles in total
// instead of a zero-padded N x N reduce), and the root scatters them into the symmetric
// matrix in profile-index space.  Entries of a tile outside the triangle are left as they are.
__global__ void __launch_bounds__(256)
distance_pack_kernel(const double *__restrict__ acc, const uint32_t *__restrict__ cnt,
                     const double *__restrict__ totals, const double *__restrict__ norm2,
                     const int32_t *__restrict__ order, uint64_t n, int metric, int do_scale,
                     uint64_t tile_begin, uint64_t tile_end, double *__restrict__ packed)
{
    const uint64_t tile = tile_begin + blockIdx.x;
    if (tile >= tile_end) return;
    uint32_t I, J;
    tile_coords(tile, n, I, J);
    double *dst = packed + uint64_t(blockIdx.x) * (TA * TB);
    for (uint32_t e = threadIdx.x; e < TA * TB; e += blockDim.x) {
        const uint64_t p = uint64_t(I) * TA + e / TB, q = uint64_t(J) * TB + e % TB;
        double v = 0.0;
        if (p < q && q < n) {
            const int32_t ip = order ? order[p] : int32_t(p), iq = order ? order[q] : int32_t(q);
            const double s = acc[p * n + q];
            if (metric == M_PROD || metric == M_SUM) v = s / double(cnt[p * n + q] + 1u);
            else if (metric == M_EUCLID) { v = sqrt(s); if (do_scale) v *= totals[iq]; }
            else v = s / (sqrt(norm2[ip]) * sqrt(norm2[iq]));
        }
        dst[e] = v;
    }
}

__global__ void __launch_bounds__(256)
distance_unpack_kernel(const double *__restrict__ packed, const int32_t *__restrict__ order, uint64_t n,
                       uint64_t tile_begin, uint64_t tile_end, double *__restrict__ out)
{
    const uint64_t tile = tile_begin + blockIdx.x;
    if (tile >= tile_end) return;
    uint32_t I, J;
    tile_coords(tile, n, I, J);
    const double *src = packed + uint64_t(blockIdx.x) * (TA * TB);
    for (uint32_t e = threadIdx.x; e < TA * TB; e += blockDim.x) {
        const uint64_t p = uint64_t(I) * TA + e / TB, q = uint64_t(J) * TB + e % TB;
        if (p >= q || q >= n) continue;
        const int32_t ip = order ? order[p] : int32_t(p), iq = order ? order[q] : int32_t(q);
        const double v = src[e];
        out[uint64_t(ip) * n + iq] = v;
        out[uint64_t(iq) * n + ip] = v;
    }
}

// d(p, p): 0 for the distances (nan when scaling a zero-total profile, as the
// reference), cosine similarity 1 (nan for an all-zero profile).
__global__ void distance_diagonal_kernel(const double *__restrict__ totals,
                                         const double *__restrict__ norm2, uint64_t n, int metric,
                                         int do_scale, double *__restrict__ out)
{
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v = 0.0;
    if (metric == M_COSINE) v = norm2[i] / (sqrt(norm2[i]) * sqrt(norm2[i]));
    else if (do_scale && !(totals[i] > 0.0)) v = nan("");
    out[i * n + i] = v;
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
static int metric_id(int metric, int pairwise, int *out)
{
    if (metric == KPAL_METRIC_MULTISET && pairwise == KPAL_PAIRWISE_PROD) *out = M_PROD;
    else if (metric == KPAL_METRIC_MULTISET && pairwise == KPAL_PAIRWISE_SUM) *out = M_SUM;
    else if (metric == KPAL_METRIC_EUCLIDEAN) *out = M_EUCLID;
    else if (metric == KPAL_METRIC_COSINE) *out = M_COSINE;
    else return bad_arg("unknown metric / pairwise selector");
    return KPAL_OK;
}

int launch_prepare(const int64_t *d_counts, uint64_t n, int k, int do_balance, int do_scale,
                   double *d_F, double *d_P, uint32_t *d_bitmap, double *d_totals,
                   double *d_norm2, unsigned long long *d_totals_i64, cudaStream_t stream)
{
    if (k < 1 || k > KPAL_MAX_K) return bad_arg("k out of range");
    if (n == 0) return KPAL_OK;
    if (n > 65535) return bad_arg("at most 65535 profiles per prepare call");
    const uint64_t d = 1ull << (2 * k), stride = prepared_stride(k);
    KPAL_CUDA(cudaMemsetAsync(d_totals_i64, 0, n * sizeof(unsigned long long), stream));
    KPAL_CUDA(cudaMemsetAsync(d_norm2, 0, n * sizeof(double), stream));
    const unsigned bx = unsigned(std::min<uint64_t>((d + 255) / 256, 64));
    profile_totals_kernel<<<dim3(bx, unsigned(n)), 256, 0, stream>>>(d_counts, d, d_totals_i64);
    KPAL_LAUNCH_CHECK("profile_totals_kernel");
    const unsigned cx = unsigned(std::min<uint64_t>((stride + 255) / 256, 64));
    profile_convert_kernel<<<dim3(cx, unsigned(n)), 256, 0, stream>>>(
        d_counts, d, stride, k, do_balance, do_scale, d_totals_i64, d_F, d_P, d_bitmap, d_totals,
        d_norm2);
    KPAL_LAUNCH_CHECK("profile_convert_kernel");
    return KPAL_OK;
}

constexpr uint64_t kSmallN = 12;   // up to 66 pairs go through distance_small_kernel

template <int METRIC>
static int launch_tiles_metric(const TileArgs &a, bool exact, unsigned grid, cudaStream_t stream)
{
    if (a.n <= kSmallN) {
        // NOTE: in the small path the NaN-counts-as-set corner of np.logical_or
        // does not matter: the result is NaN anyway.
        const dim3 g(unsigned(a.n * (a.n - 1) / 2), unsigned(a.n_slices));
        if (g.x == 0) return KPAL_OK;
        if (exact) distance_small_kernel<METRIC, true><<<g, 256, 0, stream>>>(a);
        else distance_small_kernel<METRIC, false><<<g, 256, 0, stream>>>(a);
        KPAL_LAUNCH_CHECK("distance_small_kernel");
        return KPAL_OK;
    }
    constexpr int stage_bytes = A_BYTES + B_BYTES + (METRIC == M_PROD ? B_BYTES : 0);
    constexpr int smem = stage_bytes * STAGES;
    if (exact) {
        KPAL_CUDA(cudaFuncSetAttribute(distance_tile_kernel<METRIC, true>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        distance_tile_kernel<METRIC, true><<<grid, TILE_THREADS, smem, stream>>>(a);
    } else {
        KPAL_CUDA(cudaFuncSetAttribute(distance_tile_kernel<METRIC, false>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        distance_tile_kernel<METRIC, false><<<grid, TILE_THREADS, smem, stream>>>(a);
    }
    KPAL_LAUNCH_CHECK("distance_tile_kernel");
    return KPAL_OK;
}

int launch_distance_tiles(const double *d_F, const double *d_P, const uint32_t *d_bitmap,
                          const double *d_totals, const double *d_norm2, const int32_t *d_order,
                          uint64_t n, int k, int metric, int pairwise, int do_scale, int down,
                          uint64_t tile_begin, uint64_t tile_end, double *d_acc, uint32_t *d_cnt,
                          double *d_out, cudaStream_t stream, double *d_packed)
{
    (void)down;   // the order array already encodes ascending / descending totals
    int m;
    KPAL_CHECK(metric_id(metric, pairwise, &m));
    if (k < 1 || k > KPAL_MAX_K) return bad_arg("k out of range");
    if (n < 1) return bad_arg("no profiles");
    if (m == M_PROD && !d_P) return bad_arg("multiset/prod needs the P (= x + 1) array");
    const uint64_t total_tiles = num_tiles(n);
    if (tile_end > total_tiles) tile_end = total_tiles;
    if (tile_begin >= tile_end) return KPAL_OK;

    TileArgs a;
    a.F = d_F; a.P = d_P; a.bitmap = d_bitmap; a.totals = d_totals; a.order = d_order;
    a.n = n; a.d = 1ull << (2 * k); a.stride = prepared_stride(k);
    a.tile_begin = tile_begin; a.n_tiles_range = tile_end - tile_begin;
    a.slice_len = std::min<uint64_t>(kSliceLen, a.stride);
    a.n_slices = (a.stride + a.slice_len - 1) / a.slice_len;
    a.do_scale = do_scale;
    a.acc = d_acc; a.cnt = d_cnt;

    KPAL_CUDA(cudaMemsetAsync(d_acc, 0, n * n * sizeof(double), stream));
    KPAL_CUDA(cudaMemsetAsync(d_cnt, 0, n * n * sizeof(uint32_t), stream));
    const uint64_t items = a.n_tiles_range * a.n_slices;
    if (items > 0x7fffffffull) return bad_arg("too many work items");
    const bool exact = g_exact_div;
    switch (m) {
    case M_PROD: KPAL_CHECK(launch_tiles_metric<M_PROD>(a, exact, unsigned(items), stream)); break;
    case M_SUM: KPAL_CHECK(launch_tiles_metric<M_SUM>(a, exact, unsigned(items), stream)); break;
    case M_EUCLID: KPAL_CHECK(launch_tiles_metric<M_EUCLID>(a, exact, unsigned(items), stream)); break;
    default: KPAL_CHECK(launch_tiles_metric<M_COSINE>(a, exact, unsigned(items), stream)); break;
    }
    if (d_packed) {     // multi-GPU: finished values tile by tile, for the gather onto the root
        distance_pack_kernel<<<unsigned(a.n_tiles_range), 256, 0, stream>>>(
            d_acc, d_cnt, d_totals, d_norm2, d_order, n, m, do_scale, tile_begin, tile_end, d_packed);
        KPAL_LAUNCH_CHECK("distance_pack_kernel");
        return KPAL_OK;
    }
    distance_finalize_kernel<<<unsigned(a.n_tiles_range), 256, 0, stream>>>(
        d_acc, d_cnt, d_totals, d_norm2, d_order, n, m, do_scale, tile_begin, tile_end, d_out);
    KPAL_LAUNCH_CHECK("distance_finalize_kernel");
    if (tile_begin == 0) {   // the rank that owns tile 0 also writes the diagonal
        distance_diagonal_kernel<<<unsigned((n + 255) / 256), 256, 0, stream>>>(
            d_totals, d_norm2, n, m, do_scale, d_out);
        KPAL_LAUNCH_CHECK("distance_diagonal_kernel");
    }
    return KPAL_OK;
}

// packed tiles [tile_begin, tile_end) -> symmetric out (+ the diagonal when asked to)
int launch_distance_unpack(const double *d_packed, const double *d_totals, const double *d_norm2,
                           const int32_t *d_order, uint64_t n, int metric, int pairwise, int do_scale,
                           uint64_t tile_begin, uint64_t tile_end, int diagonal, double *d_out,
                           cudaStream_t stream)
{
    int m;
    KPAL_CHECK(metric_id(metric, pairwise, &m));
    const uint64_t total_tiles = num_tiles(n);
    if (tile_end > total_tiles) tile_end = total_tiles;
    if (tile_begin < tile_end) {
        distance_unpack_kernel<<<unsigned(tile_end - tile_begin), 256, 0, stream>>>(
            d_packed, d_order, n, tile_begin, tile_end, d_out);
        KPAL_LAUNCH_CHECK("distance_unpack_kernel");
    }
    if (diagonal) {
        distance_diagonal_kernel<<<unsigned((n + 255) / 256), 256, 0, stream>>>(
            d_totals, d_norm2, n, m, do_scale, d_out);
        KPAL_LAUNCH_CHECK("distance_diagonal_kernel");
    }
    return KPAL_OK;
}

uint64_t distance_num_tiles(uint64_t n) { return num_tiles(n); }
uint64_t distance_tile_elems() { return uint64_t(TA) * TB; }
uint64_t prepared_stride_host(int k) { return prepared_stride(k); }

}  // namespace kpal
