// Euclidean distance and cosine similarity through an EXACT integer Gram matrix on the
// 5th-generation tensor cores (tcgen05, accumulators in tensor memory).
//
// Replaces, for these two vector functions, the per-pair passes of metrics.euclidean /
// metrics.cosine_similarity (reference kpal/metrics.py:126-147, with the scale step of
// kpal/kdistlib.py:150-157 and kpal/metrics.py:49-86).  Both are functions of three sums per pair:
//
//   G_pq = sum_k x_p[k] x_q[k]        n_p = G_pp = sum_k x_p[k]^2        S_p = sum_k x_p[k]
//
//   cosine(p, q)    = G_pq / (sqrt(n_p) sqrt(n_q))                       (scaling cancels)
//   euclidean(p, q) = sqrt(n_p + n_q - 2 G_pq)                           (unscaled)
//   scaled, A = the profile with the smaller total, s = S_B / S_A (kpal/metrics.py:64-70):
//       sum_k (s x_A - x_B)^2 = (S_B^2 n_A - 2 S_A S_B G + S_A^2 n_B) / S_A^2
//       (with `down` both sides are divided by s: / S_B^2 instead)
//
// The counts are integers, so G is computed EXACTLY: the counts go to the tensor cores as
// unsigned 8-bit operands (tcgen05.mma kind::i8, 32-bit integer accumulators in TMEM), the
// numerators above are evaluated in 128-bit integer arithmetic, and the only roundings are
// the final conversion to double, one division and one square root.  This is more accurate
// than the element-wise fp64 form of distance.cu (and than the reference's own float
// pipeline); the tolerance the tests state for it is the same 1e-9 relative.
//
// Exactness conditions, checked on the device by the prepare kernel:
//   * every count <= 255 (one 8-bit limb per count; k-mer profiles at the BASELINE sizes
//     have counts in the tens) -- otherwise the caller takes the fp64 tile kernel;
//   * a 32-bit accumulator never overflows: by Cauchy-Schwarz G_pq <= max n_p, and all
//     partial sums are non-negative, so max_p n_p < 2^31 makes one accumulation over the
//     whole profile exact; otherwise the profile is cut into chunks of 32768 elements
//     (255^2 x 32768 < 2^31) whose results are added in 64-bit integers.
//
// Kernel: one CTA per 128 x 256 tile of the upper triangle (and K range), warp-specialised:
//   warp 0    TMA producer: 2-D tensor-map loads (cp.async.bulk.tensor, 128-byte swizzle) of a
//             128-row and a 256-row box of 128 profile elements into a 4-stage ring;
//   warp 1    allocates 256 TMEM columns and issues tcgen05.mma (M = 128, N = 256, K = 32 per
//             instruction, 4 per stage); tcgen05.commit releases the stage / signals the epilogue;
//   warps 2-5 epilogue: tcgen05.ld of the 128 x 256 int32 tile, 64-bit integer atomic adds
//             into G (the K ranges of a tile are different CTAs).
#include "common.cuh"

#include <cuda.h>
#include <algorithm>
#include <mutex>

namespace kpal {

constexpr int GM = 128;                 // tile rows (MMA M)
constexpr int GN = 256;                 // tile columns (MMA N)
constexpr int GK = 128;                 // profile elements (= bytes) per stage: one 128-byte swizzle row
constexpr int GSTAGES = 4;
constexpr int GA_BYTES = GM * GK;       // 16 KB
constexpr int GB_BYTES = GN * GK;       // 32 KB
constexpr int GSTAGE_BYTES = GA_BYTES + GB_BYTES;
constexpr int GRAM_THREADS = 192;       // producer warp, MMA warp, four epilogue warps
constexpr uint32_t GRAM_SMEM = GSTAGES * GSTAGE_BYTES + 1024 /* alignment */ + 256 /* barriers */;
constexpr uint64_t kGramSafeChunk = 32768;      // elements whose u8 x u8 products always fit 31 bits

// ---------------------------------------------------------------------------
// per-profile pre-pass: counts -> u8 rows, exact totals and squared norms, range check
// ---------------------------------------------------------------------------
// flags[0] |= 1 when a count exceeds 255; norms_max = max_p n_p via atomicMax.
__global__ void __launch_bounds__(256)
gram_prepare_kernel(const int64_t *__restrict__ counts, uint64_t d, uint64_t dp, int k, int do_balance,
                    uint8_t *__restrict__ x8, unsigned long long *__restrict__ totals,
                    unsigned long long *__restrict__ norms, unsigned int *__restrict__ flags)
{
    const uint64_t p = blockIdx.y;
    const int64_t *row = counts + p * d;
    uint8_t *out = x8 + p * dp;
    const int shift = 32 - 2 * k;
    unsigned long long s = 0, n2 = 0;
    bool big = false;
    // four neighbouring elements per thread: one 32-bit store
    for (uint64_t i4 = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i4 < dp;
         i4 += uint64_t(gridDim.x) * blockDim.x * 4) {
        uint32_t packed = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint64_t i = i4 + j;
            long long x = 0;
            if (i < d) {
                x = row[i];
                if (do_balance) x += __ldg(row + rc_index(uint32_t(i), shift));
            }
            if (x < 0 || x > 255) { big = true; x = 0; }
            s += (unsigned long long)x;
            n2 += (unsigned long long)(x * x);
            packed |= uint32_t(x) << (8 * j);
        }
        *reinterpret_cast<uint32_t *>(out + i4) = packed;
    }
    for (int o = 16; o; o >>= 1) {
        s += __shfl_down_sync(0xffffffffu, s, o);
        n2 += __shfl_down_sync(0xffffffffu, n2, o);
    }
    if (__any_sync(0xffffffffu, big) && (threadIdx.x & 31) == 0) atomicOr(flags, 1u);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(totals + p, s);
        atomicAdd(norms + p, n2);
    }
}

__global__ void gram_norm_max_kernel(const unsigned long long *__restrict__ norms, uint64_t n,
                                     unsigned long long *__restrict__ norm_max)
{
    unsigned long long m = 0;
    for (uint64_t i = threadIdx.x; i < n; i += blockDim.x) m = max(m, norms[i]);
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_down_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(norm_max, m);
}

// ---------------------------------------------------------------------------
// PTX helpers (mbarrier, TMA tensor load, tcgen05)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t g_smem_u32(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void g_mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void g_mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void g_mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
// 2-D tensor-map load: box at (c0 = element along the profile, c1 = profile row) -> smem, completes on `bar`
__device__ __forceinline__ void g_tma_load_2d(uint32_t dst, const CUtensorMap *map, int32_t c0, int32_t c1, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// K-major operand tile with 128-byte swizzle (what the tensor map writes): rows of 128 bytes,
// 8-row swizzle atoms 1024 bytes apart.  Descriptor fields in 16-byte units.
__device__ __forceinline__ uint64_t g_umma_desc(uint32_t smem_addr)
{
    uint64_t desc = 0;
    desc |= uint64_t((smem_addr & 0x3FFFFu) >> 4);          // start address
    desc |= uint64_t(1) << 16;                              // leading byte offset (unused with swizzle): 1
    desc |= uint64_t(1024 >> 4) << 32;                      // stride byte offset: 8 rows x 128 B
    desc |= uint64_t(1) << 46;                              // descriptor version (Blackwell)
    desc |= uint64_t(2) << 61;                              // SWIZZLE_128B
    return desc;
}
// instruction descriptor, kind::i8: D = s32, A = B = unsigned 8 bit, both K-major, N = 256, M = 128
constexpr uint32_t kGramIdesc = (2u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) |
                                (uint32_t(GN >> 3) << 17) | (uint32_t(GM >> 4) << 24);

__device__ __forceinline__ void g_umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(kGramIdesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void g_umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// tile enumeration over the upper triangle: row blocks I of GM profiles, column blocks J of GN;
// tile (I, J) holds a pair p < q iff J >= (GM * I) / GN
__host__ __device__ inline uint64_t gram_first_col(uint64_t I) { return (uint64_t(GM) * I) / GN; }
__host__ __device__ inline uint64_t gram_tiles(uint64_t n)
{
    const uint64_t NI = (n + GM - 1) / GM, NJ = (n + GN - 1) / GN;
    uint64_t t = 0;
    for (uint64_t I = 0; I < NI; ++I) t += NJ - gram_first_col(I);
    return t;
}

struct GramArgs {
    uint64_t n;                 // profiles
    uint32_t n_kb;              // 128-element blocks per profile row (dp / 128)
    uint32_t kb_per_item;       // blocks accumulated by one CTA
    uint32_t n_splits;          // CTAs per tile
    uint64_t n_tiles;
    long long *gram;            // [n][n] int64, zeroed; entries of the tiles' rectangles are added
};

__global__ void __launch_bounds__(GRAM_THREADS, 1)
gram_u8_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const GramArgs a)
{
    extern __shared__ unsigned char gram_smem_raw[];
    // 1024-byte alignment for the 128-byte swizzle atoms
    const uint32_t base = (g_smem_u32(gram_smem_raw) + 1023u) & ~1023u;
    unsigned char *aligned = gram_smem_raw + (base - g_smem_u32(gram_smem_raw));
    uint64_t *bars = reinterpret_cast<uint64_t *>(aligned + GSTAGES * GSTAGE_BYTES);    // full[4], empty[4], tmem_full
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * GSTAGES + 1);
    const uint32_t bars_s = base + GSTAGES * GSTAGE_BYTES;
    auto full_bar = [&](uint32_t s) { return bars_s + 8u * s; };
    auto empty_bar = [&](uint32_t s) { return bars_s + 8u * (GSTAGES + s); };
    const uint32_t tmem_full_bar = bars_s + 8u * (2 * GSTAGES);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // work item -> tile (I, J) and K range
    const uint64_t item = blockIdx.x;
    uint64_t t = item % a.n_tiles;
    const uint32_t split = uint32_t(item / a.n_tiles);
    const uint64_t NJ = (a.n + GN - 1) / GN;
    uint64_t I = 0;
    while (t >= NJ - gram_first_col(I)) { t -= NJ - gram_first_col(I); ++I; }
    const uint64_t J = gram_first_col(I) + t;
    const uint32_t kb0 = split * a.kb_per_item;
    const uint32_t kb1 = min(kb0 + a.kb_per_item, a.n_kb);
    const uint32_t n_iter = kb1 > kb0 ? kb1 - kb0 : 0;

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < GSTAGES; ++s) { g_mbar_init(full_bar(s), 1); g_mbar_init(empty_bar(s), 1); }
        g_mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {            // one warp allocates the accumulator columns (and frees them at the end)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(g_smem_u32(tmem_slot)), "n"(GN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *reinterpret_cast<volatile uint32_t *>(tmem_slot);

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (uint32_t it = 0; it < n_iter; ++it) {
                const uint32_t s = it % GSTAGES;
                g_mbar_wait(empty_bar(s), ((it / GSTAGES) & 1u) ^ 1u);
                g_mbar_expect_tx(full_bar(s), GSTAGE_BYTES);
                const int32_t c0 = int32_t((kb0 + it) * GK);
                g_tma_load_2d(base + s * GSTAGE_BYTES, &map_a, c0, int32_t(I * GM), full_bar(s));
                g_tma_load_2d(base + s * GSTAGE_BYTES + GA_BYTES, &map_b, c0, int32_t(J * GN), full_bar(s));
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            for (uint32_t it = 0; it < n_iter; ++it) {
                const uint32_t s = it % GSTAGES;
                g_mbar_wait(full_bar(s), (it / GSTAGES) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t da = g_umma_desc(base + s * GSTAGE_BYTES);
                const uint64_t db = g_umma_desc(base + s * GSTAGE_BYTES + GA_BYTES);
#pragma unroll
                for (uint32_t kk = 0; kk < GK / 32; ++kk)           // 32 bytes of K per instruction: + 2 x 16 B
                    g_umma_i8(tmem_d, da + 2u * kk, db + 2u * kk, (it | kk) ? 1u : 0u);
                g_umma_commit(empty_bar(s));                        // the stage is free once these MMAs have read it
            }
            g_umma_commit(tmem_full_bar);                           // accumulator complete
        }
    } else if (n_iter) {
        // ===== epilogue: TMEM -> registers -> 64-bit adds into G =====
        g_mbar_wait(tmem_full_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t quarter = warp & 3u;                         // a warp reads TMEM lanes 32 (warp % 4) ..
        const uint64_t row = I * GM + quarter * 32 + lane;
        const uint64_t col0 = J * GN;
#pragma unroll 1
        for (uint32_t c = 0; c < GN; c += 32) {
            uint32_t v[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(tmem_d + ((quarter * 32u) << 16) + c));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row < a.n) {
                unsigned long long *dst = reinterpret_cast<unsigned long long *>(a.gram) + row * a.n + col0 + c;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (col0 + c + j < a.n && v[j]) atomicAdd(dst + j, (unsigned long long)v[j]);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(GN) : "memory");
    }
}

// ---------------------------------------------------------------------------
// G, n, S -> distances (kpal/metrics.py:126-147 after kpal/kdistlib.py:150-157), symmetric output
// ---------------------------------------------------------------------------
__device__ __forceinline__ double u128_to_double(unsigned __int128 v)
{
    return double((unsigned long long)(v >> 64)) * 18446744073709551616.0 + double((unsigned long long)v);
}

__global__ void __launch_bounds__(256)
gram_finalize_kernel(const long long *__restrict__ gram, const unsigned long long *__restrict__ totals,
                     const unsigned long long *__restrict__ norms, uint64_t n, int cosine, int do_scale, int down,
                     double *__restrict__ out)
{
    const uint64_t q = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint64_t p = blockIdx.y;
    if (q >= n || p > q) return;
    double v;
    if (p == q) {
        // d(p, p): 0 (nan when scaling a zero-total profile, as the reference); cosine 1 (nan for an all-zero profile)
        if (cosine) v = double(norms[p]) / (sqrt(double(norms[p])) * sqrt(double(norms[p])));
        else v = (do_scale && totals[p] == 0) ? nan("") : 0.0;
        out[p * n + p] = v;
        return;
    }
    const unsigned long long g = (unsigned long long)gram[p * n + q];
    const unsigned long long np = norms[p], nq = norms[q], sp = totals[p], sq = totals[q];
    if (cosine) {
        v = double(g) / (sqrt(double(np)) * sqrt(double(nq)));                   // kpal/metrics.py:147
    } else if (!do_scale) {
        v = sqrt(double(np + nq - 2 * g));                                       // exact integer under the root
    } else if (sp == 0 || sq == 0) {
        v = nan("");                                                             // the reference divides by a zero total
    } else {
        // A = the smaller total (ties: no scaling at all, kpal/metrics.py:67-70)
        const bool p_is_a = sp < sq;
        const unsigned long long sa = p_is_a ? sp : sq, sb = p_is_a ? sq : sp;
        const unsigned long long na = p_is_a ? np : nq, nb = p_is_a ? nq : np;
        // S_B^2 n_A + S_A^2 n_B - 2 S_A S_B G  >= 0, exact in 128 bits
        const unsigned __int128 num = (unsigned __int128)sb * sb * na + (unsigned __int128)sa * sa * nb -
                                      (unsigned __int128)2 * sa * sb * g;
        v = sqrt(u128_to_double(num)) / double(down ? sb : sa);
    }
    out[p * n + q] = v;
    out[q * n + p] = v;
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled()
{
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        cudaGetLastError();
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

uint64_t gram_row_stride(int k)                 // bytes per u8 profile row: 4^k rounded up to whole stages
{
    const uint64_t d = 1ull << (2 * k);
    return (d + GK - 1) / GK * GK;
}

int launch_gram_prepare(const int64_t *d_counts, uint64_t n, int k, int do_balance, uint8_t *d_x8,
                        unsigned long long *d_totals, unsigned long long *d_norms, unsigned int *d_flags,
                        cudaStream_t stream)
{
    if (k < 1 || k > KPAL_MAX_K) return bad_arg("k out of range");
    if (n == 0) return KPAL_OK;
    if (n > 65535) return bad_arg("at most 65535 profiles per prepare call");
    const uint64_t d = 1ull << (2 * k), dp = gram_row_stride(k);
    KPAL_CUDA(cudaMemsetAsync(d_totals, 0, n * 8, stream));
    KPAL_CUDA(cudaMemsetAsync(d_norms, 0, n * 8, stream));
    const unsigned bx = unsigned(std::min<uint64_t>((dp / 4 + 255) / 256, 64));
    gram_prepare_kernel<<<dim3(bx, unsigned(n)), 256, 0, stream>>>(d_counts, d, dp, k, do_balance, d_x8, d_totals,
                                                                   d_norms, d_flags);
    KPAL_LAUNCH_CHECK("gram_prepare_kernel");
    return KPAL_OK;
}

// x8: [n][gram_row_stride(k)] u8 (rows zero-padded), totals / norms: exact sums, gram: [n][n] int64 scratch.
// norm_max: the largest norm of the set (host value; decides the accumulation chunk).
int launch_gram_distances(const uint8_t *d_x8, const unsigned long long *d_totals, const unsigned long long *d_norms,
                          unsigned long long norm_max, uint64_t n, int k, int metric, int do_scale, int down,
                          long long *d_gram, double *d_out, cudaStream_t stream)
{
    if (metric != KPAL_METRIC_EUCLIDEAN && metric != KPAL_METRIC_COSINE)
        return bad_arg("the Gram form serves the euclidean and cosine vector functions");
    if (k < 1 || k > KPAL_MAX_K) return bad_arg("k out of range");
    if (n < 1) return bad_arg("no profiles");
    EncodeTiledFn encode = encode_tiled();
    if (!encode) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return KPAL_ECUDA; }
    const uint64_t dp = gram_row_stride(k);

    CUtensorMap map_a, map_b;
    const cuuint64_t dims[2] = {dp, n};
    const cuuint64_t strides[1] = {dp};                       // bytes between rows
    const cuuint32_t elem_strides[2] = {1, 1};
    const cuuint32_t box_a[2] = {GK, GM}, box_b[2] = {GK, GN};
    void *gptr = const_cast<uint8_t *>(d_x8);
    if (encode(&map_a, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, gptr, dims, strides, box_a, elem_strides,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
        encode(&map_b, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, gptr, dims, strides, box_b, elem_strides,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed for the profile rows");
        return KPAL_ECUDA;
    }

    GramArgs a;
    a.n = n;
    a.n_kb = uint32_t(dp / GK);
    a.n_tiles = gram_tiles(n);
    // K ranges: one accumulation is exact while max norm < 2^31; else chunks of 32768 elements.
    // More ranges than that only to fill the GPU (items >= ~2 waves of one CTA per SM).
    uint32_t max_kb = norm_max < (1ull << 31) ? a.n_kb : uint32_t(kGramSafeChunk / GK);
    uint32_t splits = (a.n_kb + max_kb - 1) / max_kb;
    const uint64_t want_items = uint64_t(sm_count()) * 2;
    while (a.n_tiles * splits < want_items && (a.n_kb + splits - 1) / splits > 64) splits *= 2;
    a.kb_per_item = (a.n_kb + splits - 1) / splits;
    a.n_splits = (a.n_kb + a.kb_per_item - 1) / a.kb_per_item;
    a.gram = d_gram;
    const uint64_t items = a.n_tiles * a.n_splits;
    if (items > 0x7fffffffull) return bad_arg("too many work items");

    KPAL_CUDA(cudaMemsetAsync(d_gram, 0, n * n * sizeof(long long), stream));
    KPAL_CUDA(cudaFuncSetAttribute(gram_u8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(GRAM_SMEM)));
    gram_u8_kernel<<<unsigned(items), GRAM_THREADS, GRAM_SMEM, stream>>>(map_a, map_b, a);
    KPAL_LAUNCH_CHECK("gram_u8_kernel");
    const dim3 grid(unsigned((n + 255) / 256), unsigned(n));
    gram_finalize_kernel<<<grid, 256, 0, stream>>>(d_gram, d_totals, d_norms, n, metric == KPAL_METRIC_COSINE, do_scale,
                                                   down, d_out);
    KPAL_LAUNCH_CHECK("gram_finalize_kernel");
    return KPAL_OK;
}

int launch_gram_norm_max(const unsigned long long *d_norms, uint64_t n, unsigned long long *d_norm_max,
                         cudaStream_t stream)
{
    KPAL_CUDA(cudaMemsetAsync(d_norm_max, 0, 8, stream));
    gram_norm_max_kernel<<<1, 256, 0, stream>>>(d_norms, n, d_norm_max);
    KPAL_LAUNCH_CHECK("gram_norm_max_kernel");
    return KPAL_OK;
}

}  // namespace kpal
