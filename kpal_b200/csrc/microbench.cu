// Device micro-benchmarks that set the realistic ceilings for the two hot
// kernels (run on the GPU box: `kpal_b200/csrc/microbench`):
//   1. global RED.ADD throughput into tables of 1 MB .. 1 GB (u32 / u64)
//   2. shared-memory atomic throughput (spread addresses)
//   3. fp64 pipe: DFMA rate, MUFU.RCP64H rate + accuracy, IEEE division rate
// Output: one JSON object per line.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
    fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

template <typename T>
__global__ void red_kernel(T *table, uint32_t mask, int per_thread)
{
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u;
#pragma unroll 8
    for (int i = 0; i < per_thread; ++i) {
        s = hash32(s + i);
        atomicAdd(table + (s & mask), T(1));
    }
}

__global__ void smem_atomic_kernel(uint32_t *out, int per_thread, uint32_t mask)
{
    extern __shared__ uint32_t h[];
    for (uint32_t i = threadIdx.x; i <= mask; i += blockDim.x) h[i] = 0;
    __syncthreads();
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u;
#pragma unroll 8
    for (int i = 0; i < per_thread; ++i) {
        s = hash32(s + i);
        atomicAdd(&h[s & mask], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = h[0];
}

// returning shared atomics (what a block-level multisplit needs for ranking)
__global__ void smem_atomic_ret_kernel(uint32_t *out, int per_thread, uint32_t mask)
{
    extern __shared__ uint32_t h[];
    for (uint32_t i = threadIdx.x; i <= mask; i += blockDim.x) h[i] = 0;
    __syncthreads();
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u, acc = 0;
#pragma unroll 8
    for (int i = 0; i < per_thread; ++i) {
        s = hash32(s + i);
        acc += atomicAdd(&h[s & mask], 1u);
    }
    __syncthreads();
    if (acc == 0xdeadbeef || threadIdx.x == 0) out[blockIdx.x] = h[0] + acc;
}

__global__ void dfma_kernel(double *out, int iters)
{
    double a[8];
    for (int j = 0; j < 8; ++j) a[j] = threadIdx.x * 1e-3 + j;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = fma(a[j], m, c);
    double s = 0;
    for (int j = 0; j < 8; ++j) s += a[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mode 0: MUFU.RCP64H + 1 Newton step ; 1: IEEE 1.0/x ; 2: seed only
template <int MODE>
__global__ void rcp_kernel(double *out, int iters)
{
    double a[8];
    for (int j = 0; j < 8; ++j) a[j] = 1.0 + threadIdx.x * 1e-3 + j;
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            double u = a[j], q;
            if (MODE == 1) q = 1.0 / u;
            else {
                asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(u));
                if (MODE == 0) { double e = fma(-u, q, 1.0); q = fma(q, e, q); }
            }
            a[j] = q + 1.5;
        }
    double s = 0;
    for (int j = 0; j < 8; ++j) s += a[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the multiset/prod inner op as used by distance_tile_kernel (6 fp64 + 1 MUFU)
__global__ void prod_term_kernel(double *out, int iters)
{
    double acc[8], fa[8];
    for (int j = 0; j < 8; ++j) { acc[j] = 0; fa[j] = (threadIdx.x + j) * 1e-4; }
    const double fb = 3e-4, rb = 0.25, t = 1e-6;
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const double n = fa[j] - fb, u = fa[j] + t;
            double q;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(u));
            const double e = fma(-u, q, 1.0);
            q = fma(q, e, q);
            acc[j] = fma(fabs(n) * q, rb, acc[j]);
            fa[j] += 1e-9;
        }
    double s = 0;
    for (int j = 0; j < 8; ++j) s += acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void rcp_accuracy_kernel(double *max_err1, double *max_err0, int n)
{
    double m1 = 0, m0 = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t h1 = hash32(i), h2 = hash32(i ^ 0x9e3779b9u);
        // random mantissa, exponent in [-60, 60]
        double u = ldexp(1.0 + (double(h1) * 4294967296.0 + h2) / 18446744073709551616.0,
                         int(h1 % 121) - 60);
        double q;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(u));
        const double exact = 1.0 / u;
        m0 = fmax(m0, fabs(q - exact) / exact);
        const double e = fma(-u, q, 1.0);
        q = fma(q, e, q);
        m1 = fmax(m1, fabs(q - exact) / exact);
    }
    // crude max-reduce through atomics on the bit pattern (values are >= 0)
    atomicMax(reinterpret_cast<unsigned long long *>(max_err1), __double_as_longlong(m1));
    atomicMax(reinterpret_cast<unsigned long long *>(max_err0), __double_as_longlong(m0));
}

template <typename F>
static float time_ms(F &&f, int reps = 5)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f();                                    // warm-up
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(a));
        f();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, sms, prop.clockRate);

    // 1. global RED
    void *table; CK(cudaMalloc(&table, size_t(1) << 31));
    for (int bits = 32; bits <= 64; bits += 32)
        for (int lg = 18; lg <= 28; lg += 2) {       // elements
            const size_t bytes = (size_t(1) << lg) * (bits / 8);
            if (bytes > (size_t(1) << 31)) continue;
            CK(cudaMemset(table, 0, bytes));
            const int per_thread = 64, threads = 256, blocks = sms * 64;
            const double n = double(per_thread) * threads * blocks;
            float ms = time_ms([&] {
                if (bits == 32) red_kernel<uint32_t><<<blocks, threads>>>((uint32_t *)table, (1u << lg) - 1, per_thread);
                else red_kernel<unsigned long long><<<blocks, threads>>>((unsigned long long *)table, (1u << lg) - 1, per_thread);
            });
            printf("{\"bench\": \"global_red\", \"counter_bits\": %d, \"table_mb\": %.1f, \"gatomics_per_s\": %.2f}\n",
                   bits, bytes / 1048576.0, n / ms * 1e-6);
        }
    CK(cudaFree(table));

    // 2. shared atomics
    uint32_t *o32; CK(cudaMalloc(&o32, sms * 64 * 4));
    for (int lg = 8; lg <= 14; lg += 2) {
        const int per_thread = 256, threads = 256, blocks = sms * 8;
        const size_t smem = (size_t(4) << lg);
        CK(cudaFuncSetAttribute(smem_atomic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        float ms = time_ms([&] { smem_atomic_kernel<<<blocks, threads, smem>>>(o32, per_thread, (1u << lg) - 1); });
        printf("{\"bench\": \"smem_atomic\", \"bins\": %d, \"gatomics_per_s\": %.2f}\n", 1 << lg,
               double(per_thread) * threads * blocks / ms * 1e-6);
    }

    for (int lg = 9; lg <= 13; lg += 4) {
        const int per_thread = 256, threads = 256, blocks = sms * 8;
        const size_t smem = (size_t(4) << lg);
        float ms = time_ms([&] { smem_atomic_ret_kernel<<<blocks, threads, smem>>>(o32, per_thread, (1u << lg) - 1); });
        printf("{\"bench\": \"smem_atomic_returning\", \"bins\": %d, \"gatomics_per_s\": %.2f}\n", 1 << lg,
               double(per_thread) * threads * blocks / ms * 1e-6);
    }

    // 3. fp64
    double *od; CK(cudaMalloc(&od, size_t(sms) * 16 * 256 * 8));
    {
        const int iters = 4096, threads = 256, blocks = sms * 8;
        const double ops = double(iters) * 8 * threads * blocks;
        float ms = time_ms([&] { dfma_kernel<<<blocks, threads>>>(od, iters); });
        printf("{\"bench\": \"dfma\", \"tflops\": %.2f, \"gdfma_per_s\": %.1f}\n", 2 * ops / ms * 1e-9, ops / ms * 1e-6);
        ms = time_ms([&] { rcp_kernel<0><<<blocks, threads>>>(od, iters); });
        printf("{\"bench\": \"rcp64h_newton1\", \"grcp_per_s\": %.1f}\n", ops / ms * 1e-6);
        ms = time_ms([&] { rcp_kernel<2><<<blocks, threads>>>(od, iters); });
        printf("{\"bench\": \"rcp64h_seed_only\", \"grcp_per_s\": %.1f}\n", ops / ms * 1e-6);
        ms = time_ms([&] { rcp_kernel<1><<<blocks, threads>>>(od, iters); });
        printf("{\"bench\": \"ieee_div\", \"gdiv_per_s\": %.1f}\n", ops / ms * 1e-6);
        ms = time_ms([&] { prod_term_kernel<<<blocks, threads>>>(od, iters); });
        printf("{\"bench\": \"multiset_prod_term\", \"gterms_per_s\": %.1f}\n", ops / ms * 1e-6);
    }
    {
        double *errs; CK(cudaMalloc(&errs, 16)); CK(cudaMemset(errs, 0, 16));
        rcp_accuracy_kernel<<<sms * 8, 256>>>(errs, errs + 1, 1 << 26);
        double h[2]; CK(cudaMemcpy(h, errs, 16, cudaMemcpyDeviceToHost));
        printf("{\"bench\": \"rcp64h_accuracy\", \"max_rel_err_seed\": %.3e, \"max_rel_err_newton1\": %.3e}\n", h[1], h[0]);
    }
    return 0;
}
