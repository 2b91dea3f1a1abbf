// Microbenchmark behind the design of the CIC deposit (csrc/paint.cu): issue rates of the ways a thread can add a float to a
// mesh cell on sm_100a.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atomics atomics.cu ; run: ./atomics
//   red.global.add.f32 / .v2.f32 / .v4.f32 at pseudo-random addresses, footprint L2-resident (64 MB) and DRAM-sized (8 GB);
//   shared-memory atomicAdd(float) (compiles to an ATOMS.CAST.SPIN compare-and-swap loop: there is no native float add on
//   shared memory) and atomicAdd(int) (native ATOMS.ADD) on a 48 KB tile;
//   FFMA against FFMA2 (packed f32x2) issue throughput, for the FFT butterflies.
// Prints operations per second and cycles per lane-operation per SM at the clock read from the device.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s; }

template <int VEC>
__global__ void red_kernel(float *mesh, size_t nvec, int nper)
{
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    for (int i = 0; i < nper; i++) {
        const size_t a = ((size_t) lcg(s) * 4096u + (lcg(s) >> 20)) % nvec;
        if (VEC == 1) atomicAdd(mesh + a, 1.0f);
        else if (VEC == 2) atomicAdd(reinterpret_cast<float2 *>(mesh) + a, make_float2(1.f, 1.f));
        else atomicAdd(reinterpret_cast<float4 *>(mesh) + a, make_float4(1.f, 1.f, 0.f, 0.f));
    }
}

// coherent variant: the lanes of a warp hit neighbouring float4s of a few rows, like Lagrangian neighbours in the CIC deposit
__global__ void red_coherent_kernel(float *mesh, size_t nvec, int nper)
{
    uint32_t s = (blockIdx.x * blockDim.x + (threadIdx.x >> 5)) * 2654435761u + 777u;     // one stream per warp
    const int lane = threadIdx.x & 31;
    for (int i = 0; i < nper; i++) {
        const size_t base = ((size_t) lcg(s) * 4096u + (lcg(s) >> 20)) % (nvec - 4096);
        const size_t a = base + (lane & 7) + (size_t) (lane >> 3) * 520;                  // 8 adjacent float4 in each of 4 rows
        atomicAdd(reinterpret_cast<float4 *>(mesh) + a, make_float4(1.f, 1.f, 0.f, 0.f));
    }
}

template <bool ISFLOAT>
__global__ void smem_kernel(float *out, int nper)
{
    __shared__ float tile[12288];
    for (int i = threadIdx.x; i < 12288; i += blockDim.x) tile[i] = 0.f;
    __syncthreads();
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 99u;
    for (int i = 0; i < nper; i++) {
        const int a = lcg(s) % 12288;
        if (ISFLOAT) atomicAdd(tile + a, 1.0f);
        else atomicAdd(reinterpret_cast<int *>(tile) + a, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = tile[5];
}

template <bool PACKED>
__global__ void fma_kernel(float2 *out, int nper)
{
    float2 a[8];
    #pragma unroll
    for (int k = 0; k < 8; k++) a[k] = make_float2(threadIdx.x * 1e-3f + k, 1.0f);
    const float2 b = make_float2(1.0001f, 0.9999f), c = make_float2(1e-4f, -1e-4f);
    for (int i = 0; i < nper; i++) {
        #pragma unroll
        for (int k = 0; k < 8; k++) {
            if (PACKED) a[k] = __ffma2_rn(a[k], b, c);
            else { a[k].x = fmaf(a[k].x, b.x, c.x); a[k].y = fmaf(a[k].y, b.y, c.y); }
        }
    }
    float2 r = make_float2(0.f, 0.f);
    #pragma unroll
    for (int k = 0; k < 8; k++) { r.x += a[k].x; r.y += a[k].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <typename F> static float time_ms(F f)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();                                     // warm-up
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main()
{
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int nsm = p.multiProcessorCount;
    const double ghz = clk_khz * 1e-6;
    printf("%s, %d SMs, %.3f GHz (nominal)\n", p.name, nsm, ghz);
    const size_t big = (size_t) 8 << 30;
    float *mesh = nullptr;
    CK(cudaMalloc(&mesh, big));
    CK(cudaMemset(mesh, 0, big));
    const int grid = nsm * 8, block = 256, nper = 2048;
    const double lanes = (double) grid * block * nper;
    struct { const char *name; size_t bytes; } foot[2] = { { "64 MB (L2)", (size_t) 64 << 20 }, { "8 GB (DRAM)", big } };
    for (int f = 0; f < 2; f++) {
        const size_t nf = foot[f].bytes / 4;
        float ms;
        ms = time_ms([&]() { red_kernel<1><<<grid, block>>>(mesh, nf, nper); });
        printf("red.global.add.f32     random, %-12s: %8.2f Gop/s  %6.2f cyc/lane/SM\n", foot[f].name, lanes / ms * 1e-6, ms * 1e-3 * ghz * 1e9 * nsm / lanes);
        ms = time_ms([&]() { red_kernel<2><<<grid, block>>>(mesh, nf / 2, nper); });
        printf("red.global.add.v2.f32  random, %-12s: %8.2f Gop/s  %6.2f cyc/lane/SM\n", foot[f].name, lanes / ms * 1e-6, ms * 1e-3 * ghz * 1e9 * nsm / lanes);
        ms = time_ms([&]() { red_kernel<4><<<grid, block>>>(mesh, nf / 4, nper); });
        printf("red.global.add.v4.f32  random, %-12s: %8.2f Gop/s  %6.2f cyc/lane/SM\n", foot[f].name, lanes / ms * 1e-6, ms * 1e-3 * ghz * 1e9 * nsm / lanes);
        ms = time_ms([&]() { red_coherent_kernel<<<grid, block>>>(mesh, nf / 4, nper); });
        printf("red.global.add.v4.f32  4 rows x 8 adjacent, %-12s: %8.2f Gop/s  %6.2f cyc/lane/SM\n", foot[f].name, lanes / ms * 1e-6, ms * 1e-3 * ghz * 1e9 * nsm / lanes);
    }
    float *out = nullptr;
    CK(cudaMalloc(&out, sizeof(float) * grid * block * 2));
    float ms = time_ms([&]() { smem_kernel<true><<<grid, block>>>(out, nper); });
    printf("shared atomicAdd(float) random in 48 KB  : %8.2f Gop/s  %6.2f cyc/lane/SM\n", lanes / ms * 1e-6, ms * 1e-3 * ghz * 1e9 * nsm / lanes);
    ms = time_ms([&]() { smem_kernel<false><<<grid, block>>>(out, nper); });
    printf("shared atomicAdd(int)   random in 48 KB  : %8.2f Gop/s  %6.2f cyc/lane/SM\n", lanes / ms * 1e-6, ms * 1e-3 * ghz * 1e9 * nsm / lanes);
    const int nfma = 4096;
    const double fl = (double) grid * block * nfma * 8;
    ms = time_ms([&]() { fma_kernel<false><<<grid, block>>>((float2 *) out, nfma); });
    printf("FFMA  (2 per complex): %8.2f G complex-fma/s  %6.3f cyc per warp-complex-fma per SMSP\n", fl / ms * 1e-6, ms * 1e-3 * ghz * 1e9 * nsm * 4 / (fl / 32));
    ms = time_ms([&]() { fma_kernel<true><<<grid, block>>>((float2 *) out, nfma); });
    printf("FFMA2 (1 per complex): %8.2f G complex-fma/s  %6.3f cyc per warp-complex-fma per SMSP\n", fl / ms * 1e-6, ms * 1e-3 * ghz * 1e9 * nsm * 4 / (fl / 32));
    CK(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}
