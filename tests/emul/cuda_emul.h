// Minimal CPU stand-ins for the CUDA execution model, shared by the kernel emulations in this directory: every CUDA thread of
// a CTA is an OS thread, __syncthreads() a pthread barrier, shared memory an ordinary array, the asynchronous copies of
// csrc/fft_reg.cuh synchronous (FPM_EMULATE).  The kernel SOURCE that runs here is the one nvcc compiles for the device.
#pragma once
#include <pthread.h>
#include <sched.h>
#include <string.h>
#include <stdint.h>
#include <functional>
#include <thread>
#include <vector>

#define FPM_EMULATE 1
struct FpmEmulDim3 { unsigned x = 1, y = 1, z = 1; };
static thread_local FpmEmulDim3 threadIdx;
static FpmEmulDim3 blockIdx, blockDim, gridDim;
static pthread_barrier_t fpm_emul_barrier;
unsigned char *fpm_emul_dyn_smem = nullptr;
#define __syncthreads() pthread_barrier_wait(&fpm_emul_barrier)
#define __ldg(p) (*(p))
#define __global__
#define __grid_constant__
#define __launch_bounds__(...)
#define __shared__ static
static inline size_t __cvta_generic_to_shared(const void *p) { return (size_t) p; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmaf_rn(float a, float b, float c) { return __builtin_fmaf(a, b, c); }
static inline double __drcp_rn(double a) { return 1.0 / a; }

// runs kernel(args...) as a grid of `grid` CTAs of `block` threads with `smem_bytes` of dynamic shared memory, CTA after CTA
template <typename Kernel>
static void fpm_emul_launch(unsigned grid, unsigned block, size_t smem_bytes, Kernel kernel)
{
    std::vector<unsigned char> smem(smem_bytes + 1024);
    fpm_emul_dyn_smem = (unsigned char *) (((uintptr_t) smem.data() + 1023) & ~(uintptr_t) 1023);
    gridDim.x = grid; blockDim.x = block;
    for (unsigned b = 0; b < grid; b++) {
        blockIdx.x = b;
        pthread_barrier_init(&fpm_emul_barrier, NULL, block);
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < block; t++) pool.emplace_back([&, t]() { threadIdx.x = t; kernel(); });
        for (auto &th : pool) th.join();
        pthread_barrier_destroy(&fpm_emul_barrier);
    }
}
