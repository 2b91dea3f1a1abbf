// Minimal CPU stand-ins for the CUDA execution model, shared by the kernel emulations in this directory: every CUDA thread of
// a CTA is an OS thread, __syncthreads() a pthread barrier, shared memory an ordinary array, the asynchronous copies of
// csrc/fft_reg.cuh synchronous (FPM_EMULATE).  The kernel SOURCE that runs here is the one nvcc compiles for the device.
#pragma once
#include <pthread.h>
#include <sched.h>
#include <string.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <functional>
#include <thread>
#include <vector>

#define FPM_EMULATE 1
struct FpmEmulDim3 { unsigned x = 1, y = 1, z = 1; };
static thread_local FpmEmulDim3 threadIdx;
static FpmEmulDim3 blockIdx, blockDim, gridDim;
static pthread_barrier_t fpm_emul_barrier;
#ifdef FPM_EMUL_EXTERN_SMEM                      // several translation units in one library (emul_lib/): one of them defines it
extern unsigned char *fpm_emul_dyn_smem;
#else
unsigned char *fpm_emul_dyn_smem = nullptr;     // declared extern by csrc/common.cuh under FPM_EMULATE
#endif
static bool fpm_emul_sequential = false;        // CUDA threads run one after the other (kernels without barriers only)
static inline void fpm_emul_barrier_wait()
{
    if (fpm_emul_sequential) { fprintf(stderr, "cuda_emul: __syncthreads() in a kernel launched sequentially\n"); abort(); }
    pthread_barrier_wait(&fpm_emul_barrier);
}
#define __syncthreads() fpm_emul_barrier_wait()
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

// ---- warp primitives: the 32 lanes of a warp meet at a per-warp barrier around an exchange buffer (the kernels call them
// with warp-uniform control flow, like the hardware requires for *_sync with a full mask)
static pthread_barrier_t fpm_emul_warp_barrier[32];
static double fpm_emul_shfl_buf[32][32];
template <typename T> static T fpm_emul_shfl(T v, int src)
{
    static_assert(sizeof(T) <= sizeof(double), "shuffle of at most 8 bytes");
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    memcpy(&fpm_emul_shfl_buf[warp][lane], &v, sizeof(T));
    pthread_barrier_wait(&fpm_emul_warp_barrier[warp]);
    T r = v;
    if (src >= 0 && src < 32) memcpy(&r, &fpm_emul_shfl_buf[warp][src], sizeof(T));
    pthread_barrier_wait(&fpm_emul_warp_barrier[warp]);
    return r;
}
template <typename T> static T __shfl_down_sync(unsigned, T v, int d) { return fpm_emul_shfl(v, (int) (threadIdx.x % 32) + d); }
template <typename T> static T __shfl_up_sync(unsigned, T v, int d) { return fpm_emul_shfl(v, (int) (threadIdx.x % 32) - d); }
template <typename T> static T __shfl_xor_sync(unsigned, T v, int x) { return fpm_emul_shfl(v, (int) (threadIdx.x % 32) ^ x); }
static inline void __syncwarp() { pthread_barrier_wait(&fpm_emul_warp_barrier[threadIdx.x / 32]); }
static pthread_mutex_t fpm_emul_atomic_lock = PTHREAD_MUTEX_INITIALIZER;
static inline double atomicAdd(double *p, double v)
{
    pthread_mutex_lock(&fpm_emul_atomic_lock);
    const double o = *p; *p = o + v;
    pthread_mutex_unlock(&fpm_emul_atomic_lock);
    return o;
}

// The CUDA threads of a CTA are OS threads that must all be alive at once (__syncthreads() is a real barrier); creating and joining up to
// 1024 of them for every CTA of every launch dominated the run time of the emulation, so they are kept in a pool (one per translation
// unit, like everything else in this header) and handed one CTA at a time.
#include <atomic>
#include <memory>
#include <semaphore.h>
namespace {          // internal linkage: every translation unit that includes this header needs ITS OWN copy of these member functions,
                     // because they touch that unit's thread_local threadIdx (one shared inline definition would set somebody else's)
struct FpmEmulPool {
    struct Worker { sem_t go; std::thread th; };
    std::vector<std::unique_ptr<Worker>> workers;       // worker i plays CUDA thread i of the current CTA; only the first `block` are woken
    sem_t done;
    const std::function<void()> *job = nullptr;
    std::atomic<unsigned> remaining{0};
    std::atomic<bool> stop{false};
    FpmEmulPool() { sem_init(&done, 0, 0); }
    void run(unsigned id)
    {
        Worker *w = workers[id].get();
        for (;;) {
            while (sem_wait(&w->go) != 0) { }
            if (stop.load()) return;
            threadIdx.x = id;
            (*job)();
            if (remaining.fetch_sub(1) == 1) sem_post(&done);
        }
    }
    void cta(unsigned block, const std::function<void()> &k)
    {
        while (workers.size() < block) {
            const unsigned id = (unsigned) workers.size();
            workers.emplace_back(new Worker);
            sem_init(&workers[id]->go, 0, 0);
        }
        for (unsigned id = 0; id < block; id++)
            if (!workers[id]->th.joinable()) workers[id]->th = std::thread([this, id]() { run(id); });
        job = &k;
        remaining.store(block);
        for (unsigned id = 0; id < block; id++) sem_post(&workers[id]->go);
        while (sem_wait(&done) != 0) { }
    }
    ~FpmEmulPool()
    {
        stop.store(true);
        for (auto &w : workers) if (w->th.joinable()) { sem_post(&w->go); w->th.join(); }
    }
};
}
static FpmEmulPool fpm_emul_pool;

// runs kernel(args...) as a grid of `grid` CTAs of `block` threads with `smem_bytes` of dynamic shared memory, CTA after CTA
template <typename Kernel>
static void fpm_emul_launch(unsigned grid, unsigned block, size_t smem_bytes, Kernel kernel)
{
    std::vector<unsigned char> smem(smem_bytes + 1024);
    fpm_emul_dyn_smem = (unsigned char *) (((uintptr_t) smem.data() + 1023) & ~(uintptr_t) 1023);
    gridDim.x = grid; blockDim.x = block;
    const std::function<void()> body = [&]() { kernel(); };
    for (unsigned b = 0; b < grid; b++) {
        blockIdx.x = b;
        pthread_barrier_init(&fpm_emul_barrier, NULL, block);
        for (unsigned w = 0; w < (block + 31) / 32 && w < 32; w++)
            pthread_barrier_init(&fpm_emul_warp_barrier[w], NULL, (w + 1) * 32 <= block ? 32 : block - w * 32);
        fpm_emul_pool.cta(block, body);
        pthread_barrier_destroy(&fpm_emul_barrier);
        for (unsigned w = 0; w < (block + 31) / 32 && w < 32; w++) pthread_barrier_destroy(&fpm_emul_warp_barrier[w]);
    }
}
