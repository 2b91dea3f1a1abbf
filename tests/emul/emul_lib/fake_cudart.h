// TEST INFRASTRUCTURE ONLY (tests/emul/emul_lib): stand-ins for the CUDA runtime so that the WHOLE library -- the host parts of
// csrc/*.cu (launchers, C ABI) and the C host layer on top -- runs on the CPU with the kernel sources emulated by cuda_emul.h.
// "Device" memory is host memory, a stream is executed at once, an event is a wall-clock stamp.  One device, 4 "SMs".
#pragma once
#include <cuda_runtime.h>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

// ---- device atomics the kernels use (threads of a CTA may be OS threads: real atomics)
static inline float atomicAdd(float *p, float v)
{
    unsigned int *u = reinterpret_cast<unsigned int *>(p), o = __atomic_load_n(u, __ATOMIC_RELAXED), n;
    float of;
    do { memcpy(&of, &o, 4); const float nf = of + v; memcpy(&n, &nf, 4); } while (!__atomic_compare_exchange_n(u, &o, n, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    return of;
}
static inline void atomicAdd(float2 *p, float2 v) { atomicAdd(&p->x, v.x); atomicAdd(&p->y, v.y); }
static inline void atomicAdd(float4 *p, float4 v) { atomicAdd(&p->x, v.x); atomicAdd(&p->y, v.y); atomicAdd(&p->z, v.z); atomicAdd(&p->w, v.w); }
static inline int atomicExch(int *p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline int atomicMin(int *p, int v) { int o = __atomic_load_n(p, __ATOMIC_RELAXED); while (v < o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) { } return o; }
static inline int atomicMax(int *p, int v) { int o = __atomic_load_n(p, __ATOMIC_RELAXED); while (v > o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) { } return o; }
static inline unsigned int atomicAdd(unsigned int *p, unsigned int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline void __threadfence_system() { __sync_synchronize(); }
static inline void __threadfence() { __sync_synchronize(); }
static inline void sincospif(float x, float *s, float *c) { *s = sinf((float) M_PI * x); *c = cosf((float) M_PI * x); }
static inline void sincospi(double x, double *s, double *c) { *s = sin(M_PI * x); *c = cos(M_PI * x); }

// cuda_runtime.h only has this overload for kernels under nvcc
template <typename T> static inline cudaError_t cudaFuncSetAttribute(T *, cudaFuncAttribute, int) { return cudaSuccess; }

// ---- kernel launches: tests/emul/emul_lib/build.py rewrites  k<<<g, b, s, st>>>(args)  into  FPM_EMUL_LAUNCH((k), g, b, s, st, args)
#define FPM_EMUL_UNPAREN(...) __VA_ARGS__
template <typename Kernel>
static void fpm_emul_launch_auto(const char *name, unsigned grid, unsigned block, size_t smem, Kernel kernel)
{
    // kernels with barriers or warp shuffles need one OS thread per CUDA thread; all others run their threads one after the other
    const std::string n(name);
    const bool threaded = n.find("fft_") != std::string::npos || n.find("powerspectrum") != std::string::npos || n.find("summary") != std::string::npos || n.find("_tile_kernel") != std::string::npos;
    if (threaded) { fpm_emul_sequential = false; fpm_emul_launch(grid, block, smem, kernel); return; }
    fpm_emul_sequential = true;
    gridDim.x = grid; blockDim.x = block;
    for (unsigned b = 0; b < grid; b++) { blockIdx.x = b; for (unsigned t = 0; t < block; t++) { threadIdx.x = t; kernel(); } }
    fpm_emul_sequential = false;
}
#define FPM_EMUL_LAUNCH(k, g, b, s, st, ...) fpm_emul_launch_auto(#k, (unsigned) (g), (unsigned) (b), (size_t) (s), [&]() { FPM_EMUL_UNPAREN k(__VA_ARGS__); })
