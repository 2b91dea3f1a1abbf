// fastpm_b200 -- shared device/host definitions for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stddef.h>

#define FPM_MAX_RANKS 8
#define FPM_MAX_STAGES 16

// ---------------------------------------------------------------- dynamic shared memory of a kernel
// (FPM_EMULATE: the kernel sources are also compiled for the CPU by tests/emul/, where shared memory is an ordinary array)
#ifndef FPM_EMULATE
#define FPM_DYN_SMEM(name, alignment) extern __shared__ __align__(alignment) unsigned char name[]
#else
extern unsigned char *fpm_emul_dyn_smem;
#define FPM_DYN_SMEM(name, alignment) unsigned char *name = fpm_emul_dyn_smem
#endif

// ---------------------------------------------------------------- errors
extern "C" void fpm_set_error(const char *fmt, ...);
extern unsigned long long fpm_launch_counter;      // kernels launched by this library (bench "gpu_launches")

#define FPM_CUDA_OK(expr)                                                                    \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            fpm_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return -1;                                                                       \
        }                                                                                    \
    } while (0)

extern int fpm_debug_sync;                          // FASTPM_B200_DEBUG_SYNC=1: synchronise after every launch (fault isolation)
#define FPM_CHECK_LAUNCH()                                                                   \
    do {                                                                                     \
        fpm_launch_counter++;                                                                \
        cudaError_t _e = cudaGetLastError();                                                 \
        if (_e == cudaSuccess && fpm_debug_sync) _e = cudaDeviceSynchronize();               \
        if (_e != cudaSuccess) {                                                             \
            fpm_set_error("%s:%d: kernel launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return -1;                                                                       \
        }                                                                                    \
    } while (0)

// ---------------------------------------------------------------- per-kernel-class CUDA-event timing
// Off by default.  When enabled (fpm_prof_enable), every launch of this library is bracketed by two events on
// the launching stream; fpm_prof_get sums the elapsed times per class.  bench.py uses it for the roofline line.
enum FpmKernelClass {
    FPM_K_PAINT = 0, FPM_K_READOUT, FPM_K_FFT_TILE, FPM_K_FFT_Z, FPM_K_KICK, FPM_K_DRIFT, FPM_K_KSPACE,
    FPM_K_PK, FPM_K_SUMMARY, FPM_K_OTHER,
    // round 2: what used to hide between the classes above on several GPUs
    FPM_K_MEMSET, FPM_K_BARRIER, FPM_K_HALO, FPM_K_MIGRATE, FPM_K_PUSH, FPM_K_COUNT
};
// which code path served a launch (tests assert that the benched kernels are the ones a parity run went through)
enum FpmPathCounter {
    FPM_PATH_FFT_TMA = 0, FPM_PATH_FFT_TILE_GENERIC, FPM_PATH_FFT_ZROW, FPM_PATH_FFT_Z_GENERIC, FPM_PATH_FFT_TMA_MULTI,
    FPM_PATH_PAINT_BRICKS, FPM_PATH_READOUT_BRICKS, FPM_PATH_PK_ROWS, FPM_PATH_STAGED_TRANSPOSE, FPM_PATH_PAINT_TILES,
    FPM_PATH_READOUT_TILES, FPM_PATH_READOUT3, FPM_PATH_COUNT
};
extern unsigned long long fpm_path_counter[FPM_PATH_COUNT];
extern int fpm_prof_on;
void fpm_prof_begin(int cls, cudaStream_t st);
void fpm_prof_end(int cls, cudaStream_t st);
#define FPM_TIMED(cls, st, stmt) do { if (fpm_prof_on) fpm_prof_begin(cls, st); stmt; if (fpm_prof_on) fpm_prof_end(cls, st); } while (0)

// ---------------------------------------------------------------- mesh geometry
// One PM mesh of Nmesh^3 cells over a periodic box, x-slab decomposed over `nranks` GPUs.
//
// Real-space layout (this rank):   real[xl][y][z],  xl in [0, nxl (+1 halo plane)), pitch_r floats per row
// k-space layout (this rank):      cplx[kyl][kx][kz], kyl in [0, nyl), pitch_c complex per row, kz in [0, N/2]
//   (the "transposed-out" order the reference asks PFFT for, pmpfft.c:198-203, with ky slowest so the
//    slab all-to-all is folded into the y-pass / inverse x-pass stores)
// pitch_c = N/2+1 rounded up to 16 complex (128 B) so every tile row is sector aligned; pitch_r = 2*pitch_c.
struct FpmGeom {
    int n;              // Nmesh
    int nranks, rank;
    int nxl, x0;        // local x planes [x0, x0+nxl)
    int nyl, y0;        // local ky planes [y0, y0+nyl) in k-space
    int pitch_c;        // complex elements per row
    int pitch_r;        // floats per row (= 2*pitch_c)
    double boxsize;
    double cellsize, inv_cellsize;
};

// ---------------------------------------------------------------- complex helpers
__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by -i : (x + iy)(-i) = y - ix
__device__ __forceinline__ float2 cmul_mi(float2 a) { return make_float2(a.y, -a.x); }
__device__ __forceinline__ float2 cmul_pi(float2 a) { return make_float2(-a.y, a.x); }
