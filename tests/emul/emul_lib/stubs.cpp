// TEST INFRASTRUCTURE ONLY: the entry points of csrc/fft_tma.cu, csrc/fft_zrow.cu (TMA / bulk-copy fast path: Nmesh >= 512 only,
// emulated kernel by kernel in tests/emul/tma_emul.cpp, zrow_emul.cpp) that the emulated library links against but never reaches
// with the small meshes of its tests.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstddef>
struct TmaPassArgs;
static void unreached(const char *n) { fprintf(stderr, "emul_lib: %s reached\n", n); abort(); }
int fpm_fft_tma_supported(int) { return 0; }
int fpm_fft_tma_tile_k(int) { return 16; }
int fpm_fft_tma_pass(int, const float2 *, int, int, const TmaPassArgs &, cudaStream_t) { unreached("fpm_fft_tma_pass"); return -1; }
int fpm_fft_zrow_supported(int, size_t) { return 0; }
int fpm_fft_zrow_pass(int, const float *, float *, size_t, int, float, const float2 *, const float2 *, int, cudaStream_t) { unreached("fpm_fft_zrow_pass"); return -1; }
