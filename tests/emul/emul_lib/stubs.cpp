// TEST INFRASTRUCTURE ONLY: the entry points of csrc/fft_tma.cu, csrc/fft_zrow.cu (TMA / bulk-copy fast path: Nmesh >= 512 only,
// emulated kernel by kernel in tests/emul/tma_emul.cpp, zrow_emul.cpp) and csrc/comm.cu (several GPUs) that the emulated
// one-rank library links against but never reaches with the small meshes of its tests.
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
int fpm_xbarrier_on(cudaStream_t) { unreached("fpm_xbarrier_on"); return -1; }
#define STUB(name) extern "C" int name() { unreached(#name); return -1; }
STUB(fpm_ipc_get_handle) STUB(fpm_xbarrier_init) STUB(fpm_xbarrier_set_peers) STUB(fpm_xbarrier) STUB(fpm_r2c_dist) STUB(fpm_c2r_dist)
STUB(fpm_halo_add_from) STUB(fpm_halo_fetch_from) STUB(fpm_migrate_init) STUB(fpm_migrate_classify) STUB(fpm_migrate_pack_column)
STUB(fpm_migrate_holes) STUB(fpm_migrate_fill_column) STUB(fpm_migrate_append_column) STUB(fpm_mesh_set_stage)
extern "C" void *fpm_ipc_open() { unreached("fpm_ipc_open"); return NULL; }
extern "C" void fpm_migrate_destroy() { }
