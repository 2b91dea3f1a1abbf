// fastpm_b200 -- multi-GPU device plumbing for the x-slab decomposition (one process per GPU):
//   * CUDA-IPC mapping of peer mesh buffers, so that the transposing FFT passes store straight into the peer that
//     owns the destination planes (the slab all-to-all of PFFT, pmpfft.c:281-303, fused into the pass) and halo
//     planes are pulled with plain loads over NVLink;
//   * a cross-GPU barrier kernel on flags in peer memory (system-scope fences), stream ordered, no host sync;
//   * one-plane mesh halos replacing the particle ghosts of pmghosts.c (support 2 needs planes [x0, x0+nxl]);
//   * particle migration after drift (fastpm_store_decompose, store.c:486-657): classify by owner slab, pack the
//     leavers per destination column by column, fill the holes from the tail, append what the peers packed for us.
#include "common.cuh"
#include "mesh.cuh"
#include "../../include/fastpm_b200.h"
#include <string.h>
#include <stdlib.h>
#include <vector>
#include <map>

cudaStream_t fpm_internal_stream(void);
static cudaStream_t comm_stream(void) { return fpm_internal_stream(); }

// Bytes this rank moved over NVLink since start-up, counted where the transfers are issued: [0] slab-transpose pushes by the copy
// engines (fft.cu), [1] halo planes pulled, [2] migrating particles pulled, [3] rows stored straight into peers by the transposing
// FFT pass when no staging mesh is used.  (nvidia-smi nvlink -gt d reports N/A on these B200 boxes.)
unsigned long long fpm_comm_bytes[4] = { 0, 0, 0, 0 };
extern "C" int fpm_comm_byte_counts(uint64_t *out4) { for (int i = 0; i < 4; i++) out4[i] = fpm_comm_bytes[i]; return 0; }

// ------------------------------------------------------------------ IPC
// cudaMalloc sub-allocates small blocks out of larger ones and an IPC handle always names the WHOLE underlying
// allocation: a pointer is therefore published as (handle of its allocation, offset inside it), and a handle is
// opened once per process and remembered.
#include <cuda.h>
#include <cudaTypedefs.h>
extern "C" int fpm_ipc_get_handle(void *dev_ptr, void *handle64, uint64_t *offset)
{
    static PFN_cuMemGetAddressRange_v3020 get_range = nullptr;
    if (!get_range) {
        cudaDriverEntryPointQueryResult q; void *p = nullptr;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            fpm_set_error("cuMemGetAddressRange is not available from the driver"); return -1;
        }
        get_range = (PFN_cuMemGetAddressRange_v3020) p;
    }
    CUdeviceptr base = 0; size_t size = 0;
    if (get_range(&base, &size, (CUdeviceptr) dev_ptr) != CUDA_SUCCESS) { fpm_set_error("cuMemGetAddressRange failed"); return -1; }
    cudaIpcMemHandle_t h;
    FPM_CUDA_OK(cudaIpcGetMemHandle(&h, (void *) base));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &h, 64);
    *offset = (uint64_t) ((CUdeviceptr) dev_ptr - base);
    return 0;
}

struct OpenedHandle { unsigned char h[64]; void *base; };
static std::vector<OpenedHandle> g_opened;

extern "C" void *fpm_ipc_open(const void *handle64, uint64_t offset)
{
    for (size_t i = 0; i < g_opened.size(); i++)
        if (!memcmp(g_opened[i].h, handle64, 64)) return (unsigned char *) g_opened[i].base + offset;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void *p = NULL;
    int dev_before = -1, dev_after = -1;
    cudaGetDevice(&dev_before);
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { fpm_set_error("cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return NULL; }
    cudaGetDevice(&dev_after);
    if (getenv("FASTPM_B200_DEBUG_IPC")) {
        cudaPointerAttributes at; memset(&at, 0, sizeof(at));
        cudaPointerGetAttributes(&at, p);
        fprintf(stderr, "fpm_ipc_open: device before %d, after %d; mapped %p lives on device %d (type %d)\n", dev_before, dev_after, p, at.device, (int) at.type);
    }
    if (dev_before >= 0) cudaSetDevice(dev_before);      // mapping a peer's memory must not leave another device current
    OpenedHandle o; memcpy(o.h, handle64, 64); o.base = p;
    g_opened.push_back(o);
    return (unsigned char *) p + offset;
}

// ------------------------------------------------------------------ cross-GPU barrier
struct XBarrier {
    int nranks, rank;
    unsigned long long *flags;                       // [nranks] local, written by the peers
    unsigned long long *peer_flags[FPM_MAX_RANKS];   // peer_flags[d] = rank d's flags array (mapped)
    unsigned long long epoch;
};
static XBarrier g_bar = { 1, 0, NULL, { NULL }, 0 };

__global__ void xbarrier_kernel(unsigned long long *const *peer_flags_dev, volatile unsigned long long *my_flags, int me, int n,
                                unsigned long long epoch)
{
    const int d = threadIdx.x;
    if (d < n) {
        __threadfence_system();
        // announce: my slot in rank d's array
        volatile unsigned long long *dst = peer_flags_dev[d] + me;
        *dst = epoch;
        __threadfence_system();
        while (my_flags[d] < epoch) { }
        __threadfence_system();
    }
}

static unsigned long long **g_peer_flags_dev = NULL;

extern "C" int fpm_xbarrier_init(int nranks, int rank, void *local_flags)
{
    g_bar.nranks = nranks; g_bar.rank = rank; g_bar.epoch = 0;
    g_bar.flags = (unsigned long long *) local_flags;
    FPM_CUDA_OK(cudaMemset(g_bar.flags, 0, sizeof(unsigned long long) * FPM_MAX_RANKS));
    return 0;
}

extern "C" int fpm_xbarrier_set_peers(void *const *peer_flag_ptrs)
{
    for (int d = 0; d < g_bar.nranks; d++)
        g_bar.peer_flags[d] = (d == g_bar.rank) ? g_bar.flags : (unsigned long long *) peer_flag_ptrs[d];
    if (!g_peer_flags_dev) FPM_CUDA_OK(cudaMalloc(&g_peer_flags_dev, sizeof(void *) * FPM_MAX_RANKS));
    FPM_CUDA_OK(cudaMemcpy(g_peer_flags_dev, g_bar.peer_flags, sizeof(void *) * FPM_MAX_RANKS, cudaMemcpyHostToDevice));
    return 0;
}

int fpm_xbarrier_on(cudaStream_t st)
{
    if (g_bar.nranks <= 1) return 0;
    g_bar.epoch++;
    FPM_TIMED(FPM_K_BARRIER, st, (xbarrier_kernel<<<1, 32, 0, st>>>(g_peer_flags_dev, g_bar.flags, g_bar.rank, g_bar.nranks, g_bar.epoch)));
    FPM_CHECK_LAUNCH();
    return 0;
}
extern "C" int fpm_xbarrier(void) { return fpm_xbarrier_on(comm_stream()); }

static int mesh_barrier(FpmMesh *m, cudaStream_t st) { (void) m; return fpm_xbarrier_on(st); }

// ------------------------------------------------------------------ distributed transforms
// peers[d] = rank d's buffer (peers[rank] = the local one); the FFT code (fft.cu) puts a barrier before and after
// the transposing pass through m->barrier.
// local staging mesh for the slab transposes of this mesh (fft.cu: staged_transpose); NULL = direct peer stores
extern "C" int fpm_mesh_set_stage(fpm_mesh *m, float *stage) { m->stage = stage; return 0; }
extern "C" int fpm_mesh_set_stage2(fpm_mesh *m, float *stage) { m->stage2 = stage; return 0; }

int fpm_lazy_touch(const void *p, size_t bytes);      // capi.cu: applies a deferred deconvolution of that buffer first
extern "C" int fpm_r2c_dist(fpm_mesh *m, float *real, float *const *cplx_peers, double scale)
{
    if (fpm_lazy_touch(real, 0) || fpm_lazy_touch(cplx_peers[m->geom.rank], 0)) return -1;
    m->barrier = mesh_barrier;
    return fpm_fft_r2c(m, real, real, cplx_peers, (float) scale, comm_stream());
}

extern "C" int fpm_c2r_dist(fpm_mesh *m, const float *cplx, float *const *real_peers, const fpm_transfer *kernel)
{
    if (fpm_lazy_touch(cplx, 0) || fpm_lazy_touch(real_peers[m->geom.rank], 0)) return -1;
    m->barrier = mesh_barrier;
    FpmTransferSpec s;
    if (kernel && kernel->active) {
        s.active = kernel->active; s.potorder = kernel->potorder; s.negate = kernel->negate; s.ngrad = kernel->ngrad;
        s.graddir[0] = kernel->graddir[0]; s.graddir[1] = kernel->graddir[1]; s.gradorder = kernel->gradorder;
        s.zero_selfconj = kernel->zero_selfconj; s.scale = kernel->scale;
        return fpm_fft_c2r(m, cplx, real_peers, NULL, &s, comm_stream());
    }
    return fpm_fft_c2r(m, cplx, real_peers, NULL, NULL, comm_stream());
}

static void to_spec(const fpm_transfer *kernel, FpmTransferSpec *s)
{
    s->active = kernel->active; s->potorder = kernel->potorder; s->negate = kernel->negate; s->ngrad = kernel->ngrad;
    s->graddir[0] = kernel->graddir[0]; s->graddir[1] = kernel->graddir[1]; s->gradorder = kernel->gradorder;
    s->zero_selfconj = kernel->zero_selfconj; s->scale = kernel->scale;
}
// the two halves of fpm_c2r_dist (fft.cu: fpm_fft_c2r_begin / _finish); `set` selects the staging mesh and the events
extern "C" int fpm_c2r_dist_begin(fpm_mesh *m, const float *cplx, float *const *real_peers, const fpm_transfer *kernel, int set)
{
    if (fpm_lazy_touch(cplx, 0) || fpm_lazy_touch(real_peers[m->geom.rank], 0)) return -1;
    m->barrier = mesh_barrier;
    FpmTransferSpec s;
    if (kernel && kernel->active) { to_spec(kernel, &s); return fpm_fft_c2r_begin(m, cplx, real_peers, &s, set, comm_stream()); }
    return fpm_fft_c2r_begin(m, cplx, real_peers, NULL, set, comm_stream());
}
extern "C" int fpm_c2r_dist_finish(fpm_mesh *m, float *const *real_peers, int set)
{
    m->barrier = mesh_barrier;
    return fpm_fft_c2r_finish(m, real_peers, NULL, set, comm_stream());
}

// ------------------------------------------------------------------ halo planes
// dst = src (one mesh plane pulled from the neighbour with plain loads over NVLink): the copy engines are busy pushing the slab
// transposes of the next force component (pipelined inverse transforms), where a cudaMemcpyAsync of this plane queued for ~0.6 ms
__global__ void __launch_bounds__(256) halo_copy_kernel(float *__restrict__ dst, const float *__restrict__ src, size_t n4)
{
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    float4 *d = reinterpret_cast<float4 *>(dst);
    const float4 *s = reinterpret_cast<const float4 *>(src);
    for (; i < n4; i += stride) d[i] = s[i];
}

__global__ void __launch_bounds__(256) halo_add_kernel(float *__restrict__ dst, const float *__restrict__ src, size_t n4)
{
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    float4 *d = reinterpret_cast<float4 *>(dst);
    const float4 *s = reinterpret_cast<const float4 *>(src);
    for (; i < n4; i += stride) {
        float4 a = d[i]; const float4 b = s[i];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        d[i] = a;
    }
}

// paint epilogue: my plane 0 += the halo plane (index nxl) of the previous rank.  Barriers on both sides: every rank
// has finished painting before anybody pulls, and everybody has pulled before the canvas is written again.
extern "C" int fpm_halo_add_from(const fpm_mesh *m, float *canvas_local, const float *canvas_prev_rank)
{
    const FpmGeom &g = m->geom;
    const size_t plane = (size_t) g.n * g.pitch_r;
    cudaStream_t st = comm_stream();
    if (fpm_xbarrier_on(st)) return -1;
    FPM_TIMED(FPM_K_HALO, st, (halo_add_kernel<<<148 * 8, 256, 0, st>>>(canvas_local, canvas_prev_rank + (size_t) g.nxl * plane, plane / 4)));
    fpm_comm_bytes[1] += plane * sizeof(float);
    FPM_CHECK_LAUNCH();
    if (fpm_xbarrier_on(st)) return -1;
    return 0;
}

// readout prologue: my halo plane (index nxl) = plane 0 of the next rank.  Barriers on both sides: every rank has finished the
// transform before anybody pulls, and everybody has pulled before any rank writes its mesh again -- the owner may reuse the block
// for something purely local right away (the PGD potential after the last force component: without the second barrier a slow
// neighbour pulled a plane that was already being overwritten; found with 4 emulated ranks, tests/test_cpu_full_emulation.py).
extern "C" int fpm_halo_fetch_from(const fpm_mesh *m, float *canvas_local, const float *canvas_next_rank)
{
    const FpmGeom &g = m->geom;
    const size_t plane = (size_t) g.n * g.pitch_r;
    cudaStream_t st = comm_stream();
    if (fpm_xbarrier_on(st)) return -1;
    FPM_TIMED(FPM_K_HALO, st, (halo_copy_kernel<<<148 * 4, 256, 0, st>>>(canvas_local + (size_t) g.nxl * plane, canvas_next_rank, plane / 4)));
    FPM_CHECK_LAUNCH();
    fpm_comm_bytes[1] += plane * sizeof(float);
    if (fpm_xbarrier_on(st)) return -1;
    return 0;
}

// The same for the wider windows (linear excepted: it has the CIC reach): every rank keeps the planes outside its slab in a halo
// block, hl planes below the slab followed by hr planes above it (paint.cu: window_plane).  After the deposit my first hr planes
// receive what the previous rank put above its slab and my last hl planes what the next rank put below its own; before the gather
// my halo block is filled with the next rank's first hr planes and the previous rank's last hl planes.
extern "C" int fpm_halo_add_wide_from(const fpm_mesh *m, float *canvas_local, const float *halo_prev_rank, const float *halo_next_rank, int hl, int hr)
{
    const FpmGeom &g = m->geom;
    const size_t plane = (size_t) g.n * g.pitch_r;
    if (hl > g.nxl || hr > g.nxl) { fpm_set_error("halo of %d + %d planes on slabs of %d planes", hl, hr, g.nxl); return -1; }
    cudaStream_t st = comm_stream();
    if (fpm_xbarrier_on(st)) return -1;
    if (hr > 0) FPM_TIMED(FPM_K_HALO, st, (halo_add_kernel<<<148 * 8, 256, 0, st>>>(canvas_local, halo_prev_rank + (size_t) hl * plane, (size_t) hr * plane / 4)));
    if (hl > 0) FPM_TIMED(FPM_K_HALO, st, (halo_add_kernel<<<148 * 8, 256, 0, st>>>(canvas_local + (size_t) (g.nxl - hl) * plane, halo_next_rank, (size_t) hl * plane / 4)));
    fpm_comm_bytes[1] += (size_t) (hl + hr) * plane * sizeof(float);
    FPM_CHECK_LAUNCH();
    if (fpm_xbarrier_on(st)) return -1;
    return 0;
}

extern "C" int fpm_halo_fetch_wide_from(const fpm_mesh *m, float *halo_local, const float *canvas_prev_rank, const float *canvas_next_rank, int hl, int hr)
{
    const FpmGeom &g = m->geom;
    const size_t plane = (size_t) g.n * g.pitch_r;
    if (hl > g.nxl || hr > g.nxl) { fpm_set_error("halo of %d + %d planes on slabs of %d planes", hl, hr, g.nxl); return -1; }
    cudaStream_t st = comm_stream();
    if (fpm_xbarrier_on(st)) return -1;
    if (hl > 0) FPM_TIMED(FPM_K_HALO, st, (halo_copy_kernel<<<148 * 4, 256, 0, st>>>(halo_local, canvas_prev_rank + (size_t) (g.nxl - hl) * plane, (size_t) hl * plane / 4)));
    if (hr > 0) FPM_TIMED(FPM_K_HALO, st, (halo_copy_kernel<<<148 * 4, 256, 0, st>>>(halo_local + (size_t) hl * plane, canvas_next_rank, (size_t) hr * plane / 4)));
    FPM_CHECK_LAUNCH();
    fpm_comm_bytes[1] += (size_t) (hl + hr) * plane * sizeof(float);
    if (fpm_xbarrier_on(st)) return -1;
    return 0;
}

// ------------------------------------------------------------------ particle migration
// owner slab of a position: floor(x * inv_cell) mod N, divided by the slab thickness (pm_pos_to_rank, pmpfft.c:344-368)
// wrap_bad != NULL: fastpm_store_wrap (store.c:447-475) folded in -- same operations as wrap_kernel (particles.cu), positions
// written back only where they changed
__global__ void __launch_bounds__(256) classify_kernel(const FpmGeom g, double *__restrict__ x, long long np,
        int *__restrict__ send_count, int *__restrict__ send_idx, int cap, unsigned char *__restrict__ leaver, int *__restrict__ overflow,
        int *__restrict__ wrap_bad)
{
    long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (; i < np; i += stride) {
        if (wrap_bad) {
            const double L = g.boxsize;
            #pragma unroll
            for (int d = 0; d < 3; d++) {
                const double xi = x[3 * i + d];
                if (xi >= 0 && xi < L) continue;
                const double nwrap = (double) abs((int) (xi / L));
                double x1 = remainder(xi, L);
                while (x1 < 0) x1 += L;
                while (x1 > L) x1 -= L;
                if (nwrap > 10000) atomicExch(wrap_bad, 1);
                x[3 * i + d] = x1;
            }
        }
        int ix = (int) floor(x[3 * i] * g.inv_cellsize);
        ix %= g.n; if (ix < 0) ix += g.n;
        const int dest = ix / g.nxl;
        unsigned char lv = 0;
        if (dest != g.rank) {
            const int slot = atomicAdd(&send_count[dest], 1);
            if (slot < cap) { send_idx[(size_t) dest * cap + slot] = (int) i; lv = 1; }
            else atomicExch(overflow, 1);
        }
        leaver[i] = lv;
    }
}

// pack[dest][slot] = col[send_idx[dest][slot]] for one column of `elsize`-byte elements (elsize multiple of 4)
__global__ void __launch_bounds__(256) pack_kernel(const int *__restrict__ send_idx, int count, const unsigned int *__restrict__ col,
        unsigned int *__restrict__ pack, int words)
{
    long long w = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long) count * words, stride = (long long) gridDim.x * blockDim.x;
    for (; w < total; w += stride) {
        const int s = (int) (w / words), k = (int) (w - (long long) s * words);
        pack[w] = col[(size_t) send_idx[s] * words + k];
    }
}

// holes: leavers among the first np_stay slots; movers: stayers beyond np_stay.  Equal in number.
__global__ void __launch_bounds__(256) holes_kernel(const unsigned char *__restrict__ leaver, long long np, long long np_stay,
        int *__restrict__ holes, int *__restrict__ movers, int *__restrict__ counters)
{
    long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (; i < np; i += stride) {
        if (i < np_stay) { if (leaver[i]) holes[atomicAdd(&counters[0], 1)] = (int) i; }
        else { if (!leaver[i]) movers[atomicAdd(&counters[1], 1)] = (int) i; }
    }
}

__global__ void __launch_bounds__(256) fill_kernel(const int *__restrict__ holes, const int *__restrict__ movers, int count,
        unsigned int *__restrict__ col, int words)
{
    long long w = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long) count * words, stride = (long long) gridDim.x * blockDim.x;
    for (; w < total; w += stride) {
        const int s = (int) (w / words), k = (int) (w - (long long) s * words);
        col[(size_t) holes[s] * words + k] = col[(size_t) movers[s] * words + k];
    }
}

struct Migrate {
    int cap;                 // per-destination capacity (particles)
    int *d_send_count, *d_send_idx, *d_overflow, *d_holes, *d_movers, *d_counters;
    unsigned char *d_leaver; long long leaver_cap;
    unsigned char *d_pack;   // [nranks][cap][row bytes of all migrating columns]
    size_t pack_bytes_per_dest;
};
static Migrate g_mig = { 0 };

extern "C" int fpm_migrate_init(int nranks, int cap, long long np_upper, size_t row_bytes, void *pack)
{
    g_mig.cap = cap;
    g_mig.leaver_cap = np_upper;
    g_mig.pack_bytes_per_dest = (size_t) cap * row_bytes;
    FPM_CUDA_OK(cudaMalloc(&g_mig.d_send_count, sizeof(int) * FPM_MAX_RANKS));
    FPM_CUDA_OK(cudaMalloc(&g_mig.d_send_idx, sizeof(int) * (size_t) nranks * cap));
    FPM_CUDA_OK(cudaMalloc(&g_mig.d_overflow, sizeof(int)));
    FPM_CUDA_OK(cudaMalloc(&g_mig.d_holes, sizeof(int) * (size_t) nranks * cap));
    FPM_CUDA_OK(cudaMalloc(&g_mig.d_movers, sizeof(int) * (size_t) nranks * cap));
    FPM_CUDA_OK(cudaMalloc(&g_mig.d_counters, sizeof(int) * 2));
    FPM_CUDA_OK(cudaMalloc(&g_mig.d_leaver, (size_t) np_upper));
    g_mig.d_pack = (unsigned char *) pack;       // [nranks][cap][row_bytes], in the symmetric arena (host/comm.c)
    return 0;
}

extern "C" void fpm_migrate_destroy(void)
{
    if (!g_mig.d_send_count) return;
    cudaStreamSynchronize(comm_stream());
    cudaFree(g_mig.d_send_count); cudaFree(g_mig.d_send_idx); cudaFree(g_mig.d_overflow); cudaFree(g_mig.d_holes);
    cudaFree(g_mig.d_movers); cudaFree(g_mig.d_counters); cudaFree(g_mig.d_leaver);
    memset(&g_mig, 0, sizeof(g_mig));
}

// step 1: classify; returns the per-destination counts on the host (synchronises the stream)
extern "C" int *fpm_wrap_flag_device(void);
extern "C" int fpm_wrap_flag_fetch(void);
extern "C" int fpm_migrate_classify(const fpm_mesh *m, double *x, int64_t np, int *send_count_host, int wrap, int *overflow_host)
{
    cudaStream_t st = comm_stream();
    const int G = m->geom.nranks;
    FPM_CUDA_OK(cudaMemsetAsync(g_mig.d_send_count, 0, sizeof(int) * FPM_MAX_RANKS, st));
    FPM_CUDA_OK(cudaMemsetAsync(g_mig.d_overflow, 0, sizeof(int), st));
    if (np > g_mig.leaver_cap) { fpm_set_error("migrate: np exceeds the allocated particle capacity"); return -1; }
    if (np > 0) {
        int *bad = wrap ? fpm_wrap_flag_device() : NULL;
        if (wrap && !bad) { fpm_set_error("migrate: no wrap flag"); return -1; }
        FPM_TIMED(FPM_K_MIGRATE, st, (classify_kernel<<<148 * 8, 256, 0, st>>>(m->geom, x, np, g_mig.d_send_count, g_mig.d_send_idx, g_mig.cap, g_mig.d_leaver, g_mig.d_overflow, bad)));
        FPM_CHECK_LAUNCH();
        if (wrap && fpm_wrap_flag_fetch()) return -1;
    }
    int tmp[FPM_MAX_RANKS + 1];
    FPM_CUDA_OK(cudaMemcpyAsync(tmp, g_mig.d_send_count, sizeof(int) * FPM_MAX_RANKS, cudaMemcpyDeviceToHost, st));
    FPM_CUDA_OK(cudaMemcpyAsync(tmp + FPM_MAX_RANKS, g_mig.d_overflow, sizeof(int), cudaMemcpyDeviceToHost, st));
    FPM_CUDA_OK(cudaStreamSynchronize(st));
    // more leavers for one destination than a pack buffer holds: this round moves the first `cap` of them (the kernel marked only
    // those), the caller runs another round for the rest (host/comm.c)
    *overflow_host = tmp[FPM_MAX_RANKS] ? 1 : 0;
    for (int d = 0; d < G; d++) send_count_host[d] = tmp[d] < g_mig.cap ? tmp[d] : g_mig.cap;
    return 0;
}

// step 2: pack one column for every destination: pack region of dest d, column offset `col_off` (bytes per particle before it)
extern "C" int fpm_migrate_pack_column(const fpm_mesh *m, const void *col, int elsize, const int *send_count_host, size_t col_off_bytes)
{
    cudaStream_t st = comm_stream();
    const int G = m->geom.nranks, words = elsize / 4;
    for (int d = 0; d < G; d++) {
        const int cnt = send_count_host[d];
        if (d == m->geom.rank || cnt == 0) continue;
        unsigned char *dst = g_mig.d_pack + (size_t) d * g_mig.pack_bytes_per_dest + col_off_bytes * g_mig.cap;
        FPM_TIMED(FPM_K_MIGRATE, st, (pack_kernel<<<64, 256, 0, st>>>(g_mig.d_send_idx + (size_t) d * g_mig.cap, cnt, (const unsigned int *) col, (unsigned int *) dst, words)));
        FPM_CHECK_LAUNCH();
    }
    return 0;
}

// step 3: find holes and movers (once), then fill one column at a time
extern "C" int fpm_migrate_holes(int64_t np, int64_t np_stay, int *nholes_host)
{
    cudaStream_t st = comm_stream();
    FPM_CUDA_OK(cudaMemsetAsync(g_mig.d_counters, 0, sizeof(int) * 2, st));
    if (np > 0) {
        FPM_TIMED(FPM_K_MIGRATE, st, (holes_kernel<<<148 * 8, 256, 0, st>>>(g_mig.d_leaver, np, np_stay, g_mig.d_holes, g_mig.d_movers, g_mig.d_counters)));
        FPM_CHECK_LAUNCH();
    }
    int c[2];
    FPM_CUDA_OK(cudaMemcpyAsync(c, g_mig.d_counters, sizeof(int) * 2, cudaMemcpyDeviceToHost, st));
    FPM_CUDA_OK(cudaStreamSynchronize(st));
    if (c[0] != c[1]) { fpm_set_error("migrate: %d holes but %d movers", c[0], c[1]); return -1; }
    *nholes_host = c[0];
    return 0;
}

extern "C" int fpm_migrate_fill_column(void *col, int elsize, int nholes)
{
    if (nholes <= 0) return 0;
    FPM_TIMED(FPM_K_MIGRATE, comm_stream(), (fill_kernel<<<64, 256, 0, comm_stream()>>>(g_mig.d_holes, g_mig.d_movers, nholes, (unsigned int *) col, elsize / 4)));
    FPM_CHECK_LAUNCH();
    return 0;
}

// step 4: append what rank `src_rank` packed for me: `count` elements of this column from its pack region
extern "C" int fpm_migrate_append_column(void *col, int elsize, int64_t at, const void *peer_pack_base, int my_rank, int count, size_t col_off_bytes)
{
    if (count <= 0) return 0;
    const unsigned char *src = (const unsigned char *) peer_pack_base + (size_t) my_rank * g_mig.pack_bytes_per_dest + col_off_bytes * g_mig.cap;
    if (fpm_prof_on) fpm_prof_begin(FPM_K_MIGRATE, comm_stream());
    FPM_CUDA_OK(cudaMemcpyAsync((unsigned char *) col + (size_t) at * elsize, src, (size_t) count * elsize, cudaMemcpyDeviceToDevice, comm_stream()));
    fpm_comm_bytes[2] += (unsigned long long) count * elsize;
    if (fpm_prof_on) fpm_prof_end(FPM_K_MIGRATE, comm_stream());
    return 0;
}
