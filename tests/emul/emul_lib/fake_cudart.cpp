// TEST INFRASTRUCTURE ONLY: definitions of the CUDA runtime entry points the library calls (see fake_cudart.h).
#include <cuda_runtime.h>
#include <cuda.h>
#include <fcntl.h>
#include <signal.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

unsigned char *fpm_emul_dyn_smem = nullptr;          // the one definition for all translation units (cuda_emul.h)

static std::map<void *, size_t> g_blocks;
static std::mutex g_lock;
static size_t g_used = 0;
static const size_t g_total = (size_t) 64 << 30;
static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// Large blocks (the symmetric arena of a multi-process run) live in POSIX shared memory so that the other ranks -- other
// processes on this host -- can map them: the stand-in for CUDA IPC over NVLink.
static const size_t SHM_MIN = (size_t) 16 << 20;
struct ShmInfo { char name[48]; size_t size; };
static std::map<void *, ShmInfo> g_shm;
static int g_dev = 0, g_shm_seq = 0;
static void shm_cleanup()
{
    for (auto &kv : g_shm) shm_unlink(kv.second.name);
}
static void shm_cleanup_on_abort(int sig)          // fastpm_raise aborts: do not leave the segments behind
{
    shm_cleanup();
    signal(sig, SIG_DFL);
    raise(sig);
}
static CUresult fake_cuMemGetAddressRange(CUdeviceptr *base, size_t *size, CUdeviceptr p)
{
    std::lock_guard<std::mutex> l(g_lock);
    auto it = g_blocks.upper_bound((void *) p);
    if (it == g_blocks.begin()) return CUDA_ERROR_INVALID_VALUE;
    --it;
    if ((const char *) p >= (const char *) it->first + it->second) return CUDA_ERROR_INVALID_VALUE;
    *base = (CUdeviceptr) it->first; *size = it->second;
    return CUDA_SUCCESS;
}

extern "C" {
cudaError_t cudaGetDeviceCount(int *n) { *n = 8; return cudaSuccess; }
cudaError_t cudaSetDevice(int d) { if (d < 0 || d >= 8) return cudaErrorInvalidDevice; g_dev = d; return cudaSuccess; }
cudaError_t cudaGetDevice(int *d) { *d = g_dev; return cudaSuccess; }
cudaError_t cudaGetDriverEntryPoint(const char *symbol, void **fn, unsigned long long, cudaDriverEntryPointQueryResult *q)
{
    const bool ok = !strcmp(symbol, "cuMemGetAddressRange");
    *fn = ok ? (void *) fake_cuMemGetAddressRange : NULL;
    if (q) *q = ok ? cudaDriverEntryPointSuccess : cudaDriverEntryPointSymbolNotFound;
    return ok ? cudaSuccess : cudaErrorSymbolNotFound;
}
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p)
{
    std::lock_guard<std::mutex> l(g_lock);
    auto it = g_shm.find(p);
    if (it == g_shm.end()) return cudaErrorInvalidValue;      // only shared-memory blocks can be exported
    memset(h, 0, sizeof(*h));
    memcpy(h, &it->second, sizeof(ShmInfo));
    return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned int)
{
    ShmInfo info;
    memcpy(&info, &h, sizeof(info));
    const int fd = shm_open(info.name, O_RDWR, 0600);
    if (fd < 0) return cudaErrorInvalidValue;
    void *q = mmap(NULL, info.size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (q == MAP_FAILED) return cudaErrorMemoryAllocation;
    *p = q;
    return cudaSuccess;
}
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr a, int) { *v = a == cudaDevAttrMultiProcessorCount ? 4 : 0; return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA runtime error"; }
cudaError_t cudaMemGetInfo(size_t *fr, size_t *tot) { *tot = g_total; *fr = g_total - g_used; return cudaSuccess; }
cudaError_t cudaMalloc(void **p, size_t n)
{
    void *q = NULL;
    ShmInfo info;
    memset(&info, 0, sizeof(info));
    const char *e = getenv("FPM_EMUL_SHM_MIN_MB");          // tests of a tiny arena lower the threshold
    if (n >= (e ? (size_t) atoi(e) << 20 : SHM_MIN)) {
        std::lock_guard<std::mutex> l(g_lock);
        snprintf(info.name, sizeof(info.name), "/fpm_emul_%d_%d", (int) getpid(), g_shm_seq++);
        info.size = n;
        const int fd = shm_open(info.name, O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0 || ftruncate(fd, (off_t) n) != 0) { if (fd >= 0) { close(fd); shm_unlink(info.name); } return cudaErrorMemoryAllocation; }
        q = mmap(NULL, n, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        if (q == MAP_FAILED) { shm_unlink(info.name); return cudaErrorMemoryAllocation; }
        if (g_shm.empty()) { atexit(shm_cleanup); signal(SIGABRT, shm_cleanup_on_abort); signal(SIGTERM, shm_cleanup_on_abort); }
        g_shm[q] = info;
    } else {
        if (posix_memalign(&q, 1024, n ? n : 1) != 0) return cudaErrorMemoryAllocation;
        memset(q, 0xA5, n);                            // uninitialised device memory is not zero
    }
    std::lock_guard<std::mutex> l(g_lock);
    g_blocks[q] = n; g_used += n; *p = q;
    return cudaSuccess;
}
cudaError_t cudaFree(void *p)
{
    if (!p) return cudaSuccess;
    std::lock_guard<std::mutex> l(g_lock);
    auto it = g_blocks.find(p);
    if (it == g_blocks.end()) return cudaErrorInvalidValue;
    auto sh = g_shm.find(p);
    if (sh != g_shm.end()) { munmap(p, it->second); shm_unlink(sh->second.name); g_shm.erase(sh); }
    else free(p);
    g_used -= it->second; g_blocks.erase(it);
    return cudaSuccess;
}
cudaError_t cudaMallocAsync(void **p, size_t n, cudaStream_t) { return cudaMalloc(p, n); }
cudaError_t cudaFreeAsync(void *p, cudaStream_t) { return cudaFree(p); }
cudaError_t cudaHostAlloc(void **p, size_t n, unsigned int) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t)
{ for (size_t i = 0; i < h; i++) memmove((char *) d + i * dp, (const char *) s + i * sp, w); return cudaSuccess; }
cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned int) { *s = (cudaStream_t) malloc(8); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned int) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (cudaEvent_t) calloc(1, sizeof(double)); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned int) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { *(double *) e = now_ms(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float) (*(double *) b - *(double *) a); return cudaSuccess; }
cudaError_t cudaFuncSetAttribute(const void *, cudaFuncAttribute, int) { return cudaSuccess; }
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *p)
{
    memset(a, 0, sizeof(*a));
    std::lock_guard<std::mutex> l(g_lock);
    auto it = g_blocks.upper_bound((void *) p);
    bool dev = false;
    if (it != g_blocks.begin()) { --it; dev = (const char *) p < (const char *) it->first + it->second; }
    a->type = dev ? cudaMemoryTypeDevice : cudaMemoryTypeUnregistered;
    a->devicePointer = dev ? (void *) p : NULL; a->hostPointer = dev ? NULL : (void *) p;
    return cudaSuccess;
}
}
