/* fastpm_b200 host layer -- host-scalar collectives of a one-node run through POSIX shared memory.
 *
 * The reference reduces its host scalars (particle counts, total mass, P(k) bins, summary sums) with MPI_Allreduce /
 * MPI_Allgather (e.g. solver.c:421, gravity.c:311, powerspectrum.c:113-115, store.c:531-560).  Here the ranks are the
 * processes of one NVSwitch box, so these small host buffers go through one shared segment instead of the launcher's
 * callbacks (gloo through Python: ~0.3 ms a call on 8 ranks, several calls per step): every rank copies its
 * contribution into its slot, one sense-free barrier on a monotone counter, every rank reduces all slots in rank order
 * (the same order everywhere: identical results on every rank, and run-to-run).  Slots are double-buffered by the parity
 * of the call number, so that one barrier per call is enough: a slot is rewritten two calls later, and the barrier of the
 * call in between cannot be passed before every rank has finished reading.
 *
 * The launcher's callbacks are used once, to agree on the segment and on whether all ranks share a host; they remain
 * the fallback (FASTPM_B200_HOST_COLL=callbacks, ranks on several hosts, or shm_open failing). */
#define _GNU_SOURCE
#include "internal.h"
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#define SHM_MAXR 8
#define SHM_SLOT 65536                 /* bytes per rank and parity; larger payloads go in pieces */

typedef struct {
    volatile uint64_t arrived;         /* total arrivals since the segment was made */
    char pad[120];
    unsigned char slot[2][SHM_MAXR][SHM_SLOT];
} ShmSeg;

static ShmSeg *g_seg = NULL;
static int g_shm_rank = 0, g_shm_size = 1;
static uint64_t g_calls = 0;           /* collectives done so far (the same on every rank) */
static double g_timeout = 900.0;

int fpm_shm_active(void) { return g_seg != NULL; }

static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

static void shm_barrier(void)
{
    g_calls++;
    __atomic_add_fetch(&g_seg->arrived, 1, __ATOMIC_ACQ_REL);
    const uint64_t want = g_calls * (uint64_t) g_shm_size;
    double t0 = 0;
    for (unsigned spin = 0; __atomic_load_n(&g_seg->arrived, __ATOMIC_ACQUIRE) < want; spin++) {
        if ((spin & 63) == 63) sched_yield();
        if ((spin & 0xfffff) == 0xfffff) {
            const double t = now_s();
            if (t0 == 0) t0 = t;
            else if (t - t0 > g_timeout) fastpm_raise(-1, "host collective %llu: rank %d waited %g s for its peers (one of them died?)\n", (unsigned long long) g_calls, g_shm_rank, g_timeout);
        }
    }
}

/* type: 0 double, 1 int64; op: 0 sum, 1 min, 2 max (the callbacks' convention, fastpm_b200_api.h) */
void fpm_shm_allreduce(void *v, int n, int type, int op)
{
    const int per = SHM_SLOT / 8;
    for (int done = 0; done < n; done += per) {
        const int m = n - done < per ? n - done : per;
        const int par = (int) (g_calls & 1);
        memcpy(g_seg->slot[par][g_shm_rank], (char *) v + (size_t) done * 8, (size_t) m * 8);
        shm_barrier();
        if (type == 0) {
            double *out = (double *) v + done;
            for (int i = 0; i < m; i++) {
                double acc = ((const double *) g_seg->slot[par][0])[i];
                for (int r = 1; r < g_shm_size; r++) {
                    const double x = ((const double *) g_seg->slot[par][r])[i];
                    acc = op == 0 ? acc + x : (op == 1 ? (x < acc ? x : acc) : (x > acc ? x : acc));
                }
                out[i] = acc;
            }
        } else {
            int64_t *out = (int64_t *) v + done;
            for (int i = 0; i < m; i++) {
                int64_t acc = ((const int64_t *) g_seg->slot[par][0])[i];
                for (int r = 1; r < g_shm_size; r++) {
                    const int64_t x = ((const int64_t *) g_seg->slot[par][r])[i];
                    acc = op == 0 ? acc + x : (op == 1 ? (x < acc ? x : acc) : (x > acc ? x : acc));
                }
                out[i] = acc;
            }
        }
    }
}

void fpm_shm_allgather(const void *send, int nbytes, void *recv)
{
    for (int done = 0; done < nbytes; done += SHM_SLOT) {
        const int m = nbytes - done < SHM_SLOT ? nbytes - done : SHM_SLOT;
        const int par = (int) (g_calls & 1);
        memcpy(g_seg->slot[par][g_shm_rank], (const char *) send + done, (size_t) m);
        shm_barrier();
        for (int r = 0; r < g_shm_size; r++) memcpy((char *) recv + (size_t) r * nbytes + done, g_seg->slot[par][r], (size_t) m);
    }
}

/* Called by every rank with the launcher's all-gather callback.  Returns 1 when the segment is in use. */
int fpm_shm_setup(int rank, int size, fpm_host_allgather_fn gather, void *userdata)
{
    if (g_seg) { munmap((void *) g_seg, sizeof(ShmSeg)); g_seg = NULL; }
    g_calls = 0;
    const char *mode = getenv("FASTPM_B200_HOST_COLL");
    if (size <= 1 || size > SHM_MAXR || gather == NULL || (mode && strcmp(mode, "callbacks") == 0)) return 0;
    const char *to = getenv("FASTPM_B200_HOST_COLL_TIMEOUT");
    if (to && atof(to) > 0) g_timeout = atof(to);

    /* round 1: host names and the name rank 0 proposes; rank 0 creates the segment before it answers */
    struct { char host[64]; char name[56]; int64_t ok; } mine, all[SHM_MAXR];
    memset(&mine, 0, sizeof(mine));
    gethostname(mine.host, sizeof(mine.host) - 1);
    int fd = -1;
    if (rank == 0) {
        snprintf(mine.name, sizeof(mine.name), "/fastpm_b200_%ld_%lx", (long) getpid(), (unsigned long) (now_s() * 1e6));
        fd = shm_open(mine.name, O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd >= 0 && ftruncate(fd, sizeof(ShmSeg)) != 0) { close(fd); shm_unlink(mine.name); fd = -1; }
        mine.ok = fd >= 0;
    }
    gather(&mine, (int) sizeof(mine), all, userdata);
    int usable = all[0].ok != 0;
    for (int r = 1; r < size; r++) if (strncmp(all[r].host, all[0].host, sizeof(mine.host)) != 0) usable = 0;

    /* round 2: everybody maps it (or reports that it could not) */
    ShmSeg *seg = NULL;
    if (usable) {
        if (rank != 0) fd = shm_open(all[0].name, O_RDWR, 0600);
        if (fd >= 0) {
            void *p = mmap(NULL, sizeof(ShmSeg), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
            if (p != MAP_FAILED) seg = (ShmSeg *) p;
        }
    }
    if (fd >= 0) close(fd);
    mine.ok = seg != NULL;
    gather(&mine, (int) sizeof(mine), all, userdata);
    if (rank == 0 && all[0].name[0] && fd >= 0) shm_unlink(all[0].name);      /* the mappings keep it alive; nothing is left behind */
    int everybody = 1;
    for (int r = 0; r < size; r++) if (!all[r].ok) everybody = 0;
    if (!everybody) {
        if (seg) munmap(seg, sizeof(ShmSeg));
        return 0;
    }
    g_seg = seg; g_shm_rank = rank; g_shm_size = size;          /* ftruncate zero-filled it: arrived == 0 */
    return 1;
}

/* ---- a one-node run without a launcher: the parent of the rank processes makes the segment before it forks them
 * (fastpm_b200_run -n N, lua_front/run.c), every rank attaches by name; no callbacks are involved at all */
int fastpm_b200_local_segment_create(char *name_out, size_t cap)
{
    snprintf(name_out, cap, "/fastpm_b200_%ld_%lx", (long) getpid(), (unsigned long) (now_s() * 1e6));
    const int fd = shm_open(name_out, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0) return -1;
    if (ftruncate(fd, sizeof(ShmSeg)) != 0) { close(fd); shm_unlink(name_out); return -1; }
    close(fd);
    return 0;
}
int fastpm_b200_local_segment_unlink(const char *name) { return shm_unlink(name); }

int fpm_shm_attach(int rank, int size, const char *name)
{
    if (g_seg) { munmap((void *) g_seg, sizeof(ShmSeg)); g_seg = NULL; }
    g_calls = 0;
    if (size < 1 || size > SHM_MAXR) return -1;
    const int fd = shm_open(name, O_RDWR, 0600);
    if (fd < 0) return -1;
    void *p = mmap(NULL, sizeof(ShmSeg), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return -1;
    g_seg = (ShmSeg *) p; g_shm_rank = rank; g_shm_size = size;
    return 0;
}
