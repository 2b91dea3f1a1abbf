/* fastpm_b200 host layer -- the communicator standing in for MPI_Comm: one process per GPU, x-slab decomposition.
 *
 * Host scalars (counts, sums, P(k) bins) travel through two callbacks supplied by the launcher -- all-reduce and
 * all-gather on host buffers, which bench.py / the tests implement with torch.distributed (gloo or NCCL).  Mesh and
 * particle data never touch the host: peers' buffers are mapped with CUDA IPC and the kernels read / write them over
 * NVLink (csrc/comm.cu).  This file keeps the registry of mapped buffers and sequences the exchanges:
 *   fpm_dist_r2c / fpm_dist_c2r   the slab transforms, their transposing pass storing into the owner of each plane
 *   fpm_halo_add / fpm_halo_fetch one mesh plane to / from the x-neighbour (replaces pm_ghosts_*, pmghosts.c:112-307)
 *   fastpm_store_decompose        particle migration (store.c:486-657) for FastPMTargetPM
 */
#include "internal.h"

/* device side, csrc/comm.cu */
int fpm_ipc_get_handle(void *dev_ptr, void *handle64, uint64_t *offset);
void *fpm_ipc_open(const void *handle64, uint64_t offset);
int fpm_xbarrier_init(int nranks, int rank, void *local_flags);
int fpm_xbarrier_set_peers(void *const *peer_flag_ptrs);
int fpm_xbarrier(void);
int fpm_mesh_set_stage(fpm_mesh *m, float *stage);
int fpm_mesh_set_stage2(fpm_mesh *m, float *stage);
int fpm_c2r_dist_begin(fpm_mesh *m, const float *cplx, float *const *real_peers, const fpm_transfer *kernel, int set);
int fpm_c2r_dist_finish(fpm_mesh *m, float *const *real_peers, int set);
int fpm_r2c_dist(fpm_mesh *m, float *real, float *const *cplx_peers, double scale);
int fpm_c2r_dist(fpm_mesh *m, const float *cplx, float *const *real_peers, const fpm_transfer *kernel);
int fpm_halo_add_from(const fpm_mesh *m, float *canvas_local, const float *canvas_prev_rank);
int fpm_halo_fetch_from(const fpm_mesh *m, float *canvas_local, const float *canvas_next_rank);
int fpm_halo_add_wide_from(const fpm_mesh *m, float *canvas_local, const float *halo_prev_rank, const float *halo_next_rank, int hl, int hr);
int fpm_halo_fetch_wide_from(const fpm_mesh *m, float *halo_local, const float *canvas_prev_rank, const float *canvas_next_rank, int hl, int hr);
int fpm_migrate_init(int nranks, int cap, long long np_upper, size_t row_bytes, void *pack);
int fpm_migrate_classify(const fpm_mesh *m, double *x, int64_t np, int *send_count_host, int wrap, int *overflow_host);
int fpm_migrate_pack_column(const fpm_mesh *m, const void *col, int elsize, const int *send_count_host, size_t col_off_bytes);
int fpm_migrate_holes(int64_t np, int64_t np_stay, int *nholes_host);
int fpm_migrate_fill_column(void *col, int elsize, int nholes);
int fpm_migrate_append_column(void *col, int elsize, int64_t at, const void *peer_pack_base, int my_rank, int count, size_t col_off_bytes);

#define MAXR 8
static int g_rank = 0, g_size = 1;
static fpm_host_allreduce_fn g_allreduce = NULL;
static fpm_host_allgather_fn g_allgather = NULL;
static void *g_cb_data = NULL;

int fpm_comm_rank(MPI_Comm comm) { (void) comm; return g_rank; }
int fpm_comm_size(MPI_Comm comm) { (void) comm; return g_size; }

/* host/shmcoll.c: the same collectives through a shared segment when all ranks sit on one host */
int fpm_shm_setup(int rank, int size, fpm_host_allgather_fn gather, void *userdata);
int fpm_shm_active(void);
static int g_callbacks_only = 0;            /* self test of the callbacks */
#define SHM_ON() (!g_callbacks_only && fpm_shm_active())
void fpm_shm_allreduce(void *v, int n, int type, int op);
void fpm_shm_allgather(const void *send, int nbytes, void *recv);

int fastpm_b200_host_collectives_shared(void) { return g_size > 1 && SHM_ON(); }

void fpm_comm_allreduce_double(MPI_Comm comm, double *v, int n, int op)
{
    (void) comm;
    if (g_size == 1) return;
    if (SHM_ON()) { fpm_shm_allreduce(v, n, 0, op); return; }
    if (!g_allreduce) fastpm_raise(-1, "multi-rank run without an all-reduce callback\n");
    g_allreduce(v, n, 0, op, g_cb_data);
}
void fpm_comm_allreduce_i64(MPI_Comm comm, int64_t *v, int n, int op)
{
    (void) comm;
    if (g_size == 1) return;
    if (SHM_ON()) { fpm_shm_allreduce(v, n, 1, op); return; }
    if (!g_allreduce) fastpm_raise(-1, "multi-rank run without an all-reduce callback\n");
    g_allreduce(v, n, 1, op, g_cb_data);
}
void fpm_comm_barrier(MPI_Comm comm) { int64_t z = 0; fpm_comm_allreduce_i64(comm, &z, 1, 0); }

static void allgather(const void *send, int nbytes, void *recv)
{
    if (g_size == 1) { memcpy(recv, send, nbytes); return; }
    if (SHM_ON()) { fpm_shm_allgather(send, nbytes, recv); return; }
    if (!g_allgather) fastpm_raise(-1, "multi-rank run without an all-gather callback\n");
    g_allgather(send, nbytes, recv, g_cb_data);
}

/* ------------------------------------------------------------------ the symmetric arena (host/support.c)
 * Every device buffer of a multi-GPU run lives in one arena per process; arena_peer[r] is rank r's arena as mapped
 * into this process with CUDA IPC at start-up.  All ranks allocate the same sizes in the same order, so the peer
 * copies of a local buffer sit at the same offset: no per-call handle exchange, nothing can go stale. */
int fastpm_b200_arena_init(size_t bytes);
void *fastpm_b200_arena_base(void);
size_t fastpm_b200_arena_size(void);
int fastpm_b200_arena_contains(const void *p);
static char *arena_peer[MAXR];

static void peers_of(void *local, void *peers[MAXR])
{
    if (!fastpm_b200_arena_contains(local))
        fastpm_raise(-1, "multi-GPU exchange on a buffer (%p) that was not allocated from the symmetric arena (pm_alloc / fastpm_memory_alloc)\n", local);
    const size_t off = (size_t) ((char *) local - (char *) fastpm_b200_arena_base());
    if (getenv("FASTPM_B200_CHECK_SYMMETRY")) {
        uint64_t mine = off, all[MAXR];
        allgather(&mine, 8, all);
        for (int r = 0; r < g_size; r++)
            if (all[r] != mine) fastpm_raise(-1, "asymmetric allocation: offset %zu here, %zu on rank %d\n", off, (size_t) all[r], r);
    }
    for (int r = 0; r < g_size; r++) peers[r] = arena_peer[r] + off;
}

/* ------------------------------------------------------------------ set-up by the launcher */
/* the collectives only (no device arena): for host-side tools of a multi-process run, e.g. the snapshot writers of host/io.c */
void fastpm_b200_comm_init_host(int rank, int size, fpm_host_allreduce_fn allreduce, fpm_host_allgather_fn allgather_cb, void *userdata)
{
    if (size > MAXR) fastpm_raise(-1, "at most %d slabs (one node) are supported\n", MAXR);
    g_rank = rank; g_size = size; g_allreduce = allreduce; g_allgather = allgather_cb; g_cb_data = userdata;
    fpm_shm_setup(rank, size, allgather_cb, userdata);
}

static void comm_init_device(int rank, int size);
static void *g_xbarrier_flags = NULL;

/* the end of a multi-rank program that goes on to libfastpm_cleanup (which insists that every block has been returned): every rank
 * has finished its device work, then the communicator's own block goes back */
void fastpm_b200_comm_finalize(void)
{
    if (g_size <= 1 || !g_xbarrier_flags) return;
    FPM_MUST(fpm_sync());
    fpm_comm_barrier(MPI_COMM_WORLD);
    fastpm_memory_free(_libfastpm_get_gmem(), g_xbarrier_flags);
    g_xbarrier_flags = NULL;
}

void fastpm_b200_comm_init(int rank, int size, fpm_host_allreduce_fn allreduce, fpm_host_allgather_fn allgather_cb, void *userdata)
{
    libfastpm_init();
    if (size > MAXR) fastpm_raise(-1, "at most %d slabs (one node) are supported\n", MAXR);
    g_rank = rank; g_size = size; g_allreduce = allreduce; g_allgather = allgather_cb; g_cb_data = userdata;
    fpm_shm_setup(rank, size, allgather_cb, userdata);
    if (size == 1) return;
    comm_init_device(rank, size);
}

/* The same for the ranks of a one-node run that have no launcher to supply callbacks (fastpm_b200_run -n N): they attach to the
 * shared segment their parent process made (fastpm_b200_local_segment_create) and exchange everything through it. */
int fpm_shm_attach(int rank, int size, const char *name);
void fastpm_b200_comm_init_local(int rank, int size, const char *segment)
{
    libfastpm_init();
    if (size > MAXR) fastpm_raise(-1, "at most %d slabs (one node) are supported\n", MAXR);
    g_rank = rank; g_size = size; g_allreduce = NULL; g_allgather = NULL; g_cb_data = NULL;
    if (size == 1) return;
    if (fpm_shm_attach(rank, size, segment) != 0) fastpm_raise(-1, "rank %d cannot attach to the shared segment %s\n", rank, segment);
    comm_init_device(rank, size);
}

static void comm_init_device(int rank, int size)
{
    /* arena: FASTPM_B200_ARENA_GB, else 85 % of what is free now; the smallest over ranks so that offsets stay in range everywhere */
    size_t free_b = 0, total_b = 0;
    if (fpm_device_mem_info(&free_b, &total_b) != 0) fastpm_raise(-1, "%s\n", fpm_last_error());
    const char *e = getenv("FASTPM_B200_ARENA_GB"), *ef = getenv("FASTPM_B200_ARENA_FRAC");
    double arena_frac = ef ? atof(ef) : 0.85;             /* capacity runs (BASELINE configs[4]) raise it: little else is allocated */
    if (!(arena_frac > 0.05 && arena_frac <= 0.97)) arena_frac = 0.85;
    int64_t want = e ? (int64_t) (atof(e) * 1073741824.0) : (int64_t) (arena_frac * free_b);
    fpm_comm_allreduce_i64(MPI_COMM_WORLD, &want, 1, 1);
    if (fastpm_b200_arena_init((size_t) want) != 0) fastpm_raise(-1, "arena of %lld bytes: %s\n", (long long) want, fpm_last_error());
    unsigned char h[72], hall[MAXR * 72];
    uint64_t off = 0;
    if (fpm_ipc_get_handle(fastpm_b200_arena_base(), h, &off) != 0) fastpm_raise(-1, "%s\n", fpm_last_error());
    memcpy(h + 64, &off, 8);
    allgather(h, 72, hall);
    for (int r = 0; r < size; r++) {
        if (r == rank) { arena_peer[r] = fastpm_b200_arena_base(); continue; }
        uint64_t roff; memcpy(&roff, hall + 72 * r + 64, 8);
        arena_peer[r] = fpm_ipc_open(hall + 72 * r, roff);
        if (!arena_peer[r]) fastpm_raise(-1, "mapping rank %d's arena: %s\n", r, fpm_last_error());
    }
    void *flags = fastpm_memory_alloc(_libfastpm_get_gmem(), "xbarrier flags", 4096, FASTPM_MEMORY_FLOATING), *peers[MAXR];
    g_xbarrier_flags = flags;
    if (fpm_xbarrier_init(size, rank, flags) != 0) fastpm_raise(-1, "%s\n", fpm_last_error());
    peers_of(flags, peers);
    if (fpm_xbarrier_set_peers(peers) != 0) fastpm_raise(-1, "%s\n", fpm_last_error());
    fpm_comm_barrier(MPI_COMM_WORLD);
}

/* ------------------------------------------------------------------ mesh exchanges */
void fpm_halo_add(PM *pm, FastPMFloat *canvas)
{
    void *peers[MAXR];
    if (pm->whalo) {            /* a window wider than CIC: what the neighbours put into their halo blocks (pm->whalo, host/gravity.c) */
        peers_of(pm->whalo, peers);
        FPM_MUST(fpm_halo_add_wide_from(pm->mesh, canvas, peers[(g_rank - 1 + g_size) % g_size], peers[(g_rank + 1) % g_size], pm->whl, pm->whr));
        return;
    }
    peers_of(canvas, peers);
    FPM_MUST(fpm_halo_add_from(pm->mesh, canvas, peers[(g_rank - 1 + g_size) % g_size]));
}

void fpm_halo_fetch(PM *pm, FastPMFloat *canvas)
{
    void *peers[MAXR];
    peers_of(canvas, peers);
    if (pm->whalo) {
        FPM_MUST(fpm_halo_fetch_wide_from(pm->mesh, pm->whalo, peers[(g_rank - 1 + g_size) % g_size], peers[(g_rank + 1) % g_size], pm->whl, pm->whr));
        return;
    }
    FPM_MUST(fpm_halo_fetch_from(pm->mesh, canvas, peers[(g_rank + 1) % g_size]));
}

/* one extra mesh per PM: the slab transposes are staged locally and pushed by the copy engines (csrc/fft.cu) */
size_t fastpm_b200_arena_largest_free(void);
static void ensure_stage(PM *pm)
{
    if (pm->stage || pm->stage_off || getenv("FASTPM_B200_NO_STAGE")) return;
    /* capacity runs (BASELINE.json config 5: 4096^3 mesh on 8 GPUs) have no room for a third mesh: the transposes then store
     * straight into the peers.  Every rank sees the same arena state, so the decision is the same everywhere. */
    const size_t need = sizeof(FastPMFloat) * pm->allocsize;
    if (fastpm_b200_arena_largest_free() < need + need / 8) { pm->stage_off = 1; return; }
    pm->stage = fastpm_memory_alloc(pm->mem, "FFT transpose staging", need, FASTPM_MEMORY_FLOATING);
    FPM_MUST(fpm_mesh_set_stage(pm->mesh, pm->stage));
}

/* Pipelined inverse transforms (host/gravity.c): the second staging mesh, taken once and only when every rank has room for it
 * beside one more canvas (all ranks see the same arena state, the decision is the same everywhere).  Returns 1 when available. */
int fpm_dist_pipeline_ready(PM *pm)
{
    ensure_stage(pm);
    if (!pm->stage || getenv("FASTPM_B200_NO_PIPELINE")) return 0;
    const size_t need = sizeof(FastPMFloat) * pm->allocsize;
    if (!pm->stage2) {
        if (pm->stage2_off) return 0;
        if (fastpm_b200_arena_largest_free() < 2 * (need + need / 8)) { pm->stage2_off = 1; return 0; }      /* stage 2 and the second canvas */
        pm->stage2 = fastpm_memory_alloc(pm->mem, "FFT transpose staging (second set)", need, FASTPM_MEMORY_FLOATING);
        FPM_MUST(fpm_mesh_set_stage2(pm->mesh, pm->stage2));
    }
    return fastpm_b200_arena_largest_free() >= need + need / 8;                                               /* the second canvas */
}

void fpm_dist_c2r_begin(PM *pm, const FastPMFloat *cplx, FastPMFloat *real, const fpm_transfer *kernel, int set)
{
    void *peers[MAXR];
    if (real == cplx) fastpm_raise(-1, "distributed c2r is out of place\n");
    peers_of(real, peers);
    FPM_MUST(fpm_c2r_dist_begin(pm->mesh, cplx, (float *const *) peers, kernel, set));
}

void fpm_dist_c2r_finish(PM *pm, FastPMFloat *real, int set)
{
    void *peers[MAXR];
    peers_of(real, peers);
    FPM_MUST(fpm_c2r_dist_finish(pm->mesh, (float *const *) peers, set));
}

void fpm_dist_r2c(PM *pm, FastPMFloat *real, FastPMFloat *cplx, double scale)
{
    void *peers[MAXR];
    ensure_stage(pm);
    if (real == cplx) fastpm_raise(-1, "distributed r2c is out of place\n");
    peers_of(cplx, peers);
    FPM_MUST(fpm_r2c_dist(pm->mesh, real, (float *const *) peers, scale));
}

void fpm_dist_c2r(PM *pm, const FastPMFloat *cplx, FastPMFloat *real, const fpm_transfer *kernel)
{
    void *peers[MAXR];
    ensure_stage(pm);
    if (real == cplx) fastpm_raise(-1, "distributed c2r is out of place\n");
    peers_of(real, peers);
    FPM_MUST(fpm_c2r_dist(pm->mesh, cplx, (float *const *) peers, kernel));
}

/* the calls the force / IC code uses: one rank or many */
void fpm_mesh_r2c(PM *pm, FastPMFloat *real, FastPMFloat *cplx, double scale)
{
    if (pm->NTask > 1) fpm_dist_r2c(pm, real, cplx, scale);
    else FPM_MUST(fpm_r2c(pm->mesh, real, cplx, scale));
}
void fpm_mesh_c2r(PM *pm, const FastPMFloat *cplx, FastPMFloat *real, const fpm_transfer *kernel)
{
    if (pm->NTask > 1) fpm_dist_c2r(pm, cplx, real, kernel);
    else FPM_MUST(fpm_c2r(pm->mesh, cplx, real, kernel));
}
/* gather from a real field that was just transformed: the +1 plane comes from the x-neighbour */
void fpm_mesh_readout(PM *pm, FastPMFloat *canvas, const double *x, int64_t np, float *out, int stride, double prescale)
{
    if (pm->NTask > 1) fpm_halo_fetch(pm, canvas);
    FPM_MUST(fpm_readout(pm->mesh, canvas, x, np, out, stride, prescale));
}

/* ------------------------------------------------------------------ particle migration */
static void *pack_local = NULL, *pack_peers[MAXR];
static int mig_cap = 0;
static size_t mig_row_bytes = 0;

void fpm_migrate_destroy(void);
/* called when the particle store goes away (fastpm_solver_destroy): the pack buffers are sized for that store */
void fpm_comm_release_migration(void)
{
    if (!pack_local) return;
    fpm_comm_barrier(MPI_COMM_WORLD);           /* nobody is still pulling from our pack buffers */
    fpm_migrate_destroy();
    fastpm_memory_free(_libfastpm_get_gmem(), pack_local);
    pack_local = NULL; mig_cap = 0; mig_row_bytes = 0;
}

typedef struct { void *ptr; int elsize; } MigCol;

/* the largest number of exchange rounds any fastpm_store_decompose of this process needed (diagnostics, tests) */
static int g_migrate_rounds_max = 0;
int fastpm_b200_migrate_rounds_max(void) { return g_migrate_rounds_max; }

/* set by fastpm_decompose (host/solver.c) around its call: the force evaluation that follows recomputes ACC, so the column need not
 * travel.  Any other caller of the public fastpm_store_decompose gets every column moved, like the reference (store.c:302-323). */
int fpm_decompose_skip_acc = 0;

static int migrating_columns(FastPMStore *p, MigCol *cols)
{
    int n = 0;
    for (int ci = 0; ci < 32; ci++) {
        if (!p->columns[ci]) continue;
        if (fpm_decompose_skip_acc && p->_column_info[ci].attribute == COLUMN_ACC) continue;
        if (p->_column_info[ci].elsize % 4 != 0) fastpm_raise(-1, "column %s cannot migrate (element size %zu)\n", p->_column_info[ci].name, p->_column_info[ci].elsize);
        cols[n].ptr = p->columns[ci]; cols[n].elsize = (int) p->_column_info[ci].elsize; n++;
    }
    return n;
}

int fastpm_store_decompose(FastPMStore *p, fastpm_store_target_func target_func, void *data, MPI_Comm comm)
{
    (void) comm;
    if (g_size == 1) return 0;                            /* one slab owns every particle */
    fpm_store_flush(p);
    if (target_func != (fastpm_store_target_func) FastPMTargetPM)
        fastpm_raise(-1, "fastpm_b200: fastpm_store_decompose supports the PM target (x-slabs) only\n");
    PM *pm = data;
    MigCol cols[32];
    const int ncol = migrating_columns(p, cols);
    size_t row = 0;
    for (int j = 0; j < ncol; j++) row += cols[j].elsize;
    if (!pack_local) {
        /* pack buffers: g_size destinations x this fraction of np_upper.  A step moves ~1 % of a slab across each face; more than a
         * buffer holds goes in a further round (below), so the fraction only trades memory against rounds: 8 % on 2 GPUs, 4 % on 4,
         * 2 % on 8 (16 % of a store in total in every case) */
        double frac = 0.16 / g_size;
        if (frac > 0.08) frac = 0.08;
        const char *e = getenv("FASTPM_B200_MIGRATE_FRAC");
        if (e) frac = atof(e);
        mig_cap = (int) (frac * p->np_upper);
        if (mig_cap < 4096) mig_cap = 4096;
        const char *ec = getenv("FASTPM_B200_MIGRATE_CAP");            /* tests: a buffer of a few particles forces several rounds */
        if (ec && atoi(ec) > 0) mig_cap = atoi(ec);
        if ((size_t) mig_cap > p->np_upper) mig_cap = (int) p->np_upper;
        /* sized for every allocated column (the solver's own calls leave ACC behind, other callers move it too) */
        size_t row_all = 0;
        for (int ci = 0; ci < 32; ci++) if (p->columns[ci]) row_all += p->_column_info[ci].elsize;
        mig_row_bytes = row_all;
        pack_local = fastpm_memory_alloc(p->mem, "migration pack buffers", (size_t) g_size * mig_cap * row_all, FASTPM_MEMORY_FLOATING);
        if (fpm_migrate_init(g_size, mig_cap, (long long) p->np_upper, row_all, pack_local) != 0) fastpm_raise(-1, "%s\n", fpm_last_error());
        peers_of(pack_local, pack_peers);
    }
    if (row > mig_row_bytes) fastpm_raise(-1, "fastpm_b200: particle columns were added between decompositions\n");

    /* The reference moves any fraction of the store in one exchange (store.c:486-657) and fails only when a rank would hold more
     * than np_upper.  The pack buffers here hold `mig_cap` particles per destination: when more than that leave for one slab
     * (thin slabs, a restart from a snapshot that every rank read an even share of) the exchange runs in rounds -- each round
     * moves up to mig_cap particles per destination, the rest stay put and are classified again.  Every rank takes the same
     * number of rounds (the overflow flag is all-reduced), so nobody waits in a collective the others never enter. */
    int wrap = (fpm_pending_wrap == p);                    /* fastpm_decompose left the periodic wrap to the classification pass */
    if (wrap) fpm_pending_wrap = NULL;
    for (int round = 0; ; round++) {
        if (round + 1 > g_migrate_rounds_max) g_migrate_rounds_max = round + 1;
        int send[MAXR], all[MAXR * MAXR], overflow = 0;
        memset(send, 0, sizeof(send));
        if (fpm_migrate_classify(pm->mesh, (double *) p->x, (int64_t) p->np, send, wrap, &overflow) != 0) fastpm_raise(-1, "%s\n", fpm_last_error());
        wrap = 0;
        allgather(send, sizeof(int) * MAXR, all);
        int64_t nsend = 0, nrecv = 0;
        for (int r = 0; r < g_size; r++) { if (r != g_rank) { nsend += send[r]; nrecv += all[r * MAXR + g_rank]; } }
        const int64_t np_stay = (int64_t) p->np - nsend;
        int64_t flags[2] = { (np_stay + nrecv > (int64_t) p->np_upper) ? 1 : 0, overflow };
        fpm_comm_allreduce_i64(MPI_COMM_WORLD, flags, 2, 2);
        if (flags[0]) return -1;                           /* "Out of particle storage space", solver.c:585-590 */

        size_t off = 0;
        for (int j = 0; j < ncol; j++) {
            FPM_MUST(fpm_migrate_pack_column(pm->mesh, cols[j].ptr, cols[j].elsize, send, off));
            off += cols[j].elsize;
        }
        int nholes = 0;
        if (fpm_migrate_holes((int64_t) p->np, np_stay, &nholes) != 0) fastpm_raise(-1, "%s\n", fpm_last_error());
        for (int j = 0; j < ncol; j++) FPM_MUST(fpm_migrate_fill_column(cols[j].ptr, cols[j].elsize, nholes));
        FPM_MUST(fpm_xbarrier());                          /* every rank has packed */
        int64_t at = np_stay;
        for (int r = 0; r < g_size; r++) {
            if (r == g_rank) continue;
            const int cnt = all[r * MAXR + g_rank];
            off = 0;
            for (int j = 0; j < ncol; j++) {
                FPM_MUST(fpm_migrate_append_column(cols[j].ptr, cols[j].elsize, at, pack_peers[r], g_rank, cnt, off));
                off += cols[j].elsize;
            }
            at += cnt;
        }
        FPM_MUST(fpm_xbarrier());                          /* every rank has pulled: pack buffers may be reused */
        p->np = (size_t) at;
        if (!flags[1]) break;                              /* nobody had more leavers than a pack buffer holds */
        if (round > 4096) fastpm_raise(-1, "fastpm_b200: particle migration does not converge\n");
    }
    return 0;
}


/* ------------------------------------------------------------------ host-only self test of the callback plumbing
 * (no device: used by the world_size-2 gloo test on CPU).  Returns 0 when the collectives behave as the host layer
 * assumes: sum / min / max all-reduce on doubles and int64, rank-ordered all-gather of fixed-size records. */
int fastpm_b200_comm_selftest(int rank, int size, fpm_host_allreduce_fn allreduce, fpm_host_allgather_fn allgather_cb, void *userdata)
{
    int save_rank = g_rank, save_size = g_size;
    fpm_host_allreduce_fn sa = g_allreduce; fpm_host_allgather_fn sg = g_allgather; void *sd = g_cb_data;
    g_rank = rank; g_size = size; g_allreduce = allreduce; g_allgather = allgather_cb; g_cb_data = userdata;
    int bad = 0;
    for (int pass = 0; pass < 2; pass++) {
        /* pass 0: the launcher's callbacks; pass 1: the shared segment of host/shmcoll.c (set up through those callbacks) */
        g_callbacks_only = pass == 0;
        if (pass == 1 && !(size > 1 && fpm_shm_setup(rank, size, allgather_cb, userdata))) break;
        double v[3] = { rank + 1.0, rank + 1.0, rank + 1.0 };
        fpm_comm_allreduce_double(MPI_COMM_WORLD, &v[0], 1, 0);
        fpm_comm_allreduce_double(MPI_COMM_WORLD, &v[1], 1, 1);
        fpm_comm_allreduce_double(MPI_COMM_WORLD, &v[2], 1, 2);
        if (v[0] != size * (size + 1) / 2.0 || v[1] != 1.0 || v[2] != (double) size) bad |= 1 << (4 * pass);
        int64_t n = 1000 + rank;
        fpm_comm_allreduce_i64(MPI_COMM_WORLD, &n, 1, 0);
        if (n != 1000 * (int64_t) size + size * (size - 1) / 2) bad |= 2 << (4 * pass);
        uint64_t mine = 0xabc00000ull + rank, all[MAXR];
        allgather(&mine, 8, all);
        for (int r = 0; r < size; r++) if (all[r] != 0xabc00000ull + r) bad |= 4 << (4 * pass);
        /* a payload of several slots (the P(k) bins of a large mesh), twice, so that both parities are reused */
        enum { NBIG = 20000 };
        double *big = (double *) malloc(NBIG * sizeof(double));
        for (int rep = 0; rep < 2; rep++) {
            for (int i = 0; i < NBIG; i++) big[i] = (double) (i % 97) * (rank + 1) + rep;
            fpm_comm_allreduce_double(MPI_COMM_WORLD, big, NBIG, 0);
            for (int i = 0; i < NBIG; i++) if (big[i] != (double) (i % 97) * (size * (size + 1) / 2) + rep * size) { bad |= 8 << (4 * pass); break; }
        }
        free(big);
    }
    g_callbacks_only = 0;
    /* the slab owner rule used by the migration kernel (pm_pos_to_rank, pmpfft.c:344-368) */
    g_rank = save_rank; g_size = save_size; g_allreduce = sa; g_allgather = sg; g_cb_data = sd;
    fpm_shm_setup(save_rank, save_size, sg, sd);
    return bad;
}

/* owner slab of an x coordinate for an Nmesh-cell box split over `size` slabs (host restatement for tests) */
int fastpm_b200_slab_owner(double x, double boxsize, int nmesh, int size)
{
    int ix = (int) floor(x * (1.0 / (boxsize / nmesh)));
    ix %= nmesh; if (ix < 0) ix += nmesh;
    return ix / (nmesh / size);
}
