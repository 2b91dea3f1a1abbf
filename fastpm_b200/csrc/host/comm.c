/* fastpm_b200 host layer -- communicator table standing in for MPI_Comm (one process per GPU).
 * Rank 0 / size 1 unless fastpm_b200_comm_init() joined the process to an x-slab decomposition; the
 * multi-GPU exchanges (FFT transpose through peer memory, halo planes, particle migration) are filled in
 * by the distributed build (see DESIGN.md section "multi-GPU"). */
#include "internal.h"

static int g_rank = 0, g_size = 1;

int fpm_comm_rank(MPI_Comm comm) { (void) comm; return g_rank; }
int fpm_comm_size(MPI_Comm comm) { (void) comm; return g_size; }

typedef void (*fpm_host_allreduce_fn)(void *buf, int count, int is_int64, int op, void *userdata);
static fpm_host_allreduce_fn g_allreduce = NULL;
static void *g_allreduce_data = NULL;

/* the launcher (bench.py under torchrun) installs the host all-reduce it already has (torch.distributed) */
void fastpm_b200_comm_init(int rank, int size, fpm_host_allreduce_fn allreduce, void *userdata)
{
    g_rank = rank; g_size = size; g_allreduce = allreduce; g_allreduce_data = userdata;
}

void fpm_comm_allreduce_double(MPI_Comm comm, double *v, int n, int op)
{
    (void) comm;
    if (g_size == 1) return;
    if (!g_allreduce) fastpm_raise(-1, "multi-rank run without an all-reduce callback\n");
    g_allreduce(v, n, 0, op, g_allreduce_data);
}
void fpm_comm_allreduce_i64(MPI_Comm comm, int64_t *v, int n, int op)
{
    (void) comm;
    if (g_size == 1) return;
    if (!g_allreduce) fastpm_raise(-1, "multi-rank run without an all-reduce callback\n");
    g_allreduce(v, n, 1, op, g_allreduce_data);
}
void fpm_comm_barrier(MPI_Comm comm) { int64_t z = 0; fpm_comm_allreduce_i64(comm, &z, 1, 0); }

void fpm_halo_add(PM *pm, FastPMFloat *canvas) { (void) pm; (void) canvas; fastpm_raise(-1, "multi-GPU halo exchange is not wired in this build\n"); }
void fpm_halo_fetch(PM *pm, FastPMFloat *canvas) { (void) pm; (void) canvas; fastpm_raise(-1, "multi-GPU halo exchange is not wired in this build\n"); }
void fpm_dist_r2c(PM *pm, FastPMFloat *real, FastPMFloat *cplx, double scale) { (void) pm; (void) real; (void) cplx; (void) scale; fastpm_raise(-1, "multi-GPU FFT is not wired in this build\n"); }
void fpm_dist_c2r(PM *pm, const FastPMFloat *cplx, FastPMFloat *real, const fpm_transfer *kernel) { (void) pm; (void) cplx; (void) real; (void) kernel; fastpm_raise(-1, "multi-GPU FFT is not wired in this build\n"); }

int fastpm_store_decompose(FastPMStore *p, fastpm_store_target_func target_func, void *data, MPI_Comm comm)
{
    (void) p; (void) target_func; (void) data;
    if (fpm_comm_size(comm) == 1) return 0;            /* one slab owns every particle */
    fastpm_raise(-1, "multi-GPU particle migration is not wired in this build\n");
    return -1;
}
