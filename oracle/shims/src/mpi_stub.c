/* ORACLE / TEST INFRASTRUCTURE ONLY.  One-rank MPI stand-in (see mpi.h). */
#include <mpi.h>
#include <string.h>
#include <stdlib.h>
#include <stdio.h>
#include <time.h>

#define MAX_DERIVED 256
static size_t derived_size[MAX_DERIVED];
static int derived_used[MAX_DERIVED];

static size_t type_size(MPI_Datatype t)
{
    if (t >= MPI_SHIM_DERIVED_BASE) return derived_size[t - MPI_SHIM_DERIVED_BASE];
    return (size_t)(t & 0xff);
}

int MPI_Init(int *argc, char ***argv) { (void) argc; (void) argv; return 0; }
int MPI_Finalize(void) { return 0; }
/* one rank; a test may ask the reference what it would do AS another rank where that only selects a branch or a seed (the rand
 * column of store.c:694-720): ref_set_fake_rank() */
static int fake_rank = 0;
void ref_set_fake_rank(int r) { fake_rank = r; }
int MPI_Comm_rank(MPI_Comm comm, int *rank) { (void) comm; *rank = fake_rank; return 0; }
int MPI_Comm_size(MPI_Comm comm, int *size) { (void) comm; *size = 1; return 0; }
int MPI_Comm_free(MPI_Comm *comm) { *comm = MPI_COMM_NULL; return 0; }
int MPI_Comm_dup(MPI_Comm comm, MPI_Comm *out) { *out = comm; return 0; }
int MPI_Cart_sub(MPI_Comm comm, const int remain_dims[], MPI_Comm *newcomm)
{ (void) remain_dims; *newcomm = comm; return 0; }
int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm *newcomm) { (void) color; (void) key; *newcomm = comm; return 0; }
int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype t, MPI_Op op, int root, MPI_Comm comm)
{
    (void) op; (void) root; (void) comm;
    if (sendbuf != MPI_IN_PLACE) memcpy(recvbuf, sendbuf, count * type_size(t));
    return 0;
}
int MPI_Gather(const void *sendbuf, int sendcount, MPI_Datatype st, void *recvbuf, int recvcount, MPI_Datatype rt, int root, MPI_Comm comm)
{
    (void) recvcount; (void) rt; (void) root; (void) comm;
    if (sendbuf != MPI_IN_PLACE) memcpy(recvbuf, sendbuf, sendcount * type_size(st));
    return 0;
}
int MPI_Gatherv(const void *sendbuf, int sendcount, MPI_Datatype st, void *recvbuf, const int *recvcounts, const int *displs,
                MPI_Datatype rt, int root, MPI_Comm comm)
{
    (void) recvcounts; (void) root; (void) comm;
    if (sendbuf != MPI_IN_PLACE) memcpy((char *) recvbuf + displs[0] * type_size(rt), sendbuf, sendcount * type_size(st));
    return 0;
}
int MPI_Scatterv(const void *sendbuf, const int *sendcounts, const int *displs, MPI_Datatype st, void *recvbuf, int recvcount,
                 MPI_Datatype rt, int root, MPI_Comm comm)
{
    (void) recvcount; (void) rt; (void) root; (void) comm;
    if (recvbuf != MPI_IN_PLACE) memcpy(recvbuf, (const char *) sendbuf + displs[0] * type_size(st), sendcounts[0] * type_size(st));
    return 0;
}
int MPI_Barrier(MPI_Comm comm) { (void) comm; return 0; }
int MPI_Bcast(void *buf, int count, MPI_Datatype t, int root, MPI_Comm comm)
{ (void) buf; (void) count; (void) t; (void) root; (void) comm; return 0; }

int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype t, MPI_Op op, MPI_Comm comm)
{
    (void) op; (void) comm;
    if (sendbuf != MPI_IN_PLACE) memcpy(recvbuf, sendbuf, count * type_size(t));
    return 0;
}
int MPI_Allgather(const void *sendbuf, int sendcount, MPI_Datatype st, void *recvbuf, int recvcount, MPI_Datatype rt, MPI_Comm comm)
{
    (void) recvcount; (void) rt; (void) comm;
    if (sendbuf != MPI_IN_PLACE) memcpy(recvbuf, sendbuf, sendcount * type_size(st));
    return 0;
}
int MPI_Alltoall(const void *sendbuf, int sendcount, MPI_Datatype st, void *recvbuf, int recvcount, MPI_Datatype rt, MPI_Comm comm)
{
    (void) recvcount; (void) rt; (void) comm;
    memcpy(recvbuf, sendbuf, sendcount * type_size(st));
    return 0;
}
int MPI_Alltoallv(const void *sendbuf, const int *sendcounts, const int *sdispls, MPI_Datatype st,
                  void *recvbuf, const int *recvcounts, const int *rdispls, MPI_Datatype rt, MPI_Comm comm)
{
    (void) recvcounts; (void) comm;
    memcpy((char *) recvbuf + rdispls[0] * type_size(rt),
           (const char *) sendbuf + sdispls[0] * type_size(st),
           sendcounts[0] * type_size(st));
    return 0;
}
int MPI_Sendrecv(const void *sendbuf, int sendcount, MPI_Datatype st, int dest, int sendtag,
                 void *recvbuf, int recvcount, MPI_Datatype rt, int source, int recvtag,
                 MPI_Comm comm, MPI_Status *status)
{
    (void) dest; (void) sendtag; (void) recvcount; (void) rt; (void) source; (void) recvtag; (void) comm; (void) status;
    memcpy(recvbuf, sendbuf, sendcount * type_size(st));
    return 0;
}

/* self send/recv pairing: one pending slot per tag is plenty for a 1-rank run */
static struct { const void *buf; size_t n; int tag; int live; } pending_send[64];
static struct { void *buf; size_t n; int tag; int live; } pending_recv[64];
static void try_match(void)
{
    for (int i = 0; i < 64; i++) if (pending_send[i].live)
        for (int j = 0; j < 64; j++) if (pending_recv[j].live && pending_recv[j].tag == pending_send[i].tag) {
            size_t n = pending_send[i].n < pending_recv[j].n ? pending_send[i].n : pending_recv[j].n;
            memcpy(pending_recv[j].buf, pending_send[i].buf, n);
            pending_send[i].live = pending_recv[j].live = 0;
            break;
        }
}
int MPI_Isend(const void *buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm comm, MPI_Request *req)
{
    (void) dest; (void) comm;
    for (int i = 0; i < 64; i++) if (!pending_send[i].live) {
        pending_send[i].buf = buf; pending_send[i].n = count * type_size(t);
        pending_send[i].tag = tag; pending_send[i].live = 1; break;
    }
    *req = 1; try_match(); return 0;
}
int MPI_Irecv(void *buf, int count, MPI_Datatype t, int source, int tag, MPI_Comm comm, MPI_Request *req)
{
    (void) source; (void) comm;
    for (int i = 0; i < 64; i++) if (!pending_recv[i].live) {
        pending_recv[i].buf = buf; pending_recv[i].n = count * type_size(t);
        pending_recv[i].tag = tag; pending_recv[i].live = 1; break;
    }
    *req = 2; try_match(); return 0;
}
int MPI_Waitall(int count, MPI_Request reqs[], MPI_Status statuses[])
{ (void) count; (void) reqs; (void) statuses; try_match(); return 0; }

int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype *newtype)
{
    for (int i = 0; i < MAX_DERIVED; i++) if (!derived_used[i]) {
        derived_used[i] = 1; derived_size[i] = count * type_size(oldtype);
        *newtype = MPI_SHIM_DERIVED_BASE + i; return 0;
    }
    fprintf(stderr, "mpi shim: out of derived datatypes\n"); abort();
}
int MPI_Type_commit(MPI_Datatype *t) { (void) t; return 0; }
int MPI_Type_free(MPI_Datatype *t)
{
    if (*t >= MPI_SHIM_DERIVED_BASE) derived_used[*t - MPI_SHIM_DERIVED_BASE] = 0;
    *t = 0; return 0;
}
int MPI_Type_get_extent(MPI_Datatype t, MPI_Aint *lb, MPI_Aint *extent)
{ *lb = 0; *extent = (MPI_Aint) type_size(t); return 0; }
int MPI_Abort(MPI_Comm comm, int code) { (void) comm; fprintf(stderr, "MPI_Abort(%d)\n", code); abort(); }
double MPI_Wtime(void)
{
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
