/* ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product.
 *
 * Single-rank MPI stand-in so that the reference's own libfastpm sources
 * (/root/reference/libfastpm/*.c) compile unmodified without an MPI install.
 * Only the entry points the hot path links against are provided
 * (list measured with `nm -u`).  Semantics: a
 * communicator always has exactly one rank, rank id 0.
 */
#ifndef ORACLE_SHIM_MPI_H
#define ORACLE_SHIM_MPI_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
typedef ptrdiff_t MPI_Aint;

#define MPI_COMM_WORLD 1
#define MPI_COMM_NULL 0
#define MPI_SUCCESS 0
#define MPI_IN_PLACE ((void *) 1)
#define MPI_STATUS_IGNORE ((MPI_Status *) 0)
#define MPI_STATUSES_IGNORE ((MPI_Status *) 0)

/* basic datatypes: handle value == size in bytes + tag in the high bits */
#define MPI_BYTE      0x101
#define MPI_CHAR      0x201
#define MPI_INT       0x304
#define MPI_FLOAT     0x404
#define MPI_LONG      0x508
#define MPI_LONG_LONG 0x608
#define MPI_DOUBLE    0x708
#define MPI_DOUBLE_INT 0x810   /* struct {double; int;} padded to 16 */
#define MPI_UNSIGNED_LONG 0x908
#define MPI_UNSIGNED  0xa04
#define MPI_DATATYPE_NULL 0
#define MPI_SHIM_DERIVED_BASE 0x10000

enum { MPI_SUM = 1, MPI_MIN, MPI_MAX, MPI_LAND, MPI_LOR, MPI_MINLOC, MPI_MAXLOC };

int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Comm_free(MPI_Comm *comm);
int MPI_Comm_dup(MPI_Comm comm, MPI_Comm *out);
int MPI_Cart_sub(MPI_Comm comm, const int remain_dims[], MPI_Comm *newcomm);
int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm *newcomm);
int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype t, MPI_Op op, int root, MPI_Comm comm);
int MPI_Gather(const void *sendbuf, int sendcount, MPI_Datatype st, void *recvbuf, int recvcount, MPI_Datatype rt, int root, MPI_Comm comm);
int MPI_Gatherv(const void *sendbuf, int sendcount, MPI_Datatype st, void *recvbuf, const int *recvcounts, const int *displs,
                MPI_Datatype rt, int root, MPI_Comm comm);
int MPI_Scatterv(const void *sendbuf, const int *sendcounts, const int *displs, MPI_Datatype st, void *recvbuf, int recvcount,
                 MPI_Datatype rt, int root, MPI_Comm comm);
int MPI_Barrier(MPI_Comm comm);
int MPI_Bcast(void *buf, int count, MPI_Datatype t, int root, MPI_Comm comm);
int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype t, MPI_Op op, MPI_Comm comm);
int MPI_Allgather(const void *sendbuf, int sendcount, MPI_Datatype st, void *recvbuf, int recvcount, MPI_Datatype rt, MPI_Comm comm);
int MPI_Alltoall(const void *sendbuf, int sendcount, MPI_Datatype st, void *recvbuf, int recvcount, MPI_Datatype rt, MPI_Comm comm);
int MPI_Alltoallv(const void *sendbuf, const int *sendcounts, const int *sdispls, MPI_Datatype st,
                  void *recvbuf, const int *recvcounts, const int *rdispls, MPI_Datatype rt, MPI_Comm comm);
int MPI_Sendrecv(const void *sendbuf, int sendcount, MPI_Datatype st, int dest, int sendtag,
                 void *recvbuf, int recvcount, MPI_Datatype rt, int source, int recvtag,
                 MPI_Comm comm, MPI_Status *status);
int MPI_Isend(const void *buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Irecv(void *buf, int count, MPI_Datatype t, int source, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Waitall(int count, MPI_Request reqs[], MPI_Status statuses[]);
int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype *newtype);
int MPI_Type_commit(MPI_Datatype *t);
int MPI_Type_free(MPI_Datatype *t);
int MPI_Type_get_extent(MPI_Datatype t, MPI_Aint *lb, MPI_Aint *extent);
int MPI_Abort(MPI_Comm comm, int code);
double MPI_Wtime(void);

#ifdef __cplusplus
}
#endif
#endif
